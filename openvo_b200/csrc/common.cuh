// Shared declarations for the openvo_b200 CUDA kernels (sm_100a only).
#pragma once
#ifdef OVO_EMU
#include "cuda_emu.h"  // tests/emu: CPU execution of the kernels for the no-GPU test tier (never the product path)
#else
#include <cuda_runtime.h>
#define OVO_DYN_SMEM(type, name)                                  \
    extern __shared__ __align__(16) unsigned char name##_raw[]; \
    type* name = reinterpret_cast<type*>(name##_raw)
namespace ovo {
void prof_pre(const char* tag, cudaStream_t st);   // counts the launch; records a start event when profiling is on
void prof_post(cudaStream_t st);
}
#define OVO_LAUNCH(kern, grid, block, smem, st, ...)          \
    do {                                                      \
        ovo::prof_pre(#kern, st);                             \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);         \
        ovo::prof_post(st);                                   \
    } while (0)
#endif
#include <cstdint>
#include <cstddef>
#include <cstdio>

// Debug build (-DOVO_BOUNDS, openvo_b200.build.build_variant("bounds", ["OVO_BOUNDS"])): every computed shared / global
// offset of the SGBM kernels is checked on the device; a violation prints its site and traps (the stream then reports an error).
// The GPU parity tests are run once per round against that variant (tools/bounds_check.sh) in place of compute-sanitizer.
#if defined(OVO_BOUNDS) && defined(OVO_EMU)
#include <cstdlib>
#define OVO_DEVCHECK(cond)                                                         \
    do {                                                                           \
        if (!(cond)) {                                                             \
            fprintf(stderr, "OVO_BOUNDS %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::abort();                                                          \
        }                                                                          \
    } while (0)
#elif defined(OVO_BOUNDS)
#define OVO_DEVCHECK(cond)                                                                                         \
    do {                                                                                                           \
        if (!(cond)) {                                                                                             \
            printf("OVO_BOUNDS %s:%d block (%d,%d,%d) thread %d: %s\n", __FILE__, __LINE__, (int)blockIdx.x, (int)blockIdx.y, \
                   (int)blockIdx.z, (int)threadIdx.x, #cond);                                                      \
            __trap();                                                                                              \
        }                                                                                                          \
    } while (0)
#else
#define OVO_DEVCHECK(cond) ((void)0)
#endif

namespace ovo {

// ---- error plumbing -------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define OVO_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ovo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)
#define OVO_LAUNCH_CHECK() OVO_CUDA(cudaGetLastError())

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- StereoSGBM ------------------------------------------------------------------------------------------
// Parameters after OpenCV's own normalisation: P2 = max(P2, P1+1), disp12 = d12>0 ? d12 : 1,
// ftzero = max(preFilterCap, 15) | 1.  Dp = D rounded up to 64/128/256 (one warp owns one cost vector).
struct SgbmDims {
    int W, H, D, Dp, W1;
    int bs, P1, P2, uniq, disp12, ftzero, speckleWin, speckleDiff;
    int mode;  // 0 = MODE_SGBM (5 directions, what the reference uses); 1 = MODE_HH (8 directions, opt-in extension)
};

struct SgbmWorkspace {          // per frame, device pointers
    uint32_t* prep;             // [2 img][2 type][H][W]  byte-packed (v, vmin, vmax, 0)
    int16_t* C;                 // [H][W1][Dp] aggregated BT cost
    int16_t* Lv;                // [3 or 6][H][W1][Dp] paths from (x-1,y-1), (x,y-1), (x+1,y-1) (+ the three from row y+1 in MODE_HH)
    int16_t* ckpt;              // [H][2][ceil((W1-W1/2)/K)][Dp] state of the two horizontal sweeps every K cells
    int16_t* raw;               // [H][W] disparity after WTA + LR check
    int16_t* med;               // [H][W] after median
    int32_t* label;             // [H][W] speckle CCL labels
    int32_t* csize;             // [H][W] component sizes
};

size_t sgbm_workspace_bytes(const SgbmDims& d);
void sgbm_carve(const SgbmDims& d, uint8_t* base, SgbmWorkspace* ws);
// left/right: device u8 [nb][H][pitch]; disp_out: device i16 [nb][H][W]; ws0: frame 0's workspace, frame f's is
// ws_stride*f bytes further
int sgbm_launch(const SgbmDims& d, const SgbmWorkspace* ws0, size_t ws_stride, int nb, const uint8_t* left, const uint8_t* right,
                int pitch, size_t frame_stride, int16_t* disp_out, cudaStream_t st);

// ---- ORB -------------------------------------------------------------------------------------------------
const int ORB_NLEVELS = 8;
struct OrbLevel {
    int w, h;          // level size
    int off;           // pixel offset of this level inside a concatenated pyramid plane
    int nfeat;         // feature budget
    float scale, inv;  // scale_l, 1/scale_l (float32)
    int row_off;       // offset of this level's first row inside per-row arrays
    int mw, moff;      // candidate bit mask: 32-bit words per row, word offset of the level
};
struct OrbDims {
    int W, H, nfeatures;
    int total_px, total_rows, total_mwords, cand_cap, kp_cap;
    OrbLevel lv[ORB_NLEVELS];
    int fast_tiles[ORB_NLEVELS + 1], blur_tiles[ORB_NLEVELS + 1];  // first flat tile index of each level (+ total)
};
void orb_make_dims(int W, int H, int nfeatures, OrbDims* d);

struct OrbWorkspace {           // per frame, device pointers
    uint8_t *pyr, *maskpyr, *score, *blur;  // [total_px] each
    int32_t* row_offset;                     // [total_rows] candidates before the row, over all levels
    uint32_t* candmask;                      // [total_mwords] one bit per pixel: NMS + border + mask survivor
    int32_t* lvl_count;                      // [64]: [0..7] candidates per level, [8..15] level offsets, [16] total; survivors of the
                                             // first retainBest (FAST score >= the level's boundary score): [17..24] per level,
                                             // [25..32] offsets, [33] total, [34..41] the levels' boundary scores
    int32_t* cand_xy;                        // [cand_cap] packed (y<<16 | x)
    uint8_t* cand_score;                     // [cand_cap] FAST score
    float* cand_harris;                      // [cand_cap] Harris response (written for survivors only)
    float* harris_dense;                     // [cand_cap] Harris responses of the survivors, in candidate order (what the host reads)
    int32_t* surv_id;                        // [cand_cap] candidate ids of the survivors, in candidate order
    int32_t* sel;                            // [kp_cap] selected candidate ids (level-major, final order)
};
size_t orb_workspace_bytes(const OrbDims& d);
void orb_carve(const OrbDims& d, uint8_t* base, OrbWorkspace* ws);
// host-side exact KeyPointsFilter::retainBest emulation (libstdc++ introselect order), host_select.cpp
// lvl_count: [64] as above; scores: FAST score of every candidate; harris_dense: Harris response of the survivors;
// out_sel: [kp_cap]; returns number selected (-1 on capacity overflow, -2 if the device's survivor set disagrees)
int orb_host_select(const OrbDims& d, const int32_t* lvl_count, const uint8_t* scores, const float* harris_dense, int32_t* out_sel);

// ---- matcher / pose ----------------------------------------------------------------------------------------
// nn_out [nq][4] = (idx0, d0, idx1, d1), ties -> lowest train index
int knn2_launch(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* nn_out, uint32_t* scratch, cudaStream_t st);
size_t knn2_scratch_bytes(int nq_cap, int nt_cap);

struct GatherParams {
    double Q[16];
    double thr;
    int roi_x0, roi_y0;  // crop origin in the full image (bug-compatible B1 slices)
    int cw, ch;          // cropped size
    int disp_pitch;      // elements
};
// ratio test + ordered compaction + fused reproject/bilinear lookup.  matches_out [nq][3] (q, t, dist),
// pts1/pts2 [nq][3] f32, counts_out[0] = number of matches, counts_out[1] = lookups with no usable tap
int match_gather_launch(const GatherParams& p, const int32_t* nn, int nq, const float* kp1, const float* kp2,
                        const float* disp1, const float* disp2, int32_t* matches_out, float* pts1, float* pts2,
                        int32_t* counts_out, const int32_t* nn_rev, cudaStream_t st);
// Umeyama with OpenCV's scale leak.  m_dev: device count; out [16] f64: 3x4 [R|t], scale, angle,
// |t|, flag(ok=1)
int umeyama_launch(const float* pts1, const float* pts2, const int32_t* m_dev, int m_cap, double* out, cudaStream_t st);

}  // namespace ovo
