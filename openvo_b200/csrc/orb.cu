// ORB (FAST-9/16 + NMS, Harris ranking, intensity-centroid orientation, rBRIEF-256 over an 8-level pyramid) for
// sm_100a — stages 1+2 of the openVO hot path.
//
// Replaces cv2.ORB_create(nfeatures).detectAndCompute(img, mask) as called by the reference at
// src/openVO/stereo_odometer.py:117 (object built at :22).  Semantics: SURVEY.md Appendix A.1-A.2 (bit-exact incl.
// keypoint order; oracle = oracle/orb_restate.cpp).
//
// Phase 1 (device): pyramid (INTER_LINEAR_EXACT, chained) of image and mask -> FAST score plane -> NMS + mask + border
//   test, emitted in raster order by a count / scan / emit triple -> Harris response of every candidate -> float32
//   separable Gaussian with OpenCV's AVX2 FMA-body / scalar-tail split.
// Host: the two KeyPointsFilter::retainBest passes (libstdc++ introselect permutation, host_select.cpp).
// Phase 2 (device): one warp per kept keypoint: IC angle (integer moments, warp reduction) and the 256 rBRIEF tests.
// All float arithmetic uses explicit __f*_rn intrinsics so that nvcc contracts nothing that OpenCV did not.
#include "common.cuh"
#include <cmath>

namespace ovo {

namespace {

constexpr int kEdge = 31, kFastT = 20, kHalfPatch = 15;

// u_max of the circular patch per |v|: {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3}, packed 4 bits each below
// global (not __constant__) memory: it is copied to shared memory with coalesced word loads; constant memory would serialise the
// 32 different addresses of a warp
__device__ __align__(16) const signed char c_pattern[1024] = {
#include "orb_pattern.inc"
};
// 7-tap sigma=2 Gaussian, float32: k0..k3 (symmetric)
__constant__ float c_gk[4] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95c0p-3f};

template <typename T>
__device__ __forceinline__ T* fptr(T* p, size_t stride_bytes, int f) {
    return (T*)((const uint8_t*)p + stride_bytes * (size_t)f);
}

// ---- pyramid ------------------------------------------------------------------------------------------------------
// tab: per destination index (src index | c1 << 16), c1 = weight of src[index+1] in 1/256
__global__ void __launch_bounds__(256) k_orb_resize(OrbDims d, OrbWorkspace ws, size_t ws_stride, int level,
                                                    const int32_t* __restrict__ xtab, const int32_t* __restrict__ ytab, int has_mask) {
    // a thread produces 4 horizontally adjacent pixels of one row: one division and one row-table look-up per thread, and the
    // source bytes of neighbouring outputs come from the same cache lines
    const OrbLevel L = d.lv[level], S = d.lv[level - 1];
    const int f = blockIdx.y, W4 = (L.w + 3) >> 2;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= W4 * L.h) return;
    const int y = t / W4, x0 = (t - y * W4) * 4;
    const int ty = ytab[y];
    const int j0 = ty & 0xFFFF, cy = ty >> 16, j1 = min(j0 + 1, S.h - 1);
    int i0[4], i1[4], cx[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int tx = xtab[min(x0 + e, L.w - 1)];
        i0[e] = tx & 0xFFFF; cx[e] = tx >> 16; i1[e] = min(i0[e] + 1, S.w - 1);
    }
    for (int pl = 0; pl < 1 + has_mask; pl++) {
        uint8_t* base = fptr(pl ? ws.maskpyr : ws.pyr, ws_stride, f);
        const uint8_t* r0 = base + S.off + (size_t)j0 * S.w;
        const uint8_t* r1 = base + S.off + (size_t)j1 * S.w;
        uint8_t* out = base + L.off + (size_t)y * L.w + x0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (x0 + e >= L.w) break;
            const uint32_t h0 = r0[i0[e]] * (256 - cx[e]) + r0[i1[e]] * cx[e];
            const uint32_t h1 = r1[i0[e]] * (256 - cx[e]) + r1[i1[e]] * cx[e];
            uint32_t v = (h0 * (256 - cy) + h1 * cy + (1u << 15)) >> 16;
            if (pl) v = v > 254 ? v : 0;  // THRESH_TOZERO(254) on the mask levels
            out[e] = (uint8_t)v;
        }
    }
}

__global__ void __launch_bounds__(256) k_orb_copy_level0(OrbDims d, OrbWorkspace ws, size_t ws_stride, const uint8_t* __restrict__ img,
                                                         int pitch, size_t frame_stride, const uint8_t* __restrict__ mask, int mask_pitch,
                                                         size_t mask_frame_stride) {
    const int f = blockIdx.y, n = d.W * d.H;
#pragma unroll
    for (int e = 0; e < 4; e++) {  // 4 pixels per thread: 1024 consecutive pixels per CTA
        const int i = (blockIdx.x * 4 + e) * 256 + threadIdx.x;
        if (i >= n) break;
        const int y = i / d.W, x = i - y * d.W;
        fptr(ws.pyr, ws_stride, f)[i] = img[frame_stride * f + (size_t)y * pitch + x];
        if (mask) fptr(ws.maskpyr, ws_stride, f)[i] = mask[mask_frame_stride * f + (size_t)y * mask_pitch + x];
    }
}

// ---- FAST-9/16 score ----------------------------------------------------------------------------------------------------
constexpr int kFastTX = 32, kFastTY = 16;  // tile of the fused FAST + NMS kernel (one 32-bit candidate word per tile row)

// FAST-9/16 score of the pixel at (cx, cy) of a shared-memory tile: the largest threshold for which it is still a corner
// = max over the 16 arcs of (min over 9 consecutive ring differences), for the brighter and the darker polarity, minus 1;
// 0 when that is below the detection threshold.  Both polarities ride in one register: with u = centre - ring + 256 (> 0)
// the word u * (1 - 2^16) + (512 << 16) holds (centre - ring) + 256 in its low half and (ring - centre) + 256 in its high
// half, so one multiply-add per ring pixel packs them and the arc minima / maxima are packed unsigned 16-bit operations.
__device__ __forceinline__ int fast_score(const uint8_t (*tile)[kFastTX + 8], int cx, int cy) {
    const uint32_t ck = (uint32_t)(tile[cy][cx] + 256) * 0xFFFF0001u + (512u << 16);
    auto pk = [&](int dy, int dx) -> uint32_t { return (uint32_t)tile[cy + dy][cx + dx] * 0x0000FFFFu + ck; };
    uint32_t w[16];
    w[0] = pk(3, 0);   w[1] = pk(3, 1);   w[2] = pk(2, 2);    w[3] = pk(1, 3);
    w[4] = pk(0, 3);   w[5] = pk(-1, 3);  w[6] = pk(-2, 2);   w[7] = pk(-3, 1);
    w[8] = pk(-3, 0);  w[9] = pk(-3, -1); w[10] = pk(-2, -2); w[11] = pk(-1, -3);
    w[12] = pk(0, -3); w[13] = pk(1, -3); w[14] = pk(2, -2);  w[15] = pk(3, -1);
    uint32_t m2[16], m4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m2[k] = __vminu2(w[k], w[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) m4[k] = __vminu2(m2[k], m2[(k + 2) & 15]);
    uint32_t best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {  // min over ring[k .. k+8] = min(m4[k], m4[k+4], ring[k+8])
        const uint32_t a = __vimin3_u16x2(m4[k], m4[(k + 4) & 15], w[(k + 8) & 15]);
        const uint32_t b = __vimin3_u16x2(m4[k + 1], m4[(k + 5) & 15], w[(k + 9) & 15]);
        best = __vimax3_u16x2(best, a, b);
    }
    const int m = (int)max(best & 0xFFFFu, best >> 16) - 256;
    return m > kFastT ? m - 1 : 0;
}

// Tiles of every level are numbered consecutively (level-major, then row-major), so one grid covers the whole pyramid
// without empty CTAs.  OrbDims carries the per-level prefix for the two tile shapes in use (FAST, blur).
__device__ __forceinline__ void orb_flat_tile(const int (&off)[ORB_NLEVELS + 1], const OrbDims& d, int TX, int t, int& level, int& tx, int& ty) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < ORB_NLEVELS; k++) l += (t >= off[k]) ? 1 : 0;
    const int nx = (d.lv[l].w + TX - 1) / TX;
    t -= off[l];
    level = l;
    ty = t / nx;
    tx = t - ty * nx;
}

// FAST score + 3x3 NMS + border + mask, fused per 32x16 tile.  Every pixel of the tile plus a one-pixel apron is scored
// (dense, no corner queue: with packed arithmetic the score costs less than a separate corner test); then a pixel is a
// candidate when its score strictly beats its 8 neighbours, it lies inside the 31-pixel border and the mask is set.
// Output: one bit per pixel (a 32-bit word per tile row, bit i = column x0 + i) and the score byte of each candidate.
__global__ void __launch_bounds__(256) k_orb_fast_nms(OrbDims d, OrbWorkspace ws, size_t ws_stride, int has_mask) {
    constexpr int TX = kFastTX, TY = kFastTY, SW = TX + 2, SH = TY + 2;
    __shared__ uint8_t tile[TY + 8][kFastTX + 8];
    __shared__ uint8_t sc[SH][SW + 2];
    int level, bx, by;
    orb_flat_tile(d.fast_tiles, d, TX, blockIdx.x, level, bx, by);
    const int f = blockIdx.y;
    const OrbLevel L = d.lv[level];
    const int x0 = bx * TX, y0 = by * TY;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* cm = fptr(ws.candmask, ws_stride, f) + L.moff;
    // candidates live in [kEdge, w - kEdge) x [kEdge, h - kEdge): tiles outside only clear their words
    const bool live = x0 < L.w - kEdge && x0 + TX > kEdge && y0 < L.h - kEdge && y0 + TY > kEdge;
    if (!live) {
        if (tid < TY && y0 + tid < L.h) cm[(size_t)(y0 + tid) * L.mw + bx] = 0u;
        return;
    }
    const uint8_t* img = fptr(ws.pyr, ws_stride, f) + L.off;
    if (x0 >= 4 && x0 + TX + 4 <= L.w && y0 >= 4 && y0 + TY + 4 <= L.h) {  // interior tile: no clamping
        const uint8_t* base = img + (size_t)(y0 - 4) * L.w + (x0 - 4);
        for (int i = tid; i < (TY + 8) * (TX + 8); i += 256) {
            const int ty = i / (TX + 8), tx = i - ty * (TX + 8);
            tile[ty][tx] = base[ty * L.w + tx];
        }
    } else {
        for (int i = tid; i < (TY + 8) * (TX + 8); i += 256) {
            const int ty = i / (TX + 8), tx = i - ty * (TX + 8);
            const int gx = min(max(x0 + tx - 4, 0), L.w - 1), gy = min(max(y0 + ty - 4, 0), L.h - 1);
            tile[ty][tx] = img[(size_t)gy * L.w + gx];
        }
    }
    __syncthreads();
    // scores outside [kEdge-1, w-kEdge] x [kEdge-1, h-kEdge] are never read by a candidate: left at 0
    for (int i = tid; i < SW * SH; i += 256) {
        const int sy = i / SW, sx = i - sy * SW;
        const int x = x0 - 1 + sx, y = y0 - 1 + sy;
        int v = 0;
        if (x >= kEdge - 1 && x <= L.w - kEdge && y >= kEdge - 1 && y <= L.h - kEdge) v = fast_score(tile, sx + 3, sy + 3);
        sc[sy][sx] = (uint8_t)v;
    }
    __syncthreads();
    const uint8_t* mk = has_mask ? fptr(ws.maskpyr, ws_stride, f) + L.off : nullptr;
    uint8_t* out = fptr(ws.score, ws_stride, f) + L.off;
#pragma unroll
    for (int rr = 0; rr < TY / 8; rr++) {
        const int ty = wid + rr * 8, x = x0 + lane, y = y0 + ty;
        if (y >= L.h) continue;  // warp-uniform
        bool c = false;
        const int s = sc[ty + 1][lane + 1];
        if (s != 0 && x >= kEdge && x < L.w - kEdge && y >= kEdge && y < L.h - kEdge) {
            const uint8_t *r0 = &sc[ty][lane], *r1 = &sc[ty + 1][lane], *r2 = &sc[ty + 2][lane];
            c = r0[0] < s && r0[1] < s && r0[2] < s && r1[0] < s && r1[2] < s && r2[0] < s && r2[1] < s && r2[2] < s;
            if (c && mk) c = mk[(size_t)y * L.w + x] != 0;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, c);
        if (lane == 0) cm[(size_t)y * L.mw + bx] = bal;
        if (c) out[(size_t)y * L.w + x] = (uint8_t)s;
    }
}

// one CTA per frame: candidates per row (popcount of the row's mask words) and their exclusive scan over all rows of all
// levels (rows are numbered level-major, so raster order inside a level is kept).  row_offset[r] = candidates before row r;
// lvl_count[0..7] = per-level totals, lvl_count[8..15] = level base offsets, lvl_count[16] = total
__global__ void __launch_bounds__(1024) k_orb_scan(OrbDims d, OrbWorkspace ws, size_t ws_stride) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int f = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t* cm = fptr(ws.candmask, ws_stride, f);
    int32_t* ro = fptr(ws.row_offset, ws_stride, f);
    int32_t* lvl = fptr(ws.lvl_count, ws_stride, f);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int r0 = 0; r0 < d.total_rows; r0 += 1024) {
        const int r = r0 + threadIdx.x;
        int v = 0;
        if (r < d.total_rows) {
            int l = 0;
#pragma unroll
            for (int k = 1; k < ORB_NLEVELS; k++) l += (r >= d.lv[k].row_off) ? 1 : 0;
            const OrbLevel L = d.lv[l];
            const int y = r - L.row_off;
            if (y >= kEdge && y < L.h - kEdge) {
                const uint32_t* w = cm + L.moff + (size_t)y * L.mw;
                for (int i = 0; i < L.mw; i++) v += __popc(w[i]);
            }
        }
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int before = carry + (wid ? warp_sums[wid - 1] : 0) + incl - v;
        if (r < d.total_rows) ro[r] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x < ORB_NLEVELS) {
        const int l = threadIdx.x;
        const int b0 = ro[d.lv[l].row_off], b1 = l + 1 < ORB_NLEVELS ? ro[d.lv[l + 1].row_off] : carry;
        lvl[l] = b1 - b0;
        lvl[ORB_NLEVELS + l] = b0;
    }
    if (threadIdx.x == 0) lvl[2 * ORB_NLEVELS] = carry;
}

// candidates in raster order: one warp per row walks the row's mask words
__global__ void __launch_bounds__(256) k_orb_emit(OrbDims d, OrbWorkspace ws, size_t ws_stride) {
    const int lane = threadIdx.x & 31, f = blockIdx.y;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= d.total_rows) return;
    int l = 0;
#pragma unroll
    for (int k = 1; k < ORB_NLEVELS; k++) l += (r >= d.lv[k].row_off) ? 1 : 0;
    const OrbLevel L = d.lv[l];
    const int y = r - L.row_off;
    if (y < kEdge || y >= L.h - kEdge) return;
    const uint32_t* w = fptr(ws.candmask, ws_stride, f) + L.moff + (size_t)y * L.mw;
    const uint8_t* sc = fptr(ws.score, ws_stride, f) + L.off + (size_t)y * L.w;
    int32_t* cxy = fptr(ws.cand_xy, ws_stride, f);
    uint8_t* cscore = fptr(ws.cand_score, ws_stride, f);
    int base = fptr(ws.row_offset, ws_stride, f)[r];
    for (int i0 = 0; i0 < L.mw; i0 += 32) {
        uint32_t bits = i0 + lane < L.mw ? w[i0 + lane] : 0u;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int idx = base + incl - cnt;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int x = (i0 + lane) * 32 + b;
            if (idx < d.cand_cap) {
                cxy[idx] = (y << 16) | x;
                cscore[idx] = sc[x];
            }
            idx++;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__device__ __forceinline__ int cand_level(const int32_t* lvl, int i) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < ORB_NLEVELS; k++) l += (i >= lvl[ORB_NLEVELS + k]) ? 1 : 0;
    return l;
}

// ---- survivors of the first retainBest pass (A.1.5) -----------------------------------------------------------------------
// KeyPointsFilter::retainBest(2 * n_l) on the FAST score keeps every candidate whose score reaches the (2 n_l)-th largest one
// (all ties at the boundary stay).  That SET needs no ordering: a 256-bin histogram of the 8-bit scores gives the boundary,
// an ordered compaction the survivors' ids.  Only they get a Harris response, and only their responses travel to the host,
// which still produces the survivors' ORDER (libstdc++'s introselect permutation, host_select.cpp).
// Pass A (one CTA per frame and level): boundary score of the level and the number of survivors.
__global__ void __launch_bounds__(1024) k_orb_survivors_count(OrbDims d, OrbWorkspace ws, size_t ws_stride) {
    __shared__ int hist[256];
    __shared__ int s_amb;
    const int f = blockIdx.y, l = blockIdx.x;
    int32_t* lvl = fptr(ws.lvl_count, ws_stride, f);
    const uint8_t* score = fptr(ws.cand_score, ws_stride, f);
    const int base = lvl[ORB_NLEVELS + l], n = min(lvl[l], max(d.cand_cap - base, 0)), keep = 2 * d.lv[l].nfeat;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    if (n > keep && keep > 0)
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[score[base + i]], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int amb = 0, cnt = n;             // n <= keep: nothing is cut
        if (keep == 0) { amb = 256; cnt = 0; }  // retainBest(0): nothing survives
        else if (n > keep) {
            cnt = 0;
            for (amb = 255; amb > 0; amb--) {
                cnt += hist[amb];
                if (cnt >= keep) break;
            }
            if (amb == 0) cnt += hist[0];
        }
        s_amb = amb;
        lvl[17 + l] = cnt;
        lvl[34 + l] = amb;
    }
}

// Pass B (one CTA per frame and level): ordered compaction of the level's survivors behind those of the lower levels.
__global__ void __launch_bounds__(1024) k_orb_survivors(OrbDims d, OrbWorkspace ws, size_t ws_stride) {
    __shared__ int warp_sums[32];
    const int f = blockIdx.y, l = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t* lvl = fptr(ws.lvl_count, ws_stride, f);
    const uint8_t* score = fptr(ws.cand_score, ws_stride, f);
    int32_t* surv = fptr(ws.surv_id, ws_stride, f);
    const int base = lvl[ORB_NLEVELS + l], n = min(lvl[l], max(d.cand_cap - base, 0)), amb = lvl[34 + l];
    int out0 = 0;
    for (int k = 0; k < l; k++) out0 += lvl[17 + k];
    int done = 0;  // survivors of the chunks before this one
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool p = i < n && (int)score[base + i] >= amb;
        const uint32_t bal = __ballot_sync(0xffffffffu, p);
        if (lane == 0) warp_sums[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            int s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        if (p) surv[out0 + done + (wid ? warp_sums[wid - 1] : 0) + __popc(bal & ((1u << lane) - 1))] = base + i;
        done += warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        lvl[25 + l] = out0;
        if (l == ORB_NLEVELS - 1) lvl[33] = out0 + done;
    }
}

// ---- Harris response of the survivors (A.1.6) ------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_orb_harris(OrbDims d, OrbWorkspace ws, size_t ws_stride) {
    const int f = blockIdx.y;
    const int32_t* lvl = fptr(ws.lvl_count, ws_stride, f);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= lvl[33]) return;
    const int i = fptr(ws.surv_id, ws_stride, f)[j];
    const OrbLevel L = d.lv[cand_level(lvl, i)];
    const uint8_t* img = fptr(ws.pyr, ws_stride, f) + L.off;
    const int xy = fptr(ws.cand_xy, ws_stride, f)[i];
    const int x0 = xy & 0xFFFF, y0 = xy >> 16;
    // 7x7 window of 3x3 Sobel responses = a 9x9 patch: three rows of 9 pixels slide down in registers (81 loads, not 392)
    int a = 0, b = 0, c = 0;
    const uint8_t* p = img + (size_t)(y0 - 4) * L.w + (x0 - 4);
    int rm[9], r0[9], rp[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { rm[k] = p[k]; r0[k] = p[L.w + k]; }
    p += 2 * L.w;
#pragma unroll
    for (int row = 0; row < 7; row++) {
#pragma unroll
        for (int k = 0; k < 9; k++) rp[k] = p[k];
        p += L.w;
#pragma unroll
        for (int k = 1; k <= 7; k++) {
            const int Ix = 2 * (r0[k + 1] - r0[k - 1]) + (rm[k + 1] - rm[k - 1]) + (rp[k + 1] - rp[k - 1]);
            const int Iy = 2 * (rp[k] - rm[k]) + (rp[k - 1] - rm[k - 1]) + (rp[k + 1] - rm[k + 1]);
            a += Ix * Ix;
            b += Iy * Iy;
            c += Ix * Iy;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) { rm[k] = r0[k]; r0[k] = rp[k]; }
    }
    const float scale = __fdiv_rn(1.f, 4.f * 7.f * 255.f);
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float det = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
    const float tr = __fadd_rn(fa, fb);
    const float resp = __fmul_rn(__fsub_rn(det, __fmul_rn(__fmul_rn(0.04f, tr), tr)), s4);
    fptr(ws.cand_harris, ws_stride, f)[i] = resp;
    fptr(ws.harris_dense, ws_stride, f)[j] = resp;
}

// ---- float32 separable Gaussian with the AVX2 engine's body / tail split (A.2.1) -------------------------------------------
constexpr int kBlurTX = 64, kBlurTY = 16;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(256) k_orb_blur(OrbDims d, OrbWorkspace ws, size_t ws_stride) {
    __shared__ __align__(4) uint8_t src[kBlurTY + 6][kBlurTX + 8];
    __shared__ float rows[kBlurTY + 6][kBlurTX];
    int level, bx, by;
    orb_flat_tile(d.blur_tiles, d, kBlurTX, blockIdx.x, level, bx, by);
    const int f = blockIdx.y;
    const OrbLevel L = d.lv[level];
    const int x0 = bx * kBlurTX, y0 = by * kBlurTY;
    const uint8_t* img = fptr(ws.pyr, ws_stride, f) + L.off;
    if (x0 >= 3 && x0 + kBlurTX + 3 <= L.w && y0 >= 3 && y0 + kBlurTY + 3 <= L.h) {  // interior tile: no border reflection
        const uint8_t* base = img + (size_t)(y0 - 3) * L.w + (x0 - 3);
        for (int i = threadIdx.x; i < (kBlurTY + 6) * (kBlurTX + 6); i += 256) {
            const int ty = i / (kBlurTX + 6), tx = i - ty * (kBlurTX + 6);
            src[ty][tx] = base[ty * L.w + tx];
        }
    } else {
        for (int i = threadIdx.x; i < (kBlurTY + 6) * (kBlurTX + 6); i += 256) {
            const int ty = i / (kBlurTX + 6), tx = i - ty * (kBlurTX + 6);
            src[ty][tx] = img[(size_t)reflect101(y0 + ty - 3, L.h) * L.w + reflect101(x0 + tx - 3, L.w)];
        }
    }
    __syncthreads();
    const float k0 = c_gk[0], k1 = c_gk[1], k2 = c_gk[2], k3 = c_gk[3];
    const int wbody = (L.w / 32) * 32, wcol = (L.w / 4) * 4;
    // row pass: a thread produces 4 adjacent outputs of a row from 10 source bytes (three aligned words)
    for (int i = threadIdx.x; i < (kBlurTY + 6) * (kBlurTX / 4); i += 256) {
        const int ty = i / (kBlurTX / 4), tx = (i - ty * (kBlurTX / 4)) * 4;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&src[ty][tx]);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        float p[10];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            p[j] = (float)((w0 >> (8 * j)) & 0xFFu);
            p[4 + j] = (float)((w1 >> (8 * j)) & 0xFFu);
        }
        p[8] = (float)(w2 & 0xFFu);
        p[9] = (float)((w2 >> 8) & 0xFFu);
#pragma unroll
        for (int o = 0; o < 4; o++) {
            float acc;
            if (x0 + tx + o < wbody) {
                acc = __fmaf_rn(k0, p[o], 0.f);
                acc = __fmaf_rn(k1, p[o + 1], acc); acc = __fmaf_rn(k2, p[o + 2], acc); acc = __fmaf_rn(k3, p[o + 3], acc);
                acc = __fmaf_rn(k2, p[o + 4], acc); acc = __fmaf_rn(k1, p[o + 5], acc); acc = __fmaf_rn(k0, p[o + 6], acc);
            } else {
                acc = __fmul_rn(k0, p[o]);
                acc = __fadd_rn(acc, __fmul_rn(k1, p[o + 1])); acc = __fadd_rn(acc, __fmul_rn(k2, p[o + 2]));
                acc = __fadd_rn(acc, __fmul_rn(k3, p[o + 3])); acc = __fadd_rn(acc, __fmul_rn(k2, p[o + 4]));
                acc = __fadd_rn(acc, __fmul_rn(k1, p[o + 5])); acc = __fadd_rn(acc, __fmul_rn(k0, p[o + 6]));
            }
            rows[ty][tx + o] = acc;
        }
    }
    __syncthreads();
    // column pass: a thread produces 4 consecutive outputs of a column from 10 row sums
    uint8_t* out = fptr(ws.blur, ws_stride, f) + L.off;
    {
        const int tx = threadIdx.x % kBlurTX, tyb = (threadIdx.x / kBlurTX) * 4;
        const int x = x0 + tx;
        float r[10];
#pragma unroll
        for (int j = 0; j < 10; j++) r[j] = rows[tyb + j][tx];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int y = y0 + tyb + o;
            if (x >= L.w || y >= L.h) continue;
            float c = __fmul_rn(k3, r[o + 3]);
            const float s1 = __fadd_rn(r[o + 4], r[o + 2]);
            const float s2 = __fadd_rn(r[o + 5], r[o + 1]);
            const float s3 = __fadd_rn(r[o + 6], r[o]);
            if (x < wcol) {
                c = __fmaf_rn(k2, s1, c); c = __fmaf_rn(k1, s2, c); c = __fmaf_rn(k0, s3, c);
            } else {
                c = __fadd_rn(c, __fmul_rn(k2, s1)); c = __fadd_rn(c, __fmul_rn(k1, s2)); c = __fadd_rn(c, __fmul_rn(k0, s3));
            }
            const int v = __float2int_rn(c);
            out[(size_t)y * L.w + x] = (uint8_t)min(max(v, 0), 255);
        }
    }
}

// ---- phase 2: orientation + descriptor, one warp per kept keypoint -----------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
    const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(128) k_orb_describe(OrbDims d, OrbWorkspace ws, size_t ws_stride, const int32_t* __restrict__ n_sel,
                                                      float* __restrict__ kp_out, uint8_t* __restrict__ desc_out) {
    // the sampling pattern is indexed per lane: constant memory would serialise the 32 different addresses
    __shared__ __align__(16) signed char s_pattern[1024];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        reinterpret_cast<uint32_t*>(s_pattern)[i] = reinterpret_cast<const uint32_t*>(c_pattern)[i];
    __syncthreads();
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= n_sel[f]) return;
    const int32_t* lvl = fptr(ws.lvl_count, ws_stride, f);
    const int ci = fptr(ws.sel, ws_stride, f)[k];
    const int level = cand_level(lvl, ci);
    const OrbLevel L = d.lv[level];
    const int xy = fptr(ws.cand_xy, ws_stride, f)[ci];
    const int x0 = xy & 0xFFFF, y0 = xy >> 16;
    const uint8_t* img = fptr(ws.pyr, ws_stride, f) + L.off;
    // IC angle (A.1.7): lane v+15 sums row v
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - kHalfPatch;
        const int um = (int)((0x3689abcddeeeffffull >> (4 * abs(v))) & 15ull);  // c_umax[|v|], 4 bits each (a per-lane constant-memory index would serialise)
        const uint8_t* r = img + (size_t)(y0 + v) * L.w + x0;
        int rs = 0;
        for (int u = -um; u <= um; u++) {
            const int p = r[u];
            m10 += u * p;
            rs += p;
        }
        m01 = v * rs;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    const float ptx = __fmul_rn((float)x0, L.scale), pty = __fmul_rn((float)y0, L.scale);
    if (lane == 0) {
        float* o = kp_out + ((size_t)f * d.kp_cap + k) * 6;
        o[0] = ptx; o[1] = pty; o[2] = __fmul_rn(31.f, L.scale); o[3] = angle;
        o[4] = fptr(ws.cand_harris, ws_stride, f)[ci];
        o[5] = (float)level;
    }
    // rBRIEF (A.2.2): lane i produces descriptor byte i
    const float ang = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
    const float ca = (float)cos((double)ang), sa = (float)sin((double)ang);
    const int cx = __float2int_rn(__fmul_rn(ptx, L.inv)), cy = __float2int_rn(__fmul_rn(pty, L.inv));
    const uint8_t* bl = fptr(ws.blur, ws_stride, f) + L.off;
    uint32_t byte = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const signed char* q = s_pattern + (lane * 8 + b) * 4;
        int val[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = (float)q[2 * e], py = (float)q[2 * e + 1];
            const float rx = __fsub_rn(__fmul_rn(px, ca), __fmul_rn(py, sa));
            const float ry = __fadd_rn(__fmul_rn(px, sa), __fmul_rn(py, ca));
            val[e] = bl[(size_t)(cy + __float2int_rn(ry)) * L.w + (cx + __float2int_rn(rx))];
        }
        byte |= (val[0] < val[1] ? 1u : 0u) << b;
    }
    desc_out[((size_t)f * d.kp_cap + k) * 32 + lane] = (uint8_t)byte;
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------------
void orb_make_dims(int W, int H, int nfeatures, OrbDims* d) {
    // SURVEY.md A.1.1 — float32 arithmetic exactly as OpenCV's ORB_Impl does it
    d->W = W; d->H = H; d->nfeatures = nfeatures;
    const float scaleFactor = 1.2f;
    int off = 0, rows = 0, mwords = 0;
    for (int l = 0; l < ORB_NLEVELS; l++) {
        OrbLevel& L = d->lv[l];
        L.scale = (float)std::pow((double)scaleFactor, (double)l);
        L.inv = 1.0f / L.scale;
        L.w = (int)lrintf((float)W * L.inv);
        L.h = (int)lrintf((float)H * L.inv);
        L.off = off; L.row_off = rows;
        L.mw = (L.w + 31) / 32; L.moff = mwords;
        off += (L.w * L.h + 15) / 16 * 16;
        rows += L.h;
        mwords += L.mw * L.h;
    }
    d->total_px = off; d->total_rows = rows; d->total_mwords = mwords;
    d->fast_tiles[0] = d->blur_tiles[0] = 0;
    for (int l = 0; l < ORB_NLEVELS; l++) {
        const OrbLevel& L = d->lv[l];
        d->fast_tiles[l + 1] = d->fast_tiles[l] + ((L.w + kFastTX - 1) / kFastTX) * ((L.h + kFastTY - 1) / kFastTY);
        d->blur_tiles[l + 1] = d->blur_tiles[l] + ((L.w + kBlurTX - 1) / kBlurTX) * ((L.h + kBlurTY - 1) / kBlurTY);
    }
    float factor = (float)(1.0 / (double)scaleFactor);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)ORB_NLEVELS));
    int sum = 0;
    for (int l = 0; l < ORB_NLEVELS - 1; l++) {
        d->lv[l].nfeat = (int)lrintf(nd);
        sum += d->lv[l].nfeat;
        nd *= factor;
    }
    d->lv[ORB_NLEVELS - 1].nfeat = nfeatures - sum > 0 ? nfeatures - sum : 0;
    d->cand_cap = off / 4 + 1024;  // strict 3x3 NMS leaves at most one candidate per 2x2 block
    // retainBest keeps every tie at its boundary, so a frame can yield more than nfeatures keypoints (cv2 returns them all):
    // room for 25 % + 128 extra; beyond that ovo_orb_detect_finish reports an error instead of silently truncating
    d->kp_cap = nfeatures + nfeatures / 4 + 128;
}

size_t orb_workspace_bytes(const OrbDims& d) {
    size_t b = 0;
    b += 4 * align_up((size_t)d.total_px + 64, 256);
    b += align_up((size_t)d.total_rows * 4, 256);
    b += align_up((size_t)d.total_mwords * 4, 256);
    b += align_up(64 * 4, 256);
    b += 4 * align_up((size_t)d.cand_cap * 4, 256) + align_up((size_t)d.cand_cap, 256);
    b += align_up((size_t)d.kp_cap * 4, 256);
    return b;
}

void orb_carve(const OrbDims& d, uint8_t* base, OrbWorkspace* ws) {
    uint8_t* p = base;
    const size_t px = align_up((size_t)d.total_px + 64, 256);
    ws->pyr = p; p += px;
    ws->maskpyr = p; p += px;
    ws->score = p; p += px;
    ws->blur = p; p += px;
    ws->row_offset = (int32_t*)p; p += align_up((size_t)d.total_rows * 4, 256);
    ws->candmask = (uint32_t*)p; p += align_up((size_t)d.total_mwords * 4, 256);
    ws->lvl_count = (int32_t*)p; p += align_up(64 * 4, 256);
    ws->cand_xy = (int32_t*)p; p += align_up((size_t)d.cand_cap * 4, 256);
    ws->cand_harris = (float*)p; p += align_up((size_t)d.cand_cap * 4, 256);
    ws->harris_dense = (float*)p; p += align_up((size_t)d.cand_cap * 4, 256);
    ws->surv_id = (int32_t*)p; p += align_up((size_t)d.cand_cap * 4, 256);
    ws->cand_score = p; p += align_up((size_t)d.cand_cap, 256);
    ws->sel = (int32_t*)p; p += align_up((size_t)d.kp_cap * 4, 256);
}

// resize tables (host): for level l >= 1, xtab at tab + tab_off[l][0] (w_l entries), ytab at tab + tab_off[l][1]
void orb_make_resize_tables(const OrbDims& d, int32_t* tab, int* tab_off /*[8][2]*/, int* total) {
    int off = 0;
    for (int l = 1; l < ORB_NLEVELS; l++) {
        for (int ax = 0; ax < 2; ax++) {
            const int s = ax == 0 ? d.lv[l - 1].w : d.lv[l - 1].h, t = ax == 0 ? d.lv[l].w : d.lv[l].h;
            tab_off[2 * l + ax] = off;
            const double scale = 1.0 / ((double)t / (double)s);
            for (int v = 0; v < t; v++) {
                const double fv = scale * (v + 0.5) - 0.5;
                int i = (int)std::floor(fv);
                int c = (int)std::nearbyint((fv - i) * 256.0);
                if (i < 0) { i = 0; c = 0; }
                if (i >= s - 1) { i = s - 1; c = 0; }
                if (tab) tab[off + v] = i | (c << 16);
            }
            off += t;
        }
    }
    *total = off;
}

int orb_phase1_launch(const OrbDims& d, const OrbWorkspace* ws0, size_t ws_stride, const int32_t* tab_dev, const int* tab_off, int nb,
                      const uint8_t* img, int pitch, size_t frame_stride, const uint8_t* mask, int mask_pitch,
                      size_t mask_frame_stride, cudaStream_t st) {
    const OrbWorkspace& ws = *ws0;
    const int has_mask = mask != nullptr;
    {
        dim3 grid(cdiv(d.W * d.H, 1024), nb);
        OVO_LAUNCH(k_orb_copy_level0, grid, dim3(256), 0, st, d, ws, ws_stride, img, pitch, frame_stride, mask, mask_pitch, mask_frame_stride);
        OVO_LAUNCH_CHECK();
    }
    for (int l = 1; l < ORB_NLEVELS; l++) {
        dim3 grid(cdiv(((d.lv[l].w + 3) / 4) * d.lv[l].h, 256), nb);
        OVO_LAUNCH(k_orb_resize, grid, dim3(256), 0, st, d, ws, ws_stride, l, tab_dev + tab_off[2 * l], tab_dev + tab_off[2 * l + 1], has_mask);
        OVO_LAUNCH_CHECK();
    }
    {
        OVO_LAUNCH(k_orb_fast_nms, dim3(d.fast_tiles[ORB_NLEVELS], nb), dim3(256), 0, st, d, ws, ws_stride, has_mask);
        OVO_LAUNCH_CHECK();
        OVO_LAUNCH(k_orb_scan, dim3(nb), dim3(1024), 0, st, d, ws, ws_stride);
        OVO_LAUNCH_CHECK();
        OVO_LAUNCH(k_orb_emit, dim3(cdiv(d.total_rows, 8), nb), dim3(256), 0, st, d, ws, ws_stride);
        OVO_LAUNCH_CHECK();
    }
    {
        OVO_LAUNCH(k_orb_survivors_count, dim3(ORB_NLEVELS, nb), dim3(1024), 0, st, d, ws, ws_stride);
        OVO_LAUNCH_CHECK();
        OVO_LAUNCH(k_orb_survivors, dim3(ORB_NLEVELS, nb), dim3(1024), 0, st, d, ws, ws_stride);
        OVO_LAUNCH_CHECK();
        dim3 grid(cdiv(d.cand_cap, 128), nb);
        OVO_LAUNCH(k_orb_harris, grid, dim3(128), 0, st, d, ws, ws_stride);
        OVO_LAUNCH_CHECK();
    }
    {
        OVO_LAUNCH(k_orb_blur, dim3(d.blur_tiles[ORB_NLEVELS], nb), dim3(256), 0, st, d, ws, ws_stride);
        OVO_LAUNCH_CHECK();
    }
    return 0;
}

int orb_phase2_launch(const OrbDims& d, const OrbWorkspace* ws0, size_t ws_stride, int nb, int max_sel, const int32_t* n_sel_dev,
                      float* kp_out, uint8_t* desc_out, cudaStream_t st) {
    if (max_sel <= 0) return 0;
    dim3 grid(cdiv(max_sel, 4), nb);
    OVO_LAUNCH(k_orb_describe, grid, dim3(128), 0, st, d, *ws0, ws_stride, n_sel_dev, kp_out, desc_out);
    OVO_LAUNCH_CHECK();
    return 0;
}

}  // namespace ovo
