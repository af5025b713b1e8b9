// StereoSGBM (MODE_SGBM, 5 directions) for sm_100a — stage 4 of the openVO hot path.
//
// Replaces cv2.StereoSGBM.compute as called by the reference at src/openVO/stereo_camera.py:51 (object built at
// :23-27).  Semantics: SURVEY.md Appendix A.4 (bit-exact; oracle = oracle/sgbm_restate.cpp).
//
// Data flow (per frame, all in HBM/L2; D padded to Dp in {64,128,256} so that one warp owns one cost vector):
//   k_sgbm_prep   images -> byte-packed (v, vmin, vmax) planes for the Sobel-x-clipped and the raw rows (A.4.1)
//   k_sgbm_cost   prep   -> C[y][x1][d] int16 : Birchfield-Tomasi cost summed over the blockSize^2 window (A.4.2)
//   k_sgbm_vert   C      -> Lv[3][y][x1][d]   : paths from (x-1,y-1), (x,y-1), (x+1,y-1); one warp per scan line
//   k_sgbm_horiz  C, Lv  -> raw disparity     : paths from (x-1,y) and (x+1,y) run towards each other by two warps
//                                               per row, parking their state every 4 cells; whoever reaches a cell
//                                               second replays the other's path from the checkpoint, owns the complete
//                                               5-path sum and does WTA / uniqueness / sub-pixel / disp2; then the LR check
//   k_median3, k_ccl_*   -> 3x3 median and speckle filter (connected components, union-find)
// All cost arithmetic is packed 2 x u16 per register on the DPX pipe (VIADDMNMX.U16x2 / VIMNMX.U16x2); the per-cell
// minimum is one CREDUX.MIN.  No tensor cores: nothing here is a contraction.
#include "common.cuh"

namespace ovo {

namespace {

constexpr uint32_t kMaxC2 = 0x7FFF7FFFu;  // MAX_COST in both halves
constexpr int kInv = -16;                 // (minDisparity - 1) * 16
constexpr uint32_t kD2Init = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t bcast16(uint32_t v) { return v * 0x10001u; }

template <typename T>
__device__ __forceinline__ T* frame_ptr(T* p, size_t stride_bytes, int f) {
    return (T*)((const uint8_t*)p + stride_bytes * (size_t)f);
}

// ------------------------------------------------------------------------------------------------------------
// A.4.1 row preparation.  One word per pixel and row type: byte0 = v, byte1 = min(vl, vr, v), byte2 = max(...).
// ------------------------------------------------------------------------------------------------------------
constexpr int kEPT = 4;  // pixels per thread of the small per-pixel kernels (CTA = 256 threads = 1024 consecutive pixels)

__global__ void __launch_bounds__(256) k_sgbm_prep(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right, int pitch,
                                                   size_t frame_stride, SgbmDims d, SgbmWorkspace ws, size_t ws_stride) {
    // one thread = 4 adjacent pixels of a row: 3 rows x 8 bytes are loaded once and shared by the 6 gradients they need
    const int f = blockIdx.y >> 1, im = blockIdx.y & 1;
    const int W = d.W, H = d.H, ftzero = d.ftzero;
    const int QW = (W + 3) >> 2;
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= QW * H) return;
    const int y = q / QW, x0 = (q - y * QW) * 4;
    const uint8_t* I = (im ? right : left) + frame_stride * (size_t)f;
    const uint8_t* rows[3] = {I + (size_t)y * pitch, I + (size_t)max(y - 1, 0) * pitch, I + (size_t)min(y + 1, H - 1) * pitch};
    int a[3][8];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 8; j++) a[k][j] = rows[k][min(max(x0 - 2 + j, 0), W - 1)];
    int G[6], R[6];  // clipped Sobel-x + ftzero, and the raw value, at x0-1 .. x0+4 (ftzero on the border columns)
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const int xx = x0 - 1 + j;
        const bool border = xx <= 0 || xx >= W - 1;
        const int v = 2 * (a[0][j + 2] - a[0][j]) + (a[1][j + 2] - a[1][j]) + (a[2][j + 2] - a[2][j]);
        G[j] = border ? ftzero : min(max(v, -ftzero), ftzero) + ftzero;
        R[j] = border ? ftzero : a[0][j + 1];
    }
    uint2* out = reinterpret_cast<uint2*>(frame_ptr(ws.prep, ws_stride, f)) + (size_t)im * H * W + (size_t)y * W;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int x = x0 + p;
        if (x >= W) break;
        uint2 o;
        {
            const int c = G[p + 1], vl = x > 0 ? (c + G[p]) >> 1 : c, vr = x < W - 1 ? (c + G[p + 2]) >> 1 : c;
            o.x = (uint32_t)c | ((uint32_t)min(min(vl, vr), c) << 8) | ((uint32_t)max(max(vl, vr), c) << 16);
        }
        {
            const int c = R[p + 1], vl = x > 0 ? (c + R[p]) >> 1 : c, vr = x < W - 1 ? (c + R[p + 2]) >> 1 : c;
            o.y = (uint32_t)c | ((uint32_t)min(min(vl, vr), c) << 8) | ((uint32_t)max(max(vl, vr), c) << 16);
        }
        out[x] = o;
    }
}

// ------------------------------------------------------------------------------------------------------------
// A.4.1 + A.4.2 cost volume.  A thread owns one disparity pair (d, d+1) of one unit = (TX columns) x (RS rows); it
// walks the rows of the strip, and inside a row the TX + 2*SW2 columns, keeping the horizontal window in registers
// and the vertical window as a ring of horizontal sums in (thread-private, conflict-free) shared memory.
// ------------------------------------------------------------------------------------------------------------
constexpr int kCostThreads = 128;
#ifndef OVO_COST_RS
#define OVO_COST_RS 32
#endif
constexpr int kCostRS = OVO_COST_RS;  // rows per unit (the vertical window adds 2*SW2 halo rows)

__device__ __forceinline__ uint32_t bt_pair(uint32_t lw, uint32_t rw0, uint32_t rw1) {
    // lw: left word at x; rw0 / rw1: right words at x-d and x-d-1
    const uint32_t v = __byte_perm(rw0, rw1, 0x7430), vmin = __byte_perm(rw0, rw1, 0x7531), vmax = __byte_perm(rw0, rw1, 0x7632);
    const uint32_t u = __byte_perm(lw, 0, 0x4040), umin = __byte_perm(lw, 0, 0x4141), umax = __byte_perm(lw, 0, 0x4242);
    const uint32_t c0 = __vmaxu2(vmin, u) - __vminu2(vmax, u);  // max(0, u - vmax, vmin - u), both halves
    const uint32_t c1 = __vmaxu2(umin, v) - __vminu2(umax, v);  // max(0, v - umax, umin - v)
    return __vminu2(c0, c1);
}

// one row of a unit: horizontal sums of TX columns, folded into the vertical ring / running sums
template <int SW2, int TX, bool EDGE, bool PAD>
__device__ __forceinline__ void cost_row(const uint2* __restrict__ Lrow, const uint2* __restrict__ Rrow, int x0, int d0, int D, int W1,
                                         bool pad, uint32_t* slot, bool have_old, uint32_t (&vs)[TX]) {
    constexpr int BS = 2 * SW2 + 1;
    uint32_t win[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) win[i] = 0;
    uint32_t hs = 0;
    // interior tiles: all TX + 2*SW2 columns are in range, so every address is a row base plus a compile-time offset and
    // the right-image word of disparity d+1 is the previous column's word of disparity d
    const uint2* Lb = Lrow + x0 + D;
    const uint2* Rb = Rrow + x0 + D - d0;
    uint2 rw[2];  // right words of this and of the previous column, alternating so that no copy is needed
    rw[0] = rw[1] = make_uint2(0, 0);
    if (!EDGE && (!PAD || !pad)) rw[1] = __ldg(Rb - SW2 - 1);
#pragma unroll
    for (int j = -SW2; j < TX + SW2; j++) {
        uint32_t pix = 0;
        if (!PAD || !pad) {
            uint2 lw, r0, r1;
            if (EDGE) {
                const int xx = min(max(x0 + j, 0), W1 - 1);
                lw = __ldg(Lrow + xx + D);
                r0 = __ldg(Rrow + xx + D - d0);
                r1 = __ldg(Rrow + xx + D - d0 - 1);
            } else {
                lw = __ldg(Lb + j);
                rw[(j + SW2) & 1] = __ldg(Rb + j);
                r0 = rw[(j + SW2) & 1];
                r1 = rw[((j + SW2) & 1) ^ 1];
            }
            const uint32_t cg = bt_pair(lw.x, r0.x, r1.x);
            const uint32_t cr = bt_pair(lw.y, r0.y, r1.y);
            pix = cg + ((cr >> 2) & 0x3FFF3FFFu);
        }
        hs = hs + pix - win[0];
#pragma unroll
        for (int i = 0; i < BS - 1; i++) win[i] = win[i + 1];
        win[BS - 1] = pix;
        if (j >= SW2) {
            const int c = j - SW2;
            const uint32_t old = have_old ? slot[c * kCostThreads] : 0u;
            vs[c] = vs[c] - old + hs;
            slot[c * kCostThreads] = hs;
        }
    }
}

#ifndef OVO_COST_MINB
#define OVO_COST_MINB 5
#endif
template <int SW2, int TX, bool PAD>
__global__ void __launch_bounds__(kCostThreads, OVO_COST_MINB) k_sgbm_cost(SgbmDims d, SgbmWorkspace ws, size_t ws_stride) {
    constexpr int BS = 2 * SW2 + 1;
    __shared__ uint32_t ring[BS * TX * kCostThreads];
    const int npairs = d.Dp >> 1;
    const int tid = threadIdx.y * npairs + threadIdx.x;
    const int n_xt = (d.W1 + TX - 1) / TX, n_ys = (d.H + kCostRS - 1) / kCostRS;
    const int unit = blockIdx.x * blockDim.y + threadIdx.y;
    if (unit >= n_xt * n_ys) return;
    const int x0 = (unit % n_xt) * TX, y0 = (unit / n_xt) * kCostRS;
    const int f = blockIdx.z;
    const int W = d.W, H = d.H, D = d.D, W1 = d.W1;
    const int d0 = 2 * threadIdx.x;
    const bool pad = d0 >= D;
    const uint2* prep = reinterpret_cast<const uint2*>(frame_ptr(ws.prep, ws_stride, f));
    const size_t plane = (size_t)H * W;
    uint32_t* Cw = reinterpret_cast<uint32_t*>(frame_ptr(ws.C, ws_stride, f));
    const int yend = min(y0 + kCostRS, H);
    const bool edge = x0 - SW2 < 0 || x0 + TX + SW2 > W1;

    uint32_t vs[TX];
#pragma unroll
    for (int c = 0; c < TX; c++) vs[c] = 0;

    int k = 0;  // rows accumulated so far
    for (int r = y0 - SW2; r < yend + SW2; r++, k++) {
        const int yc = min(max(r, 0), H - 1);
        const uint2* Lrow = prep + (size_t)yc * W;
        const uint2* Rrow = Lrow + plane;
        uint32_t* slot = ring + (size_t)(k % BS) * TX * kCostThreads + tid;
        if (edge) cost_row<SW2, TX, true, PAD>(Lrow, Rrow, x0, d0, D, W1, pad, slot, k >= BS, vs);
        else cost_row<SW2, TX, false, PAD>(Lrow, Rrow, x0, d0, D, W1, pad, slot, k >= BS, vs);
        const int y = r - SW2;
        if (y >= y0) {
            uint32_t* out = Cw + ((size_t)y * W1 + x0) * npairs + threadIdx.x;
#pragma unroll
            for (int c = 0; c < TX; c++)
                if (x0 + c < W1) out[(size_t)c * npairs] = vs[c];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// A.4.3 one path step on a warp-wide cost vector.  Lane l owns d in [2*NPR*l, 2*NPR*(l+1)), two per register.
// A predecessor outside the image is the all-zero vector with min 0, for which the step formula yields L = C, so path
// (re)starts are just a state reset followed by the ordinary step.
// ------------------------------------------------------------------------------------------------------------
template <int NPR>
struct PathState {
    uint32_t L[NPR];
    uint32_t m;  // min over d of L (warp-uniform)
};

template <int NPR>
__device__ __forceinline__ void path_reset(PathState<NPR>& s) {
#pragma unroll
    for (int r = 0; r < NPR; r++) s.L[r] = 0;
    s.m = 0;
}

template <int NPR>
__device__ __forceinline__ uint32_t vec_min(const uint32_t (&L)[NPR]) {  // min over d (warp-uniform)
    uint32_t t = L[0];
#pragma unroll
    for (int r = 1; r < NPR; r++) t = __vminu2(t, L[r]);
    return __reduce_min_sync(0xffffffffu, min(t & 0xFFFFu, t >> 16));
}

template <int NPR, bool PAD>
__device__ __forceinline__ void path_step(PathState<NPR>& s, const uint32_t (&c)[NPR], const uint32_t (&padmask)[NPR],
                                          uint32_t P1P1, uint32_t P2, bool lane_first, bool lane_last) {
    uint32_t below = __shfl_up_sync(0xffffffffu, s.L[NPR - 1], 1);  // neighbour lane's top pair
    uint32_t above = __shfl_down_sync(0xffffffffu, s.L[0], 1);      // neighbour lane's bottom pair
    if (lane_first) below = kMaxC2;                                  // Lp[-1] = MAX_COST
    if (lane_last) above = kMaxC2;                                   // Lp[D]  = MAX_COST
    const uint32_t mP2 = bcast16(s.m + P2), mm = bcast16(s.m);
    uint32_t out[NPR];
#pragma unroll
    for (int r = 0; r < NPR; r++) {
        const uint32_t dm1 = __funnelshift_l(r == 0 ? below : s.L[r - 1], s.L[r], 16);       // Lp[d-1]
        const uint32_t dp1 = __funnelshift_r(s.L[r], r == NPR - 1 ? above : s.L[r + 1], 16); // Lp[d+1]
        uint32_t t = __viaddmin_u16x2(dm1, P1P1, s.L[r]);
        t = __viaddmin_u16x2(dp1, P1P1, t);
        t = __vminu2(t, mP2);
        out[r] = c[r] + (t - mm);
        if (PAD) out[r] |= padmask[r];
    }
#pragma unroll
    for (int r = 0; r < NPR; r++) s.L[r] = out[r];
    s.m = vec_min<NPR>(s.L);
}

template <int NPR>
__device__ __forceinline__ void make_padmask(uint32_t (&padmask)[NPR], int lane, int D) {
#pragma unroll
    for (int r = 0; r < NPR; r++) padmask[r] = (2 * (NPR * lane + r) >= D) ? kMaxC2 : 0u;
}

template <int NPR>
__device__ __forceinline__ void ldv(uint32_t (&v)[NPR], const uint32_t* p) {
    if constexpr (NPR == 4) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (NPR == 2) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = *p;
    }
}
template <int NPR>
__device__ __forceinline__ void stv(uint32_t* p, const uint32_t (&v)[NPR]) {
    if constexpr (NPR == 4) *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
    else if constexpr (NPR == 2) *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1]);
    else *p = v[0];
}

// ------------------------------------------------------------------------------------------------------------
// Paths 1..3 (top -> bottom).  One warp per scan line; a diagonal line that leaves the image on one side re-enters
// on the other with a fresh (zero) predecessor, so every warp does exactly H steps.  Cell offsets are 32-bit word
// offsets inside one frame's volume (< 2^32 words up to 4K / 256 disparities).
// ------------------------------------------------------------------------------------------------------------
#ifndef OVO_VERT_PF4
#define OVO_VERT_PF4 4
#endif
template <int NPR>
__host__ __device__ constexpr int vert_pf() { return NPR == 4 ? OVO_VERT_PF4 : 8; }  // register prefetch depth (cells)

template <int NPR, bool PAD, int DIR, bool UP>
__device__ __forceinline__ void vert_line(const SgbmDims& d, const uint32_t* __restrict__ C, uint32_t* __restrict__ Lout, int line,
                                          int lane) {
    constexpr int WPC = 32 * NPR;  // words per cell
    constexpr int kVertPF = vert_pf<NPR>();
    constexpr int STEP = DIR == 0 ? 1 : (DIR == 2 ? -1 : 0);
    const int W1 = d.W1, H = d.H;
    const int xreset = DIR == 0 ? 0 : (DIR == 2 ? W1 - 1 : -1);
    const ptrdiff_t rowstride = (ptrdiff_t)W1 * WPC;
    const ptrdiff_t rowadv = UP ? -rowstride : rowstride;  // UP: MODE_HH's second pass walks the rows bottom -> top
    // pointer / column of the next row of this scan line (wraps around the image, which is where the path restarts)
    auto next_row = [&](int& x, auto*& p) {
        p += rowadv + STEP * WPC;
        if (STEP != 0) {
            x += STEP;
            if (STEP > 0 && x == W1) { x = 0; p -= rowstride; }
            if (STEP < 0 && x < 0) { x = W1 - 1; p += rowstride; }
        }
    };
    uint32_t padmask[NPR];
    make_padmask<NPR>(padmask, lane, d.D);
    const uint32_t P1P1 = bcast16(d.P1), P2 = d.P2;
    const bool lane_first = lane == 0, lane_last = lane == 31;

    uint32_t cbuf[kVertPF][NPR];
    int xpf = line, x = line;
    const size_t row0 = UP ? (size_t)(H - 1) * rowstride : 0;
    const uint32_t* ppf = C + row0 + (size_t)line * WPC;
    uint32_t* pl = Lout + row0 + (size_t)line * WPC;
#pragma unroll
    for (int i = 0; i < kVertPF; i++) {
        if (i < H) ldv<NPR>(cbuf[i], ppf);
        next_row(xpf, ppf);
    }
    PathState<NPR> s;
    path_reset<NPR>(s);
    const ptrdiff_t adv = rowadv + STEP * WPC;
    int y = 0;
    for (; y + 2 * kVertPF <= H; y += kVertPF) {  // every step and every prefetch of this group is in range
        // a diagonal line wraps at most once per image width: groups that neither wrap nor restart advance by a constant
        const bool clean = STEP == 0 || (STEP > 0 ? (x != 0 && x + 2 * kVertPF < W1) : (x != W1 - 1 && x - 2 * kVertPF >= 0));
        if (clean) {
#pragma unroll
            for (int i = 0; i < kVertPF; i++) {
                uint32_t c[NPR];
#pragma unroll
                for (int r = 0; r < NPR; r++) c[r] = cbuf[i][r];
                ldv<NPR>(cbuf[i], ppf);
                ppf += adv;
                path_step<NPR, PAD>(s, c, padmask, P1P1, P2, lane_first, lane_last);
                stv<NPR>(pl, s.L);
                pl += adv;
            }
            x += STEP * kVertPF;
            xpf += STEP * kVertPF;
        } else {
#pragma unroll
            for (int i = 0; i < kVertPF; i++) {
                uint32_t c[NPR];
#pragma unroll
                for (int r = 0; r < NPR; r++) c[r] = cbuf[i][r];
                ldv<NPR>(cbuf[i], ppf);
                next_row(xpf, ppf);
                if (STEP != 0 && x == xreset) path_reset<NPR>(s);
                path_step<NPR, PAD>(s, c, padmask, P1P1, P2, lane_first, lane_last);
                stv<NPR>(pl, s.L);
                next_row(x, pl);
            }
        }
    }
    for (; y < H; y += kVertPF) {
#pragma unroll
        for (int i = 0; i < kVertPF; i++) {
            if (y + i < H) {
                uint32_t c[NPR];
#pragma unroll
                for (int r = 0; r < NPR; r++) c[r] = cbuf[i][r];
                if (y + i + kVertPF < H) ldv<NPR>(cbuf[i], ppf);
                next_row(xpf, ppf);
                if (STEP != 0 && x == xreset) path_reset<NPR>(s);
                path_step<NPR, PAD>(s, c, padmask, P1P1, P2, lane_first, lane_last);
                stv<NPR>(pl, s.L);
                next_row(x, pl);
            }
        }
    }
}

template <int NPR, bool PAD>
__global__ void __launch_bounds__(256) k_sgbm_vert(SgbmDims d, SgbmWorkspace ws, size_t ws_stride) {
    const int lane = threadIdx.x & 31;
    // direction is the fastest-varying block coordinate: the three directions of the same columns are resident together and
    // share their reads of C through L2
    const int ndir = d.mode ? 6 : 3;
    const int dir = blockIdx.x % ndir, f = blockIdx.y;
    const int line = (blockIdx.x / ndir) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (line >= d.W1) return;
    const uint32_t* C = reinterpret_cast<const uint32_t*>(frame_ptr(ws.C, ws_stride, f)) + lane * NPR;
    uint32_t* Lout = reinterpret_cast<uint32_t*>(frame_ptr(ws.Lv, ws_stride, f)) + (size_t)dir * d.H * d.W1 * (32 * NPR) + lane * NPR;
    if (dir == 0) vert_line<NPR, PAD, 0, false>(d, C, Lout, line, lane);
    else if (dir == 1) vert_line<NPR, PAD, 1, false>(d, C, Lout, line, lane);
    else if (dir == 2) vert_line<NPR, PAD, 2, false>(d, C, Lout, line, lane);
    else if (dir == 3) vert_line<NPR, PAD, 0, true>(d, C, Lout, line, lane);   // from (x-1, y+1)
    else if (dir == 4) vert_line<NPR, PAD, 1, true>(d, C, Lout, line, lane);   // from (x,   y+1)
    else vert_line<NPR, PAD, 2, true>(d, C, Lout, line, lane);                 // from (x+1, y+1)
}

// ------------------------------------------------------------------------------------------------------------
// Paths 0 and 4 + selection (A.4.4) + LR check (A.4.5).  One CTA (two warps) per row.  Warp 0 runs path 0 left -> right,
// warp 1 runs path 4 right -> left.  Phase 1: each warp advances its own path over its half of the row reading only C
// and parks its state every K cells (a checkpoint, 1/K of a volume).  Phase 2: each warp continues into the other
// half; per K-cell segment it replays the other warp's path forward from the checkpoint (registers only), then walks
// the segment in its own direction, where S = sat(L1 + L2 + L3 + replayed + own) is complete and selection runs.
// Every volume is read once per use (C twice, Lv once) and nothing volume-sized is written.  disp2 is
// order-independent: min cost, ties to the larger x (= larger d), which is what the reference's right-to-left sweep keeps.
// ------------------------------------------------------------------------------------------------------------
#ifndef OVO_HOR_MINB
#define OVO_HOR_MINB 8
#endif
template <int NPR>
__host__ __device__ constexpr int horiz_seg() { return 4; }  // K: cells per checkpoint segment (= cells per batched selection)

template <int NPR>
__device__ __forceinline__ uint32_t half_of(const uint32_t (&S)[NPR], int k) {  // k = 2*r + h, compile-time after unrolling
    return (k & 1) ? (S[k >> 1] >> 16) : (S[k >> 1] & 0xFFFFu);
}

// Per-cell selection, serial part only: argmin, the runner-up over |d-best| > 1 (for the uniqueness test) and the two
// neighbours of the minimum are reduced here and parked in shared memory; the uniqueness decision, the sub-pixel
// division and the disp2 update are order-independent and run data-parallel over the row afterwards.
template <int NPR, bool PAD>
__device__ __forceinline__ void wta_cell(const uint32_t (&S)[NPR], int lane, const SgbmDims& d, int x1, uint32_t* selA,
                                         uint32_t* selB, uint16_t* selBest) {
    const int D = d.D;
    const int dd0 = 2 * NPR * lane;
    // first d minimising S: the smallest (S << 9 | d)
    uint32_t kbest = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 2 * NPR; k++) {
        const uint32_t key = half_of<NPR>(S, k) * 512u + (uint32_t)(dd0 + k);
        if (!PAD || dd0 + k < D) kbest = min(kbest, key);
    }
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, kbest);
    const int minS = (int)(kmin >> 9), best = (int)(kmin & 511u);
    const int fac = 100 - d.uniq;
    uint32_t m2 = 0xFFFFu;
    if (fac > 0) {  // exists d, |d-best|>1, S[d]*fac < minS*100  <=>  (min over those d) * fac < minS*100
#pragma unroll
        for (int k = 0; k < 2 * NPR; k++) {
            const bool far = (uint32_t)(dd0 + k - best + 1) > 2u && (!PAD || dd0 + k < D);
            m2 = min(m2, far ? half_of<NPR>(S, k) : 0xFFFFu);
        }
        m2 = __reduce_min_sync(0xffffffffu, m2);
    } else {        // uniquenessRatio >= 100: evaluate the predicate as written; park 0 = reject, 0xFFFF = accept
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 2 * NPR; k++)
            if ((!PAD || dd0 + k < D) && (int)half_of<NPR>(S, k) * fac < minS * 100 && abs(dd0 + k - best) > 1) bad = true;
        m2 = __any_sync(0xffffffffu, bad) ? 0u : 0xFFFFu;
    }
    auto at = [&](int dd) -> uint32_t {  // S[dd], dd warp-uniform
        const int r = (dd % (2 * NPR)) >> 1;
        uint32_t w = S[0];
#pragma unroll
        for (int rr = 1; rr < NPR; rr++)
            if (r == rr) w = S[rr];
        w = __shfl_sync(0xffffffffu, w, dd / (2 * NPR));
        return (dd & 1) ? (w >> 16) : (w & 0xFFFFu);
    };
    const uint32_t sm1 = at(max(best - 1, 0)), sp1 = at(min(best + 1, D - 1));
    if (lane == 0) {
        selA[x1] = (uint32_t)minS | (m2 << 16);
        selB[x1] = sm1 | (sp1 << 16);
        selBest[x1] = (uint16_t)best;
    }
}

// Selection for the K = 4 cells of a segment at once (uniquenessRatio < 100, the usual case).  The warp has parked the
// four S vectors in shared memory; 8 lanes share a cell, each scanning Dp/8 consecutive disparities, so the reductions,
// the neighbour look-ups and the stores are paid once per four cells.  Padded disparities hold MAX_COST and the larger
// d, so they can neither win the argmin nor lower the runner-up.
template <int NPR>
__device__ __forceinline__ void wta_batch(const uint32_t* svec, int lane, const SgbmDims& d, int xo, int dirx, int cnt, uint32_t* selA,
                                          uint32_t* selB, uint16_t* selBest) {
    constexpr int WPC = 32 * NPR, NW = 4 * NPR;  // words per cell, words per lane
    const int g = lane >> 3, q = lane & 7;
    uint32_t w[NW];
#pragma unroll
    for (int i = 0; i < NW; i += 4) {
        const uint4 t = *reinterpret_cast<const uint4*>(svec + g * WPC + q * NW + i);
        w[i] = t.x; w[i + 1] = t.y; w[i + 2] = t.z; w[i + 3] = t.w;
    }
    // first d minimising S: the smallest (S << 9 | d)
    uint32_t kbest = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        kbest = min(kbest, (w[i] & 0xFFFFu) * 512u + (uint32_t)(2 * i));
        kbest = min(kbest, (w[i] >> 16) * 512u + (uint32_t)(2 * i + 1));
    }
    kbest += (uint32_t)(q * 2 * NW);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) kbest = min(kbest, __shfl_xor_sync(0xffffffffu, kbest, o));
    const int minS = (int)(kbest >> 9), best = (int)(kbest & 511u);
    // runner-up over |d - best| > 1: bit c of `near` marks this lane's c-th disparity as one of best-1, best, best+1
    const uint32_t sh = (uint32_t)(best - q * 2 * NW + 1);  // bit of best+1, plus 2; huge (wrapped) when best lies far below
    const uint32_t near = sh < 34u ? (uint32_t)((7ull << sh) >> 2) : 0u;  // 64-bit: a lane covers up to 32 disparities
    uint32_t m2 = 0xFFFFu;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        if (!(near & (1u << (2 * i)))) m2 = min(m2, w[i] & 0xFFFFu);
        if (!(near & (1u << (2 * i + 1)))) m2 = min(m2, w[i] >> 16);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    if (q == 0 && g < cnt) {
        const uint16_t* sh16 = reinterpret_cast<const uint16_t*>(svec + g * WPC);
        const uint32_t sm1 = sh16[max(best - 1, 0)], sp1 = sh16[min(best + 1, 2 * WPC - 1)];  // only used when 0 < best < D-1
        const int x1 = xo - dirx * g;
        selA[x1] = (uint32_t)minS | (m2 << 16);
        selB[x1] = sm1 | (sp1 << 16);
        selBest[x1] = (uint16_t)best;
    }
}

template <int NPR>
struct HorizRow {          // per-thread view of one row
    const uint32_t* C;     // + row offset + lane
    const uint32_t* Lv;    // Lv[0] of the row; Lv[v] is v * vol words further
    size_t vol;
    uint32_t* ck_own;      // checkpoints of this warp's phase-1 sweep (segment j at j * 32*NPR words)
    const uint32_t* ck_oth;
    int xa, n1, n2;        // first cell and length of the own phase-1 sweep; length of the other warp's
    uint32_t *selA, *selB;
    uint16_t* selBest;
    uint32_t* svec;        // this warp's shared-memory slots for the K cost vectors of a segment
};

#ifndef OVO_HOR_PF1
#define OVO_HOR_PF1 2
#endif
template <int NPR, bool PAD, int DIRX>
__device__ __forceinline__ void horiz_phase1(const SgbmDims& d, const HorizRow<NPR>& R, PathState<NPR>& s, int lane) {
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>(), DS = DIRX * WPC;
    constexpr int PF = OVO_HOR_PF1 * K;  // register prefetch depth (cells): this phase is light, so it needs a long run-ahead
    uint32_t padmask[NPR];
    make_padmask<NPR>(padmask, lane, d.D);
    const uint32_t P1P1 = bcast16(d.P1), P2 = d.P2;
    const bool lane_first = lane == 0, lane_last = lane == 31;
    const int n1 = R.n1;
    uint32_t cb[PF][NPR];
    const uint32_t* pc = R.C + (ptrdiff_t)R.xa * WPC;
#pragma unroll
    for (int i = 0; i < PF; i++)
        if (i < n1) ldv<NPR>(cb[i], pc + i * DS);
    uint32_t* pk = R.ck_own;  // slot of the checkpoint taken before step k (slot 0, the all-zero start, is never written)
    for (int k = 0; k < n1; k += PF) {
        pc += PF * DS;  // now points at the cell PF ahead of step k
        if (k + 2 * PF <= n1) {
#pragma unroll
            for (int i = 0; i < PF; i++) {
                if (i % K == 0) {
                    if (k + i) stv<NPR>(pk, s.L);
                    pk += WPC;
                }
                uint32_t c[NPR];
#pragma unroll
                for (int r = 0; r < NPR; r++) c[r] = cb[i][r];
                ldv<NPR>(cb[i], pc + i * DS);
                path_step<NPR, PAD>(s, c, padmask, P1P1, P2, lane_first, lane_last);
            }
        } else {
#pragma unroll
            for (int i = 0; i < PF; i++) {
                if (k + i < n1) {
                    if (i % K == 0) {
                        if (k + i) stv<NPR>(pk, s.L);
                        pk += WPC;
                    }
                    uint32_t c[NPR];
#pragma unroll
                    for (int r = 0; r < NPR; r++) c[r] = cb[i][r];
                    if (k + i + PF < n1) ldv<NPR>(cb[i], pc + i * DS);
                    path_step<NPR, PAD>(s, c, padmask, P1P1, P2, lane_first, lane_last);
                }
            }
        }
    }
}

template <int NPR, int DIRX, bool FULL>
__device__ __forceinline__ void horiz_load_lv(const HorizRow<NPR>& R, ptrdiff_t o0, int cnt, uint32_t (&lv)[3][horiz_seg<NPR>()][NPR]) {
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>(), DS = DIRX * WPC;
#pragma unroll
    for (int i = K - 1; i >= 0; i--) {  // the cell consumed first is requested first
        if (FULL || i < cnt) {
            ldv<NPR>(lv[0][i], R.Lv + o0 - i * DS);
            ldv<NPR>(lv[1][i], R.Lv + R.vol + o0 - i * DS);
            ldv<NPR>(lv[2][i], R.Lv + 2 * R.vol + o0 - i * DS);
        }
    }
}

// One K-cell segment of phase 2.  On entry cb / ckv / lv hold C, the other warp's checkpoint and Lv[0..2] of segment j
// (requested one segment earlier); on exit they hold those of segment j-1.  FULL: all K cells exist (only the segment
// next to the rendezvous can be short).
template <int NPR, bool PAD, bool HH, bool BATCH, int DIRX, bool FULL>
__device__ __forceinline__ void horiz_segment(const SgbmDims& d, const HorizRow<NPR>& R, PathState<NPR>& s, int lane, int j, int cnt,
                                              uint32_t (&cb)[horiz_seg<NPR>()][NPR], uint32_t (&ckv)[NPR],
                                              uint32_t (&lv)[3][horiz_seg<NPR>()][NPR], const uint32_t (&padmask)[NPR],
                                              uint32_t P1P1, uint32_t P2) {
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>(), DS = DIRX * WPC;
    const bool lane_first = lane == 0, lane_last = lane == 31;
    constexpr bool batched = BATCH;  // selection of the K cells at once; the odd uniquenessRatio >= 100 goes cell by cell
    // the other warp's cell k (counted along ITS sweep) sits at x = xo - DIRX * k
    const int xo = (DIRX > 0 ? d.W1 - 1 : 0) - DIRX * (j * K);
    const ptrdiff_t o0 = (ptrdiff_t)xo * WPC;  // cell i of the segment is at o0 - i * DS
    uint32_t sv[K][NPR];
    {
        // replay the other warp's path over the segment
        PathState<NPR> o;
        if (j == 0) {
            path_reset<NPR>(o);
        } else {
#pragma unroll
            for (int r = 0; r < NPR; r++) o.L[r] = ckv[r];
            o.m = vec_min<NPR>(o.L);
        }
#pragma unroll
        for (int i = 0; i < K; i++) {
            if (FULL || i < cnt) {
                path_step<NPR, PAD>(o, cb[i], padmask, P1P1, P2, lane_first, lane_last);
#pragma unroll
                for (int r = 0; r < NPR; r++) sv[i][r] = o.L[r];
            }
        }
#pragma unroll
        for (int i = K - 1; i >= 0; i--) {
            if (FULL || i < cnt) {
#pragma unroll
                for (int r = 0; r < NPR; r++)
                    sv[i][r] = __viaddmin_u16x2(__viaddmin_u16x2(__viaddmin_u16x2(lv[0][i][r], lv[1][i][r], kMaxC2), lv[2][i][r], kMaxC2),
                                                sv[i][r], kMaxC2);
            }
        }
    }
    if (HH) {  // MODE_HH: the three bottom-up paths Lv[3..5]
#pragma unroll
        for (int v = 3; v < 6; v++) {
#pragma unroll
            for (int i = K - 1; i >= 0; i--) {
                if (FULL || i < cnt) {
                    uint32_t u[NPR];
                    ldv<NPR>(u, R.Lv + (size_t)v * R.vol + o0 - i * DS);
#pragma unroll
                    for (int r = 0; r < NPR; r++) sv[i][r] = __viaddmin_u16x2(sv[i][r], u[r], kMaxC2);
                }
            }
        }
    }
    // request everything the next segment (always a full one) needs; it arrives while this one is being walked
    uint32_t cn[K][NPR], ckn[NPR];
#pragma unroll
    for (int r = 0; r < NPR; r++) ckn[r] = 0;
    if (j > 0) {
        const uint32_t* pn = R.C + o0 + K * DS;
#pragma unroll
        for (int i = 0; i < K; i++) ldv<NPR>(cn[i], pn - i * DS);
        if (j > 1) ldv<NPR>(ckn, R.ck_oth + (size_t)(j - 1) * WPC);
        horiz_load_lv<NPR, DIRX, true>(R, o0 + K * DS, K, lv);
    }
    // own path over the segment, in the own direction (= the other's, reversed)
#pragma unroll
    for (int i = K - 1; i >= 0; i--) {
        if (FULL || i < cnt) {
            path_step<NPR, PAD>(s, cb[i], padmask, P1P1, P2, lane_first, lane_last);
            uint32_t S[NPR];
#pragma unroll
            for (int r = 0; r < NPR; r++) S[r] = __viaddmin_u16x2(sv[i][r], s.L[r], kMaxC2);
            if (batched) stv<NPR>(R.svec + i * WPC + lane * NPR, S);
            else wta_cell<NPR, PAD>(S, lane, d, xo - DIRX * i, R.selA, R.selB, R.selBest);
        }
    }
    if (batched) {
        __syncwarp();
        wta_batch<NPR>(R.svec, lane, d, xo, DIRX, FULL ? K : cnt, R.selA, R.selB, R.selBest);
        __syncwarp();
    }
    if (j > 0) {
#pragma unroll
        for (int i = 0; i < K; i++)
#pragma unroll
            for (int r = 0; r < NPR; r++) cb[i][r] = cn[i][r];
#pragma unroll
        for (int r = 0; r < NPR; r++) ckv[r] = ckn[r];
    }
}

template <int NPR, bool PAD, bool HH, bool BATCH, int DIRX>
__device__ __forceinline__ void horiz_phase2(const SgbmDims& d, const HorizRow<NPR>& R, PathState<NPR>& s, int lane) {
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>(), DS = DIRX * WPC;
    const int n2 = R.n2;
    if (n2 <= 0) return;
    uint32_t padmask[NPR];
    make_padmask<NPR>(padmask, lane, d.D);
    const uint32_t P1P1 = bcast16(d.P1), P2 = d.P2;
    int j = (n2 + K - 1) / K - 1;
    int cnt = n2 - j * K;
    uint32_t cb[K][NPR], ckv[NPR], lv[3][K][NPR];
#pragma unroll
    for (int r = 0; r < NPR; r++) ckv[r] = 0;
    {
        const ptrdiff_t o0 = (ptrdiff_t)((DIRX > 0 ? d.W1 - 1 : 0) - DIRX * (j * K)) * WPC;
#pragma unroll
        for (int i = 0; i < K; i++)
            if (i < cnt) ldv<NPR>(cb[i], R.C + o0 - i * DS);
        if (j > 0) ldv<NPR>(ckv, R.ck_oth + (size_t)j * WPC);
        horiz_load_lv<NPR, DIRX, false>(R, o0, cnt, lv);
    }
    if (cnt < K) {
        horiz_segment<NPR, PAD, HH, BATCH, DIRX, false>(d, R, s, lane, j, cnt, cb, ckv, lv, padmask, P1P1, P2);
        j--;
    }
    for (; j >= 0; j--) horiz_segment<NPR, PAD, HH, BATCH, DIRX, true>(d, R, s, lane, j, K, cb, ckv, lv, padmask, P1P1, P2);
}

template <int NPR, bool PAD, bool HH, bool BATCH>
__global__ void __launch_bounds__(64, NPR == 4 ? 6 : OVO_HOR_MINB) k_sgbm_horiz(SgbmDims d, SgbmWorkspace ws, size_t ws_stride) {
    OVO_DYN_SMEM(uint32_t, hsm);
    uint32_t* d2key = hsm;                                        // [W]
    uint32_t* selA = hsm + d.W;                                   // [W1] minS | runner-up << 16
    uint32_t* selB = selA + d.W1;                                 // [W1] S[best-1] | S[best+1] << 16
    int16_t* disp1s = reinterpret_cast<int16_t*>(selB + d.W1);    // [W]
    uint16_t* selBest = reinterpret_cast<uint16_t*>(disp1s + d.W);  // [W1]
    const int y = blockIdx.x, f = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int W = d.W, W1 = d.W1, H = d.H;
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>();
    __shared__ __align__(16) uint32_t svec_s[2][K * WPC];
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        d2key[i] = kD2Init;
        disp1s[i] = (int16_t)kInv;
    }
    __syncthreads();

    const int mid = W1 >> 1;
    const int sph = (W1 - mid + K - 1) / K;  // checkpoint slots per half row
    HorizRow<NPR> R;
    const size_t rowoff = (size_t)y * W1 * WPC + lane * NPR;
    R.vol = (size_t)H * W1 * WPC;
    R.C = reinterpret_cast<const uint32_t*>(frame_ptr(ws.C, ws_stride, f)) + rowoff;
    R.Lv = reinterpret_cast<const uint32_t*>(frame_ptr(ws.Lv, ws_stride, f)) + rowoff;
    uint32_t* ck = reinterpret_cast<uint32_t*>(frame_ptr(ws.ckpt, ws_stride, f)) + (size_t)y * 2 * sph * WPC + lane * NPR;
    R.ck_own = ck + (size_t)wid * sph * WPC;
    R.ck_oth = ck + (size_t)(1 - wid) * sph * WPC;
    R.xa = wid == 0 ? 0 : W1 - 1;
    R.n1 = wid == 0 ? mid : W1 - mid;
    R.n2 = W1 - R.n1;
    R.selA = selA; R.selB = selB; R.selBest = selBest;
    R.svec = svec_s[wid];

    PathState<NPR> s;
    path_reset<NPR>(s);
    if (wid == 0) horiz_phase1<NPR, PAD, 1>(d, R, s, lane);
    else horiz_phase1<NPR, PAD, -1>(d, R, s, lane);
    __syncthreads();
    if (wid == 0) horiz_phase2<NPR, PAD, HH, BATCH, 1>(d, R, s, lane);
    else horiz_phase2<NPR, PAD, HH, BATCH, -1>(d, R, s, lane);
    __syncthreads();
    // ---- uniqueness, sub-pixel refinement and disp2 (A.4.4), data-parallel over the row
    {
        const int D = d.D, fac = 100 - d.uniq;
        for (int x1 = threadIdx.x; x1 < W1; x1 += blockDim.x) {
            const uint32_t a = selA[x1], b = selB[x1];
            const int minS = (int)(a & 0xFFFFu), m2 = (int)(a >> 16), best = selBest[x1];
            const bool reject = fac > 0 ? (m2 * fac < minS * 100) : (m2 == 0);
            if (reject) continue;
            const int x = x1 + D;
            if (minS < 32767) atomicMin(&d2key[x - best], ((uint32_t)minS << 16) | (uint32_t)(0xFFFF - best));
            int dsp = best * 16;
            if (best > 0 && best < D - 1) {
                const int sm1 = (int)(b & 0xFFFFu), sp1 = (int)(b >> 16);
                const int den = max(sm1 + sp1 - 2 * minS, 1);
                dsp += ((sm1 - sp1) * 16 + den) / (2 * den);
            }
            disp1s[x] = (int16_t)dsp;
        }
    }
    __syncthreads();
    // ---- LR check (A.4.5)
    int16_t* out = frame_ptr(ws.raw, ws_stride, f) + (size_t)y * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        int d1 = disp1s[x];
        if (d1 != kInv) {
            const int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
            const int _x = x - _d, x_ = x - d_;
            auto d2at = [&](int xx) -> int {
                const uint32_t kk = d2key[xx];
                return kk == kD2Init ? kInv : (int)(0xFFFFu - (kk & 0xFFFFu));
            };
            bool badl = false, badr = false;
            if (_x >= 0 && _x < W) { const int v = d2at(_x); badl = v >= 0 && abs(v - _d) > d.disp12; }
            if (x_ >= 0 && x_ < W) { const int v = d2at(x_); badr = v >= 0 && abs(v - d_) > d.disp12; }
            if (badl && badr) d1 = kInv;
        }
        out[x] = (int16_t)d1;
    }
}

// ------------------------------------------------------------------------------------------------------------
// A.4.6 post filters
// ------------------------------------------------------------------------------------------------------------
// packed compare-exchange: both 16-bit halves (two adjacent pixels) at once
__device__ __forceinline__ void cswap2(uint32_t& a, uint32_t& b) { const uint32_t t = __vmins2(a, b); b = __vmaxs2(a, b); a = t; }

__global__ void __launch_bounds__(256) k_median3(const int16_t* __restrict__ src, size_t src_stride, int16_t* __restrict__ dst,
                                                 size_t dst_stride, int W, int H) {
    // a thread filters two horizontally adjacent pixels: low half = pixel x, high half = pixel x + 1
    const int f = blockIdx.y, PW = (W + 1) >> 1, n = PW * H;
    const uint16_t* s = reinterpret_cast<const uint16_t*>(frame_ptr(src, src_stride, f));
    int16_t* o = frame_ptr(dst, dst_stride, f);
#pragma unroll
    for (int e = 0; e < kEPT; e++) {
        const int i = (blockIdx.x * kEPT + e) * 256 + threadIdx.x;
        if (i >= n) break;
        const int y = i / PW, x = (i - y * PW) * 2;
        const int xm = max(x - 1, 0), x1 = min(x + 1, W - 1), x2 = min(x + 2, W - 1);
        uint32_t v[9];
#pragma unroll
        for (int dy = -1; dy <= 1; dy++) {
            const uint16_t* r = s + (size_t)min(max(y + dy, 0), H - 1) * W;
            const uint32_t a = r[xm], b = r[x], c = r[x1], d = r[x2];
            v[(dy + 1) * 3 + 0] = a | (b << 16);   // left neighbours of (x, x+1)
            v[(dy + 1) * 3 + 1] = b | (c << 16);   // the pixels themselves
            v[(dy + 1) * 3 + 2] = c | (d << 16);   // right neighbours
        }
        // 19-exchange median-of-9 network
        cswap2(v[1], v[2]); cswap2(v[4], v[5]); cswap2(v[7], v[8]); cswap2(v[0], v[1]); cswap2(v[3], v[4]); cswap2(v[6], v[7]);
        cswap2(v[1], v[2]); cswap2(v[4], v[5]); cswap2(v[7], v[8]); cswap2(v[0], v[3]); cswap2(v[5], v[8]); cswap2(v[4], v[7]);
        cswap2(v[3], v[6]); cswap2(v[1], v[4]); cswap2(v[2], v[5]); cswap2(v[4], v[7]); cswap2(v[4], v[2]); cswap2(v[6], v[4]);
        cswap2(v[4], v[2]);
        o[(size_t)y * W + x] = (int16_t)(v[4] & 0xFFFFu);
        if (x + 1 < W) o[(size_t)y * W + x + 1] = (int16_t)(v[4] >> 16);
    }
}

__device__ __forceinline__ int ccl_find(volatile int32_t* L, int i) {
    int p = L[i];
    while (p != i) { i = p; p = L[i]; }
    return i;
}
__device__ __forceinline__ void ccl_union(int32_t* L, int a, int b) {
    bool done;
    do {
        a = ccl_find(L, a);
        b = ccl_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = old == b; b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = old == a; a = old; }
        else done = true;
    } while (!done);
}

// Speckle filter = connected components over 4-neighbour edges |a-b| <= maxDiff (order-independent, so any labelling
// algorithm gives OpenCV's result).  Runs first: every pixel points at the start of its horizontal run (one warp per row,
// ballot + clz), then only the non-redundant vertical edges are united (an edge is redundant when the pixel to the left
// closes a 4-cycle of edges), then sizes are accumulated with warp-aggregated atomics.
__global__ void __launch_bounds__(256) k_ccl_rows(const int16_t* __restrict__ img, int32_t* label, int32_t* csize, int W, int H,
                                                  int maxDiff, size_t img_stride, size_t ws_stride) {
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), f = blockIdx.y;
    if (y >= H) return;
    const int16_t* s = frame_ptr(img, img_stride, f) + (size_t)y * W;
    int32_t* L = frame_ptr(label, ws_stride, f) + (size_t)y * W;
    int32_t* C = frame_ptr(csize, ws_stride, f) + (size_t)y * W;
    int carry_start = -1;   // start x of the run that reaches the end of the previous chunk (-1: none)
    int prev_last = kInv;   // value of the previous chunk's last pixel
    for (int x0 = 0; x0 < W; x0 += 32) {
        const int x = x0 + lane;
        const int v = x < W ? (int)s[x] : kInv;
        int left = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) left = prev_last;
        const bool valid = v != kInv;
        const bool link = valid && left != kInv && abs(v - left) <= maxDiff && x > 0;
        const uint32_t starts = __ballot_sync(0xffffffffu, valid && !link);
        const uint32_t below = starts & (0xFFFFFFFFu >> (31 - lane));
        const int start_x = below ? x0 + 31 - __clz((int)below) : carry_start;
        if (x < W) {
            L[x] = valid ? y * W + start_x : -1;
            C[x] = 0;
        }
        const int last_start = __shfl_sync(0xffffffffu, valid ? start_x : -1, 31);
        carry_start = last_start;
        prev_last = __shfl_sync(0xffffffffu, v, 31);
    }
}
__global__ void __launch_bounds__(256) k_ccl_vmerge(const int16_t* __restrict__ img, int32_t* label, int W, int H, int maxDiff,
                                                    size_t img_stride, size_t ws_stride) {
    const int f = blockIdx.y, n = W * H;
    const int16_t* s = frame_ptr(img, img_stride, f);
#pragma unroll
    for (int e = 0; e < kEPT; e++) {
        const int i = W + (blockIdx.x * kEPT + e) * 256 + threadIdx.x;  // rows 1 .. H-1
        if (i >= n) break;
        const int x = i % W;
        const int v = s[i], u = s[i - W];
        if (v == kInv || u == kInv || abs(v - u) > maxDiff) continue;
        if (x > 0) {
            const int vl = s[i - 1], ul = s[i - W - 1];
            if (vl != kInv && ul != kInv && abs(v - vl) <= maxDiff && abs(u - ul) <= maxDiff && abs(vl - ul) <= maxDiff) continue;
        }
        ccl_union(frame_ptr(label, ws_stride, f), i, i - W);
    }
}
__global__ void __launch_bounds__(256) k_ccl_count(int32_t* label, int32_t* csize, int n, size_t ws_stride) {
    const int f = blockIdx.y;
    int32_t* L = frame_ptr(label, ws_stride, f);
#pragma unroll
    for (int e = 0; e < kEPT; e++) {  // every thread runs all rounds: the warp votes below need the full warp
        const int i = (blockIdx.x * kEPT + e) * 256 + threadIdx.x;
        int r = -1;
        if (i < n && L[i] >= 0) {
            r = ccl_find(L, i);
            L[i] = r;  // only ever lowers a label towards its root: concurrent finds stay valid
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, r);
        if (r >= 0 && (threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&frame_ptr(csize, ws_stride, f)[r], __popc(peers));
    }
}
__global__ void __launch_bounds__(256) k_ccl_apply(const int16_t* __restrict__ img, const int32_t* __restrict__ label,
                                                   const int32_t* __restrict__ csize, int16_t* __restrict__ out, int n, int maxSize,
                                                   size_t img_stride, size_t ws_stride, size_t out_stride) {
    const int f = blockIdx.y;
    const int32_t* L = frame_ptr(label, ws_stride, f);
#pragma unroll
    for (int e = 0; e < kEPT; e++) {
        const int i = (blockIdx.x * kEPT + e) * 256 + threadIdx.x;
        if (i >= n) break;
        const int16_t v = frame_ptr(img, img_stride, f)[i];
        int16_t o = v;
        if (v != kInv) {
            int r = L[i];
            while (L[r] != r) r = L[r];
            if (frame_ptr(csize, ws_stride, f)[r] <= maxSize) o = (int16_t)kInv;
        }
        frame_ptr(out, out_stride, f)[i] = o;
    }
}

template <int NPR, bool PAD>
int launch_paths(const SgbmDims& d, const SgbmWorkspace& ws, size_t ws_stride, int nb, cudaStream_t st) {
    dim3 gv((d.mode ? 6 : 3) * cdiv(d.W1, 8), nb);
    { auto k_sgbm_vert_t = k_sgbm_vert<NPR, PAD>; OVO_LAUNCH(k_sgbm_vert_t, gv, dim3(256), 0, st, d, ws, ws_stride); }
    OVO_LAUNCH_CHECK();
    dim3 gh(d.H, nb);
    const size_t smem = (size_t)d.W * 6 + (size_t)d.W1 * 10 + 16;
    auto go = [&](auto k_sgbm_horiz_t) -> int {
        if (smem > 40 * 1024) OVO_CUDA(cudaFuncSetAttribute(k_sgbm_horiz_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OVO_LAUNCH(k_sgbm_horiz_t, gh, dim3(64), smem, st, d, ws, ws_stride);
        return 0;
    };
    int rc;
    if (d.uniq < 100) rc = d.mode ? go(k_sgbm_horiz<NPR, PAD, true, true>) : go(k_sgbm_horiz<NPR, PAD, false, true>);
    else rc = d.mode ? go(k_sgbm_horiz<NPR, PAD, true, false>) : go(k_sgbm_horiz<NPR, PAD, false, false>);
    if (rc) return rc;
    OVO_LAUNCH_CHECK();
    return 0;
}

template <int SW2, int TX>
int launch_cost(const SgbmDims& d, const SgbmWorkspace& ws, size_t ws_stride, int nb, cudaStream_t st) {
    const int npairs = d.Dp / 2, upb = kCostThreads / npairs;
    const int units = cdiv(d.W1, TX) * cdiv(d.H, kCostRS);
    dim3 grid(cdiv(units, upb), 1, nb), block(npairs, upb);
    if (d.D == d.Dp) { auto k_sgbm_cost_t = k_sgbm_cost<SW2, TX, false>; OVO_LAUNCH(k_sgbm_cost_t, grid, block, 0, st, d, ws, ws_stride); }
    else { auto k_sgbm_cost_t = k_sgbm_cost<SW2, TX, true>; OVO_LAUNCH(k_sgbm_cost_t, grid, block, 0, st, d, ws, ws_stride); }
    OVO_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// checkpoints of the horizontal sweeps: per row 2 x ceil((W1 - W1/2) / K) cost vectors (K = 4 for Dp = 256, else 8)
static size_t ckpt_bytes(const SgbmDims& d) {
    const int K = d.Dp == 256 ? horiz_seg<4>() : horiz_seg<2>();
    const int sph = (d.W1 - (d.W1 >> 1) + K - 1) / K;
    return align_up((size_t)d.H * 2 * sph * d.Dp * 2, 256);
}

size_t sgbm_workspace_bytes(const SgbmDims& d) {
    const size_t vol = align_up((size_t)d.H * d.W1 * d.Dp * 2, 256);
    const size_t img = align_up((size_t)d.H * d.W * 4, 256);
    return align_up(4 * img /*prep*/ + (d.mode ? 7 : 4) * vol /*C + Lv[3]*/ + ckpt_bytes(d) + 4 * img /*raw, med (i16) + label, csize (i32) -> 2*0.5+2 = 3 img*/, 256);
}

void sgbm_carve(const SgbmDims& d, uint8_t* base, SgbmWorkspace* ws) {
    const size_t vol = align_up((size_t)d.H * d.W1 * d.Dp * 2, 256);
    const size_t img = align_up((size_t)d.H * d.W * 4, 256);
    uint8_t* p = base;
    ws->prep = reinterpret_cast<uint32_t*>(p); p += 4 * img;
    ws->C = reinterpret_cast<int16_t*>(p); p += vol;
    ws->Lv = reinterpret_cast<int16_t*>(p); p += (d.mode ? 6 : 3) * vol;
    ws->ckpt = reinterpret_cast<int16_t*>(p); p += ckpt_bytes(d);
    ws->raw = reinterpret_cast<int16_t*>(p); p += img / 2;
    ws->med = reinterpret_cast<int16_t*>(p); p += img / 2;
    ws->label = reinterpret_cast<int32_t*>(p); p += img;
    ws->csize = reinterpret_cast<int32_t*>(p); p += img;
}

int sgbm_launch(const SgbmDims& d, const SgbmWorkspace* ws0, size_t ws_stride, int nb, const uint8_t* left, const uint8_t* right,
                int pitch, size_t frame_stride, int16_t* disp_out, cudaStream_t st) {
    const SgbmWorkspace& ws = *ws0;
    {
        dim3 grid(cdiv(((d.W + 3) / 4) * d.H, 256), nb * 2);
        OVO_LAUNCH(k_sgbm_prep, grid, dim3(256), 0, st, left, right, pitch, frame_stride, d, ws, ws_stride);
        OVO_LAUNCH_CHECK();
    }
    int rc;
    switch (d.bs) {
        case 3: rc = launch_cost<1, 16>(d, ws, ws_stride, nb, st); break;
        case 5: rc = launch_cost<2, 16>(d, ws, ws_stride, nb, st); break;
        case 7: rc = launch_cost<3, 8>(d, ws, ws_stride, nb, st); break;
        case 9: rc = launch_cost<4, 8>(d, ws, ws_stride, nb, st); break;
        case 11: rc = launch_cost<5, 8>(d, ws, ws_stride, nb, st); break;
        default: set_error("blockSize %d unsupported (3,5,7,9,11)", d.bs); return 1;
    }
    if (rc) return rc;
    switch (d.Dp) {
        case 64: rc = d.D == 64 ? launch_paths<1, false>(d, ws, ws_stride, nb, st) : launch_paths<1, true>(d, ws, ws_stride, nb, st); break;
        case 128: rc = d.D == 128 ? launch_paths<2, false>(d, ws, ws_stride, nb, st) : launch_paths<2, true>(d, ws, ws_stride, nb, st); break;
        case 256: rc = d.D == 256 ? launch_paths<4, false>(d, ws, ws_stride, nb, st) : launch_paths<4, true>(d, ws, ws_stride, nb, st); break;
        default: set_error("padded disparity range %d unsupported", d.Dp); return 1;
    }
    if (rc) return rc;
    const int n = d.W * d.H;
    const size_t out_stride = (size_t)n * 2;
    dim3 gimg(cdiv(((d.W + 1) / 2) * d.H, 256 * kEPT), nb);
    if (d.speckleWin <= 0) {
        OVO_LAUNCH(k_median3, gimg, dim3(256), 0, st, ws.raw, ws_stride, disp_out, out_stride, d.W, d.H);
        OVO_LAUNCH_CHECK();
        return 0;
    }
    OVO_LAUNCH(k_median3, gimg, dim3(256), 0, st, ws.raw, ws_stride, ws.med, ws_stride, d.W, d.H);
    OVO_LAUNCH_CHECK();
    dim3 glin(cdiv(n, 256 * kEPT), nb);
    OVO_LAUNCH(k_ccl_rows, dim3(cdiv(d.H, 8), nb), dim3(256), 0, st, ws.med, ws.label, ws.csize, d.W, d.H, d.speckleDiff, ws_stride, ws_stride);
    OVO_LAUNCH_CHECK();
    if (d.H > 1) {
        OVO_LAUNCH(k_ccl_vmerge, dim3(cdiv(n - d.W, 256 * kEPT), nb), dim3(256), 0, st, ws.med, ws.label, d.W, d.H, d.speckleDiff, ws_stride, ws_stride);
        OVO_LAUNCH_CHECK();
    }
    OVO_LAUNCH(k_ccl_count, glin, dim3(256), 0, st, ws.label, ws.csize, n, ws_stride);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_ccl_apply, glin, dim3(256), 0, st, ws.med, ws.label, ws.csize, disp_out, n, d.speckleWin, ws_stride, ws_stride, out_stride);
    OVO_LAUNCH_CHECK();
    return 0;
}

}  // namespace ovo
