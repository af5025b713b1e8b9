// StereoSGBM (MODE_SGBM, 5 directions) for sm_100a — stage 4 of the openVO hot path.
//
// Replaces cv2.StereoSGBM.compute as called by the reference at src/openVO/stereo_camera.py:51 (object built at
// :23-27).  Semantics: SURVEY.md Appendix A.4 (bit-exact; oracle = oracle/sgbm_restate.cpp).
//
// Data flow (per frame, all in HBM/L2; D padded to Dp in {64,128,256}):
//   k_sgbm_prep   images -> byte-packed (v, vmin, vmax) planes for the Sobel-x-clipped and the raw rows (A.4.1)
//   k_sgbm_cost   prep   -> C[y][x1][word] int16x2 : Birchfield-Tomasi cost summed over the blockSize^2 window (A.4.2)
//   k_sgbm_vsum   C      -> Sv[y][x1][word] = sat(L1 + L2 + L3): the three top-down paths (from (x-1,y-1), (x,y-1), (x+1,y-1)) of
//                                               a band of 16 rows x a tile of columns per CTA; a warp carries the path state of a
//                                               few adjacent scan lines in registers and follows them through the band; the three
//                                               directions of a cell meet in a three-row shared-memory ring (V stores, the first
//                                               diagonal adds, the second adds and writes Sv), so only ONE volume leaves the chip.
//                                               Diagonal lines that enter a tile from the side are recomputed from the band's top
//                                               row (a 16-column halo); path state crosses bands through a small L2-resident buffer
//   k_sgbm_vert   C      -> Lv[6][y][x1][word] : MODE_HH only (opt-in): one volume per direction, one warp per scan line
//   k_sgbm_horiz  C, Sv  -> raw disparity     : paths from (x-1,y) and (x+1,y) run towards each other by two warps
//                                               per row, parking their state every 4 cells; whoever reaches a cell
//                                               second replays the other's path from the checkpoint, owns the complete
//                                               5-path sum and does WTA / uniqueness / sub-pixel / disp2; then the LR check
//   k_median3, k_ccl_*   -> 3x3 median and speckle filter (connected components, union-find)
// Cost-vector layout ("half-split"): 32-bit word k of a cell (k < Dp/2) holds disparity k in its low and disparity k + Dp/2 in
// its high half.  The d-1 / d+1 neighbours of a word are then simply the previous / next WORD, so the path recurrence needs
// no funnel shifts; only the first and the last word of a vector need a byte permute (MAX_COST enters there).
// All cost arithmetic is packed 2 x u16 per register on the DPX pipe (VIADDMNMX.U16x2 / VIMNMX.U16x2); the per-cell
// minimum is one CREDUX.MIN.  No tensor cores: nothing here is a contraction.
#include <cstdlib>
#include <cstring>
#include <type_traits>

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
// A.4.1 + A.4.2 cost volume.  A thread owns one word (disparities k and k + Dp/2) of one unit = (TX columns) x (RS rows); it
// walks the rows of the strip, and inside a row the TX + 2*SW2 columns, keeping the horizontal window in registers
// and the vertical window as a ring of horizontal sums in (thread-private, conflict-free) shared memory.
// ------------------------------------------------------------------------------------------------------------
constexpr int kCostThreads = 128;
#ifndef OVO_COST_RS
#define OVO_COST_RS 32
#endif
constexpr int kCostRS = OVO_COST_RS;  // rows per unit (the vertical window adds 2*SW2 halo rows)

__device__ __forceinline__ uint32_t bt_pair(uint32_t lw, uint32_t rw0, uint32_t rw1) {
    // lw: left word at x; rw0 / rw1: right words at x - dlo and x - dhi
    const uint32_t v = __byte_perm(rw0, rw1, 0x7430), vmin = __byte_perm(rw0, rw1, 0x7531), vmax = __byte_perm(rw0, rw1, 0x7632);
    const uint32_t u = __byte_perm(lw, 0, 0x4040), umin = __byte_perm(lw, 0, 0x4141), umax = __byte_perm(lw, 0, 0x4242);
    const uint32_t c0 = __vmaxu2(vmin, u) - __vminu2(vmax, u);  // max(0, u - vmax, vmin - u), both halves
    const uint32_t c1 = __vmaxu2(umin, v) - __vminu2(umax, v);  // max(0, v - umax, umin - v)
    return __vminu2(c0, c1);
}

// one row of a unit: horizontal sums of TX columns, folded into the vertical ring / running sums.  The thread owns word k of
// every cell: disparities dlo = k (low half) and dhi = k + Dp/2 (high half); a padded half (d >= D) is computed from a valid
// address and ignored downstream (the path kernels force MAX_COST there).
template <int SW2, int TX, bool EDGE, bool PAD>
__device__ __forceinline__ void cost_row(const uint2* __restrict__ Lrow, const uint2* __restrict__ Rrow, int x0, int dlo, int dhi, int D,
                                         int W1, bool pad, uint32_t* slot, bool have_old, uint32_t (&vs)[TX]) {
    constexpr int BS = 2 * SW2 + 1;
    uint32_t win[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) win[i] = 0;
    uint32_t hs = 0;
    // interior tiles: all TX + 2*SW2 columns are in range, so every address is a row base plus a compile-time offset
    const uint2* Lb = Lrow + x0 + D;
    const uint2* Rb0 = Rrow + x0 + D - dlo;
    const uint2* Rb1 = Rrow + x0 + D - dhi;
#pragma unroll
    for (int j = -SW2; j < TX + SW2; j++) {
        uint32_t pix = 0;
        if (!PAD || !pad) {
            uint2 lw, r0, r1;
            OVO_DEVCHECK(x0 + j + D - dhi >= 0 || EDGE);
            OVO_DEVCHECK((EDGE ? min(x0 + j, W1 - 1) : x0 + j) + D < W1 + D && dlo <= dhi && dhi < D);
            if (EDGE) {
                const int xx = min(max(x0 + j, 0), W1 - 1);
                lw = __ldg(Lrow + xx + D);
                r0 = __ldg(Rrow + xx + D - dlo);
                r1 = __ldg(Rrow + xx + D - dhi);
            } else {
                lw = __ldg(Lb + j);
                r0 = __ldg(Rb0 + j);
                r1 = __ldg(Rb1 + j);
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
    const int dlo = threadIdx.x;                                  // word k: disparities k and k + Dp/2
    const bool pad = dlo >= D;                                    // both halves padded
    const int dhi = dlo + npairs < D ? dlo + npairs : dlo;        // a padded high half reads a valid address; its value is ignored
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
        if (edge) cost_row<SW2, TX, true, PAD>(Lrow, Rrow, x0, dlo, dhi, D, W1, pad, slot, k >= BS, vs);
        else cost_row<SW2, TX, false, PAD>(Lrow, Rrow, x0, dlo, dhi, D, W1, pad, slot, k >= BS, vs);
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
// A.4.3 one path step on a cost vector held by LPC lanes (LPC = 32: one cell per warp; 8 or 16: 4 or 2 adjacent cells per
// warp), NPR words per lane: lane q of a cell owns words [q*NPR, (q+1)*NPR), word k = (disparity k | disparity k + Dp/2 << 16).
// A predecessor outside the image is the all-zero vector with min 0, for which the step formula yields L = C, so path
// (re)starts are just a state reset followed by the ordinary step.
// ------------------------------------------------------------------------------------------------------------
template <int NPR>
struct PathState {
    uint32_t L[NPR];
    uint32_t mm;  // min over d of L, in both halves (uniform over the cell's lanes)
};

template <int NPR>
__device__ __forceinline__ void path_reset(PathState<NPR>& s) {
#pragma unroll
    for (int r = 0; r < NPR; r++) s.L[r] = 0;
    s.mm = 0;
}

// lane-constant operands of path_step
template <int LPC>
struct PathLane {
    int src_below, src_above;      // lanes holding word k-1 of my first word / word k+1 of my last word (wrapping inside the cell)
    uint32_t sel_below, sel_above; // byte-permute selectors: identity, except at the two ends of the vector where MAX_COST enters
    uint32_t gmask;                // the lanes of my cell
    __device__ __forceinline__ void init(int lane) {
        const int q = lane & (LPC - 1), base = lane & ~(LPC - 1);
        src_below = base | ((q - 1) & (LPC - 1));
        src_above = base | ((q + 1) & (LPC - 1));
        // word -1 = (MAX_COST, disparity Dp/2 - 1 = low half of the last word); word Dp/2 = (disparity Dp/2 = high half of word 0, MAX_COST)
        sel_below = q == 0 ? 0x1054u : 0x3210u;
        sel_above = q == LPC - 1 ? 0x7632u : 0x3210u;
        gmask = LPC == 32 ? 0xffffffffu : (((1u << (LPC & 31)) - 1u) << base);
    }
};

template <int LPC, int NPR>
__device__ __forceinline__ uint32_t vec_min(const uint32_t (&L)[NPR], const PathLane<LPC>& pl) {  // min over d, broadcast to both halves
    uint32_t t = L[0];
#pragma unroll
    for (int r = 1; r < NPR; r++) t = __vminu2(t, L[r]);
    t = __vminu2(t, __byte_perm(t, 0, 0x1032));  // both halves = min(lo, hi): 32-bit order == 16-bit order from here on
    if constexpr (LPC == 32) {
        return __reduce_min_sync(0xffffffffu, t);
    } else {
        // several cells per warp: REDUX wants one mask for the whole warp (per-cell masks take ptxas' divergent slow path), so
        // the cell's lanes run a butterfly instead
#pragma unroll
        for (int o = 1; o < LPC; o <<= 1) t = min(t, __shfl_xor_sync(0xffffffffu, t, o));
        return t;
    }
}

template <int LPC, int NPR, bool PAD>
__device__ __forceinline__ void path_step(PathState<NPR>& s, const uint32_t (&c)[NPR], const uint32_t (&padmask)[NPR],
                                          uint32_t P1P1, uint32_t P2P2, const PathLane<LPC>& pl) {
    uint32_t below = __shfl_sync(0xffffffffu, s.L[NPR - 1], pl.src_below);
    uint32_t above = __shfl_sync(0xffffffffu, s.L[0], pl.src_above);
    below = __byte_perm(below, kMaxC2, pl.sel_below);  // Lp[-1] = MAX_COST
    above = __byte_perm(above, kMaxC2, pl.sel_above);  // Lp[Dp] = MAX_COST
    const uint32_t mP2 = s.mm + P2P2;
    uint32_t out[NPR];
#pragma unroll
    for (int r = 0; r < NPR; r++) {
        const uint32_t dm1 = r == 0 ? below : s.L[r - 1];        // Lp[d-1]
        const uint32_t dp1 = r == NPR - 1 ? above : s.L[r + 1];  // Lp[d+1]
        // min(Lp[d-1] + P1, Lp[d+1] + P1, Lp[d], m + P2): the two neighbours share their "+ P1" (a plain add: no half can carry,
        // MAX_COST + P1 < 2^16), then one three-way minimum — two half-rate DPX instructions per word instead of three
        const uint32_t t = __vimin3_u16x2(__vminu2(dm1, dp1) + P1P1, s.L[r], mP2);
        out[r] = c[r] + (t - s.mm);
        if (PAD) out[r] |= padmask[r];
    }
#pragma unroll
    for (int r = 0; r < NPR; r++) s.L[r] = out[r];
    s.mm = vec_min<LPC, NPR>(s.L, pl);
}

template <int LPC, int NPR>
__device__ __forceinline__ void make_padmask(uint32_t (&padmask)[NPR], int lane, int D) {
    const int q = lane & (LPC - 1);
#pragma unroll
    for (int r = 0; r < NPR; r++) {
        const int k = q * NPR + r;
        padmask[r] = (k >= D ? 0x7FFFu : 0u) | (k + LPC * NPR >= D ? 0x7FFF0000u : 0u);
    }
}

// a + b per 16-bit half, saturating at MAX_COST unless the parameters guarantee that no sum of path costs can reach it
// (launch_paths: 5 * (blockSize^2 * (2 * ftzero + 63) + P2) <= 32000); the plain add is a full-rate instruction, the
// saturating one a half-rate DPX instruction
template <bool SAT>
__device__ __forceinline__ uint32_t sum16(uint32_t a, uint32_t b) {
    return SAT ? __viaddmin_u16x2(a, b, kMaxC2) : a + b;
}

template <int NPR>
__device__ __forceinline__ void ldv(uint32_t (&v)[NPR], const uint32_t* p) {
    if constexpr (NPR % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NPR; i += 4) {
            const uint4 t = *reinterpret_cast<const uint4*>(p + i);
            v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        }
    } else if constexpr (NPR == 2) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = *p;
    }
}
template <int NPR>
__device__ __forceinline__ void stv(uint32_t* p, const uint32_t (&v)[NPR]) {
    if constexpr (NPR % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NPR; i += 4) *reinterpret_cast<uint4*>(p + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if constexpr (NPR == 2) {
        *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1]);
    } else {
        *p = v[0];
    }
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
    make_padmask<32, NPR>(padmask, lane, d.D);
    const uint32_t P1P1 = bcast16(d.P1), P2P2 = bcast16(d.P2);
    PathLane<32> plane;
    plane.init(lane);

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
                path_step<32, NPR, PAD>(s, c, padmask, P1P1, P2P2, plane);
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
                path_step<32, NPR, PAD>(s, c, padmask, P1P1, P2P2, plane);
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
                path_step<32, NPR, PAD>(s, c, padmask, P1P1, P2P2, plane);
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
// Sv = sat(L1 + L2 + L3) for a band of BR rows x a tile of B columns per CTA (MODE_SGBM).  A "line group" is CPW = 32 / LPC
// adjacent scan lines of one direction carried by one warp (LPC lanes per cell, NPR words per lane): V groups stay on their
// columns, D1 groups (path from (x-1, y-1)) move one column to the right per row, D3 groups (from (x+1, y-1)) one to the left.
// A tile has 16 V groups (B = 16 * CPW columns) and 16 + BR / CPW groups per diagonal direction: the extra ones start in the
// BR columns beside the tile and are only there to carry the state of the lines that enter the tile further down (the halo is
// recomputed instead of exchanged between CTAs, so a band is ONE launch without inter-CTA synchronisation).  Every warp owns two
// groups of the same direction and walks the band's rows at its own pace.  The three directions of a cell meet in a three-row
// shared-memory ring: the V warps store L2 of row r, the D1 warps add L1, the D3 warps add L3 and write the finished sum to
// Sv.  The hand-overs are producer / consumer named barriers (bar.arrive by the warps that wrote, bar.sync by the warps that
// read; one barrier per edge V->D1, D1->D3, D3->V and ring row), never a CTA-wide barrier, so the warps of one direction may
// run up to a few rows ahead of the next one and their shared-memory and arithmetic phases overlap instead of alternating.
// The state of the lines at the last row of the band goes to a small buffer (3 x W1 vectors per frame and parity) from which
// the next band's launch starts, so path state never leaves registers inside a band and Sv is the only volume written.
// ------------------------------------------------------------------------------------------------------------
template <int LPC>
__host__ __device__ constexpr int vs_cpw() { return 32 / LPC; }
// VG: V groups per tile (tile width B = VG * CPW columns); BR: rows per band = columns of halo on each side
template <int LPC, int VG>
__host__ __device__ constexpr int vs_tile() { return VG * vs_cpw<LPC>(); }
template <int LPC, int BR, int VG>
__host__ __device__ constexpr int vs_diag_groups() { return VG + BR / vs_cpw<LPC>(); }
constexpr int kVsSlots = 2;  // line groups per warp (both of the warp's direction)
template <int LPC, int BR, int VG>
__host__ __device__ constexpr int vs_groups() { return VG + 2 * vs_diag_groups<LPC, BR, VG>(); }
template <int LPC, int BR, int VG>
__host__ __device__ constexpr int vs_warps() { return vs_groups<LPC, BR, VG>() / kVsSlots; }
#ifndef OVO_VS_MINB
#define OVO_VS_MINB 2
#endif
template <int LPC, int BR, int VG>
__host__ __device__ constexpr int vs_ctas_per_sm() { return vs_warps<LPC, BR, VG>() <= 16 ? OVO_VS_MINB : 1; }

// ---- bulk-copy (TMA 1-D) + mbarrier helpers: the C strip of a row is brought into shared memory by the copy engine ----
#ifndef OVO_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    // bounded spin: a lost transaction traps instead of hanging the device
#pragma unroll 1
    for (int spin = 0; spin < (1 << 26); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
        if (ok) return;
    }
    __trap();
}
#endif

// producer / consumer named barriers (ids 1..15; 0 is __syncthreads): count = threads of the warps that arrive + of those that wait
#ifndef OVO_EMU
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
#else
__device__ __forceinline__ void nbar_arrive(int id, int count) { ovo_emu::named_barrier(id, count, false); }
__device__ __forceinline__ void nbar_sync(int id, int count) { ovo_emu::named_barrier(id, count, true); }
#endif

constexpr size_t kStripSlackFront = 8 * 1024, kStripSlackBack = 32 * 1024;  // see sgbm_workspace_bytes
#ifndef OVO_VS_RING
#define OVO_VS_RING 3
#endif
#ifndef OVO_VS_AHEAD
#define OVO_VS_AHEAD 1
#endif
constexpr int kVsRing = OVO_VS_RING;    // rows of the shared-memory ring in which the three directions of a cell meet: the V warps may
                                        // be this many rows ahead of the D3 warps
constexpr int kVsAhead = OVO_VS_AHEAD;  // the strip of row r + kVsAhead is requested when the first V warp starts row r
constexpr int kVsStrips = kVsRing + kVsAhead;  // C strips in shared memory: rows r - kVsRing + 1 .. r in use, kVsAhead more in flight
static_assert(3 * kVsRing <= 15, "named barriers 1 .. 15");

template <int NPR>
struct VsSlot {
    PathState<NPR> s;
    int x;          // column of this lane's cell at the band's first row
    int scol;       // word offset of this lane's words inside a strip row: (x - (x0 - BR)) * DH + q * NPR
    int32_t off;    // word offset of this lane's words in Sv for the current row (a frame's volume is < 2^31 words)
};

// shared-memory position (in words) of chunk i (4 words) of lane q inside a cell of the ring: for a fixed i the lanes of a cell
// hit consecutive 16-byte chunks (conflict-free 128-bit accesses)
template <int LPC>
__device__ __forceinline__ int vs_chunk(int q, int i) { return (i * LPC + q) * 4; }

// The three directions of a cell meet in the ring: V (KIND 0) stores, D1 (1) adds, D3 (2) adds and writes the sum to Sv.
template <int KIND, int LPC, int NPR, bool SAT>
__device__ __forceinline__ void vs_meet(const PathState<NPR>& st, uint32_t* cell, int q, uint32_t* __restrict__ sv) {
    if (KIND == 0) {
#pragma unroll
        for (int i = 0; i < NPR / 4; i++)
            *reinterpret_cast<uint4*>(cell + vs_chunk<LPC>(q, i)) = make_uint4(st.L[4 * i], st.L[4 * i + 1], st.L[4 * i + 2], st.L[4 * i + 3]);
    } else {
        uint32_t a[NPR];
#pragma unroll
        for (int i = 0; i < NPR / 4; i++) {
            const uint4 v = *reinterpret_cast<const uint4*>(cell + vs_chunk<LPC>(q, i));
            a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < NPR; k++) a[k] = sum16<SAT>(a[k], st.L[k]);
        if (KIND == 1) {
#pragma unroll
            for (int i = 0; i < NPR / 4; i++)
                *reinterpret_cast<uint4*>(cell + vs_chunk<LPC>(q, i)) = make_uint4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
        } else {
            stv<NPR>(sv, a);
        }
    }
}

template <int LPC, int NPR, bool PAD, int BR, int VG, bool SAT>
__global__ void __launch_bounds__(32 * vs_warps<LPC, BR, VG>(), vs_ctas_per_sm<LPC, BR, VG>())
    k_sgbm_vsum(SgbmDims d, SgbmWorkspace ws, size_t ws_stride, int y0, int parity) {
    constexpr int CPW = vs_cpw<LPC>(), B = vs_tile<LPC, VG>(), DH = LPC * NPR, NDG = vs_diag_groups<LPC, BR, VG>();
    constexpr int SW = B + 2 * BR;  // cells of a C strip: the tile and the halo on both sides
    constexpr int WV = VG / kVsSlots, WD = NDG / kVsSlots;  // warps per direction
    static_assert(NPR % 4 == 0, "128-bit accesses");
    static_assert(VG % kVsSlots == 0 && NDG % kVsSlots == 0, "a warp's groups are of one direction");
    static_assert((size_t)BR * DH * 4 <= kStripSlackFront && (size_t)(B + BR) * DH * 4 <= kStripSlackBack, "strip overhang vs workspace slack");
    OVO_DYN_SMEM(uint32_t, smem);
    uint32_t* strips = smem;                           // [kVsStrips][SW][DH]  (first: the bulk copies want 16-byte alignment)
    uint32_t* ring = smem + kVsStrips * SW * DH;       // [kVsRing][B][DH]
#ifndef OVO_EMU
    uint64_t* bar = reinterpret_cast<uint64_t*>(ring + kVsRing * B * DH);  // [kVsStrips]
#endif
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = lane & (LPC - 1), g = lane / LPC;
    const int f = blockIdx.y, x0 = blockIdx.x * B;
    const int W1 = d.W1, H = d.H;
    const int R = min(BR, H - y0);
    const size_t vol = (size_t)H * W1 * DH;  // words
    const uint32_t* __restrict__ C = reinterpret_cast<const uint32_t*>(frame_ptr(ws.C, ws_stride, f));
    uint32_t* Sv = reinterpret_cast<uint32_t*>(frame_ptr(ws.Lv, ws_stride, f));
    uint32_t* bb = Sv + vol;  // band-state buffer [2 parity][3 kinds][W1][DH], right behind Sv
    const uint32_t* bb_in = bb + (size_t)(parity ^ 1) * 3 * W1 * DH;
    uint32_t* bb_out = bb + (size_t)parity * 3 * W1 * DH;
    // strip of band row rn: cells [x0 - BR, x0 + B + BR) of image row y0 + rn (running over the row ends into the neighbouring
    // rows; the frame's workspace surrounds C, so the source is always valid memory, and those cells are never stored)
    const uint32_t* strip_src = C + ((ptrdiff_t)y0 * W1 + (x0 - BR)) * DH;
#ifdef OVO_EMU
    __shared__ volatile int emu_landed[kVsStrips];  // image row whose strip sits in each slot (stands in for the mbarrier phase)
#endif
    auto fetch = [&](int rn) {  // thread 0 only
        if (rn >= R) return;
        const uint32_t* src = strip_src + (ptrdiff_t)rn * W1 * DH;
        uint32_t* dst = strips + (rn % kVsStrips) * SW * DH;
        // the copy may run over the ends of C, never out of the frame's workspace (sgbm_workspace_bytes)
        OVO_DEVCHECK(src >= reinterpret_cast<const uint32_t*>(frame_ptr(ws.prep, ws_stride, f)) &&
                     src + SW * DH <= reinterpret_cast<const uint32_t*>(frame_ptr(ws.prep, ws_stride, f)) + ws_stride / 4);
        OVO_DEVCHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#ifndef OVO_EMU
        mbar_expect_tx(&bar[rn % kVsStrips], SW * DH * 4);
        bulk_g2s(dst, src, SW * DH * 4, &bar[rn % kVsStrips]);
#else
        for (int i = 0; i < SW * DH; i++) dst[i] = src[i];  // the copy engine, emulated: done at once
        emu_landed[rn % kVsStrips] = rn;
#endif
    };
#ifndef OVO_EMU
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kVsStrips; i++) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
#else
    if (threadIdx.x == 0)
        for (int i = 0; i < kVsStrips; i++) emu_landed[i] = -1;
#endif
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kVsAhead; i++) fetch(i);
    }

    uint32_t padmask[NPR];
    make_padmask<LPC, NPR>(padmask, lane, d.D);
    const uint32_t P1P1 = bcast16(d.P1), P2P2 = bcast16(d.P2);
    PathLane<LPC> pl;
    pl.init(lane);
    const bool edge = x0 - BR <= 0 || x0 + B + BR >= W1;  // diagonal lines of this tile can start at the image border
    const int mine_lo = BR * DH + q * NPR;                // scol of the tile's first column for this lane

    // One direction: KIND 0 = V (dx 0), 1 = D1 (dx +1), 2 = D3 (dx -1); gw = the warp's index among the warps of its direction.
    // Barrier ids (i = ring row): 1 + i: V stored ring row i;  1 + kVsRing + i: D1 added to it;  1 + 2 kVsRing + i: D3 is done with it
    // (and with the strip of the same image row, whose slot the copy of row r + kVsAhead then reuses).
    auto run = [&](auto kind_t, auto edge_t, int gw) {
        constexpr int KIND = decltype(kind_t)::value;
        constexpr bool EDGE = decltype(edge_t)::value;
        constexpr int DX = KIND == 0 ? 0 : KIND == 1 ? 1 : -1;
        constexpr int NA = (WV + WD) * 32, NB = 2 * WD * 32, NC = (WD + WV) * 32;
        VsSlot<NPR> sl[kVsSlots];
#pragma unroll
        for (int k = 0; k < kVsSlots; k++) {
            const int gi = gw * kVsSlots + k;  // group of this direction
            const int xs = KIND == 1 ? x0 - BR + gi * CPW : x0 + gi * CPW;
            sl[k].x = xs + g;
            sl[k].scol = (sl[k].x - (x0 - BR)) * DH + q * NPR;
            sl[k].off = (int32_t)(((ptrdiff_t)y0 * W1 + sl[k].x) * DH + q * NPR);
            // state of the predecessor cell (x - dx, y0 - 1), or the all-zero vector when it lies outside the image
            const int xp = sl[k].x - DX;
            const bool have = y0 > 0 && xp >= 0 && xp < W1;
            uint32_t v[NPR];
#pragma unroll
            for (int r = 0; r < NPR; r++) v[r] = 0;
            if (have) ldv<NPR>(v, bb_in + ((size_t)KIND * W1 + xp) * DH + q * NPR);
#pragma unroll
            for (int r = 0; r < NPR; r++) sl[k].s.L[r] = v[r];
            sl[k].s.mm = vec_min<LPC, NPR>(sl[k].s.L, pl);
        }
        int sbase = 0, rbase = 0, ri = 0;  // strip / ring row of the current image row, r % kVsRing
#pragma unroll 1
        for (int r = 0; r < R; r++) {
            if (KIND == 0) {
                // ring row r % kVsRing and the strip slot of row r + kVsAhead were last used for row r - kVsRing: wait until D3 is done with it
                if (r >= kVsRing) nbar_sync(1 + 2 * kVsRing + ri, NC);
                if (threadIdx.x == 0) fetch(r + kVsAhead);
            }
#ifndef OVO_EMU
            mbar_wait(&bar[r % kVsStrips], (r / kVsStrips) & 1);
#else
            ovo_emu::spin_until(&emu_landed[r % kVsStrips], r);
#endif
            uint32_t c[kVsSlots][NPR];
#pragma unroll
            for (int k = 0; k < kVsSlots; k++) {
                OVO_DEVCHECK(sl[k].scol >= 0 && sl[k].scol + NPR <= SW * DH && sbase == (r % kVsStrips) * SW * DH);
                ldv<NPR>(c[k], strips + sbase + sl[k].scol);
            }
            bool mine[kVsSlots];
#pragma unroll
            for (int k = 0; k < kVsSlots; k++) {
                mine[k] = (unsigned)(sl[k].scol - mine_lo) < (unsigned)(B * DH);
                if (EDGE && KIND != 0) {
                    const int xc = sl[k].x + DX * r;
                    // a diagonal line enters the image here: its predecessor lies outside (all-zero vector)
                    if (xc == (DX > 0 ? 0 : W1 - 1)) path_reset<NPR>(sl[k].s);
                    mine[k] = mine[k] && xc < W1;
                } else if (EDGE) {
                    mine[k] = mine[k] && sl[k].x < W1;
                }
                path_step<LPC, NPR, PAD>(sl[k].s, c[k], padmask, P1P1, P2P2, pl);
            }
            if (KIND == 1) nbar_sync(1 + ri, NA);
            if (KIND == 2) nbar_sync(1 + kVsRing + ri, NB);
#pragma unroll
            for (int k = 0; k < kVsSlots; k++) {
                if (mine[k]) {
                    OVO_DEVCHECK(rbase == (r % kVsRing) * B * DH && sl[k].scol - mine_lo >= 0 && sl[k].scol - mine_lo + DH - q * NPR <= B * DH);
                    OVO_DEVCHECK(sl[k].off >= 0 && (size_t)sl[k].off + NPR <= vol &&
                                 sl[k].off == (int32_t)(((ptrdiff_t)(y0 + r) * W1 + sl[k].x + DX * r) * DH + q * NPR));
                    vs_meet<KIND, LPC, NPR, SAT>(sl[k].s, ring + rbase + (sl[k].scol - mine_lo), q, Sv + sl[k].off);
                    if (r == R - 1) {
                        const int xc = sl[k].x + DX * r;
                        OVO_DEVCHECK(xc >= 0 && xc < W1);
                        stv<NPR>(bb_out + ((size_t)KIND * W1 + xc) * DH + q * NPR, sl[k].s.L);
                    }
                }
                sl[k].scol += DX * DH;
                sl[k].off += (W1 + DX) * DH;
            }
            if (KIND == 0) nbar_arrive(1 + ri, NA);
            if (KIND == 1) nbar_arrive(1 + kVsRing + ri, NB);
            if (KIND == 2 && r + kVsRing < R) nbar_arrive(1 + 2 * kVsRing + ri, NC);  // (nobody waits for the band's last rows)
            sbase = sbase == (kVsStrips - 1) * SW * DH ? 0 : sbase + SW * DH;
            rbase = rbase == (kVsRing - 1) * B * DH ? 0 : rbase + B * DH;
            ri = ri == kVsRing - 1 ? 0 : ri + 1;
        }
    };
    using T = std::true_type;
    using F = std::false_type;
    using K0 = std::integral_constant<int, 0>;
    using K1 = std::integral_constant<int, 1>;
    using K2 = std::integral_constant<int, 2>;
    if (wid < WV) {
        if (edge) run(K0(), T(), wid); else run(K0(), F(), wid);
    } else if (wid < WV + WD) {
        if (edge) run(K1(), T(), wid - WV); else run(K1(), F(), wid - WV);
    } else {
        if (edge) run(K2(), T(), wid - WV - WD); else run(K2(), F(), wid - WV - WD);
    }
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

// value and disparity of half h of this lane's word r (one cell per warp): word k = lane*NPR + r, d = k + h * Dp/2
template <int NPR>
__device__ __forceinline__ uint32_t half_of(const uint32_t (&S)[NPR], int r, int h) {
    return h ? (S[r] >> 16) : (S[r] & 0xFFFFu);
}

// Per-cell selection, serial part only: argmin, the runner-up over |d-best| > 1 (for the uniqueness test) and the two
// neighbours of the minimum are reduced here and parked in shared memory; the uniqueness decision, the sub-pixel
// division and the disp2 update are order-independent and run data-parallel over the row afterwards.
template <int NPR, bool PAD>
__device__ __forceinline__ void wta_cell(const uint32_t (&S)[NPR], int lane, const SgbmDims& d, int x1, uint32_t* selA,
                                         uint32_t* selB, uint16_t* selBest) {
    constexpr int DH = 32 * NPR;
    const int D = d.D;
    const int k0 = NPR * lane;
    // first d minimising S: the smallest (S << 9 | d)
    uint32_t kbest = 0xFFFFFFFFu;
#pragma unroll
    for (int r = 0; r < NPR; r++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int dd = k0 + r + h * DH;
            const uint32_t key = half_of<NPR>(S, r, h) * 512u + (uint32_t)dd;
            if (!PAD || dd < D) kbest = min(kbest, key);
        }
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, kbest);
    const int minS = (int)(kmin >> 9), best = (int)(kmin & 511u);
    const int fac = 100 - d.uniq;
    uint32_t m2 = 0xFFFFu;
    if (fac > 0) {  // exists d, |d-best|>1, S[d]*fac < minS*100  <=>  (min over those d) * fac < minS*100
#pragma unroll
        for (int r = 0; r < NPR; r++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int dd = k0 + r + h * DH;
                const bool far = (uint32_t)(dd - best + 1) > 2u && (!PAD || dd < D);
                m2 = min(m2, far ? half_of<NPR>(S, r, h) : 0xFFFFu);
            }
        m2 = __reduce_min_sync(0xffffffffu, m2);
    } else {        // uniquenessRatio >= 100: evaluate the predicate as written; park 0 = reject, 0xFFFF = accept
        bool bad = false;
#pragma unroll
        for (int r = 0; r < NPR; r++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int dd = k0 + r + h * DH;
                if ((!PAD || dd < D) && (int)half_of<NPR>(S, r, h) * fac < minS * 100 && abs(dd - best) > 1) bad = true;
            }
        m2 = __any_sync(0xffffffffu, bad) ? 0u : 0xFFFFu;
    }
    auto at = [&](int dd) -> uint32_t {  // S[dd], dd warp-uniform
        const int k = dd % DH, r = k % NPR;
        uint32_t w = S[0];
#pragma unroll
        for (int rr = 1; rr < NPR; rr++)
            if (r == rr) w = S[rr];
        w = __shfl_sync(0xffffffffu, w, k / NPR);
        return dd >= DH ? (w >> 16) : (w & 0xFFFFu);
    };
    const uint32_t sm1 = at(max(best - 1, 0)), sp1 = at(min(best + 1, D - 1));
    if (lane == 0) {
        selA[x1] = (uint32_t)minS | (m2 << 16);
        selB[x1] = sm1 | (sp1 << 16);
        selBest[x1] = (uint16_t)best;
    }
}

// Selection for the K = 4 cells of a segment at once (uniquenessRatio < 100, the usual case).  The warp has parked the
// four S vectors in shared memory; 8 lanes share a cell, each scanning Dp/16 consecutive words (= that many disparities of the
// lower and of the upper half of the range), so the reductions, the neighbour look-ups and the stores are paid once per four
// cells.  Padded disparities hold MAX_COST and the larger d, so they can neither win the argmin nor lower the runner-up.
template <int NPR>
__device__ __forceinline__ void wta_batch(const uint32_t* svec, int lane, const SgbmDims& d, int xo, int dirx, int cnt, uint32_t* selA,
                                          uint32_t* selB, uint16_t* selBest) {
    constexpr int WPC = 32 * NPR, NW = 4 * NPR;  // words per cell (= Dp/2), words per lane
    const int g = lane >> 3, q = lane & 7;
    uint32_t w[NW];
#pragma unroll
    for (int i = 0; i < NW; i += 4) {
        const uint4 t = *reinterpret_cast<const uint4*>(svec + g * WPC + q * NW + i);
        w[i] = t.x; w[i + 1] = t.y; w[i + 2] = t.z; w[i + 3] = t.w;
    }
    // first d minimising S: the smallest (S << 9 | d); word q*NW + i holds d = q*NW + i and d + WPC
    uint32_t kbest = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        kbest = min(kbest, (w[i] & 0xFFFFu) * 512u + (uint32_t)i);
        kbest = min(kbest, (w[i] >> 16) * 512u + (uint32_t)(i + WPC));
    }
    kbest += (uint32_t)(q * NW);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) kbest = min(kbest, __shfl_xor_sync(0xffffffffu, kbest, o));
    const int minS = (int)(kbest >> 9), best = (int)(kbest & 511u);
    // runner-up over |d - best| > 1: bit i of near_lo / near_hi marks the low / high half of this lane's word i as one of
    // best-1, best, best+1
    const uint32_t sl = (uint32_t)(best - q * NW + 1);        // bit of best+1 among the low halves, plus 2; huge (wrapped) when far below
    const uint32_t shh = (uint32_t)(best - WPC - q * NW + 1);  // the same among the high halves
    const uint32_t near_lo = sl < (uint32_t)(NW + 2) ? (uint32_t)((7u << sl) >> 2) : 0u;
    const uint32_t near_hi = shh < (uint32_t)(NW + 2) ? (uint32_t)((7u << shh) >> 2) : 0u;
    uint32_t m2 = 0xFFFFu;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        if (!(near_lo & (1u << i))) m2 = min(m2, w[i] & 0xFFFFu);
        if (!(near_hi & (1u << i))) m2 = min(m2, w[i] >> 16);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    if (q == 0 && g < cnt) {
        const uint16_t* sh16 = reinterpret_cast<const uint16_t*>(svec + g * WPC);
        auto at = [&](int dd) -> uint32_t { return sh16[2 * (dd % WPC) + dd / WPC]; };
        const uint32_t sm1 = at(max(best - 1, 0)), sp1 = at(min(best + 1, 2 * WPC - 1));  // only used when 0 < best < D-1
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
    int sph;               // checkpoint slots per half row
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
    make_padmask<32, NPR>(padmask, lane, d.D);
    const uint32_t P1P1 = bcast16(d.P1), P2P2 = bcast16(d.P2);
    PathLane<32> pl;
    pl.init(lane);
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
                    OVO_DEVCHECK(pk >= R.ck_own && pk - R.ck_own < (ptrdiff_t)R.sph * WPC);
                    if (k + i) stv<NPR>(pk, s.L);
                    pk += WPC;
                }
                uint32_t c[NPR];
#pragma unroll
                for (int r = 0; r < NPR; r++) c[r] = cb[i][r];
                OVO_DEVCHECK(R.xa + DIRX * (k + i + PF) >= 0 && R.xa + DIRX * (k + i + PF) < d.W1);
                ldv<NPR>(cb[i], pc + i * DS);
                path_step<32, NPR, PAD>(s, c, padmask, P1P1, P2P2, pl);
            }
        } else {
#pragma unroll
            for (int i = 0; i < PF; i++) {
                if (k + i < n1) {
                    if (i % K == 0) {
                        OVO_DEVCHECK(pk >= R.ck_own && pk - R.ck_own < (ptrdiff_t)R.sph * WPC);
                        if (k + i) stv<NPR>(pk, s.L);
                        pk += WPC;
                    }
                    uint32_t c[NPR];
#pragma unroll
                    for (int r = 0; r < NPR; r++) c[r] = cb[i][r];
                    if (k + i + PF < n1) ldv<NPR>(cb[i], pc + i * DS);
                    path_step<32, NPR, PAD>(s, c, padmask, P1P1, P2P2, pl);
                }
            }
        }
    }
}

// NV: volumes of vertical paths the row sums up: 1 = Sv (already L1 + L2 + L3, k_sgbm_vsum), 3 = Lv[0..2], 6 = MODE_HH
template <int NV>
__host__ __device__ constexpr int horiz_nreg() { return NV < 3 ? NV : 3; }  // volumes prefetched one segment ahead

template <int NPR, int NV, int DIRX, bool FULL>
__device__ __forceinline__ void horiz_load_lv(const HorizRow<NPR>& R, ptrdiff_t o0, int cnt,
                                              uint32_t (&lv)[horiz_nreg<NV>()][horiz_seg<NPR>()][NPR]) {
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>(), DS = DIRX * WPC;
#pragma unroll
    for (int i = K - 1; i >= 0; i--) {  // the cell consumed first is requested first
        if (FULL || i < cnt) {
#pragma unroll
            for (int v = 0; v < horiz_nreg<NV>(); v++) ldv<NPR>(lv[v][i], R.Lv + v * R.vol + o0 - i * DS);
        }
    }
}

// One K-cell segment of phase 2.  On entry cb / ckv / lv hold C, the other warp's checkpoint and the vertical sums of
// segment j (requested one segment earlier); on exit they hold those of segment j-1.  FULL: all K cells exist (only the
// segment next to the rendezvous can be short).
template <int NPR, bool PAD, int NV, bool BATCH, bool SAT, int DIRX, bool FULL>
__device__ __forceinline__ void horiz_segment(const SgbmDims& d, const HorizRow<NPR>& R, PathState<NPR>& s, int lane, int j, int cnt,
                                              uint32_t (&cb)[horiz_seg<NPR>()][NPR], uint32_t (&ckv)[NPR],
                                              uint32_t (&lv)[horiz_nreg<NV>()][horiz_seg<NPR>()][NPR], const uint32_t (&padmask)[NPR],
                                              uint32_t P1P1, uint32_t P2P2, const PathLane<32>& pl) {
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>(), DS = DIRX * WPC;
    constexpr bool batched = BATCH;  // selection of the K cells at once; the odd uniquenessRatio >= 100 goes cell by cell
    // the other warp's cell k (counted along ITS sweep) sits at x = xo - DIRX * k
    const int xo = (DIRX > 0 ? d.W1 - 1 : 0) - DIRX * (j * K);
    const ptrdiff_t o0 = (ptrdiff_t)xo * WPC;  // cell i of the segment is at o0 - i * DS
    OVO_DEVCHECK(xo >= 0 && xo < d.W1 && xo - DIRX * ((FULL ? K : cnt) - 1) >= 0 && xo - DIRX * ((FULL ? K : cnt) - 1) < d.W1);
    OVO_DEVCHECK(j == 0 || (xo + DIRX * K >= 0 && xo + DIRX * K < d.W1 && xo + DIRX >= 0 && xo + DIRX < d.W1 && j < R.sph));
    uint32_t sv[K][NPR];
    {
        // replay the other warp's path over the segment
        PathState<NPR> o;
        if (j == 0) {
            path_reset<NPR>(o);
        } else {
#pragma unroll
            for (int r = 0; r < NPR; r++) o.L[r] = ckv[r];
            o.mm = vec_min<32, NPR>(o.L, pl);
        }
#pragma unroll
        for (int i = 0; i < K; i++) {
            if (FULL || i < cnt) {
                path_step<32, NPR, PAD>(o, cb[i], padmask, P1P1, P2P2, pl);
#pragma unroll
                for (int r = 0; r < NPR; r++) sv[i][r] = o.L[r];
            }
        }
#pragma unroll
        for (int i = K - 1; i >= 0; i--) {
            if (FULL || i < cnt) {
#pragma unroll
                for (int v = 0; v < horiz_nreg<NV>(); v++)
#pragma unroll
                    for (int r = 0; r < NPR; r++) sv[i][r] = sum16<SAT>(sv[i][r], lv[v][i][r]);
            }
        }
    }
    if (NV > 3) {  // MODE_HH: the three bottom-up paths Lv[3..5]
#pragma unroll
        for (int v = 3; v < NV; v++) {
#pragma unroll
            for (int i = K - 1; i >= 0; i--) {
                if (FULL || i < cnt) {
                    uint32_t u[NPR];
                    ldv<NPR>(u, R.Lv + (size_t)v * R.vol + o0 - i * DS);
#pragma unroll
                    for (int r = 0; r < NPR; r++) sv[i][r] = sum16<SAT>(sv[i][r], u[r]);
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
        horiz_load_lv<NPR, NV, DIRX, true>(R, o0 + K * DS, K, lv);
    }
    // own path over the segment, in the own direction (= the other's, reversed)
#pragma unroll
    for (int i = K - 1; i >= 0; i--) {
        if (FULL || i < cnt) {
            path_step<32, NPR, PAD>(s, cb[i], padmask, P1P1, P2P2, pl);
            uint32_t S[NPR];
#pragma unroll
            for (int r = 0; r < NPR; r++) S[r] = sum16<SAT>(sv[i][r], s.L[r]);
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

template <int NPR, bool PAD, int NV, bool BATCH, bool SAT, int DIRX>
__device__ __forceinline__ void horiz_phase2(const SgbmDims& d, const HorizRow<NPR>& R, PathState<NPR>& s, int lane) {
    constexpr int WPC = 32 * NPR, K = horiz_seg<NPR>(), DS = DIRX * WPC;
    const int n2 = R.n2;
    if (n2 <= 0) return;
    uint32_t padmask[NPR];
    make_padmask<32, NPR>(padmask, lane, d.D);
    const uint32_t P1P1 = bcast16(d.P1), P2P2 = bcast16(d.P2);
    PathLane<32> pl;
    pl.init(lane);
    int j = (n2 + K - 1) / K - 1;
    int cnt = n2 - j * K;
    uint32_t cb[K][NPR], ckv[NPR], lv[horiz_nreg<NV>()][K][NPR];
#pragma unroll
    for (int r = 0; r < NPR; r++) ckv[r] = 0;
    {
        const ptrdiff_t o0 = (ptrdiff_t)((DIRX > 0 ? d.W1 - 1 : 0) - DIRX * (j * K)) * WPC;
#pragma unroll
        for (int i = 0; i < K; i++)
            if (i < cnt) ldv<NPR>(cb[i], R.C + o0 - i * DS);
        if (j > 0) ldv<NPR>(ckv, R.ck_oth + (size_t)j * WPC);
        horiz_load_lv<NPR, NV, DIRX, false>(R, o0, cnt, lv);
    }
    if (cnt < K) {
        horiz_segment<NPR, PAD, NV, BATCH, SAT, DIRX, false>(d, R, s, lane, j, cnt, cb, ckv, lv, padmask, P1P1, P2P2, pl);
        j--;
    }
    for (; j >= 0; j--) horiz_segment<NPR, PAD, NV, BATCH, SAT, DIRX, true>(d, R, s, lane, j, K, cb, ckv, lv, padmask, P1P1, P2P2, pl);
}

template <int NPR, bool PAD, int NV, bool BATCH, bool SAT>
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
    R.sph = sph;
    R.selA = selA; R.selB = selB; R.selBest = selBest;
    R.svec = svec_s[wid];

    PathState<NPR> s;
    path_reset<NPR>(s);
    if (wid == 0) horiz_phase1<NPR, PAD, 1>(d, R, s, lane);
    else horiz_phase1<NPR, PAD, -1>(d, R, s, lane);
    __syncthreads();
    if (wid == 0) horiz_phase2<NPR, PAD, NV, BATCH, SAT, 1>(d, R, s, lane);
    else horiz_phase2<NPR, PAD, NV, BATCH, SAT, -1>(d, R, s, lane);
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
            OVO_DEVCHECK(best >= 0 && best < D);
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

// the fused vertical kernel is the default for MODE_SGBM; OVO_SGBM_FUSED=0 selects the one-volume-per-direction kernel
// (k_sgbm_vert, also what MODE_HH uses) for A/B measurements
static bool use_fused_vertical() {
    static const bool on = [] {
        const char* e = getenv("OVO_SGBM_FUSED");
        return !(e && e[0] == '0');
    }();
    return on;
}

static bool fused_vertical(const SgbmDims& d) { return d.mode == 0 && use_fused_vertical(); }

template <int LPC, int NPR, bool PAD, int BR, int VG, bool SAT>
int launch_vsum(const SgbmDims& d, const SgbmWorkspace& ws, size_t ws_stride, int nb, cudaStream_t st) {
    constexpr int B = vs_tile<LPC, VG>();
    const size_t smem = ((size_t)kVsRing * B + (size_t)kVsStrips * (B + 2 * BR)) * LPC * NPR * 4 + kVsStrips * 8;
    auto k_sgbm_vsum_t = k_sgbm_vsum<LPC, NPR, PAD, BR, VG, SAT>;
    if (smem > 48 * 1024) OVO_CUDA(cudaFuncSetAttribute(k_sgbm_vsum_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(cdiv(d.W1, B), nb), block(32 * vs_warps<LPC, BR, VG>());
    for (int y0 = 0, band = 0; y0 < d.H; y0 += BR, band++) {
        OVO_LAUNCH(k_sgbm_vsum_t, grid, block, smem, st, d, ws, ws_stride, y0, band & 1);
        OVO_LAUNCH_CHECK();
    }
    return 0;
}

template <int NPR, bool PAD>
int launch_paths(const SgbmDims& d, const SgbmWorkspace& ws, size_t ws_stride, int nb, cudaStream_t st) {
    const bool fused = fused_vertical(d);
    // no sum of five path costs can reach MAX_COST: the sums need no saturation (SURVEY.md A.4 bounds C by bs^2 * (2 ftzero + 63))
    // (padded disparities then hold 5 * 0x7FFF mod 2^16 = 32763 after the sums, still above every real sum: keep a margin)
    const bool nosat = 5 * (d.bs * d.bs * (2 * d.ftzero + 63) + d.P2) <= 32000;
    if (fused) {
        int rc;
#ifndef OVO_VS_VG
#define OVO_VS_VG 16
#endif
#ifndef OVO_VS_BR
#define OVO_VS_BR 16
#endif
#ifndef OVO_VS_BR256
#define OVO_VS_BR256 8
#endif
        auto vs = [&](auto sat) -> int {
            constexpr bool SAT = decltype(sat)::value;
            if constexpr (NPR == 1) return launch_vsum<8, 4, PAD, OVO_VS_BR, OVO_VS_VG, SAT>(d, ws, ws_stride, nb, st);        // Dp = 64
            else if constexpr (NPR == 2) return launch_vsum<8, 8, PAD, OVO_VS_BR, OVO_VS_VG, SAT>(d, ws, ws_stride, nb, st);   // Dp = 128
            else return launch_vsum<16, 8, PAD, OVO_VS_BR256, 16, SAT>(d, ws, ws_stride, nb, st);                       // Dp = 256
        };
        rc = nosat ? vs(std::false_type()) : vs(std::true_type());
        if (rc) return rc;
    } else {
        dim3 gv((d.mode ? 6 : 3) * cdiv(d.W1, 8), nb);
        { auto k_sgbm_vert_t = k_sgbm_vert<NPR, PAD>; OVO_LAUNCH(k_sgbm_vert_t, gv, dim3(256), 0, st, d, ws, ws_stride); }
        OVO_LAUNCH_CHECK();
    }
    dim3 gh(d.H, nb);
    const size_t smem = (size_t)d.W * 6 + (size_t)d.W1 * 10 + 16;
    auto go = [&](auto k_sgbm_horiz_t) -> int {
        if (smem > 40 * 1024) OVO_CUDA(cudaFuncSetAttribute(k_sgbm_horiz_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OVO_LAUNCH(k_sgbm_horiz_t, gh, dim3(64), smem, st, d, ws, ws_stride);
        return 0;
    };
    int rc;
    if (d.uniq < 100)
        rc = fused ? (nosat ? go(k_sgbm_horiz<NPR, PAD, 1, true, false>) : go(k_sgbm_horiz<NPR, PAD, 1, true, true>))
                   : (d.mode ? go(k_sgbm_horiz<NPR, PAD, 6, true, true>) : go(k_sgbm_horiz<NPR, PAD, 3, true, true>));
    else
        rc = fused ? (nosat ? go(k_sgbm_horiz<NPR, PAD, 1, false, false>) : go(k_sgbm_horiz<NPR, PAD, 1, false, true>))
                   : (d.mode ? go(k_sgbm_horiz<NPR, PAD, 6, false, true>) : go(k_sgbm_horiz<NPR, PAD, 3, false, true>));
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

// bytes of the vertical-path region that follows C: ONE volume (Sv) plus the band-state buffer [2][3][W1] vectors for the fused
// kernel, three volumes (six in MODE_HH) for the one-volume-per-direction kernel
static size_t lv_bytes(const SgbmDims& d) {
    const size_t vol = align_up((size_t)d.H * d.W1 * d.Dp * 2, 256);
    if (fused_vertical(d)) return align_up((size_t)d.H * d.W1 * d.Dp * 2 + (size_t)6 * d.W1 * d.Dp * 2, 256);
    return (d.mode ? 6 : 3) * vol;
}

// The fused vertical kernel copies C by row strips that overhang the tile by BR cells on the left and up to a tile + BR cells on
// the right (k_sgbm_vsum: fetch), i.e. up to 8 KB before C and 32 KB behind it; the overhang is never used, but it must be the
// frame's own memory even for images of a few rows, hence a floor under the region in front of C and a tail behind the frame.
static size_t prep_bytes(const SgbmDims& d) {
    const size_t img = align_up((size_t)d.H * d.W * 4, 256);
    return 4 * img > kStripSlackFront ? 4 * img : kStripSlackFront;
}

size_t sgbm_workspace_bytes(const SgbmDims& d) {
    const size_t vol = align_up((size_t)d.H * d.W1 * d.Dp * 2, 256);
    const size_t img = align_up((size_t)d.H * d.W * 4, 256);
    return align_up(prep_bytes(d) + vol /*C*/ + lv_bytes(d) + ckpt_bytes(d) + 4 * img /*raw, med (i16) + label, csize (i32) -> 2*0.5+2 = 3 img*/ +
                        kStripSlackBack, 256);
}

void sgbm_carve(const SgbmDims& d, uint8_t* base, SgbmWorkspace* ws) {
    const size_t vol = align_up((size_t)d.H * d.W1 * d.Dp * 2, 256);
    const size_t img = align_up((size_t)d.H * d.W * 4, 256);
    uint8_t* p = base;
    ws->prep = reinterpret_cast<uint32_t*>(p); p += prep_bytes(d);
    ws->C = reinterpret_cast<int16_t*>(p); p += vol;
    ws->Lv = reinterpret_cast<int16_t*>(p); p += lv_bytes(d);
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
