// Brute-force Hamming 2-NN, Lowe ratio test, fused 3-D point lookup, rigid alignment and the small glue kernels —
// stages 3 and 5 of the openVO hot path (sm_100a).
//
// Replaces, in the reference:
//   matcher.knnMatch(desc1, desc2, k=2)                       ref: src/openVO/stereo_odometer.py:163   
//   [m[0] for m in matches if m[0].distance < thr*m[1].distance]   ref: :164
//   bilinear_interpolate_pixels over cv2.reprojectImageTo3D    ref: :170-175, :50-79; stereo_camera.py:52 (A.5.1-2)
//   cv2.estimateAffine3D(force_rotation=True), Rodrigues norm  ref: :204-221                          (A.5.3-4)
//   .astype(float32)/16, crop, feature_mask                    ref: stereo_camera.py:51,53-55; stereo_odometer.py:38-41
#include "common.cuh"
#include <cstring>

namespace ovo {

namespace {

// ---- bulk-copy (TMA 1-D) + mbarrier helpers -------------------------------------------------------------------------
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

// ---- Hamming 2-NN -----------------------------------------------------------------------------------------------------
// Each thread owns one query descriptor (8 words in registers); the CTA streams its slice of the train set through
// shared memory in double-buffered tiles filled by the TMA bulk-copy engine; every lane reads the same train word
// (broadcast).  A candidate is the key (distance << 20 | trainIdx): keys are unique, so "two smallest keys" is exactly
// OpenCV's stable order (ties -> lowest train index) and partial results of different slices merge exactly.
constexpr int kKnnQ = 128;     // queries per CTA
constexpr int kKnnTile = 256;  // train descriptors per stage (8 KB)
constexpr uint32_t kKeyNone = 0xFFFFFFFFu;

__device__ __forceinline__ void knn2_partial_body(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                                                  int nsplit, uint32_t* __restrict__ part) {
    __shared__ __align__(128) uint4 tile[2][kKnnTile * 2];
#ifndef OVO_EMU
    __shared__ __align__(8) uint64_t bar[2];
#endif
    if ((int)(blockIdx.x * kKnnQ) >= nq) return;  // whole CTA (batched launches are sized for the largest query set)
    const int qi = blockIdx.x * kKnnQ + threadIdx.x;
    const int split = blockIdx.y;
    const int per = (nt + nsplit - 1) / nsplit;
    const int t0 = split * per, t1 = min(nt, t0 + per);
    uint32_t a[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(q + (size_t)min(qi, nq - 1) * 32);
        const uint4 lo = qp[0], hi = qp[1];
        a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    }
    uint32_t k0 = kKeyNone, k1 = kKeyNone;
    const int ntiles = t1 > t0 ? (t1 - t0 + kKnnTile - 1) / kKnnTile : 0;
#ifndef OVO_EMU
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && ntiles > 0) {
        const uint32_t bytes = (uint32_t)min(kKnnTile, t1 - t0) * 32u;
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(&tile[0][0], t + (size_t)t0 * 32, bytes, &bar[0]);
    }
#endif
    for (int it = 0; it < ntiles; it++) {
        const int st = it & 1;
        const int j0 = t0 + it * kKnnTile;
        const int cnt = min(kKnnTile, t1 - j0);
#ifndef OVO_EMU
        if (threadIdx.x == 0 && it + 1 < ntiles) {
            const int jn = j0 + kKnnTile;
            const uint32_t bytes = (uint32_t)min(kKnnTile, t1 - jn) * 32u;
            mbar_expect_tx(&bar[st ^ 1], bytes);
            bulk_g2s(&tile[st ^ 1][0], t + (size_t)jn * 32, bytes, &bar[st ^ 1]);
        }
        mbar_wait(&bar[st], (it >> 1) & 1);
#else
        for (int i = threadIdx.x; i < cnt * 2; i += kKnnQ) tile[st][i] = reinterpret_cast<const uint4*>(t + (size_t)j0 * 32)[i];
        __syncthreads();
#endif
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            const uint4 lo = tile[st][2 * j], hi = tile[st][2 * j + 1];
            const uint32_t dist = __popc(a[0] ^ lo.x) + __popc(a[1] ^ lo.y) + __popc(a[2] ^ lo.z) + __popc(a[3] ^ lo.w) +
                                  __popc(a[4] ^ hi.x) + __popc(a[5] ^ hi.y) + __popc(a[6] ^ hi.z) + __popc(a[7] ^ hi.w);
            const uint32_t key = (dist << 20) | (uint32_t)(j0 + j);
            const uint32_t lo2 = min(key, k1);  // new second-best if key displaces k1 only
            k1 = key < k0 ? k0 : lo2;
            k0 = min(k0, key);
        }
        __syncthreads();
    }
    if (qi < nq) {
        part[((size_t)qi * nsplit + split) * 2] = k0;
        part[((size_t)qi * nsplit + split) * 2 + 1] = k1;
    }
}

struct PairItem {
    const uint8_t *q, *t;
    int nq, nt;
    const float *kp1, *kp2, *disp1, *disp2;
    int32_t *nn, *matches;
    float *pts1, *pts2;
    double* out;        // [18]: 16 rigid-transform outputs, then two int32 counts in slot 16
    uint32_t* scratch;  // 2-NN partial results
    int32_t* nn_rev;    // optional: 2-NN of the train set against the query set (cross-check), or null
};
constexpr int kMaxPairs = 16;
struct PairBatch { PairItem it[kMaxPairs]; };

__global__ void __launch_bounds__(kKnnQ) k_knn2_partial(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                                                        int nsplit, uint32_t* __restrict__ part) {
    knn2_partial_body(q, nq, t, nt, nsplit, part);
}
__global__ void __launch_bounds__(kKnnQ) k_knn2_partial_b(PairBatch b, int nsplit, int reverse) {
    const PairItem& p = b.it[blockIdx.z];
    if (reverse) { if (p.nn_rev) knn2_partial_body(p.t, p.nt, p.q, p.nq, nsplit, p.scratch); }
    else knn2_partial_body(p.q, p.nq, p.t, p.nt, nsplit, p.scratch);
}

__device__ __forceinline__ void knn2_merge_body(const uint32_t* __restrict__ part, int nq, int nsplit, int32_t* __restrict__ nn) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    uint32_t k0 = kKeyNone, k1 = kKeyNone;
    for (int s = 0; s < nsplit * 2; s++) {
        const uint32_t key = part[(size_t)qi * nsplit * 2 + s];
        const uint32_t lo2 = min(key, k1);
        k1 = key < k0 ? k0 : lo2;
        k0 = min(k0, key);
    }
    int32_t* o = nn + 4 * (size_t)qi;
    o[0] = k0 == kKeyNone ? -1 : (int)(k0 & 0xFFFFFu);
    o[1] = k0 == kKeyNone ? 0 : (int)(k0 >> 20);
    o[2] = k1 == kKeyNone ? -1 : (int)(k1 & 0xFFFFFu);
    o[3] = k1 == kKeyNone ? 0 : (int)(k1 >> 20);
}
__global__ void k_knn2_merge(const uint32_t* __restrict__ part, int nq, int nsplit, int32_t* __restrict__ nn) {
    knn2_merge_body(part, nq, nsplit, nn);
}
__global__ void k_knn2_merge_b(PairBatch b, int nsplit, int reverse) {
    const PairItem& p = b.it[blockIdx.z];
    if (reverse) { if (p.nn_rev) knn2_merge_body(p.scratch, p.nt, nsplit, p.nn_rev); }
    else knn2_merge_body(p.scratch, p.nq, nsplit, p.nn);
}

// ---- reprojection (A.5.1) ------------------------------------------------------------------------------------------------
// h = Q * [x, y, d, 1]^T accumulated left to right in float64 with every product and sum rounded (no contraction);
// out_i = f32( f64(f32(h_i)) / h_3 )
__device__ __forceinline__ void reproject_px(const double* Q, int x, int y, float d, float (&o)[3]) {
    const double X = (double)x, Y = (double)y, Dd = (double)d;
    double h[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        h[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Q[4 * i], X), __dmul_rn(Q[4 * i + 1], Y)), __dmul_rn(Q[4 * i + 2], Dd)), Q[4 * i + 3]);
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = __double2float_rn(__ddiv_rn((double)__double2float_rn(h[i]), h[3]));
}

struct QMat { double q[16]; };

__global__ void k_reproject(const float* __restrict__ disp, int pitch, int cw, int ch, int x0, int y0, QMat Q, float* __restrict__ xyz) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cw) return;
    float o[3];
    reproject_px(Q.q, x + x0, y + y0, disp[(size_t)y * pitch + x], o);
    float* p = xyz + ((size_t)y * cw + x) * 3;
    p[0] = o[0]; p[1] = o[1]; p[2] = o[2];
}

// ---- ratio test + ordered compaction + bilinear lookup (A.3, A.5.2) -------------------------------------------------------------
__device__ __forceinline__ bool isinf3(const float (&p)[3]) { return isinf(p[0]) || isinf(p[1]) || isinf(p[2]); }

// returns false when no tap was usable (the reference divides 0/0 as Python ints there -> ZeroDivisionError)
__device__ bool lookup_point(const GatherParams& P, const float* __restrict__ disp, float xf, float yf, float (&out)[3]) {
    const int fx = (int)xf, fy = (int)yf;
    const double rx = __dsub_rn((double)xf, (double)fx), ry = __dsub_rn((double)yf, (double)fy);
    const double ox = __dsub_rn(1.0, rx), oy = __dsub_rn(1.0, ry);
    // reference order: p00, p01 (y+1), p10 (x+1), p11
    const int tx[4] = {fx, fx, fx + 1, fx + 1}, ty[4] = {fy, fy + 1, fy, fy + 1};
    const double tw[4] = {__dmul_rn(ox, oy), __dmul_rn(ox, ry), __dmul_rn(rx, oy), __dmul_rn(rx, ry)};
    float num[3] = {0.f, 0.f, 0.f};
    double den = 0.0;
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (tx[k] >= P.cw || ty[k] >= P.ch) continue;
        float p[3];
        reproject_px(P.Q, tx[k] + P.roi_x0, ty[k] + P.roi_y0, disp[(size_t)ty[k] * P.disp_pitch + tx[k]], p);
        if (isinf3(p)) continue;
        const float w = __double2float_rn(tw[k]);
#pragma unroll
        for (int i = 0; i < 3; i++) num[i] = __fadd_rn(num[i], __fmul_rn(w, p[i]));
        den = __dadd_rn(den, tw[k]);
        any = true;
    }
    const float fden = __double2float_rn(den);
#pragma unroll
    for (int i = 0; i < 3; i++) out[i] = __fdiv_rn(num[i], fden);
    return any;
}

__device__ __forceinline__ void match_gather_body(const GatherParams& P, const int32_t* __restrict__ nn, int nq,
                                                  const float* __restrict__ kp1, const float* __restrict__ kp2,
                                                  const float* __restrict__ disp1, const float* __restrict__ disp2,
                                                  int32_t* __restrict__ matches, float* __restrict__ pts1, float* __restrict__ pts2,
                                                  int32_t* __restrict__ counts, const int32_t* __restrict__ nn_rev) {
    __shared__ int warp_sums[32];
    __shared__ int base_s, bad_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { base_s = 0; bad_s = 0; }
    __syncthreads();
    for (int q0 = 0; q0 < nq; q0 += 1024) {
        const int qi = q0 + threadIdx.x;
        bool keep = false;
        int ti = 0, d0 = 0;
        if (qi < nq) {
            const int4 r = reinterpret_cast<const int4*>(nn)[qi];
            ti = r.x; d0 = r.y;
            // m[0].distance < thr * m[1].distance evaluated in double (A.3)
            keep = r.x >= 0 && r.z >= 0 && (double)r.y < __dmul_rn(P.thr, (double)r.w);
            // opt-in cross-check (not in the reference): the train descriptor's own nearest neighbour must be this query
            if (keep && nn_rev != nullptr) keep = nn_rev[4 * (size_t)r.x] == qi;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(bal & ((1u << lane) - 1));
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
        const int pos = base_s + (wid ? warp_sums[wid - 1] : 0) + wpre;
        const int total = warp_sums[31];
        if (keep) {
            matches[3 * pos] = qi; matches[3 * pos + 1] = ti; matches[3 * pos + 2] = d0;
            float a[3], b[3];
            const bool oka = lookup_point(P, disp1, kp1[6 * (size_t)qi], kp1[6 * (size_t)qi + 1], a);
            const bool okb = lookup_point(P, disp2, kp2[6 * (size_t)ti], kp2[6 * (size_t)ti + 1], b);
#pragma unroll
            for (int i = 0; i < 3; i++) { pts1[3 * pos + i] = a[i]; pts2[3 * pos + i] = b[i]; }
            if (!oka || !okb) atomicAdd(&bad_s, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) base_s += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { counts[0] = base_s; counts[1] = bad_s; }
}
__global__ void __launch_bounds__(1024) k_match_gather(GatherParams P, const int32_t* __restrict__ nn, int nq,
                                                       const float* __restrict__ kp1, const float* __restrict__ kp2,
                                                       const float* __restrict__ disp1, const float* __restrict__ disp2,
                                                       int32_t* __restrict__ matches, float* __restrict__ pts1, float* __restrict__ pts2,
                                                       int32_t* __restrict__ counts, const int32_t* __restrict__ nn_rev) {
    match_gather_body(P, nn, nq, kp1, kp2, disp1, disp2, matches, pts1, pts2, counts, nn_rev);
}
__global__ void __launch_bounds__(1024) k_match_gather_b(GatherParams P, PairBatch b) {
    const PairItem& p = b.it[blockIdx.x];
    match_gather_body(P, p.nn, p.nq, p.kp1, p.kp2, p.disp1, p.disp2, p.matches, p.pts1, p.pts2, reinterpret_cast<int32_t*>(p.out + 16), p.nn_rev);
}

// ---- Umeyama (A.5.3) ---------------------------------------------------------------------------------------------------------
__device__ double block_sum(double v, double* sh) {
    // 256 threads; deterministic tree
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

__device__ void svd3(const double A[9], double U[9], double S[3], double V[9]) {
    // one-sided Jacobi (Hestenes): A V = U diag(S); columns sorted by descending S
    double u[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++) u[i] = A[i];
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double al = 0, be = 0, ga = 0;
                for (int i = 0; i < 3; i++) { al += u[3 * i + p] * u[3 * i + p]; be += u[3 * i + q] * u[3 * i + q]; ga += u[3 * i + p] * u[3 * i + q]; }
                if (ga == 0.0 || fabs(ga) <= 1e-300) continue;
                const double lim = 1e-17 * sqrt(al * be);
                if (fabs(ga) <= lim) continue;
                off = fmax(off, fabs(ga) / sqrt(al * be));
                const double zeta = (be - al) / (2.0 * ga);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
                for (int i = 0; i < 3; i++) {
                    const double up = u[3 * i + p], uq = u[3 * i + q];
                    u[3 * i + p] = c * up - s * uq; u[3 * i + q] = s * up + c * uq;
                    const double vp = v[3 * i + p], vq = v[3 * i + q];
                    v[3 * i + p] = c * vp - s * vq; v[3 * i + q] = s * vp + c * vq;
                }
            }
        if (off < 1e-16) break;
    }
    double sv[3];
    for (int j = 0; j < 3; j++) sv[j] = sqrt(u[j] * u[j] + u[3 + j] * u[3 + j] + u[6 + j] * u[6 + j]);
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2 - i; j++)
            if (sv[ord[j]] < sv[ord[j + 1]]) { const int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
    for (int j = 0; j < 3; j++) {
        const int c = ord[j];
        S[j] = sv[c];
        for (int i = 0; i < 3; i++) { U[3 * i + j] = sv[c] > 0 ? u[3 * i + c] / sv[c] : 0.0; V[3 * i + j] = v[3 * i + c]; }
    }
    // complete U if the smallest singular value vanished: third column = cross(first, second)
    if (!(S[2] > 1e-300 * S[0]) || S[2] == 0.0) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

__device__ __forceinline__ double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

__device__ __forceinline__ void umeyama_body(const float* __restrict__ src, const float* __restrict__ dst, const int32_t* __restrict__ count,
                                             int cap, double* __restrict__ out) {
    __shared__ double sh[256];
    __shared__ double mu[6];
    const int n = min(*count, cap);
    const int tid = threadIdx.x;
    if (n < 1) {
        if (tid < 16) out[tid] = tid == 15 ? 0.0 : nan("");
        return;
    }
    const double inv_n = 1.0 / (double)n;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256)
        for (int k = 0; k < 3; k++) { acc[k] += (double)src[3 * i + k]; acc[3 + k] += (double)dst[3 * i + k]; }
    for (int k = 0; k < 6; k++) {
        const double s = block_sum(acc[k], sh);
        if (tid == 0) mu[k] = s * inv_n;
    }
    __syncthreads();
    double cv[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256) {
        double a[3], b[3];
        for (int k = 0; k < 3; k++) { a[k] = (double)src[3 * i + k] - mu[k]; b[k] = (double)dst[3 * i + k] - mu[3 + k]; }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) cv[3 * r + c] += b[r] * a[c];
        cv[9] += a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    }
    double tot[10];
    for (int k = 0; k < 10; k++) tot[k] = block_sum(cv[k], sh);
    if (tid != 0) return;
    double cov[9], U[9], S[3], V[9];
    for (int k = 0; k < 9; k++) cov[k] = tot[k] * inv_n;
    svd3(cov, U, S, V);
    double sg[3] = {1, 1, 1};
    if (det3(U) * det3(V) < 0) sg[2] = -1;  // det(V^T) == det(V)
    double R[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[3 * r + c] = U[3 * r] * sg[0] * V[3 * c] + U[3 * r + 1] * sg[1] * V[3 * c + 1] + U[3 * r + 2] * sg[2] * V[3 * c + 2];
    const double scale = (S[0] * sg[0] + S[1] * sg[1] + S[2] * sg[2]) * (double)n / tot[9];
    double t[3];
    for (int r = 0; r < 3; r++) t[r] = mu[3 + r] - scale * (R[3 * r] * mu[0] + R[3 * r + 1] * mu[1] + R[3 * r + 2] * mu[2]);
    for (int r = 0; r < 3; r++) {
        out[4 * r] = R[3 * r]; out[4 * r + 1] = R[3 * r + 1]; out[4 * r + 2] = R[3 * r + 2]; out[4 * r + 3] = t[r];
    }
    out[12] = scale;
    // |Rodrigues(R)|: robust angle from the skew part and the trace
    const double sx = R[7] - R[5], sy = R[2] - R[6], sz = R[3] - R[1];
    out[13] = atan2(0.5 * sqrt(sx * sx + sy * sy + sz * sz), 0.5 * (R[0] + R[4] + R[8] - 1.0));
    out[14] = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    out[15] = (double)n;
}
__global__ void __launch_bounds__(256) k_umeyama(const float* __restrict__ src, const float* __restrict__ dst, const int32_t* __restrict__ count,
                                                 int cap, double* __restrict__ out) {
    umeyama_body(src, dst, count, cap, out);
}
__global__ void __launch_bounds__(256) k_umeyama_b(PairBatch b, int cap) {
    const PairItem& p = b.it[blockIdx.x];
    umeyama_body(p.pts1, p.pts2, reinterpret_cast<const int32_t*>(p.out + 16), cap, p.out);
}

// ---- small glue -----------------------------------------------------------------------------------------------------------------
// 4 pixels per thread (CTA = 256 threads = 1024 consecutive output pixels): these kernels are bound by CTA dispatch otherwise
__global__ void __launch_bounds__(256) k_disp_post(const int16_t* __restrict__ disp, int W, int x0, int y0, int cw, int ch, float lo,
                                                   float hi, float* __restrict__ out, uint8_t* __restrict__ mask, size_t in_stride,
                                                   size_t out_stride) {
    const int f = blockIdx.y, n = cw * ch;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int i = (blockIdx.x * 4 + e) * 256 + threadIdx.x;
        if (i >= n) break;
        const int y = i / cw, x = i - y * cw;
        const float d = __fdiv_rn((float)disp[in_stride * f + (size_t)(y + y0) * W + x + x0], 16.f);
        out[out_stride * f + i] = d;
        if (mask) mask[out_stride * f + i] = (d >= lo && d <= hi) ? 255 : 0;
    }
}

__global__ void __launch_bounds__(256) k_crop_u8(const uint8_t* __restrict__ img, int pitch, size_t frame_stride, int x0, int y0, int cw,
                                                 int ch, uint8_t* __restrict__ out) {
    const int f = blockIdx.y, n = cw * ch;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int i = (blockIdx.x * 4 + e) * 256 + threadIdx.x;
        if (i >= n) break;
        const int y = i / cw, x = i - y * cw;
        out[(size_t)f * n + i] = img[frame_stride * f + (size_t)(y + y0) * pitch + x + x0];
    }
}

// ---- rectification (SURVEY.md §8(f) n1 + n2) -----------------------------------------------------------------------------------
// cv2.remap(img, map1 (int16 x,y), map2 (uint16 = fy*32 + fx), INTER_LINEAR), BORDER_CONSTANT 0, with cv2.cvtColor(BGR2GRAY)
// fused in for 3-channel input (the reference converts first, ref: src/openVO/stereo_camera.py:44-50; per-tap conversion is the
// same arithmetic).  Weights are exact integers: (32-fx)(32-fy)*32 etc., sum 32768; dst = (sum + 2^14) >> 15.
__device__ __forceinline__ int gray_at(const uint8_t* __restrict__ img, int pitch, int ch, int x, int y) {
    const uint8_t* p = img + (size_t)y * pitch + (size_t)x * ch;
    if (ch == 1) return p[0];
    return (3735 * (int)p[0] + 19235 * (int)p[1] + 9798 * (int)p[2] + (1 << 14)) >> 15;
}

__global__ void k_rectify(const uint8_t* __restrict__ img, int pitch, size_t frame_stride, int ch, int W, int H,
                          const int16_t* __restrict__ map1, const uint16_t* __restrict__ map2, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= W) return;
    const uint8_t* I = img + frame_stride * f;
    int v;
    if (map1 == nullptr) {
        v = gray_at(I, pitch, ch, x, y);  // colour conversion only
    } else {
        const int sx = map1[2 * ((size_t)y * W + x)], sy = map1[2 * ((size_t)y * W + x) + 1];
        const int m = map2[(size_t)y * W + x];
        const int fx = m & 31, fy = (m >> 5) & 31;
        auto tap = [&](int xx, int yy) -> int { return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? gray_at(I, pitch, ch, xx, yy) : 0; };
        const int acc = tap(sx, sy) * ((32 - fx) * (32 - fy) * 32) + tap(sx + 1, sy) * (fx * (32 - fy) * 32) +
                        tap(sx, sy + 1) * ((32 - fx) * fy * 32) + tap(sx + 1, sy + 1) * (fx * fy * 32);
        v = (acc + (1 << 14)) >> 15;
    }
    out[((size_t)f * H + y) * W + x] = (uint8_t)v;
}

}  // namespace

int rectify_launch(const uint8_t* img, int pitch, size_t frame_stride, int ch, int W, int H, int nb, const int16_t* map1,
                   const uint16_t* map2, uint8_t* out, cudaStream_t st) {
    dim3 grid(cdiv(W, 128), H, nb);
    OVO_LAUNCH(k_rectify, grid, dim3(128), 0, st, img, pitch, frame_stride, ch, W, H, map1, map2, out);
    OVO_LAUNCH_CHECK();
    return 0;
}

size_t knn2_scratch_bytes(int nq_cap, int nt_cap) { return (size_t)nq_cap * 32 * 2 * 4; }

int knn2_launch(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* nn_out, uint32_t* scratch, cudaStream_t st) {
    if (nq <= 0) return 0;
    if (nt >= (1 << 20)) { set_error("knn2: train set too large (%d)", nt); return 1; }
    int nsplit = cdiv(nt, 128);
    nsplit = nsplit < 1 ? 1 : (nsplit > 32 ? 32 : nsplit);
    dim3 grid(cdiv(nq, kKnnQ), nsplit);
    OVO_LAUNCH(k_knn2_partial, grid, dim3(kKnnQ), 0, st, q, nq, t, nt, nsplit, scratch);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_knn2_merge, dim3(cdiv(nq, 128)), dim3(128), 0, st, scratch, nq, nsplit, nn_out);
    OVO_LAUNCH_CHECK();
    return 0;
}

// n <= kMaxPairs frame pairs in four launches: 2-NN partial / merge, ratio + gather, rigid alignment
int pair_batch_launch(const GatherParams& gp, int n, const void* items_host, int cap, cudaStream_t st) {
    PairBatch b;
    memset(&b, 0, sizeof(b));
    memcpy(b.it, items_host, sizeof(PairItem) * n);
    int maxq = 0, maxt = 0;
    for (int i = 0; i < n; i++) {
        maxq = b.it[i].nq > maxq ? b.it[i].nq : maxq;
        maxt = b.it[i].nt > maxt ? b.it[i].nt : maxt;
        if (b.it[i].nt >= (1 << 20)) { set_error("knn2: train set too large"); return 1; }
    }
    bool any_rev = false;
    for (int i = 0; i < n; i++) any_rev = any_rev || b.it[i].nn_rev != nullptr;
    if (any_rev && maxt > 0) {  // cross-check: train -> query 2-NN first (shares the scratch with the forward pass)
        int nsplit = cdiv(maxq, 128);
        nsplit = nsplit < 1 ? 1 : (nsplit > 32 ? 32 : nsplit);
        OVO_LAUNCH(k_knn2_partial_b, dim3(cdiv(maxt, kKnnQ), nsplit, n), dim3(kKnnQ), 0, st, b, nsplit, 1);
        OVO_LAUNCH_CHECK();
        OVO_LAUNCH(k_knn2_merge_b, dim3(cdiv(maxt, 128), 1, n), dim3(128), 0, st, b, nsplit, 1);
        OVO_LAUNCH_CHECK();
    }
    if (maxq > 0) {
        int nsplit = cdiv(maxt, 128);
        nsplit = nsplit < 1 ? 1 : (nsplit > 32 ? 32 : nsplit);
        OVO_LAUNCH(k_knn2_partial_b, dim3(cdiv(maxq, kKnnQ), nsplit, n), dim3(kKnnQ), 0, st, b, nsplit, 0);
        OVO_LAUNCH_CHECK();
        OVO_LAUNCH(k_knn2_merge_b, dim3(cdiv(maxq, 128), 1, n), dim3(128), 0, st, b, nsplit, 0);
        OVO_LAUNCH_CHECK();
    }
    OVO_LAUNCH(k_match_gather_b, dim3(n), dim3(1024), 0, st, gp, b);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_umeyama_b, dim3(n), dim3(256), 0, st, b, cap);
    OVO_LAUNCH_CHECK();
    return 0;
}
int pair_item_size() { return (int)sizeof(PairItem); }
int pair_max_batch() { return kMaxPairs; }

int match_gather_launch(const GatherParams& p, const int32_t* nn, int nq, const float* kp1, const float* kp2, const float* disp1,
                        const float* disp2, int32_t* matches_out, float* pts1, float* pts2, int32_t* counts_out, const int32_t* nn_rev,
                        cudaStream_t st) {
    OVO_LAUNCH(k_match_gather, dim3(1), dim3(1024), 0, st, p, nn, nq, kp1, kp2, disp1, disp2, matches_out, pts1, pts2, counts_out, nn_rev);
    OVO_LAUNCH_CHECK();
    return 0;
}

int umeyama_launch(const float* pts1, const float* pts2, const int32_t* m_dev, int m_cap, double* out, cudaStream_t st) {
    OVO_LAUNCH(k_umeyama, dim3(1), dim3(256), 0, st, pts1, pts2, m_dev, m_cap, out);
    OVO_LAUNCH_CHECK();
    return 0;
}

int reproject_launch(const float* disp, int pitch, int cw, int ch, int x0, int y0, const double* Q16, float* xyz, cudaStream_t st) {
    QMat Q;
    for (int i = 0; i < 16; i++) Q.q[i] = Q16[i];
    dim3 grid(cdiv(cw, 128), ch);
    OVO_LAUNCH(k_reproject, grid, dim3(128), 0, st, disp, pitch, cw, ch, x0, y0, Q, xyz);
    OVO_LAUNCH_CHECK();
    return 0;
}

int disp_post_launch(const int16_t* disp, int W, int H, int x0, int y0, int cw, int ch, float lo, float hi, float* disp_f32,
                     uint8_t* mask, int nb, cudaStream_t st) {
    dim3 grid(cdiv(cw * ch, 1024), nb);
    OVO_LAUNCH(k_disp_post, grid, dim3(256), 0, st, disp, W, x0, y0, cw, ch, lo, hi, disp_f32, mask, (size_t)W * H, (size_t)cw * ch);
    OVO_LAUNCH_CHECK();
    return 0;
}

int crop_launch(const uint8_t* img, int pitch, size_t frame_stride, int x0, int y0, int cw, int ch, int nb, uint8_t* out, cudaStream_t st) {
    dim3 grid(cdiv(cw * ch, 1024), nb);
    OVO_LAUNCH(k_crop_u8, grid, dim3(256), 0, st, img, pitch, frame_stride, x0, y0, cw, ch, out);
    OVO_LAUNCH_CHECK();
    return 0;
}

}  // namespace ovo
