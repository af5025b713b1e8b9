// Optional pose pre-filters of StereoOdometer.point_cloud_transform — SURVEY.md §8(f) row n3, off by default in the reference.
//   rigid_body_filter   ref: src/openVO/stereo_odometer.py:82-105  (greedy clique on the pairwise-distance consistency graph)
//   outlier filter      ref: src/openVO/stereo_odometer.py:189-197 (relative residual under a first Umeyama fit, median + threshold)
// Both end in an ordered compaction of the two point sets.  Arithmetic follows numpy's: float32 differences / norms for the
// consistency matrix (the points are float32 and numpy keeps that type), float64 for the residuals.
#include "common.cuh"

namespace ovo {

namespace {

// consistency bit matrix: row i, bit j = | ||p_i - p_j|| - ||q_i - q_j|| | < thr ; degree[i] = popcount of row i
__global__ void __launch_bounds__(256) k_consistency(const float* __restrict__ prev, const float* __restrict__ cur, const int32_t* __restrict__ count,
                                                     int cap, float thr, uint32_t* __restrict__ bits, int32_t* __restrict__ degree) {
    __shared__ int red[8];
    const int m = min(*count, cap), i = blockIdx.x;
    if (i >= m) return;
    const int words = (cap + 31) / 32;
    const float ax = cur[3 * i], ay = cur[3 * i + 1], az = cur[3 * i + 2];
    const float bx = prev[3 * i], by = prev[3 * i + 1], bz = prev[3 * i + 2];
    int deg = 0;
    for (int w0 = 0; w0 < words; w0 += 8) {          // 8 warps, one 32-bit word of the row each
        const int w = w0 + (threadIdx.x >> 5), j = w * 32 + (threadIdx.x & 31);
        bool ok = false;
        if (w < words && j < m) {
            const float dx = __fsub_rn(ax, cur[3 * j]), dy = __fsub_rn(ay, cur[3 * j + 1]), dz = __fsub_rn(az, cur[3 * j + 2]);
            const float ex = __fsub_rn(bx, prev[3 * j]), ey = __fsub_rn(by, prev[3 * j + 1]), ez = __fsub_rn(bz, prev[3 * j + 2]);
            const float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            const float dp = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
            ok = fabsf(__fsub_rn(dn, dp)) < thr;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (w < words && (threadIdx.x & 31) == 0) bits[(size_t)i * words + w] = bal;
        deg += (threadIdx.x & 31) == 0 ? __popc(bal) : 0;
    }
    deg = __reduce_add_sync(0xffffffffu, deg);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = deg;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < 8; k++) s += red[k];
        degree[i] = s;
    }
}

// block-wide argmax of key = value << 14 | (16383 - index): the largest value, ties to the lowest index (np.argmax)
__device__ int block_argmax(uint32_t key, uint32_t* sh) {
    key = __reduce_max_sync(0xffffffffu, key);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = key;
    __syncthreads();
    uint32_t k = threadIdx.x < 32 ? sh[threadIdx.x] : 0;
    if (threadIdx.x < 32) k = __reduce_max_sync(0xffffffffu, k);
    if (threadIdx.x == 0) sh[32] = k;
    __syncthreads();
    const uint32_t r = sh[32];
    __syncthreads();
    return (int)r;
}

// greedy clique (one CTA): start from the highest degree, repeatedly add the highest-degree vertex consistent with every member
__global__ void __launch_bounds__(1024) k_clique(const uint32_t* __restrict__ bits, const int32_t* __restrict__ degree,
                                                 const int32_t* __restrict__ count, int cap, int32_t* __restrict__ clique) {
    __shared__ uint32_t sh[33];
    const int m = min(*count, cap);
    const int words = (cap + 31) / 32;
    if (m <= 0) return;
    // m <= 16383 (key packing: degree and index take 14 bits each); each thread owns vertices tid, tid + 1024, ...
    constexpr int PER = 16;
    bool comp[PER], inq[PER];
    int deg[PER];
    for (int k = 0; k < PER; k++) {
        const int v = threadIdx.x + 1024 * k;
        comp[k] = false; inq[k] = false;
        deg[k] = v < m ? degree[v] : 0;
    }
    auto best_of = [&](bool first) {
        uint32_t key = 0;
        for (int k = 0; k < PER; k++) {
            const int v = threadIdx.x + 1024 * k;
            if (v < m && (first || (comp[k] && !inq[k]))) key = max(key, ((uint32_t)deg[k] << 14) | (uint32_t)(16383 - v));
        }
        return block_argmax(key, sh);
    };
    int key = best_of(true);
    for (int it = 0; it <= m; it++) {
        if (key == 0) break;                      // no candidate left (degrees are >= 1: every point is consistent with itself)
        const int sel = 16383 - (key & 16383);
        for (int k = 0; k < PER; k++) {
            const int v = threadIdx.x + 1024 * k;
            if (v < m) {
                const bool c = (bits[(size_t)sel * words + (v >> 5)] >> (v & 31)) & 1u;
                comp[k] = it == 0 ? c : (comp[k] && c);
                if (v == sel) inq[k] = true;
            }
        }
        key = best_of(false);
    }
    for (int k = 0; k < PER; k++) {
        const int v = threadIdx.x + 1024 * k;
        if (v < m) clique[v] = inq[k] ? 1 : 0;
    }
}

// relative residuals under T (3x4, f64): e_i = ||[q_i,1] - T [p_i,1]|| / ||[q_i,1]||
__global__ void k_residuals(const float* __restrict__ prev, const float* __restrict__ cur, const int32_t* __restrict__ count, int cap,
                            const double* __restrict__ T, double* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*count, cap)) return;
    const double p[3] = {(double)prev[3 * i], (double)prev[3 * i + 1], (double)prev[3 * i + 2]};
    const double q[3] = {(double)cur[3 * i], (double)cur[3 * i + 1], (double)cur[3 * i + 2]};
    double s = 0;
    for (int r = 0; r < 3; r++) {
        const double y = T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2] + T[4 * r + 3];
        const double dlt = q[r] - y;
        s += dlt * dlt;
    }
    // the homogeneous coordinate contributes (1 - 1)^2 = 0 to the numerator and 1 to the denominator
    err[i] = sqrt(s) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + 1.0);
}

// keep[i] = err[i] < thr + median(err)   (np.median: mean of the two middle order statistics for even n); one CTA, rank by counting
__global__ void __launch_bounds__(1024) k_median_keep(const double* __restrict__ err, const int32_t* __restrict__ count, int cap, double thr,
                                                      int32_t* __restrict__ keep) {
    __shared__ double mid[2];
    const int m = min(*count, cap);
    if (m <= 0) return;
    const int lo = (m - 1) / 2, hi = m / 2;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double e = err[i];
        int rank = 0;
        for (int j = 0; j < m; j++) {
            const double o = err[j];
            rank += (o < e || (o == e && j < i)) ? 1 : 0;
        }
        if (rank == lo) mid[0] = e;
        if (rank == hi) mid[1] = e;
    }
    __syncthreads();
    const double med = lo == hi ? mid[0] : (mid[0] + mid[1]) / 2.0;
    const double limit = thr + med;
    for (int i = threadIdx.x; i < m; i += blockDim.x) keep[i] = err[i] < limit ? 1 : 0;
}

// ordered compaction of both point sets by a 0/1 mask (one CTA); count is updated in place
__global__ void __launch_bounds__(1024) k_compact_points(float* __restrict__ a, float* __restrict__ b, const int32_t* __restrict__ mask,
                                                         int32_t* __restrict__ count, int cap, float* __restrict__ a_out,
                                                         float* __restrict__ b_out) {
    __shared__ int warp_sums[32];
    __shared__ int base_s;
    const int m = min(*count, cap);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < m; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool k = i < m && mask[i] != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, k);
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
        const int pos = base_s + (wid ? warp_sums[wid - 1] : 0) + __popc(bal & ((1u << lane) - 1));
        if (k)
            for (int c = 0; c < 3; c++) { a_out[3 * pos + c] = a[3 * i + c]; b_out[3 * pos + c] = b[3 * i + c]; }
        __syncthreads();
        if (threadIdx.x == 0) base_s += warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base_s;
}

}  // namespace

size_t filter_scratch_bytes(int cap) {
    const size_t words = (cap + 31) / 32;
    return align_up((size_t)cap * words * 4, 256) + 3 * align_up((size_t)cap * 4, 256) + align_up((size_t)cap * 8, 256) +
           2 * align_up((size_t)cap * 12, 256);
}

// prev/cur: f32 [cap][3] (compacted in place through scratch), count: device i32 (updated)
int rigid_filter_launch(float* prev, float* cur, int32_t* count, int cap, float thr, uint8_t* scratch, cudaStream_t st) {
    if (cap > 16383) { set_error("rigid_body_filter supports at most 16383 points"); return 1; }
    const size_t words = (cap + 31) / 32;
    uint32_t* bits = (uint32_t*)scratch; scratch += align_up((size_t)cap * words * 4, 256);
    int32_t* degree = (int32_t*)scratch; scratch += align_up((size_t)cap * 4, 256);
    int32_t* mask = (int32_t*)scratch; scratch += 2 * align_up((size_t)cap * 4, 256) + align_up((size_t)cap * 8, 256);
    float* a2 = (float*)scratch; scratch += align_up((size_t)cap * 12, 256);
    float* b2 = (float*)scratch;
    OVO_LAUNCH(k_consistency, dim3(cap), dim3(256), 0, st, prev, cur, count, cap, thr, bits, degree);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_clique, dim3(1), dim3(1024), 0, st, bits, degree, count, cap, mask);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_compact_points, dim3(1), dim3(1024), 0, st, prev, cur, mask, count, cap, a2, b2);
    OVO_LAUNCH_CHECK();
    OVO_CUDA(cudaMemcpyAsync(prev, a2, (size_t)cap * 12, cudaMemcpyDeviceToDevice, st));
    OVO_CUDA(cudaMemcpyAsync(cur, b2, (size_t)cap * 12, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// T: device f64 [12+] (rows of [R|t], e.g. the output of the rigid-transform kernel)
int outlier_filter_launch(float* prev, float* cur, int32_t* count, int cap, const double* T, double thr, uint8_t* scratch, cudaStream_t st) {
    const size_t words = (cap + 31) / 32;
    scratch += align_up((size_t)cap * words * 4, 256) + align_up((size_t)cap * 4, 256);
    int32_t* mask = (int32_t*)scratch; scratch += 2 * align_up((size_t)cap * 4, 256);
    double* err = (double*)scratch; scratch += align_up((size_t)cap * 8, 256);
    float* a2 = (float*)scratch; scratch += align_up((size_t)cap * 12, 256);
    float* b2 = (float*)scratch;
    OVO_LAUNCH(k_residuals, dim3(cdiv(cap, 256)), dim3(256), 0, st, prev, cur, count, cap, T, err);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_median_keep, dim3(1), dim3(1024), 0, st, err, count, cap, thr, mask);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_compact_points, dim3(1), dim3(1024), 0, st, prev, cur, mask, count, cap, a2, b2);
    OVO_LAUNCH_CHECK();
    OVO_CUDA(cudaMemcpyAsync(prev, a2, (size_t)cap * 12, cudaMemcpyDeviceToDevice, st));
    OVO_CUDA(cudaMemcpyAsync(cur, b2, (size_t)cap * 12, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace ovo
