// OPT-IN robust pose stage (north-star stage 5): batched P3P RANSAC over a fixed, seed-determined
// hypothesis schedule + a small Levenberg-Marquardt refinement of the reprojection error on the inliers.
//
// NOT the reference's behaviour: openVO's pose is cv2.estimateAffine3D / Umeyama (ref: src/openVO/stereo_odometer.py:204; SURVEY.md
// §0 D1) and that stays the default.  OpenCV's solvePnPRansac samples from cv::RNG, so bit parity with it is undefined ("parity
// unpinned"); the specification this file implements is restated in oracle/pnp_restate.py (same schedule, same P3P formulation, same
// LM) and both agree with cv2.solvePnPRansac to ~1e-9 whenever the inlier sets coincide (tests).
//
//   k_pnp_hypotheses : one thread per hypothesis — 4 correspondences by splitmix64(seed, h, ctr), P3P on the first three (law of
//                      cosines in depth ratios -> quartic in v = s3/s1, real roots in [1/16, 16] by sign scan + bisection/Newton),
//                      rigid alignment by triads, 4th point disambiguates
//   k_pnp_score      : one CTA per hypothesis — inlier count over all correspondences
//   k_pnp_refine     : one CTA — argmax (ties to the lowest h), inlier mask, 20 LM iterations with block-reduced normal equations
#include "common.cuh"
#include <cmath>

namespace ovo {

namespace {

struct PnpCam { double f, cx, cy, x0, y0; };  // pixel = f * X/Z + c (full-image coordinates); x0,y0 = crop origin of the keypoints

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ void load_corr(const float* __restrict__ pts, const int32_t* __restrict__ matches, const float* __restrict__ kp2,
                                          const PnpCam& cam, int i, double (&X)[3], double (&uv)[2]) {
    X[0] = pts[3 * i]; X[1] = pts[3 * i + 1]; X[2] = pts[3 * i + 2];
    const int t = matches[3 * i + 1];
    uv[0] = (double)kp2[6 * (size_t)t] + cam.x0;
    uv[1] = (double)kp2[6 * (size_t)t + 1] + cam.y0;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// columns e1, e2, e3 of the orthonormal frame of three points (row-major 3x3: M[r][c] = e_c[r])
__device__ void triad(const double (&P)[3][3], double (&M)[9]) {
    double e1[3], d[3], e3[3], e2[3];
    for (int k = 0; k < 3; k++) { e1[k] = P[1][k] - P[0][k]; d[k] = P[2][k] - P[0][k]; }
    double n = norm3(e1);
    for (int k = 0; k < 3; k++) e1[k] /= n;
    cross3(e1, d, e3);
    n = norm3(e3);
    for (int k = 0; k < 3; k++) e3[k] /= n;
    cross3(e3, e1, e2);
    for (int r = 0; r < 3; r++) { M[3 * r] = e1[r]; M[3 * r + 1] = e2[r]; M[3 * r + 2] = e3[r]; }
}

__device__ __forceinline__ bool project(const double* R, const double* t, const double* X, const PnpCam& cam, double& u, double& v) {
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    u = cam.f * x / z + cam.cx;
    v = cam.f * y / z + cam.cy;
    return z > 0 && isfinite(u) && isfinite(v);
}

__global__ void __launch_bounds__(128) k_pnp_hypotheses(const float* __restrict__ pts, const int32_t* __restrict__ matches,
                                                        const float* __restrict__ kp2, const int32_t* __restrict__ count, int cap, PnpCam cam,
                                                        int iters, uint64_t seed, double* __restrict__ poses, int32_t* __restrict__ valid) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= iters) return;
    const int m = min(*count, cap);
    valid[h] = 0;
    if (m < 4) return;
    int idx[4];
    {
        int n = 0;
        for (uint64_t ctr = 0; ctr <= 64 && n < 4; ctr++) {
            const int r = (int)(splitmix64(seed * 0x100000001B3ull + (uint64_t)h * 0x9E3779B1ull + ctr) % (uint64_t)m);
            bool dup = false;
            for (int k = 0; k < n; k++) dup = dup || idx[k] == r;
            if (!dup) idx[n++] = r;
        }
        if (n < 4) return;
    }
    double P[3][3], uv[4][2], X4[3], j[3][3];
    for (int k = 0; k < 3; k++) {
        double X[3];
        load_corr(pts, matches, kp2, cam, idx[k], X, uv[k]);
        for (int c = 0; c < 3; c++) P[k][c] = X[c];
        j[k][0] = (uv[k][0] - cam.cx) / cam.f; j[k][1] = (uv[k][1] - cam.cy) / cam.f; j[k][2] = 1.0;
        const double n = norm3(j[k]);
        for (int c = 0; c < 3; c++) j[k][c] /= n;
    }
    load_corr(pts, matches, kp2, cam, idx[3], X4, uv[3]);
    auto d2 = [&](int a, int b) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += (P[a][c] - P[b][c]) * (P[a][c] - P[b][c]);
        return s;
    };
    auto dot = [&](int a, int b) { return j[a][0] * j[b][0] + j[a][1] * j[b][1] + j[a][2] * j[b][2]; };
    const double a2 = d2(1, 2), b2 = d2(0, 2), c2 = d2(0, 1);
    const double ca = dot(1, 2), cb = dot(0, 2), cg = dot(0, 1);
    // u = N(v) / Dn(v), q(v) = 1 - 2 cb v + v^2 ; quartic: b2 N^2 - 2 b2 cg N Dn + (b2 - c2 q) Dn^2 = 0  (ascending coefficients)
    const double q[3] = {1.0, -2 * cb, 1.0};
    const double N[3] = {b2 + (a2 - c2) * q[0], (a2 - c2) * q[1], -b2 + (a2 - c2) * q[2]};
    const double Dn[2] = {2 * b2 * cg, -2 * b2 * ca};
    double poly[5] = {0, 0, 0, 0, 0};
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) poly[a + b] += b2 * N[a] * N[b];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 2; b++) poly[a + b] -= 2 * b2 * cg * N[a] * Dn[b];
    {
        const double e[3] = {b2 - c2 * q[0], -c2 * q[1], -c2 * q[2]};
        const double dd[3] = {Dn[0] * Dn[0], 2 * Dn[0] * Dn[1], Dn[1] * Dn[1]};
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) poly[a + b] += e[a] * dd[b];
    }
    auto pv = [&](double v) { return (((poly[4] * v + poly[3]) * v + poly[2]) * v + poly[1]) * v + poly[0]; };
    auto dpv = [&](double v) { return ((4 * poly[4] * v + 3 * poly[3]) * v + 2 * poly[2]) * v + poly[1]; };
    double Pm[9];
    triad(P, Pm);
    double best_err = 1e300, bestR[9], bestT[3];
    bool have = false;
    // sign scan on a log grid over [1/16, 16], then bisection + Newton polish
    constexpr int G = 256;
    double vprev = 1.0 / 16, fprev = pv(vprev);
    for (int g = 1; g <= G; g++) {
        const double vcur = exp2(-4.0 + 8.0 * (double)g / G), fcur = pv(vcur);
        if ((fprev <= 0) != (fcur <= 0)) {
            double lo = vprev, hi = vcur, flo = fprev;
            for (int it = 0; it < 60; it++) {
                const double mid = 0.5 * (lo + hi), fm = pv(mid);
                if ((flo <= 0) != (fm <= 0)) hi = mid; else { lo = mid; flo = fm; }
            }
            double v = 0.5 * (lo + hi);
            for (int it = 0; it < 3; it++) {
                const double dv = dpv(v);
                if (dv != 0) {
                    const double vn = v - pv(v) / dv;
                    if (vn > vprev && vn < vcur) v = vn;
                }
            }
            const double den = Dn[0] + Dn[1] * v;
            if (fabs(den) >= 1e-12 * b2) {
                const double u = (N[0] + N[1] * v + N[2] * v * v) / den;
                const double qv = 1 - 2 * cb * v + v * v;
                if (u > 0 && qv > 0) {
                    const double s1 = sqrt(b2 / qv), s[3] = {s1, u * s1, v * s1};
                    double Q[3][3], Qm[9], R[9], t[3];
                    for (int k = 0; k < 3; k++)
                        for (int c = 0; c < 3; c++) Q[k][c] = s[k] * j[k][c];
                    triad(Q, Qm);
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++) R[3 * r + c] = Qm[3 * r] * Pm[3 * c] + Qm[3 * r + 1] * Pm[3 * c + 1] + Qm[3 * r + 2] * Pm[3 * c + 2];
                    for (int r = 0; r < 3; r++) t[r] = Q[0][r] - (R[3 * r] * P[0][0] + R[3 * r + 1] * P[0][1] + R[3 * r + 2] * P[0][2]);
                    double pu, pw;
                    if (project(R, t, X4, cam, pu, pw)) {
                        const double e = (pu - uv[3][0]) * (pu - uv[3][0]) + (pw - uv[3][1]) * (pw - uv[3][1]);
                        if (e < best_err) {
                            best_err = e; have = true;
                            for (int k = 0; k < 9; k++) bestR[k] = R[k];
                            for (int k = 0; k < 3; k++) bestT[k] = t[k];
                        }
                    }
                }
            }
        }
        vprev = vcur; fprev = fcur;
    }
    if (!have) return;
    double* o = poses + 12 * (size_t)h;
    for (int k = 0; k < 9; k++) o[k] = bestR[k];
    for (int k = 0; k < 3; k++) o[9 + k] = bestT[k];
    valid[h] = 1;
}

__global__ void __launch_bounds__(128) k_pnp_score(const float* __restrict__ pts, const int32_t* __restrict__ matches, const float* __restrict__ kp2,
                                                   const int32_t* __restrict__ count, int cap, PnpCam cam, double thr2, const double* __restrict__ poses,
                                                   const int32_t* __restrict__ valid, int32_t* __restrict__ inliers) {
    __shared__ int red[4];
    const int h = blockIdx.x;
    const int m = min(*count, cap);
    if (!valid[h]) {
        if (threadIdx.x == 0) inliers[h] = -1;
        return;
    }
    const double* R = poses + 12 * (size_t)h;
    int n = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double X[3], uv[2], pu, pw;
        load_corr(pts, matches, kp2, cam, i, X, uv);
        if (project(R, R + 9, X, cam, pu, pw) && (pu - uv[0]) * (pu - uv[0]) + (pw - uv[1]) * (pw - uv[1]) < thr2) n++;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) inliers[h] = red[0] + red[1] + red[2] + red[3];
}

__device__ double block_sum256(double v, double* sh) {
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

__device__ void rodrigues_apply(const double* w, const double* R, const double* t, const double* dt, double* Rn, double* tn) {
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double E[9];
    if (th < 1e-12) {
        const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
        for (int k = 0; k < 9; k++) E[k] = K[k] + (k % 4 == 0 ? 1.0 : 0.0);
    } else {
        const double k0 = w[0] / th, k1 = w[1] / th, k2 = w[2] / th;
        const double K[9] = {0, -k2, k1, k2, 0, -k0, -k1, k0, 0};
        double KK[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) KK[3 * r + c] = K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c];
        const double s = sin(th), c1 = 1 - cos(th);
        for (int k = 0; k < 9; k++) E[k] = (k % 4 == 0 ? 1.0 : 0.0) + s * K[k] + c1 * KK[k];
    }
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rn[3 * r + c] = E[3 * r] * R[c] + E[3 * r + 1] * R[3 + c] + E[3 * r + 2] * R[6 + c];
        tn[r] = E[3 * r] * t[0] + E[3 * r + 1] * t[1] + E[3 * r + 2] * t[2] + dt[r];
    }
}

__global__ void __launch_bounds__(256) k_pnp_refine(const float* __restrict__ pts, const int32_t* __restrict__ matches, const float* __restrict__ kp2,
                                                    const int32_t* __restrict__ count, int cap, PnpCam cam, double thr2, int iters,
                                                    const double* __restrict__ poses, const int32_t* __restrict__ inliers, uint8_t* __restrict__ mask,
                                                    int lm_iters, double* __restrict__ out) {
    __shared__ double sh[256];
    __shared__ double R[9], t[3], Rn[9], tn[3], A[36], g[6];
    __shared__ int best_s;
    __shared__ double lam_s, c0_s;
    const int tid = threadIdx.x;
    const int m = min(*count, cap);
    // argmax inliers, ties to the lowest hypothesis index
    {
        long long key = -1;
        for (int h = tid; h < iters; h += 256)
            if (inliers[h] >= 0) {
                const long long k = ((long long)inliers[h] << 20) | (long long)(0xFFFFF - h);
                key = k > key ? k : key;
            }
        sh[tid] = (double)key;  // exact: < 2^53
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (tid < s) sh[tid] = fmax(sh[tid], sh[tid + s]);
            __syncthreads();
        }
        if (tid == 0) {
            const long long k = (long long)sh[0];
            best_s = k < 0 ? -1 : (int)(0xFFFFF - (k & 0xFFFFF));
        }
        __syncthreads();
    }
    const int best = best_s;
    const int n_inl = best >= 0 ? inliers[best] : 0;
    if (best < 0 || n_inl < 4) {
        if (tid < 16) out[tid] = tid == 12 ? (double)n_inl : (tid == 15 ? (double)m : nan(""));
        return;
    }
    if (tid < 9) R[tid] = poses[12 * (size_t)best + tid];
    if (tid < 3) t[tid] = poses[12 * (size_t)best + 9 + tid];
    __syncthreads();
    for (int i = tid; i < m; i += 256) {
        double X[3], uv[2], pu, pw;
        load_corr(pts, matches, kp2, cam, i, X, uv);
        mask[i] = (project(R, t, X, cam, pu, pw) && (pu - uv[0]) * (pu - uv[0]) + (pw - uv[1]) * (pw - uv[1]) < thr2) ? 1 : 0;
    }
    __syncthreads();
    auto cost_of = [&](const double* Rc, const double* tc) {
        double c = 0;
        for (int i = tid; i < m; i += 256)
            if (mask[i]) {
                double X[3], uv[2], pu, pw;
                load_corr(pts, matches, kp2, cam, i, X, uv);
                project(Rc, tc, X, cam, pu, pw);
                c += (pu - uv[0]) * (pu - uv[0]) + (pw - uv[1]) * (pw - uv[1]);
            }
        return block_sum256(c, sh);
    };
    {
        const double c = cost_of(R, t);
        if (tid == 0) { c0_s = c; lam_s = 1e-3; }
        __syncthreads();
    }
    for (int it = 0; it < lm_iters; it++) {
        double a[21], b[6];
        for (int k = 0; k < 21; k++) a[k] = 0;
        for (int k = 0; k < 6; k++) b[k] = 0;
        for (int i = tid; i < m; i += 256)
            if (mask[i]) {
                double X[3], uv[2];
                load_corr(pts, matches, kp2, cam, i, X, uv);
                const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
                const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
                const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
                const double r0 = cam.f * x / z + cam.cx - uv[0], r1 = cam.f * y / z + cam.cy - uv[1];
                const double fz = cam.f / z, fx = -cam.f * x / (z * z), fy = -cam.f * y / (z * z);
                // rows of J = dproj/dY * [-[Y]x | I]
                const double J0[6] = {fx * y, fz * z - fx * x, -fz * y, fz, 0, fx};
                const double J1[6] = {-fz * z + fy * y, -fy * x, fz * x, 0, fz, fy};
                int k = 0;
                for (int p = 0; p < 6; p++) {
                    for (int qd = p; qd < 6; qd++) a[k++] += J0[p] * J0[qd] + J1[p] * J1[qd];
                    b[p] += J0[p] * r0 + J1[p] * r1;
                }
            }
        {
            int k = 0;
            for (int p = 0; p < 6; p++)
                for (int qd = p; qd < 6; qd++) {
                    const double s = block_sum256(a[k++], sh);
                    if (tid == 0) { A[6 * p + qd] = s; A[6 * qd + p] = s; }
                }
            for (int p = 0; p < 6; p++) {
                const double s = block_sum256(b[p], sh);
                if (tid == 0) g[p] = s;
            }
        }
        __syncthreads();
        if (tid == 0) {
            // solve (A + lam diag(A)) step = -g by Gaussian elimination with partial pivoting
            double M[6][7];
            for (int p = 0; p < 6; p++) {
                for (int qd = 0; qd < 6; qd++) M[p][qd] = A[6 * p + qd] + (p == qd ? lam_s * A[6 * p + p] : 0.0);
                M[p][6] = -g[p];
            }
            bool ok = true;
            for (int c = 0; c < 6 && ok; c++) {
                int piv = c;
                for (int r = c + 1; r < 6; r++)
                    if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
                if (fabs(M[piv][c]) < 1e-300) { ok = false; break; }
                if (piv != c)
                    for (int k = 0; k < 7; k++) { const double tmp = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = tmp; }
                for (int r = c + 1; r < 6; r++) {
                    const double fct = M[r][c] / M[c][c];
                    for (int k = c; k < 7; k++) M[r][k] -= fct * M[c][k];
                }
            }
            double step[6] = {0, 0, 0, 0, 0, 0};
            if (ok)
                for (int r = 5; r >= 0; r--) {
                    double s = M[r][6];
                    for (int k = r + 1; k < 6; k++) s -= M[r][k] * step[k];
                    step[r] = s / M[r][r];
                }
            rodrigues_apply(step, R, t, step + 3, Rn, tn);
        }
        __syncthreads();
        const double c1 = cost_of(Rn, tn);
        if (tid == 0) {
            if (isfinite(c1) && c1 < c0_s) {
                for (int k = 0; k < 9; k++) R[k] = Rn[k];
                for (int k = 0; k < 3; k++) t[k] = tn[k];
                c0_s = c1;
                lam_s *= 0.1;
            } else {
                lam_s *= 10.0;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int r = 0; r < 3; r++) {
            out[4 * r] = R[3 * r]; out[4 * r + 1] = R[3 * r + 1]; out[4 * r + 2] = R[3 * r + 2]; out[4 * r + 3] = t[r];
        }
        out[12] = (double)n_inl;
        const double sx = R[7] - R[5], sy = R[2] - R[6], sz = R[3] - R[1];
        out[13] = atan2(0.5 * sqrt(sx * sx + sy * sy + sz * sz), 0.5 * (R[0] + R[4] + R[8] - 1.0));
        out[14] = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        out[15] = (double)best;
    }
}

}  // namespace

size_t pnp_scratch_bytes(int max_iters, int cap) {
    return align_up((size_t)max_iters * 12 * 8, 256) + 2 * align_up((size_t)max_iters * 4, 256) + align_up((size_t)cap, 256);
}

int pnp_ransac_launch(const float* pts, const int32_t* matches, const float* kp2, const int32_t* count, int cap, const double* Q16, int x0,
                      int y0, int iters, int max_iters, double thr_px, uint64_t seed, uint8_t* scratch, double* out, cudaStream_t st) {
    if (iters < 1 || iters > max_iters) { set_error("ransac iterations must be in [1, %d]", max_iters); return 1; }
    PnpCam cam;
    cam.f = Q16[2 * 4 + 3]; cam.cx = -Q16[0 * 4 + 3]; cam.cy = -Q16[1 * 4 + 3]; cam.x0 = x0; cam.y0 = y0;
    double* poses = (double*)scratch; scratch += align_up((size_t)max_iters * 12 * 8, 256);
    int32_t* valid = (int32_t*)scratch; scratch += align_up((size_t)max_iters * 4, 256);
    int32_t* inl = (int32_t*)scratch; scratch += align_up((size_t)max_iters * 4, 256);
    uint8_t* mask = scratch;
    OVO_LAUNCH(k_pnp_hypotheses, dim3(cdiv(iters, 128)), dim3(128), 0, st, pts, matches, kp2, count, cap, cam, iters, seed, poses, valid);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_pnp_score, dim3(iters), dim3(128), 0, st, pts, matches, kp2, count, cap, cam, thr_px * thr_px, poses, valid, inl);
    OVO_LAUNCH_CHECK();
    OVO_LAUNCH(k_pnp_refine, dim3(1), dim3(256), 0, st, pts, matches, kp2, count, cap, cam, thr_px * thr_px, iters, poses, inl, mask, 20, out);
    OVO_LAUNCH_CHECK();
    return 0;
}

}  // namespace ovo
