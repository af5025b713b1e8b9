// ORACLE (test infrastructure only — never imported by the product path).
//
// CPU restatement of cv2.StereoSGBM (MODE_SGBM, OpenCV 4.13.0), the arithmetic behind the reference call
//   ref: src/openVO/stereo_camera.py:23-27 (StereoSGBM_create, 10 positional args, no mode -> MODE_SGBM)
//   ref: src/openVO/stereo_camera.py:51    (stereoSGBM.compute(L, R))
// OpenCV is an un-vendored, un-pinned dependency of the reference (setup.cfg has no install_requires); its
// source is not on this box.  The algorithm below follows SURVEY.md Appendix A.4 and is pinned against the
// installed cv2 4.13.0 binary by tests/test_oracle_vs_cv2.py and the golden fixtures in tests/golden/.
//
// Plain scalar C++; intermediate volumes (C, S) are exposed so each CUDA sub-stage can be checked alone.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

struct SgbmP {
    int W, H, D, bs, P1, P2, uniq, disp12, ftzero, speckleWin, speckleRange;
};

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// A.4.1: gradient row (Sobel-x, clipped to +-ftzero, biased by ftzero) and raw row; first/last element = ftzero.
void prep_rows(const uint8_t* img, int W, int H, int y, int ftzero, std::vector<int>& g, std::vector<int>& r) {
    const uint8_t* row = img + (size_t)y * W;
    const uint8_t* rn = img + (size_t)std::max(y - 1, 0) * W;
    const uint8_t* rs = img + (size_t)std::min(y + 1, H - 1) * W;
    g.assign(W, ftzero);
    r.assign(W, ftzero);
    for (int x = 1; x <= W - 2; x++) {
        int v = 2 * (row[x + 1] - row[x - 1]) + (rn[x + 1] - rn[x - 1]) + (rs[x + 1] - rs[x - 1]);
        g[x] = clampi(v, -ftzero, ftzero) + ftzero;
        r[x] = row[x];
    }
}

inline void lrminmax(const std::vector<int>& v, int x, int W, int& vmin, int& vmax) {
    int c = v[x];
    int vl = x > 0 ? (c + v[x - 1]) / 2 : c;
    int vr = x < W - 1 ? (c + v[x + 1]) / 2 : c;
    vmin = std::min(std::min(vl, vr), c);
    vmax = std::max(std::max(vl, vr), c);
}

// pix[x1*D + d] for one row y (A.4.1)
void pix_row(const uint8_t* L, const uint8_t* R, const SgbmP& p, int y, std::vector<int16_t>& pix) {
    const int W = p.W, D = p.D, W1 = W - D;
    std::vector<int> gl, rl, gr, rr;
    prep_rows(L, W, p.H, y, p.ftzero, gl, rl);
    prep_rows(R, W, p.H, y, p.ftzero, gr, rr);
    pix.assign((size_t)W1 * D, 0);
    for (int pass = 0; pass < 2; pass++) {
        const std::vector<int>& u_ = pass == 0 ? gl : rl;
        const std::vector<int>& v_ = pass == 0 ? gr : rr;
        const int sh = pass == 0 ? 0 : 2;
        for (int x1 = 0; x1 < W1; x1++) {
            int x = x1 + D;
            int u = u_[x], umin, umax;
            lrminmax(u_, x, W, umin, umax);
            for (int d = 0; d < D; d++) {
                int xr = x - d;
                int v = v_[xr], vmin, vmax;
                lrminmax(v_, xr, W, vmin, vmax);
                int c0 = std::max(0, std::max(u - vmax, vmin - u));
                int c1 = std::max(0, std::max(v - umax, umin - v));
                pix[(size_t)x1 * D + d] += (int16_t)(std::min(c0, c1) >> sh);
            }
        }
    }
}

inline int16_t sat16(int v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }

// one path step (A.4.3)
inline void path_step(const int16_t* C, const int16_t* Lp, int16_t* Lout, int D, int P1, int P2, bool pred_in_image) {
    if (!pred_in_image) {
        for (int d = 0; d < D; d++) Lout[d] = C[d];
        return;
    }
    int m = Lp[0];
    for (int d = 1; d < D; d++) m = std::min(m, (int)Lp[d]);
    for (int d = 0; d < D; d++) {
        int a = Lp[d];
        int b = (d > 0 ? (int)Lp[d - 1] : 32767) + P1;
        int c = (d < D - 1 ? (int)Lp[d + 1] : 32767) + P1;
        int e = m + P2;
        Lout[d] = (int16_t)(C[d] + std::min(std::min(a, b), std::min(c, e)) - m);
    }
}

void median3(const int16_t* src, int16_t* dst, int W, int H) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int16_t v[9];
            int k = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    v[k++] = src[(size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)];
            std::nth_element(v, v + 4, v + 9);
            dst[(size_t)y * W + x] = v[4];
        }
}

// A.4.6 filterSpeckles: 4-connected components with |a-b| <= maxDiff; size <= maxSize -> newVal
void speckles(int16_t* img, int W, int H, int newVal, int maxSize, int maxDiff) {
    std::vector<int> label((size_t)W * H, 0);
    std::vector<int> stack;
    std::vector<int> comp;
    for (int i = 0; i < W * H; i++) {
        if (img[i] == newVal || label[i]) continue;
        comp.clear();
        stack.clear();
        stack.push_back(i);
        label[i] = 1;
        while (!stack.empty()) {
            int p = stack.back();
            stack.pop_back();
            comp.push_back(p);
            int px = p % W, py = p / W;
            const int nb[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
            for (auto& o : nb) {
                int qx = px + o[0], qy = py + o[1];
                if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
                int q = qy * W + qx;
                if (label[q] || img[q] == newVal) continue;
                if (std::abs((int)img[p] - (int)img[q]) > maxDiff) continue;
                label[q] = 1;
                stack.push_back(q);
            }
        }
        if ((int)comp.size() <= maxSize) {
            // defer writes: mark with label 2 and blank afterwards so later fills still see original values
            for (int q : comp) label[q] = 2;
        }
    }
    for (int i = 0; i < W * H; i++)
        if (label[i] == 2) img[i] = (int16_t)newVal;
}

}  // namespace

extern "C" {

// Full StereoSGBM.compute restatement.  Optional outputs (may be null):
//   C_out, S_out : int16 [H][W1][D] aggregated cost / final 5-path sum
//   raw_out      : int16 [H][W] disparity after LR check, before median/speckle
//   med_out      : int16 [H][W] after the 3x3 median, before the speckle filter
// Returns 0, or 1 if the parameters leave the pinned validity domain (A.4 "Validity domain").
static int sgbm_hh(const uint8_t* L, const uint8_t* R, const SgbmP& p, int16_t* disp_out);

int orc_sgbm_compute_mode(const uint8_t* L, const uint8_t* R, int W, int H, int minDisparity, int numDisparities,
                          int blockSize, int P1, int P2, int disp12MaxDiff, int preFilterCap, int uniquenessRatio,
                          int speckleWindowSize, int speckleRange, int mode, int16_t* disp_out);

int orc_sgbm_compute(const uint8_t* L, const uint8_t* R, int W, int H, int minDisparity, int numDisparities,
                     int blockSize, int P1, int P2, int disp12MaxDiff, int preFilterCap, int uniquenessRatio,
                     int speckleWindowSize, int speckleRange, int16_t* disp_out, int16_t* C_out, int16_t* S_out,
                     int16_t* raw_out, int16_t* med_out) {
    SgbmP p;
    // OpenCV's own normalisation of non-positive arguments (cv2.StereoSGBM_create's defaults are P1 = P2 = 0): P1 <= 0 -> 2,
    // P2 <= 0 -> 5, then P2 = max(P2, P1 + 1); uniquenessRatio < 0 -> 10 (pinned against cv2 in tests/test_oracle.py)
    P1 = P1 > 0 ? P1 : 2;
    P2 = std::max(P2 > 0 ? P2 : 5, P1 + 1);
    uniquenessRatio = uniquenessRatio >= 0 ? uniquenessRatio : 10;
    p.W = W; p.H = H; p.D = numDisparities; p.bs = blockSize; p.P1 = P1;
    p.P2 = P2;
    p.uniq = uniquenessRatio;
    p.disp12 = disp12MaxDiff > 0 ? disp12MaxDiff : 1;
    p.ftzero = std::max(preFilterCap, 15) | 1;
    p.speckleWin = speckleWindowSize; p.speckleRange = speckleRange;
    if (minDisparity != 0 || p.D <= 0 || p.D % 16 || W <= p.D || blockSize < 1 || !(blockSize & 1)) return 1;
    if (blockSize * blockSize * (2 * p.ftzero + 63) + p.P2 > 32767) return 1;
    const int D = p.D, W1 = W - D, SW2 = blockSize / 2, SH2 = blockSize / 2;
    const int INV = -16, MAXC = 32767;
    const size_t rowsz = (size_t)W1 * D;

    // A.4.1 + A.4.2: per-row pixel costs, horizontal box sums (clamped in the W1 domain), vertical box sums.
    std::vector<std::vector<int16_t>> hs(H);
    {
        std::vector<int16_t> pix;
        for (int y = 0; y < H; y++) {
            pix_row(L, R, p, y, pix);
            hs[y].assign(rowsz, 0);
            for (int x1 = 0; x1 < W1; x1++)
                for (int dx = -SW2; dx <= SW2; dx++) {
                    const int16_t* s = &pix[(size_t)clampi(x1 + dx, 0, W1 - 1) * D];
                    int16_t* o = &hs[y][(size_t)x1 * D];
                    for (int d = 0; d < D; d++) o[d] += s[d];
                }
        }
    }
    std::vector<int16_t> Crow(rowsz), Srow(rowsz);
    std::vector<int16_t> Lprev[3], Lcur[3], L0(rowsz);  // dirs 1..3 need the previous row
    for (int r = 0; r < 3; r++) { Lprev[r].assign(rowsz, 0); Lcur[r].assign(rowsz, 0); }
    std::vector<int16_t> disp1((size_t)W * H, (int16_t)INV);
    std::vector<int> disp2(W), disp2cost(W);
    std::vector<int16_t> Lh(D), Lh_prev(D);

    for (int y = 0; y < H; y++) {
        std::fill(Crow.begin(), Crow.end(), 0);
        for (int dy = -SH2; dy <= SH2; dy++) {
            const std::vector<int16_t>& h = hs[clampi(y + dy, 0, H - 1)];
            for (size_t i = 0; i < rowsz; i++) Crow[i] += h[i];
        }
        if (C_out) memcpy(C_out + (size_t)y * rowsz, Crow.data(), rowsz * 2);
        // forward sweep: L0 (x1-1,y), L1 (x1-1,y-1), L2 (x1,y-1), L3 (x1+1,y-1)
        for (int x1 = 0; x1 < W1; x1++) {
            const int16_t* C = &Crow[(size_t)x1 * D];
            path_step(C, x1 > 0 ? &L0[(size_t)(x1 - 1) * D] : nullptr, &L0[(size_t)x1 * D], D, P1, p.P2, x1 > 0);
            path_step(C, (x1 > 0 && y > 0) ? &Lprev[0][(size_t)(x1 - 1) * D] : nullptr, &Lcur[0][(size_t)x1 * D], D, P1, p.P2, x1 > 0 && y > 0);
            path_step(C, y > 0 ? &Lprev[1][(size_t)x1 * D] : nullptr, &Lcur[1][(size_t)x1 * D], D, P1, p.P2, y > 0);
            path_step(C, (x1 < W1 - 1 && y > 0) ? &Lprev[2][(size_t)(x1 + 1) * D] : nullptr, &Lcur[2][(size_t)x1 * D], D, P1, p.P2, x1 < W1 - 1 && y > 0);
            for (int d = 0; d < D; d++) {
                size_t i = (size_t)x1 * D + d;
                Srow[i] = sat16((int)L0[i] + Lcur[0][i] + Lcur[1][i] + Lcur[2][i]);
            }
        }
        for (int r = 0; r < 3; r++) std::swap(Lprev[r], Lcur[r]);
        // right-to-left sweep: L4 + selection (A.4.4)
        for (int x = 0; x < W; x++) { disp2[x] = INV; disp2cost[x] = MAXC; }
        int16_t* d1row = &disp1[(size_t)y * W];
        for (int x1 = W1 - 1; x1 >= 0; x1--) {
            const int16_t* C = &Crow[(size_t)x1 * D];
            path_step(C, Lh_prev.data(), Lh.data(), D, P1, p.P2, x1 < W1 - 1);
            int16_t* S = &Srow[(size_t)x1 * D];
            int best = 0, minS = MAXC + 1;
            for (int d = 0; d < D; d++) {
                S[d] = sat16((int)S[d] + Lh[d]);
                if (S[d] < minS) { minS = S[d]; best = d; }
            }
            std::swap(Lh, Lh_prev);
            bool uniq_ok = true;
            for (int d = 0; d < D; d++)
                if (S[d] * (100 - p.uniq) < minS * 100 && std::abs(best - d) > 1) { uniq_ok = false; break; }
            if (!uniq_ok) continue;
            int x = x1 + D;
            int x2 = x - best;
            if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = best; }
            int dsp;
            if (best > 0 && best < D - 1) {
                int den = std::max(S[best - 1] + S[best + 1] - 2 * S[best], 1);
                dsp = best * 16 + ((S[best - 1] - S[best + 1]) * 16 + den) / (2 * den);
            } else dsp = best * 16;
            d1row[x] = (int16_t)dsp;
        }
        if (S_out) memcpy(S_out + (size_t)y * rowsz, Srow.data(), rowsz * 2);
        // A.4.5 LR check
        for (int x = D; x < W; x++) {
            int d1 = d1row[x];
            if (d1 == INV) continue;
            int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
            int _x = x - _d, x_ = x - d_;
            if (0 <= _x && _x < W && disp2[_x] >= 0 && std::abs(disp2[_x] - _d) > p.disp12 &&
                0 <= x_ && x_ < W && disp2[x_] >= 0 && std::abs(disp2[x_] - d_) > p.disp12)
                d1row[x] = (int16_t)INV;
        }
    }
    if (raw_out) memcpy(raw_out, disp1.data(), (size_t)W * H * 2);
    std::vector<int16_t> med((size_t)W * H);
    median3(disp1.data(), med.data(), W, H);
    if (med_out) memcpy(med_out, med.data(), (size_t)W * H * 2);
    if (speckleWindowSize > 0) speckles(med.data(), W, H, INV, speckleWindowSize, 16 * speckleRange);
    memcpy(disp_out, med.data(), (size_t)W * H * 2);
    return 0;
}

// mode 0 = MODE_SGBM (what the reference uses), 1 = MODE_HH (8 paths, two passes over stored C / S volumes; SURVEY.md A.4 end;
// NOT used by the reference — opt-in extension row n4, pinned against cv2's own MODE_HH)
int orc_sgbm_compute_mode(const uint8_t* L, const uint8_t* R, int W, int H, int minDisparity, int numDisparities,
                          int blockSize, int P1, int P2, int disp12MaxDiff, int preFilterCap, int uniquenessRatio,
                          int speckleWindowSize, int speckleRange, int mode, int16_t* disp_out) {
    if (mode == 0)
        return orc_sgbm_compute(L, R, W, H, minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap,
                                uniquenessRatio, speckleWindowSize, speckleRange, disp_out, nullptr, nullptr, nullptr, nullptr);
    SgbmP p;
    // OpenCV's own normalisation of non-positive arguments (cv2.StereoSGBM_create's defaults are P1 = P2 = 0): P1 <= 0 -> 2,
    // P2 <= 0 -> 5, then P2 = max(P2, P1 + 1); uniquenessRatio < 0 -> 10 (pinned against cv2 in tests/test_oracle.py)
    P1 = P1 > 0 ? P1 : 2;
    P2 = std::max(P2 > 0 ? P2 : 5, P1 + 1);
    uniquenessRatio = uniquenessRatio >= 0 ? uniquenessRatio : 10;
    p.W = W; p.H = H; p.D = numDisparities; p.bs = blockSize; p.P1 = P1;
    p.P2 = P2;
    p.uniq = uniquenessRatio;
    p.disp12 = disp12MaxDiff > 0 ? disp12MaxDiff : 1;
    p.ftzero = std::max(preFilterCap, 15) | 1;
    p.speckleWin = speckleWindowSize; p.speckleRange = speckleRange;
    if (minDisparity != 0 || p.D <= 0 || p.D % 16 || W <= p.D || blockSize < 1 || !(blockSize & 1)) return 1;
    if (blockSize * blockSize * (2 * p.ftzero + 63) + p.P2 > 32767) return 1;
    return sgbm_hh(L, R, p, disp_out);
}

}  // extern "C"

static int sgbm_hh(const uint8_t* L, const uint8_t* R, const SgbmP& p, int16_t* disp_out) {
    const int W = p.W, H = p.H, D = p.D, W1 = W - D, SW2 = p.bs / 2, SH2 = p.bs / 2;
    const int INV = -16, MAXC = 32767;
    const size_t rowsz = (size_t)W1 * D;
    std::vector<std::vector<int16_t>> hs(H);
    {
        std::vector<int16_t> pix;
        for (int y = 0; y < H; y++) {
            pix_row(L, R, p, y, pix);
            hs[y].assign(rowsz, 0);
            for (int x1 = 0; x1 < W1; x1++)
                for (int dx = -SW2; dx <= SW2; dx++) {
                    const int16_t* s = &pix[(size_t)clampi(x1 + dx, 0, W1 - 1) * D];
                    int16_t* o = &hs[y][(size_t)x1 * D];
                    for (int d = 0; d < D; d++) o[d] += s[d];
                }
        }
    }
    std::vector<int16_t> C((size_t)H * rowsz, 0), S((size_t)H * rowsz, 0);
    for (int y = 0; y < H; y++)
        for (int dy = -SH2; dy <= SH2; dy++) {
            const std::vector<int16_t>& h = hs[clampi(y + dy, 0, H - 1)];
            int16_t* c = &C[(size_t)y * rowsz];
            for (size_t i = 0; i < rowsz; i++) c[i] += h[i];
        }
    hs.clear();
    std::vector<int16_t> disp1((size_t)W * H, (int16_t)INV);
    std::vector<int> disp2(W), disp2cost(W);
    for (int pass = 0; pass < 2; pass++) {
        const int y0 = pass == 0 ? 0 : H - 1, y1 = pass == 0 ? H : -1, dy = pass == 0 ? 1 : -1;
        const int x0 = pass == 0 ? 0 : W1 - 1, x1e = pass == 0 ? W1 : -1, dx = pass == 0 ? 1 : -1;
        std::vector<int16_t> Lprev[3], Lcur[3], L0(rowsz);
        for (int r = 0; r < 3; r++) { Lprev[r].assign(rowsz, 0); Lcur[r].assign(rowsz, 0); }
        for (int y = y0; y != y1; y += dy) {
            const int16_t* Crow = &C[(size_t)y * rowsz];
            int16_t* Srow = &S[(size_t)y * rowsz];
            const bool have_prev_row = y != y0;
            if (pass == 1) for (int x = 0; x < W; x++) { disp2[x] = INV; disp2cost[x] = MAXC; }
            for (int x1 = x0; x1 != x1e; x1 += dx) {
                const int16_t* Cc = &Crow[(size_t)x1 * D];
                const int xp = x1 - dx, xn = x1 + dx;  // previous / next column in sweep order
                const bool in_p = xp >= 0 && xp < W1, in_n = xn >= 0 && xn < W1;
                path_step(Cc, in_p ? &L0[(size_t)xp * D] : nullptr, &L0[(size_t)x1 * D], D, p.P1, p.P2, in_p);
                path_step(Cc, (in_p && have_prev_row) ? &Lprev[0][(size_t)xp * D] : nullptr, &Lcur[0][(size_t)x1 * D], D, p.P1, p.P2, in_p && have_prev_row);
                path_step(Cc, have_prev_row ? &Lprev[1][(size_t)x1 * D] : nullptr, &Lcur[1][(size_t)x1 * D], D, p.P1, p.P2, have_prev_row);
                path_step(Cc, (in_n && have_prev_row) ? &Lprev[2][(size_t)xn * D] : nullptr, &Lcur[2][(size_t)x1 * D], D, p.P1, p.P2, in_n && have_prev_row);
                int16_t* Sc = &Srow[(size_t)x1 * D];
                int best = 0, minS = MAXC + 1;
                for (int d = 0; d < D; d++) {
                    size_t i = (size_t)x1 * D + d;
                    Sc[d] = sat16((int)Sc[d] + L0[i] + Lcur[0][i] + Lcur[1][i] + Lcur[2][i]);
                    if (Sc[d] < minS) { minS = Sc[d]; best = d; }
                }
                if (pass == 0) continue;
                bool uniq_ok = true;
                for (int d = 0; d < D; d++)
                    if (Sc[d] * (100 - p.uniq) < minS * 100 && std::abs(best - d) > 1) { uniq_ok = false; break; }
                if (!uniq_ok) continue;
                int x = x1 + D, x2 = x - best;
                if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = best; }
                int dsp;
                if (best > 0 && best < D - 1) {
                    int den = std::max(Sc[best - 1] + Sc[best + 1] - 2 * Sc[best], 1);
                    dsp = best * 16 + ((Sc[best - 1] - Sc[best + 1]) * 16 + den) / (2 * den);
                } else dsp = best * 16;
                disp1[(size_t)y * W + x] = (int16_t)dsp;
            }
            for (int r = 0; r < 3; r++) std::swap(Lprev[r], Lcur[r]);
            if (pass == 1) {
                int16_t* d1row = &disp1[(size_t)y * W];
                for (int x = D; x < W; x++) {
                    int d1 = d1row[x];
                    if (d1 == INV) continue;
                    int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
                    int _x = x - _d, x_ = x - d_;
                    if (0 <= _x && _x < W && disp2[_x] >= 0 && std::abs(disp2[_x] - _d) > p.disp12 &&
                        0 <= x_ && x_ < W && disp2[x_] >= 0 && std::abs(disp2[x_] - d_) > p.disp12)
                        d1row[x] = (int16_t)INV;
                }
            }
        }
    }
    std::vector<int16_t> med((size_t)W * H);
    median3(disp1.data(), med.data(), W, H);
    if (p.speckleWin > 0) speckles(med.data(), W, H, INV, p.speckleWin, 16 * p.speckleRange);
    memcpy(disp_out, med.data(), (size_t)W * H * 2);
    return 0;
}
