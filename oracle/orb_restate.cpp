// ORACLE (test infrastructure only — never imported by the product path).
//
// CPU restatement of cv2.ORB_create(nfeatures=n).detectAndCompute(img, mask) and of
// cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) as shipped in OpenCV 4.13.0 (AVX2 dispatch), the arithmetic behind
//   ref: src/openVO/stereo_odometer.py:22   (ORB_create / BFMatcher.create)
//   ref: src/openVO/stereo_odometer.py:117  (detectAndCompute(next_img, feature_mask(next_disp)))
//   ref: src/openVO/stereo_odometer.py:163-164 (knnMatch + ratio test)
// OpenCV is an un-vendored, un-pinned dependency of the reference; the algorithm follows SURVEY.md
// Appendix A.1-A.3 and is pinned against the installed cv2 binary by tests/test_oracle_vs_cv2.py and the
// golden fixtures.  Build with -ffp-contract=off -mfma: fused multiply-adds appear ONLY where written (fmaf).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

const int kPattern[1024] = {
#include "../openvo_b200/csrc/orb_pattern.inc"
};

const int NLEVELS = 8, EDGE = 31, HALF_PATCH = 15, FAST_T = 20, HARRIS_BLOCK = 7;

struct Level {
    int w, h;
    float scale, inv;
    int nfeat;
    std::vector<uint8_t> img, mask, blur;
};

inline int rinti(float v) { return (int)lrintf(v); }  // round-half-to-even under the default rounding mode

void level_geometry(int W, int H, int nfeatures, Level* lv) {
    const float scaleFactor = 1.2f;
    for (int l = 0; l < NLEVELS; l++) {
        lv[l].scale = (float)std::pow((double)scaleFactor, (double)l);
        lv[l].inv = 1.0f / lv[l].scale;
        lv[l].w = rinti((float)W * lv[l].inv);
        lv[l].h = rinti((float)H * lv[l].inv);
    }
    float factor = (float)(1.0 / (double)scaleFactor);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)NLEVELS));
    int sum = 0;
    for (int l = 0; l < NLEVELS - 1; l++) {
        lv[l].nfeat = rinti(nd);
        sum += lv[l].nfeat;
        nd *= factor;
    }
    lv[NLEVELS - 1].nfeat = std::max(nfeatures - sum, 0);
}

// A.1.2 INTER_LINEAR_EXACT, 8-bit fixed-point coefficients
void resize_exact(const std::vector<uint8_t>& src, int sw, int sh, std::vector<uint8_t>& dst, int dw, int dh) {
    auto coeffs = [](int s, int t, std::vector<int>& idx, std::vector<int>& c1) {
        idx.resize(t); c1.resize(t);
        double scale = 1.0 / ((double)t / (double)s);
        for (int v = 0; v < t; v++) {
            double f = scale * (v + 0.5) - 0.5;
            int i = (int)std::floor(f);
            int c = (int)std::nearbyint((f - i) * 256.0);
            if (i < 0) { i = 0; c = 0; }
            if (i >= s - 1) { i = s - 1; c = 0; }
            idx[v] = i; c1[v] = c;
        }
    };
    std::vector<int> xi, xc, yi, yc;
    coeffs(sw, dw, xi, xc);
    coeffs(sh, dh, yi, yc);
    dst.resize((size_t)dw * dh);
    for (int y = 0; y < dh; y++) {
        const uint8_t* r0 = &src[(size_t)yi[y] * sw];
        const uint8_t* r1 = &src[(size_t)std::min(yi[y] + 1, sh - 1) * sw];
        for (int x = 0; x < dw; x++) {
            int i0 = xi[x], i1 = std::min(i0 + 1, sw - 1);
            uint32_t h0 = r0[i0] * (256 - xc[x]) + r0[i1] * xc[x];
            uint32_t h1 = r1[i0] * (256 - xc[x]) + r1[i1] * xc[x];
            uint32_t ver = h0 * (256 - yc[y]) + h1 * yc[y];
            dst[(size_t)y * dw + x] = (uint8_t)((ver + (1u << 15)) >> 16);
        }
    }
}

// A.1.3 FAST-9/16 score map (score = m-1 for corners, else 0)
const int RING[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                         {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

void fast_scores(const std::vector<uint8_t>& img, int w, int h, std::vector<uint8_t>& score) {
    score.assign((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int d[25];
            int c = img[(size_t)y * w + x];
            for (int k = 0; k < 16; k++) d[k] = c - img[(size_t)(y + RING[k][1]) * w + x + RING[k][0]];
            for (int k = 16; k < 25; k++) d[k] = d[k - 16];
            int m = -1000;
            for (int s = 0; s < 16; s++) {
                int a = 1000, b = 1000;
                for (int j = s; j < s + 9; j++) { a = std::min(a, d[j]); b = std::min(b, -d[j]); }
                m = std::max(m, std::max(a, b));
            }
            if (m > FAST_T) score[(size_t)y * w + x] = (uint8_t)(m - 1);
        }
}

struct Cand { int x, y; float resp; };

struct RespGreater { bool operator()(const Cand& a, const Cand& b) const { return a.resp > b.resp; } };

// A.1.5 KeyPointsFilter::retainBest — order is libstdc++'s introselect permutation
void retain_best(std::vector<Cand>& k, int n) {
    if (n >= 0 && k.size() > (size_t)n) {
        if (n == 0) { k.clear(); return; }
        std::nth_element(k.begin(), k.begin() + n - 1, k.end(), RespGreater());
        float amb = k[n - 1].resp;
        auto e = std::partition(k.begin() + n, k.end(), [amb](const Cand& c) { return c.resp >= amb; });
        k.resize(e - k.begin());
    }
}

inline int px(const Level& L, int x, int y) { return L.img[(size_t)y * L.w + x]; }

// A.1.6
float harris(const Level& L, int x0, int y0) {
    int a = 0, b = 0, c = 0;
    const int r = HARRIS_BLOCK / 2;
    for (int y = y0 - r; y <= y0 + r; y++)
        for (int x = x0 - r; x <= x0 + r; x++) {
            int Ix = 2 * (px(L, x + 1, y) - px(L, x - 1, y)) + (px(L, x + 1, y - 1) - px(L, x - 1, y - 1)) + (px(L, x + 1, y + 1) - px(L, x - 1, y + 1));
            int Iy = 2 * (px(L, x, y + 1) - px(L, x, y - 1)) + (px(L, x - 1, y + 1) - px(L, x - 1, y - 1)) + (px(L, x + 1, y + 1) - px(L, x + 1, y - 1));
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    float scale = 1.f / (4 * HARRIS_BLOCK * 255.f);
    float s4 = scale * scale * scale * scale;
    float fa = (float)a, fb = (float)b, fc = (float)c;
    float k = 0.04f;
    float ab = fa * fb, cc = fc * fc;
    float det = ab - cc;
    float tr = fa + fb;
    float ktr = k * tr;
    float ktr2 = ktr * tr;
    return (det - ktr2) * s4;
}

void make_umax(int* umax) {
    int v, v0, vmax = (int)std::floor(HALF_PATCH * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(HALF_PATCH * std::sqrt(2.f) / 2);
    for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(std::sqrt((double)HALF_PATCH * HALF_PATCH - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// A.1.7 cv::fastAtan2 (float, no FMA)
float fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

float ic_angle(const Level& L, int x0, int y0, const int* umax) {
    int m01 = 0, m10 = 0;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m10 += u * px(L, x0 + u, y0);
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = px(L, x0 + u, y0 + v), vm = px(L, x0 + u, y0 - v);
            vsum += (vp - vm);
            m10 += u * (vp + vm);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// A.2.1 float32 separable 7x7 sigma=2 blur with the AVX2 engine's FMA body / scalar tail split
void blur_level(const Level& L, std::vector<uint8_t>& out) {
    float k[7];
    {
        double e[7], s = 0;
        for (int i = 0; i < 7; i++) { e[i] = std::exp(-(double)((i - 3) * (i - 3)) / 8.0); s += e[i]; }
        for (int i = 0; i < 7; i++) k[i] = (float)(e[i] / s);
    }
    const int w = L.w, h = L.h;
    const int wb = (w / 32) * 32, wc = (w / 4) * 4;
    std::vector<float> rows((size_t)(h + 6) * w);
    for (int yy = -3; yy < h + 3; yy++) {
        const uint8_t* src = &L.img[(size_t)reflect101(yy, h) * w];
        float* dst = &rows[(size_t)(yy + 3) * w];
        for (int x = 0; x < w; x++) {
            float p[7];
            for (int i = 0; i < 7; i++) p[i] = (float)src[reflect101(x - 3 + i, w)];
            float acc;
            if (x < wb) {
                acc = 0.f;
                for (int i = 0; i < 7; i++) acc = fmaf(k[i], p[i], acc);
            } else {
                acc = k[0] * p[0];
                for (int i = 1; i < 7; i++) { float t = k[i] * p[i]; acc = acc + t; }
            }
            dst[x] = acc;
        }
    }
    out.resize((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float* r = &rows[(size_t)(y + 3) * w + x];
            float c = k[3] * r[0];
            for (int j = 1; j <= 3; j++) {
                float s = r[(ptrdiff_t)j * w] + r[-(ptrdiff_t)j * w];
                if (x < wc) c = fmaf(k[3 + j], s, c);
                else { float t = k[3 + j] * s; c = c + t; }
            }
            int v = rinti(c);
            out[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
}

// A.2.2
void rbrief(const Level& L, float ptx, float pty, float angle_deg, uint8_t* desc) {
    float ang = angle_deg * (float)(3.14159265358979323846 / 180.f);
    float a = (float)std::cos((double)ang), b = (float)std::sin((double)ang);
    int cx = rinti(ptx * L.inv), cy = rinti(pty * L.inv);
    auto val = [&](int pxx, int pyy) {
        float xa = pxx * a, yb = pyy * b, xb = pxx * b, ya = pyy * a;
        float x = xa - yb, y = xb + ya;
        int ix = rinti(x), iy = rinti(y);
        return (int)L.blur[(size_t)(cy + iy) * L.w + (cx + ix)];
    };
    memset(desc, 0, 32);
    for (int i = 0; i < 256; i++) {
        const int* q = &kPattern[i * 4];
        if (val(q[0], q[1]) < val(q[2], q[3])) desc[i >> 3] |= (uint8_t)(1 << (i & 7));
    }
}

}  // namespace

extern "C" {

// Level geometry only: out_wh [8][2], out_nfeat [8], out_scale [8]
void orc_orb_levels(int W, int H, int nfeatures, int* out_wh, int* out_nfeat, float* out_scale) {
    Level lv[NLEVELS];
    level_geometry(W, H, nfeatures, lv);
    for (int l = 0; l < NLEVELS; l++) {
        out_wh[2 * l] = lv[l].w; out_wh[2 * l + 1] = lv[l].h;
        out_nfeat[l] = lv[l].nfeat; out_scale[l] = lv[l].scale;
    }
}

// Full ORB.  kp_out: float [cap][6] = (pt.x, pt.y, size, angle, response, octave); desc_out: u8 [cap][32].
// Optional sub-stage dumps (null to skip), all as concatenated levels in level order:
//   pyr_out / maskpyr_out / fast_out / blur_out : u8, sum_l w_l*h_l bytes
//   cand_out: int32 [cand_cap][3] = (level, x, y) of FAST candidates after NMS+mask+border, raster order per
//             level; cand_resp_out: float [cand_cap][2] = (FAST score, Harris response); *n_cand_out = count
// mask may be null.  Returns the number of keypoints (<= cap) or -1 if cap is too small.
int orc_orb_detect_compute(const uint8_t* img, const uint8_t* mask, int W, int H, int nfeatures, float* kp_out,
                           uint8_t* desc_out, int cap, uint8_t* pyr_out, uint8_t* maskpyr_out, uint8_t* fast_out,
                           uint8_t* blur_out, int32_t* cand_out, float* cand_resp_out, int cand_cap,
                           int* n_cand_out) {
    Level lv[NLEVELS];
    level_geometry(W, H, nfeatures, lv);
    int umax[HALF_PATCH + 2];
    make_umax(umax);
    lv[0].img.assign(img, img + (size_t)W * H);
    if (mask) lv[0].mask.assign(mask, mask + (size_t)W * H);
    for (int l = 1; l < NLEVELS; l++) {
        resize_exact(lv[l - 1].img, lv[l - 1].w, lv[l - 1].h, lv[l].img, lv[l].w, lv[l].h);
        if (mask) {
            resize_exact(lv[l - 1].mask, lv[l - 1].w, lv[l - 1].h, lv[l].mask, lv[l].w, lv[l].h);
            for (auto& m : lv[l].mask) m = m > 254 ? m : 0;
        }
    }
    struct KP { float x, y, size, angle, resp; int octave; };
    std::vector<KP> all;
    size_t off = 0;
    int n_cand = 0;
    for (int l = 0; l < NLEVELS; l++) {
        Level& L = lv[l];
        const size_t npx = (size_t)L.w * L.h;
        std::vector<uint8_t> score;
        fast_scores(L.img, L.w, L.h, score);
        if (pyr_out) memcpy(pyr_out + off, L.img.data(), npx);
        if (maskpyr_out && mask) memcpy(maskpyr_out + off, L.mask.data(), npx);
        if (fast_out) memcpy(fast_out + off, score.data(), npx);
        std::vector<Cand> c;
        for (int y = 3; y < L.h - 3; y++)
            for (int x = 3; x < L.w - 3; x++) {
                int s = score[(size_t)y * L.w + x];
                if (!s) continue;
                bool keep = true;
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++)
                        if ((dx || dy) && score[(size_t)(y + dy) * L.w + x + dx] >= s) { keep = false; break; }
                if (!keep) continue;
                if (mask && L.mask[(size_t)y * L.w + x] == 0) continue;
                if (!(x >= EDGE && x < L.w - EDGE && y >= EDGE && y < L.h - EDGE)) continue;
                c.push_back({x, y, (float)s});
            }
        if (cand_out) {
            for (auto& k : c) {
                if (n_cand < cand_cap) {
                    cand_out[3 * n_cand] = l; cand_out[3 * n_cand + 1] = k.x; cand_out[3 * n_cand + 2] = k.y;
                    if (cand_resp_out) { cand_resp_out[2 * n_cand] = k.resp; cand_resp_out[2 * n_cand + 1] = harris(L, k.x, k.y); }
                }
                n_cand++;
            }
        }
        retain_best(c, 2 * L.nfeat);
        for (auto& k : c) k.resp = harris(L, k.x, k.y);
        retain_best(c, L.nfeat);
        for (auto& k : c) {
            KP kp;
            kp.angle = ic_angle(L, k.x, k.y, umax);
            kp.x = (float)k.x * L.scale; kp.y = (float)k.y * L.scale;
            kp.size = 31.f * L.scale; kp.resp = k.resp; kp.octave = l;
            all.push_back(kp);
        }
        blur_level(L, L.blur);
        if (blur_out) memcpy(blur_out + off, L.blur.data(), npx);
        off += npx;
    }
    if (n_cand_out) *n_cand_out = n_cand;
    if ((int)all.size() > cap) return -1;
    for (size_t i = 0; i < all.size(); i++) {
        const KP& k = all[i];
        float* o = kp_out + 6 * i;
        o[0] = k.x; o[1] = k.y; o[2] = k.size; o[3] = k.angle; o[4] = k.resp; o[5] = (float)k.octave;
        rbrief(lv[k.octave], k.x, k.y, k.angle, desc_out + 32 * i);
    }
    return (int)all.size();
}

// A.3: 2-NN Hamming, ties -> lowest train index.  out [nq][4] = (idx0, dist0, idx1, dist1); idx -1 if absent.
void orc_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out) {
    for (int i = 0; i < nq; i++) {
        int b0 = 1 << 30, b1 = 1 << 30, i0 = -1, i1 = -1;
        const uint64_t* a = (const uint64_t*)(q + 32 * (size_t)i);
        for (int j = 0; j < nt; j++) {
            const uint64_t* b = (const uint64_t*)(t + 32 * (size_t)j);
            int d = __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
                    __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
            if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
            else if (d < b1) { b1 = d; i1 = j; }
        }
        out[4 * i] = i0; out[4 * i + 1] = i0 < 0 ? 0 : b0; out[4 * i + 2] = i1; out[4 * i + 3] = i1 < 0 ? 0 : b1;
    }
}

}  // extern "C"
