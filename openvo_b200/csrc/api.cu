// C ABI of openvo_b200 (include/openvo_b200.h): context, workspace carving and the per-seam entry points.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/openvo_b200.h"
#include "common.cuh"

namespace ovo {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- launch counter + optional per-kernel CUDA-event profile (bench.py's roofline leg) -----------------------------------
#ifndef OVO_EMU
namespace {
struct ProfEntry { const char* tag; cudaEvent_t a, b; };
std::atomic<long long> g_launches{0};
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;
std::mutex g_prof_mu;
}  // namespace
void prof_pre(const char* tag, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfEntry e;
    e.tag = tag;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, st);
    g_prof.push_back(e);
}
void prof_post(cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, st);
}
#endif

// defined in orb.cu / match.cu
void orb_make_resize_tables(const OrbDims& d, int32_t* tab, int* tab_off, int* total);
int orb_phase1_launch(const OrbDims& d, const OrbWorkspace* ws0, size_t ws_stride, const int32_t* tab_dev, const int* tab_off, int nb,
                      const uint8_t* img, int pitch, size_t frame_stride, const uint8_t* mask, int mask_pitch,
                      size_t mask_frame_stride, cudaStream_t st);
int orb_phase2_launch(const OrbDims& d, const OrbWorkspace* ws0, size_t ws_stride, int nb, int max_sel, const int32_t* n_sel_dev,
                      float* kp_out, uint8_t* desc_out, cudaStream_t st);
int reproject_launch(const float* disp, int pitch, int cw, int ch, int x0, int y0, const double* Q16, float* xyz, cudaStream_t st);
int disp_post_launch(const int16_t* disp, int W, int H, int x0, int y0, int cw, int ch, float lo, float hi, float* disp_f32,
                     uint8_t* mask, int nb, cudaStream_t st);
int pair_batch_launch(const GatherParams& gp, int n, const void* items_host, int cap, cudaStream_t st);
int pair_item_size();
size_t filter_scratch_bytes(int cap);
size_t pnp_scratch_bytes(int max_iters, int cap);
int pnp_ransac_launch(const float* pts, const int32_t* matches, const float* kp2, const int32_t* count, int cap, const double* Q16, int x0,
                      int y0, int iters, int max_iters, double thr_px, uint64_t seed, uint8_t* scratch, double* out, cudaStream_t st);
static const int kPnpMaxIters = 4096;
int rigid_filter_launch(float* prev, float* cur, int32_t* count, int cap, float thr, uint8_t* scratch, cudaStream_t st);
int outlier_filter_launch(float* prev, float* cur, int32_t* count, int cap, const double* T, double thr, uint8_t* scratch, cudaStream_t st);
int rectify_launch(const uint8_t* img, int pitch, size_t frame_stride, int ch, int W, int H, int nb, const int16_t* map1,
                   const uint16_t* map2, uint8_t* out, cudaStream_t st);
int pair_max_batch();
int crop_launch(const uint8_t* img, int pitch, size_t frame_stride, int x0, int y0, int cw, int ch, int nb, uint8_t* out, cudaStream_t st);

struct Layout {
    SgbmDims sg;
    OrbDims orb;
    int x0, y0, cw, ch;
    size_t sgbm_bytes, orb_bytes, frame_bytes;  // per frame
    size_t tab_bytes, knn_bytes, nsel_bytes, filter_bytes, pnp_bytes, total;
    int tab_off[2 * ORB_NLEVELS], tab_total;
};

static int make_layout(const ovo_config* c, Layout* L) {
    if (!c) { set_error("null config"); return 1; }
    const ovo_sgbm_params& p = c->sgbm;
    if (c->width <= 0 || c->height <= 0 || c->max_batch < 1) { set_error("bad image size / batch"); return 1; }
    if (p.minDisparity != 0) { set_error("minDisparity != 0 is outside the pinned SGBM domain (SURVEY.md A.4)"); return 1; }
    if (p.numDisparities <= 0 || p.numDisparities % 16 || p.numDisparities > 256) { set_error("numDisparities must be a multiple of 16 in [16, 256]"); return 1; }
    if (c->width <= p.numDisparities) { set_error("image narrower than the disparity range"); return 1; }
    if (p.blockSize < 3 || p.blockSize > 11 || !(p.blockSize & 1)) { set_error("blockSize must be odd in [3, 11]"); return 1; }
    SgbmDims& s = L->sg;
    s.W = c->width; s.H = c->height; s.D = p.numDisparities;
    s.Dp = s.D <= 64 ? 64 : (s.D <= 128 ? 128 : 256);
    // OpenCV's own normalisation of non-positive arguments (cv2.StereoSGBM_create defaults to P1 = P2 = 0): P1 <= 0 -> 2,
    // P2 <= 0 -> 5, then P2 = max(P2, P1 + 1); uniquenessRatio < 0 -> 10
    s.W1 = s.W - s.D; s.bs = p.blockSize; s.P1 = p.P1 > 0 ? p.P1 : 2;
    const int p2 = p.P2 > 0 ? p.P2 : 5;
    s.P2 = p2 > s.P1 + 1 ? p2 : s.P1 + 1;
    s.uniq = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
    s.disp12 = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
    s.ftzero = (p.preFilterCap > 15 ? p.preFilterCap : 15) | 1;
    s.speckleWin = p.speckleWindowSize; s.speckleDiff = 16 * p.speckleRange;
    if (c->sgbm_mode != 0 && c->sgbm_mode != 1) { set_error("sgbm_mode must be 0 (MODE_SGBM) or 1 (MODE_HH)"); return 1; }
    s.mode = c->sgbm_mode;
    if (s.ftzero > 127) { set_error("preFilterCap too large"); return 1; }
    if (s.bs * s.bs * (2 * s.ftzero + 63) + s.P2 > 32767) {
        set_error("SGBM parameters leave the int16 cost domain OpenCV's result is pinned for (SURVEY.md A.4 validity domain)");
        return 1;
    }
    // reference slice semantics (bug-compatible B1): img[roi[1]:roi[3], roi[0]:roi[2]] with numpy clamping
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    auto norm = [&](int v, int n) { return clampi(v < 0 ? v + n : v, 0, n); };
    const int y0 = norm(c->roi[1], s.H), y1 = norm(c->roi[3], s.H), x0 = norm(c->roi[0], s.W), x1 = norm(c->roi[2], s.W);
    L->x0 = x0; L->y0 = y0; L->cw = x1 > x0 ? x1 - x0 : 0; L->ch = y1 > y0 ? y1 - y0 : 0;
    if (L->cw < 16 || L->ch < 16) { set_error("cropped frame %dx%d too small", L->cw, L->ch); return 1; }
    if (c->nfeatures < 1) { set_error("nfeatures must be positive"); return 1; }
    orb_make_dims(L->cw, L->ch, c->nfeatures, &L->orb);
    L->sgbm_bytes = sgbm_workspace_bytes(s);
    L->orb_bytes = align_up(orb_workspace_bytes(L->orb), 256);
    L->frame_bytes = L->sgbm_bytes + L->orb_bytes;
    orb_make_resize_tables(L->orb, nullptr, L->tab_off, &L->tab_total);
    L->tab_bytes = align_up((size_t)L->tab_total * 4, 256);
    L->knn_bytes = align_up(knn2_scratch_bytes(L->orb.kp_cap, L->orb.kp_cap), 256) * c->max_batch;
    L->nsel_bytes = align_up((size_t)c->max_batch * 4, 256);
    L->filter_bytes = align_up(filter_scratch_bytes(L->orb.kp_cap), 256);
    L->pnp_bytes = align_up(pnp_scratch_bytes(kPnpMaxIters, L->orb.kp_cap), 256);
    L->total = L->frame_bytes * c->max_batch + L->tab_bytes + L->knn_bytes + L->nsel_bytes + L->filter_bytes + L->pnp_bytes + 256;
    return 0;
}

}  // namespace ovo

using namespace ovo;

struct ovo_ctx {
    ovo_config cfg;
    Layout L;
    uint8_t* base;
    SgbmWorkspace sg0;
    OrbWorkspace orb0;
    int32_t* tab_dev;
    uint32_t* knn_scratch;
    int32_t* nsel_dev;
    uint8_t* filter_scratch;
    uint8_t* pnp_scratch;
    // pinned host staging for the keypoint selection
    int32_t* h_lvl;    // [max_batch][64]
    float* h_resp;     // [max_batch][cand_cap][2]: first half = FAST scores (bytes), second half = Harris of the survivors
    int32_t* h_sel;    // [max_batch][kp_cap]
    int32_t* h_nsel;   // [max_batch]
    long long h2d_bytes = 0, d2h_bytes = 0;  // staging traffic of the keypoint selection
    // ovo_orb_detect_finish_async: the second half of the ORB seam on a worker thread of the library
    int device = 0;
    std::thread worker;              // persistent (one CUDA per-thread initialisation, not one per batch): created on first use
    std::mutex wmu;
    std::condition_variable wcv;
    std::function<int()> wjob;       // pending job
    bool wbusy = false, wquit = false;
    int worker_rc = 0;
    char worker_err[512] = "";
#ifndef OVO_EMU
    // CUDA graphs of ovo_extract_begin, one per distinct argument set (the Python engine calls it with persistent buffers, so
    // there is one per batch size): ~70 kernel launches become one graph launch
    struct ExtractGraph {
        const void *left, *right, *disp16, *disp_f32, *mask, *img;
        int pitch, nb;
        size_t frame_stride;
        cudaGraphExec_t exec;
        long long launches;
    };
    std::vector<ExtractGraph> graphs;
#endif
};

extern "C" {

const char* ovo_last_error(void) { return g_err; }
int ovo_abi_version(void) { return OVO_ABI_VERSION; }

void ovo_transfer_bytes(ovo_ctx* c, long long* h2d, long long* d2h) {
    *h2d = c ? c->h2d_bytes : 0;
    *d2h = c ? c->d2h_bytes : 0;
}

long long ovo_launch_count(void) {
#ifndef OVO_EMU
    return g_launches.load();
#else
    return 0;
#endif
}

void ovo_profile_enable(int on) {
#ifndef OVO_EMU
    g_prof_on = on != 0;
#endif
}

int ovo_profile_read(char* names, int names_len, float* total_ms, int* counts, int max_entries) {
#ifndef OVO_EMU
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<double, int>> acc;
    std::vector<std::string> order;
    for (auto& e : g_prof) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e.a, e.b);
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
        if (!acc.count(e.tag)) order.push_back(e.tag);
        acc[e.tag].first += ms;
        acc[e.tag].second += 1;
    }
    g_prof.clear();
    std::string joined;
    int n = 0;
    for (auto& t : order) {
        if (n >= max_entries) break;
        total_ms[n] = (float)acc[t].first;
        counts[n] = acc[t].second;
        joined += t + "\n";
        n++;
    }
    snprintf(names, names_len, "%s", joined.c_str());
    return n;
#else
    return 0;
#endif
}

int ovo_cropped_size(const ovo_config* cfg, int* cw, int* ch) {
    Layout L;
    if (make_layout(cfg, &L)) return 1;
    *cw = L.cw; *ch = L.ch;
    return 0;
}

int ovo_kp_capacity(const ovo_config* cfg) {
    Layout L;
    if (make_layout(cfg, &L)) return -1;
    return L.orb.kp_cap;
}

size_t ovo_workspace_bytes(const ovo_config* cfg) {
    Layout L;
    if (make_layout(cfg, &L)) return 0;
    return L.total;
}

ovo_ctx* ovo_create(const ovo_config* cfg, void* workspace_dev, size_t workspace_bytes) {
    Layout L;
    if (make_layout(cfg, &L)) return nullptr;
    if (!workspace_dev || workspace_bytes < L.total || ((uintptr_t)workspace_dev & 255)) {
        set_error("workspace must be 256-byte aligned and at least %zu bytes", L.total);
        return nullptr;
    }
    ovo_ctx* c = new ovo_ctx();
    cudaGetDevice(&c->device);
    c->h_lvl = nullptr; c->h_resp = nullptr; c->h_sel = nullptr; c->h_nsel = nullptr;
    c->cfg = *cfg; c->L = L; c->base = (uint8_t*)workspace_dev;
    sgbm_carve(L.sg, c->base, &c->sg0);
    orb_carve(L.orb, c->base + L.sgbm_bytes, &c->orb0);
    uint8_t* p = c->base + L.frame_bytes * cfg->max_batch;
    c->tab_dev = (int32_t*)p; p += L.tab_bytes;
    c->knn_scratch = (uint32_t*)p; p += L.knn_bytes;
    c->nsel_dev = (int32_t*)p; p += L.nsel_bytes;
    c->filter_scratch = p; p += L.filter_bytes;
    c->pnp_scratch = p;
    std::vector<int32_t> tab(L.tab_total);
    int tot;
    orb_make_resize_tables(L.orb, tab.data(), c->L.tab_off, &tot);
    bool ok = cudaMemcpy(c->tab_dev, tab.data(), (size_t)tot * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    const int nb = cfg->max_batch;
    ok = ok && cudaMallocHost((void**)&c->h_lvl, (size_t)nb * 64 * 4) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&c->h_resp, (size_t)nb * L.orb.cand_cap * 8) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&c->h_sel, (size_t)nb * L.orb.kp_cap * 4) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&c->h_nsel, (size_t)nb * 4) == cudaSuccess;
    if (!ok) {
        set_error("ovo_create: CUDA failure while staging tables (%s)", cudaGetErrorString(cudaGetLastError()));
        ovo_destroy(c);  // frees whichever pinned buffers were allocated (the others are still null)
        return nullptr;
    }
    return c;
}

void ovo_destroy(ovo_ctx* c) {
    if (!c) return;
    if (c->worker.joinable()) {
        {
            std::unique_lock<std::mutex> lk(c->wmu);
            c->wcv.wait(lk, [&] { return !c->wbusy; });
            c->wquit = true;
        }
        c->wcv.notify_all();
        c->worker.join();
    }
#ifndef OVO_EMU
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
#endif
    if (c->h_lvl) cudaFreeHost(c->h_lvl);
    if (c->h_resp) cudaFreeHost(c->h_resp);
    if (c->h_sel) cudaFreeHost(c->h_sel);
    if (c->h_nsel) cudaFreeHost(c->h_nsel);
    delete c;
}

#define CHECK_CTX(c, nb)                                                            \
    if (!(c)) { set_error("null context"); return 1; }                              \
    if ((nb) < 1 || (nb) > (c)->cfg.max_batch) { set_error("batch %d outside [1, %d]", (nb), (c)->cfg.max_batch); return 1; }

int ovo_sgbm_compute(ovo_ctx* c, const uint8_t* left, const uint8_t* right, int pitch, size_t frame_stride, int nb, int16_t* disp,
                     void* stream) {
    CHECK_CTX(c, nb);
    if (pitch < c->L.sg.W) { set_error("pitch < width"); return 1; }
    // the per-frame stride of the SGBM workspace is the whole frame block
    return sgbm_launch(c->L.sg, &c->sg0, c->L.frame_bytes, nb, left, right, pitch, frame_stride, disp, (cudaStream_t)stream);
}

int ovo_disparity_post(ovo_ctx* c, const int16_t* disp, int nb, float* disp_f32, uint8_t* mask, void* stream) {
    CHECK_CTX(c, nb);
    return disp_post_launch(disp, c->L.sg.W, c->L.sg.H, c->L.x0, c->L.y0, c->L.cw, c->L.ch, c->cfg.min_valid_disparity,
                            c->cfg.max_valid_disparity, disp_f32, mask, nb, (cudaStream_t)stream);
}

int ovo_crop_left(ovo_ctx* c, const uint8_t* img, int pitch, size_t frame_stride, int nb, uint8_t* out, void* stream) {
    CHECK_CTX(c, nb);
    return crop_launch(img, pitch, frame_stride, c->L.x0, c->L.y0, c->L.cw, c->L.ch, nb, out, (cudaStream_t)stream);
}

int ovo_rectify(ovo_ctx* c, const uint8_t* img, int channels, int pitch, size_t frame_stride, int nb, const int16_t* map1,
                const uint16_t* map2, uint8_t* out, void* stream) {
    CHECK_CTX(c, nb);
    if (channels != 1 && channels != 3) { set_error("channels must be 1 (gray) or 3 (BGR)"); return 1; }
    if ((map1 == nullptr) != (map2 == nullptr)) { set_error("map1 and map2 must both be given or both be NULL"); return 1; }
    if (pitch < c->L.sg.W * channels) { set_error("pitch < width*channels"); return 1; }
    return rectify_launch(img, pitch, frame_stride, channels, c->L.sg.W, c->L.sg.H, nb, map1, map2, out, (cudaStream_t)stream);
}

int ovo_reproject_3d(ovo_ctx* c, const float* disp_f32, float* xyz, void* stream) {
    CHECK_CTX(c, 1);
    return reproject_launch(disp_f32, c->L.cw, c->L.cw, c->L.ch, c->L.x0, c->L.y0, c->cfg.Q, xyz, (cudaStream_t)stream);
}

int ovo_orb_detect_begin(ovo_ctx* c, const uint8_t* img, const uint8_t* mask, int nb, void* stream) {
    CHECK_CTX(c, nb);
    const Layout& L = c->L;
    const size_t fs = (size_t)L.cw * L.ch;
    return orb_phase1_launch(L.orb, &c->orb0, L.frame_bytes, c->tab_dev, L.tab_off, nb, img, L.cw, fs, mask, L.cw, fs, (cudaStream_t)stream);
}

int ovo_orb_detect_compute(ovo_ctx* c, const uint8_t* img, const uint8_t* mask, int nb, float* kp, uint8_t* desc, int* n_kp_host,
                           void* stream) {
    if (ovo_orb_detect_begin(c, img, mask, nb, stream)) return 1;
    return ovo_orb_detect_finish(c, nb, kp, desc, n_kp_host, stream);
}

int ovo_orb_detect_finish(ovo_ctx* c, int nb, float* kp, uint8_t* desc, int* n_kp_host, void* stream) {
    CHECK_CTX(c, nb);
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = c->L;
    const OrbDims& d = L.orb;
    for (int f = 0; f < nb; f++)
        OVO_CUDA(cudaMemcpyAsync(c->h_lvl + 64 * f, (uint8_t*)c->orb0.lvl_count + L.frame_bytes * f, 34 * 4, cudaMemcpyDeviceToHost, st));
    OVO_CUDA(cudaStreamSynchronize(st));
    c->d2h_bytes += 34 * 4 * (long long)nb;
    for (int f = 0; f < nb; f++) {
        const int total = c->h_lvl[64 * f + 16], surv = c->h_lvl[64 * f + 33];
        if (total > d.cand_cap) { set_error("ORB candidate overflow (%d > %d)", total, d.cand_cap); return 1; }
        // the 8-bit FAST score of every candidate (the introselect permutation depends on the whole array) and the Harris
        // response of the first pass's survivors only
        c->d2h_bytes += (long long)total + (long long)surv * 4;
        uint8_t* hs = (uint8_t*)(c->h_resp + (size_t)f * d.cand_cap * 2);
        float* hh = c->h_resp + (size_t)f * d.cand_cap * 2 + d.cand_cap;
        if (total > 0)
            OVO_CUDA(cudaMemcpyAsync(hs, (uint8_t*)c->orb0.cand_score + L.frame_bytes * f, (size_t)total, cudaMemcpyDeviceToHost, st));
        if (surv > 0)
            OVO_CUDA(cudaMemcpyAsync(hh, (uint8_t*)c->orb0.harris_dense + L.frame_bytes * f, (size_t)surv * 4, cudaMemcpyDeviceToHost, st));
    }
    OVO_CUDA(cudaStreamSynchronize(st));
    // retainBest x2 per level on the host (DESIGN.md "retainBest"); frames are independent -> one thread each
    auto select_one = [&](int f) {
        c->h_nsel[f] = orb_host_select(d, c->h_lvl + 64 * f, (const uint8_t*)(c->h_resp + (size_t)f * d.cand_cap * 2),
                                       c->h_resp + (size_t)f * d.cand_cap * 2 + d.cand_cap, c->h_sel + (size_t)f * d.kp_cap);
    };
    static const int max_threads = [] {
        const char* e = getenv("OVO_SELECT_THREADS");   // host threads used for the per-frame retainBest emulation (default 4)
        const int v = e ? atoi(e) : 4;
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    if (nb == 1 || max_threads == 1) {
        for (int f = 0; f < nb; f++) select_one(f);
    } else {
        std::vector<std::thread> th;
        const int nth = nb < max_threads ? nb : max_threads;
        for (int t = 0; t < nth; t++)
            th.emplace_back([&, t] { for (int f = t; f < nb; f += nth) select_one(f); });
        for (auto& x : th) x.join();
    }
    int max_sel = 0;
    for (int f = 0; f < nb; f++) {
        if (c->h_nsel[f] == -2) { set_error("ORB: the device's survivor set disagrees with the host selection (internal error)"); return 1; }
        if (c->h_nsel[f] < 0) { set_error("ORB keypoint capacity exceeded (ties at the retainBest boundary)"); return 1; }
        n_kp_host[f] = c->h_nsel[f];
        max_sel = c->h_nsel[f] > max_sel ? c->h_nsel[f] : max_sel;
        c->h2d_bytes += (long long)c->h_nsel[f] * 4 + 4;
        if (c->h_nsel[f] > 0)
            OVO_CUDA(cudaMemcpyAsync((uint8_t*)c->orb0.sel + L.frame_bytes * f, c->h_sel + (size_t)f * d.kp_cap, (size_t)c->h_nsel[f] * 4,
                                     cudaMemcpyHostToDevice, st));
    }
    OVO_CUDA(cudaMemcpyAsync(c->nsel_dev, c->h_nsel, (size_t)nb * 4, cudaMemcpyHostToDevice, st));
    return orb_phase2_launch(d, &c->orb0, L.frame_bytes, nb, max_sel, c->nsel_dev, kp, desc, st);
}

static int extract_begin_launches(ovo_ctx* c, const uint8_t* left, const uint8_t* right, int pitch, size_t frame_stride, int nb,
                                  int16_t* disp16, float* disp_f32, uint8_t* mask, uint8_t* img_crop, void* stream) {
    if (ovo_sgbm_compute(c, left, right, pitch, frame_stride, nb, disp16, stream)) return 1;
    if (ovo_disparity_post(c, disp16, nb, disp_f32, mask, stream)) return 1;
    if (ovo_crop_left(c, left, pitch, frame_stride, nb, img_crop, stream)) return 1;
    return ovo_orb_detect_begin(c, img_crop, mask, nb, stream);
}

int ovo_extract_begin(ovo_ctx* c, const uint8_t* left, const uint8_t* right, int pitch, size_t frame_stride, int nb, int16_t* disp16,
                      float* disp_f32, uint8_t* mask, uint8_t* img_crop, void* stream) {
    CHECK_CTX(c, nb);
#ifndef OVO_EMU
    static const bool use_graph = [] { const char* e = getenv("OVO_GRAPH"); return !(e && e[0] == '0'); }();
    if (use_graph && !g_prof_on && stream != nullptr) {
        cudaStream_t st = (cudaStream_t)stream;
        for (auto& g : c->graphs)
            if (g.left == left && g.right == right && g.disp16 == disp16 && g.disp_f32 == disp_f32 && g.mask == mask && g.img == img_crop &&
                g.pitch == pitch && g.nb == nb && g.frame_stride == frame_stride) {
                OVO_CUDA(cudaGraphLaunch(g.exec, st));
                g_launches.fetch_add(g.launches, std::memory_order_relaxed);
                return 0;
            }
        if (c->graphs.size() < 16) {
            // first call with this argument set: record the launches into a graph (thread-local capture: other host threads
            // driving other contexts are not affected), then replay it
            const long long before = g_launches.load();
            OVO_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            const int rc = extract_begin_launches(c, left, right, pitch, frame_stride, nb, disp16, disp_f32, mask, img_crop, stream);
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
            if (e != cudaSuccess || !graph) { set_error("ovo_extract_begin: stream capture failed (%s)", cudaGetErrorString(e)); return 1; }
            ovo_ctx::ExtractGraph g{left, right, disp16, disp_f32, mask, img_crop, pitch, nb, frame_stride, nullptr, g_launches.load() - before};
            const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) { set_error("ovo_extract_begin: graph instantiation failed (%s)", cudaGetErrorString(ei)); return 1; }
            c->graphs.push_back(g);
            OVO_CUDA(cudaGraphLaunch(g.exec, st));
            return 0;
        }
    }
#endif
    return extract_begin_launches(c, left, right, pitch, frame_stride, nb, disp16, disp_f32, mask, img_crop, stream);
}

int ovo_extract_finish(ovo_ctx* c, int nb, float* kp, uint8_t* desc, int* n_kp_host, void* stream) {
    return ovo_orb_detect_finish(c, nb, kp, desc, n_kp_host, stream);
}

int ovo_orb_detect_finish_async(ovo_ctx* c, int nb, float* kp, uint8_t* desc, int* n_kp_host, void* stream) {
    CHECK_CTX(c, nb);
#ifdef OVO_EMU
    c->worker_rc = ovo_orb_detect_finish(c, nb, kp, desc, n_kp_host, stream);
    if (c->worker_rc) snprintf(c->worker_err, sizeof(c->worker_err), "%s", ovo_last_error());
#else
    std::unique_lock<std::mutex> lk(c->wmu);
    if (c->wbusy) { set_error("ovo_orb_detect_finish_async: the previous call has not been waited for"); return 1; }
    if (!c->worker.joinable()) {
        c->worker = std::thread([c] {
            cudaSetDevice(c->device);  // a new host thread starts on device 0
            std::unique_lock<std::mutex> wl(c->wmu);
            for (;;) {
                c->wcv.wait(wl, [&] { return c->wquit || (c->wbusy && c->wjob); });
                if (c->wquit) return;
                std::function<int()> job = std::move(c->wjob);
                c->wjob = nullptr;
                wl.unlock();
                const int rc = job();
                wl.lock();
                c->worker_rc = rc;
                if (rc) snprintf(c->worker_err, sizeof(c->worker_err), "%s", ovo_last_error());  // the error string is thread-local
                c->wbusy = false;
                c->wcv.notify_all();
            }
        });
    }
    c->worker_rc = 0;
    c->wjob = [=] { return ovo_orb_detect_finish(c, nb, kp, desc, n_kp_host, stream); };
    c->wbusy = true;
    lk.unlock();
    c->wcv.notify_all();
#endif
    return 0;
}

int ovo_orb_detect_wait(ovo_ctx* c) {
    if (!c) { set_error("null context"); return 1; }
#ifndef OVO_EMU
    {
        std::unique_lock<std::mutex> lk(c->wmu);
        c->wcv.wait(lk, [&] { return !c->wbusy; });
    }
#endif
    if (c->worker_rc) { set_error("%s", c->worker_err); return 1; }
    return 0;
}

int ovo_knn2_hamming(ovo_ctx* c, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* nn, void* stream) {
    CHECK_CTX(c, 1);
    if (nq > c->L.orb.kp_cap || nt > c->L.orb.kp_cap) { set_error("descriptor count exceeds keypoint capacity"); return 1; }
    return knn2_launch(q, nq, t, nt, nn, c->knn_scratch, (cudaStream_t)stream);
}

int ovo_match_points(ovo_ctx* c, const int32_t* nn, int nq, double thr, const float* kp1, const float* kp2, const float* disp1,
                     const float* disp2, int32_t* matches, float* pts1, float* pts2, int32_t* counts, const int32_t* nn_rev, void* stream) {
    CHECK_CTX(c, 1);
    GatherParams p;
    memcpy(p.Q, c->cfg.Q, sizeof(p.Q));
    p.thr = thr; p.roi_x0 = c->L.x0; p.roi_y0 = c->L.y0; p.cw = c->L.cw; p.ch = c->L.ch; p.disp_pitch = c->L.cw;
    return match_gather_launch(p, nn, nq, kp1, kp2, disp1, disp2, matches, pts1, pts2, counts, nn_rev, (cudaStream_t)stream);
}

int ovo_pair_batch(ovo_ctx* c, int n, const ovo_pair_item* items, double thr, void* stream) {
    CHECK_CTX(c, 1);
    if (sizeof(ovo_pair_item) != (size_t)pair_item_size()) { set_error("ovo_pair_item layout mismatch"); return 1; }
    if (n < 0 || n > c->cfg.max_batch) { set_error("pair batch %d exceeds max_batch", n); return 1; }
    GatherParams p;
    memcpy(p.Q, c->cfg.Q, sizeof(p.Q));
    p.thr = thr; p.roi_x0 = c->L.x0; p.roi_y0 = c->L.y0; p.cw = c->L.cw; p.ch = c->L.ch; p.disp_pitch = c->L.cw;
    const size_t per = c->L.knn_bytes / c->cfg.max_batch;
    for (int i0 = 0; i0 < n; i0 += pair_max_batch()) {
        const int m = n - i0 < pair_max_batch() ? n - i0 : pair_max_batch();
        std::vector<ovo_pair_item> tmp(items + i0, items + i0 + m);
        for (int i = 0; i < m; i++) {
            if (tmp[i].nq > c->L.orb.kp_cap || tmp[i].nt > c->L.orb.kp_cap) { set_error("descriptor count exceeds keypoint capacity"); return 1; }
            tmp[i].scratch = (uint32_t*)((uint8_t*)c->knn_scratch + per * (size_t)(i0 + i));
        }
        if (pair_batch_launch(p, m, tmp.data(), c->L.orb.kp_cap, (cudaStream_t)stream)) return 1;
    }
    return 0;
}

int ovo_pnp_ransac(ovo_ctx* c, const float* pts1, const int32_t* matches, const float* kp2, const int32_t* count, int cap, int iters,
                   double reproj_px, unsigned long long seed, double* out, void* stream) {
    CHECK_CTX(c, 1);
    if (cap > c->L.orb.kp_cap) { set_error("point capacity exceeds keypoint capacity"); return 1; }
    return pnp_ransac_launch(pts1, matches, kp2, count, cap, c->cfg.Q, c->L.x0, c->L.y0, iters, kPnpMaxIters, reproj_px, (uint64_t)seed,
                             c->pnp_scratch, out, (cudaStream_t)stream);
}

int ovo_rigid_body_filter(ovo_ctx* c, float* pts_prev, float* pts_cur, int32_t* count, int cap, double thr, void* stream) {
    CHECK_CTX(c, 1);
    if (cap > c->L.orb.kp_cap) { set_error("point capacity exceeds keypoint capacity"); return 1; }
    return rigid_filter_launch(pts_prev, pts_cur, count, cap, (float)thr, c->filter_scratch, (cudaStream_t)stream);
}

int ovo_outlier_filter(ovo_ctx* c, float* pts_prev, float* pts_cur, int32_t* count, int cap, const double* T, double thr, void* stream) {
    CHECK_CTX(c, 1);
    if (cap > c->L.orb.kp_cap) { set_error("point capacity exceeds keypoint capacity"); return 1; }
    return outlier_filter_launch(pts_prev, pts_cur, count, cap, T, thr, c->filter_scratch, (cudaStream_t)stream);
}

int ovo_rigid_transform(ovo_ctx* c, const float* pts1, const float* pts2, const int32_t* count, int cap, double* out, void* stream) {
    CHECK_CTX(c, 1);
    return umeyama_launch(pts1, pts2, count, cap, out, (cudaStream_t)stream);
}

}  // extern "C"
