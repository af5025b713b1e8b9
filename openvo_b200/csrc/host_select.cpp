// Host half of ORB keypoint selection: the two cv::KeyPointsFilter::retainBest passes.
//
// OpenCV keeps the best k keypoints with std::nth_element + std::partition, so the ORDER of the survivors is the
// permutation libstdc++'s introselect happens to produce.  That order is observable (keypoint/descriptor order, and
// through it knnMatch's lowest-train-index tie rule), so it is part of the parity contract
// (ref call site: src/openVO/stereo_odometer.py:117).  A data-parallel selection yields the same SET but not the same
// ORDER; the permutation is therefore reproduced here with the same libstdc++ calls on (response, candidate-id) pairs
// — a few tens of thousands of 8-byte records per frame — between the two device phases.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace ovo {

namespace {
struct Rec {
    float resp;
    int32_t id;
};
struct RespGreater {
    bool operator()(const Rec& a, const Rec& b) const { return a.resp > b.resp; }
};
void retain_best(std::vector<Rec>& k, int n) {
    if (n >= 0 && k.size() > (size_t)n) {
        if (n == 0) {
            k.clear();
            return;
        }
        std::nth_element(k.begin(), k.begin() + n - 1, k.end(), RespGreater());
        const float amb = k[n - 1].resp;
        auto e = std::partition(k.begin() + n, k.end(), [amb](const Rec& r) { return r.resp >= amb; });
        k.resize(e - k.begin());
    }
}
}  // namespace

// lvl_count: [17] as written by k_orb_scan (per-level counts, per-level offsets, total); cand_resp: [n][2] =
// (FAST score, Harris response) in raster order per level.  out_sel receives candidate ids in final keypoint order.
int orb_host_select(const OrbDims& d, const int32_t* lvl_count, const float* cand_resp, int32_t* out_sel) {
    int n_out = 0;
    std::vector<Rec> k;
    for (int l = 0; l < ORB_NLEVELS; l++) {
        const int n = lvl_count[l], base = lvl_count[ORB_NLEVELS + l];
        k.resize(n);
        for (int i = 0; i < n; i++) k[i] = {cand_resp[2 * (size_t)(base + i)], base + i};
        retain_best(k, 2 * d.lv[l].nfeat);
        for (auto& r : k) r.resp = cand_resp[2 * (size_t)r.id + 1];
        retain_best(k, d.lv[l].nfeat);
        if (n_out + (int)k.size() > d.kp_cap) return -1;
        for (auto& r : k) out_sel[n_out++] = r.id;
    }
    return n_out;
}

}  // namespace ovo
