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

// lvl_count: [64] as written by k_orb_scan / k_orb_survivors (per-level candidate counts and offsets, total; per-level
// survivor counts and offsets, total); scores: FAST score of every candidate in raster order per level; harris_dense: Harris
// response of the survivors of the first pass (the candidates whose score reaches the level's boundary score), in candidate
// order — the device finds that SET with a histogram, so only the survivors' responses are computed and shipped; the ORDER
// of the survivors is produced here.  out_sel receives candidate ids in final keypoint order.
int orb_host_select(const OrbDims& d, const int32_t* lvl_count, const uint8_t* scores, const float* harris_dense, int32_t* out_sel) {
    int n_out = 0;
    std::vector<Rec> k;
    std::vector<int32_t> dense;
    for (int l = 0; l < ORB_NLEVELS; l++) {
        const int n = lvl_count[l], base = lvl_count[ORB_NLEVELS + l];
        k.resize(n);
        for (int i = 0; i < n; i++) k[i] = {(float)scores[base + i], base + i};
        retain_best(k, 2 * d.lv[l].nfeat);
        // survivors = every candidate with score >= the boundary score (retainBest keeps all ties); the device numbered them in
        // candidate order
        float amb = 0.f;
        if ((int)k.size() < n) {
            amb = 256.f;
            for (auto& r : k) amb = std::min(amb, r.resp);
        }
        dense.assign(n, -1);
        int j = lvl_count[25 + l];
        for (int i = 0; i < n; i++)
            if ((float)scores[base + i] >= amb) dense[i] = j++;
        if (j - lvl_count[25 + l] != lvl_count[17 + l] || lvl_count[17 + l] != (int)k.size()) return -2;
        for (auto& r : k) r.resp = harris_dense[dense[r.id - base]];
        retain_best(k, d.lv[l].nfeat);
        if (n_out + (int)k.size() > d.kp_cap) return -1;
        for (auto& r : k) out_sel[n_out++] = r.id;
    }
    return n_out;
}

}  // namespace ovo
