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
    // scratch kept per host thread: no allocation per frame
    static thread_local std::vector<uint32_t> a;   // first pass: (FAST score << 24 | index inside the level) — 4-byte records, integer compares
    static thread_local std::vector<Rec> k;        // second pass: (Harris response, candidate id) of the survivors
    static thread_local std::vector<int32_t> dense;
    // the comparators look at the score only, exactly like the reference's `a.response > b.response` on KeyPoint records: the same
    // comparison outcomes give the same introselect permutation
    auto greater = [](uint32_t x, uint32_t y) { return (x >> 24) > (y >> 24); };
    for (int l = 0; l < ORB_NLEVELS; l++) {
        const int n = lvl_count[l], base = lvl_count[ORB_NLEVELS + l], keep = 2 * d.lv[l].nfeat;
        a.resize(n);
        for (int i = 0; i < n; i++) a[i] = ((uint32_t)scores[base + i] << 24) | (uint32_t)i;
        uint32_t amb = 0;  // boundary score: every candidate with score >= amb survives (retainBest keeps all ties)
        if (n > keep) {
            if (keep == 0) {
                a.clear();
                amb = 256;
            } else {
                std::nth_element(a.begin(), a.begin() + keep - 1, a.end(), greater);
                amb = a[keep - 1] >> 24;
                auto e = std::partition(a.begin() + keep, a.end(), [amb](uint32_t r) { return (r >> 24) >= amb; });
                a.resize(e - a.begin());
            }
        }
        // the device numbered the survivors in candidate order
        dense.resize(n);
        int j = lvl_count[25 + l];
        for (int i = 0; i < n; i++) {  // branch-free: the entry of a non-survivor is never read
            dense[i] = j;
            j += (uint32_t)scores[base + i] >= amb ? 1 : 0;
        }
        if (j - lvl_count[25 + l] != lvl_count[17 + l] || lvl_count[17 + l] != (int)a.size()) return -2;
        k.resize(a.size());
        for (size_t t = 0; t < a.size(); t++) {
            const int i = (int)(a[t] & 0xFFFFFFu);
            k[t] = {harris_dense[dense[i]], base + i};
        }
        retain_best(k, d.lv[l].nfeat);
        if (n_out + (int)k.size() > d.kp_cap) return -1;
        for (auto& r : k) out_sel[n_out++] = r.id;
    }
    return n_out;
}

}  // namespace ovo
