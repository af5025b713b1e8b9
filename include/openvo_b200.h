/* openvo_b200 — C ABI of the B200-native openVO hot path.
 *
 * The reference (KevinSpevak/openVO) is pure Python and has no FFI of its own: its "operator interface" for the hot
 * path is the set of cv2 calls made by StereoCamera.compute_3d and StereoOdometer.update.  Each entry point below
 * replaces one of those call sites (cited as ref: file:line under /root/reference).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named *_dev is a CUDA device pointer owned by the caller (the Python host passes torch tensors'
 *     data_ptr()); *_host pointers are host memory.  The library allocates no device memory: the caller provides one
 *     workspace of ovo_workspace_bytes() bytes at ovo_create().
 *   - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous on that stream unless stated otherwise.
 *   - return value 0 = success; otherwise ovo_last_error() describes the failure (thread-local string).
 *   - `nb` = number of frames processed by the call (batch); per-frame arrays are contiguous with the stated stride.
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef OPENVO_B200_H
#define OPENVO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVO_ABI_VERSION 1
#define OVO_KP_FIELDS 6 /* pt.x, pt.y, size, angle, response, octave  (cv2.KeyPoint fields, class_id is always -1) */

typedef struct ovo_ctx ovo_ctx;

/* The ten positional arguments of cv2.StereoSGBM_create as the reference passes them
 * (ref: src/openVO/stereo_camera.py:23-27; mode is left at its default MODE_SGBM). */
typedef struct ovo_sgbm_params {
    int minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap, uniquenessRatio,
        speckleWindowSize, speckleRange;
} ovo_sgbm_params;

typedef struct ovo_config {
    int width, height;       /* full (rectified) image size fed to StereoSGBM */
    ovo_sgbm_params sgbm;
    int roi[4];              /* valid_region_left exactly as cv2.stereoRectify returned it (x, y, w, h); the library
                                applies the reference's slice img[roi[1]:roi[3], roi[0]:roi[2]]
                                (ref: src/openVO/stereo_camera.py:35-37) */
    double Q[16];            /* 4x4 reprojection matrix, row-major (ref: src/openVO/stereo_camera.py:17,52) */
    int nfeatures;           /* cv2.ORB_create(nfeatures=...) (ref: src/openVO/stereo_odometer.py:22) */
    int max_batch;           /* frames a single call may process */
    float min_valid_disparity, max_valid_disparity; /* StereoOdometer.MIN/MAX_VALID_DISPARITY (ref: stereo_odometer.py:6-7) */
    int sgbm_mode;           /* 0 = cv2.StereoSGBM MODE_SGBM, 5 directions — what the reference uses (its `mode=1` is commented out,
                                ref: src/openVO/stereo_camera.py:27); 1 = MODE_HH, 8 directions: opt-in extension,
                                changes ~10 % of the output pixels, so it is never the default */
} ovo_config;

const char* ovo_last_error(void);
int ovo_abi_version(void);

/* Cropped frame size (the reference's slice semantics applied to roi). */
int ovo_cropped_size(const ovo_config* cfg, int* cw, int* ch);
/* Keypoint capacity per frame (>= nfeatures: ties at the retainBest boundary are all kept). */
int ovo_kp_capacity(const ovo_config* cfg);

size_t ovo_workspace_bytes(const ovo_config* cfg);
ovo_ctx* ovo_create(const ovo_config* cfg, void* workspace_dev, size_t workspace_bytes);
void ovo_destroy(ovo_ctx* ctx);

/* Seam S-A — replaces stereoSGBM.compute(L, R) (ref: src/openVO/stereo_camera.py:51).
 * left/right: u8 [nb][height][pitch]; disp: i16 [nb][height][width], fixed point 1/16 px, invalid = -16. */
int ovo_sgbm_compute(ovo_ctx* ctx, const uint8_t* left_dev, const uint8_t* right_dev, int pitch, size_t frame_stride,
                     int nb, int16_t* disp_dev, void* stream);

/* §8(f) n1 + n2 — replaces cv2.cvtColor(img, COLOR_BGR2GRAY) and cv2.remap(img, map_1, map_2, INTER_LINEAR)
 * (ref: src/openVO/stereo_camera.py:44-50, 29-33).  img: u8 [nb][height][pitch], `channels` = 1 (gray) or 3 (BGR, converted per
 * tap with OpenCV's 15-bit coefficients).  map1: i16 [height][width][2] (CV_16SC2), map2: u16 [height][width] as produced by
 * cv2.initUndistortRectifyMap(..., CV_16SC2); pass both NULL for colour conversion only.  out: u8 [nb][height][width]. */
int ovo_rectify(ovo_ctx* ctx, const uint8_t* img_dev, int channels, int pitch, size_t frame_stride, int nb, const int16_t* map1_dev,
                const uint16_t* map2_dev, uint8_t* out_dev, void* stream);

/* a2 + a4 + a5 — replaces `.astype(np.float32)/16`, the crop and StereoOdometer.feature_mask
 * (ref: src/openVO/stereo_camera.py:51,54; src/openVO/stereo_odometer.py:38-41).
 * disp: i16 [nb][height][width] -> disp_f32: f32 [nb][ch][cw], mask: u8 [nb][ch][cw] (0 / 255). */
int ovo_disparity_post(ovo_ctx* ctx, const int16_t* disp_dev, int nb, float* disp_f32_dev, uint8_t* mask_dev, void* stream);

/* Crop of the left image (ref: src/openVO/stereo_camera.py:55): u8 [nb][height][pitch] -> u8 [nb][ch][cw]. */
int ovo_crop_left(ovo_ctx* ctx, const uint8_t* img_dev, int pitch, size_t frame_stride, int nb, uint8_t* out_dev, void* stream);

/* Seam S-B — replaces cv2.reprojectImageTo3D(disparity, Q) + crop (ref: src/openVO/stereo_camera.py:52-53).
 * Only needed when the caller reads StereoOdometer.current_3d: the pose path reprojects on the fly.
 * disp_f32: f32 [ch][cw] (cropped) -> xyz: f32 [ch][cw][3]. */
int ovo_reproject_3d(ovo_ctx* ctx, const float* disp_f32_dev, float* xyz_dev, void* stream);

/* Seam S-D — replaces orb.detectAndCompute(img, mask) (ref: src/openVO/stereo_odometer.py:117).
 * img, mask: u8 [nb][ch][cw] (mask may be NULL).  kp: f32 [nb][kp_capacity][OVO_KP_FIELDS] in cv2's order;
 * desc: u8 [nb][kp_capacity][32]; n_kp_host: int [nb].
 * SYNCHRONOUS on `stream`: KeyPointsFilter::retainBest's ordering is libstdc++'s introselect permutation, which is
 * reproduced on the host between the two device phases (see DESIGN.md "retainBest"). */
int ovo_orb_detect_compute(ovo_ctx* ctx, const uint8_t* img_dev, const uint8_t* mask_dev, int nb, float* kp_dev,
                           uint8_t* desc_dev, int* n_kp_host, void* stream);
/* The same call in two halves, for drivers that keep several batches in flight from one host thread: `begin` only queues
 * the detection phase on `stream` (asynchronous); `finish` waits for it, runs the host-side retainBest and queues the
 * descriptor phase.  begin + finish on the same ctx / stream == ovo_orb_detect_compute. */
int ovo_orb_detect_begin(ovo_ctx* ctx, const uint8_t* img_dev, const uint8_t* mask_dev, int nb, void* stream);
int ovo_orb_detect_finish(ovo_ctx* ctx, int nb, float* kp_dev, uint8_t* desc_dev, int* n_kp_host, void* stream);
/* `finish` on a worker thread of the library: returns at once; the worker waits for the detection phase on `stream`, runs the
 * host-side retainBest the moment it lands and queues the descriptor phase, while the calling thread drives other contexts.
 * ovo_orb_detect_wait joins it and reports its status; the caller must not use `stream` (nor n_kp_host) in between. */
int ovo_orb_detect_finish_async(ovo_ctx* ctx, int nb, float* kp_dev, uint8_t* desc_dev, int* n_kp_host, void* stream);
int ovo_orb_detect_wait(ovo_ctx* ctx);

/* Whole-frame entry — the per-frame part of StereoOdometer.update that does not depend on the previous frame
 * (ref: src/openVO/stereo_odometer.py:116-117 = stereo.compute_3d + orb.detectAndCompute with feature_mask), in two halves like
 * the ORB seam: `begin` queues ovo_sgbm_compute + ovo_disparity_post + ovo_crop_left + ovo_orb_detect_begin on `stream` and
 * returns at once; `finish` == ovo_orb_detect_finish.  left/right: u8 [nb][height][pitch] rectified gray; disp16: i16
 * [nb][height][width] scratch; disp_f32 / mask / img_crop: [nb][ch][cw].  On a non-default stream the ~70 launches of `begin` are
 * recorded once per distinct argument set into a CUDA graph and replayed as one launch afterwards (OVO_GRAPH=0 disables this), so
 * callers should pass persistent buffers and copy the products they keep. */
int ovo_extract_begin(ovo_ctx* ctx, const uint8_t* left_dev, const uint8_t* right_dev, int pitch, size_t frame_stride, int nb,
                      int16_t* disp16_dev, float* disp_f32_dev, uint8_t* mask_dev, uint8_t* img_crop_dev, void* stream);
int ovo_extract_finish(ovo_ctx* ctx, int nb, float* kp_dev, uint8_t* desc_dev, int* n_kp_host, void* stream);

/* Seam S-E — replaces matcher.knnMatch(desc1, desc2, k=2) (ref: src/openVO/stereo_odometer.py:163).
 * nn: i32 [nq][4] = (trainIdx0, dist0, trainIdx1, dist1); ties resolve to the lowest train index. */
int ovo_knn2_hamming(ovo_ctx* ctx, const uint8_t* q_desc_dev, int nq, const uint8_t* t_desc_dev, int nt, int32_t* nn_dev,
                     void* stream);

/* a8 + Seam S-F — replaces the ratio-test list-comprehension and the bilinear_interpolate_pixels loop of
 * StereoOdometer.point_clouds (ref: src/openVO/stereo_odometer.py:164,170-175,50-79), with the 3-D image evaluated
 * lazily from the cropped disparity (reprojectImageTo3D fused in).
 * matches: i32 [nq][3] = (queryIdx, trainIdx, distance) in query order; pts1/pts2: f32 [nq][3];
 * counts: i32 [2] = (number of matches, number of lookups whose four taps were all unusable).
 * nn_rev_dev: NULL for the reference's behaviour; otherwise the 2-NN table of the TRAIN set against the QUERY set (ovo_knn2_hamming
 * with the arguments swapped) and a match (q, t) is kept only if t's nearest neighbour is q — the left-right cross-check the reference
 * leaves as "TODO crosscheck" (ref: src/openVO/stereo_odometer.py:21), an opt-in extension. */
int ovo_match_points(ovo_ctx* ctx, const int32_t* nn_dev, int nq, double match_threshold, const float* kp1_dev,
                     const float* kp2_dev, const float* disp1_f32_dev, const float* disp2_f32_dev, int32_t* matches_dev,
                     float* pts1_dev, float* pts2_dev, int32_t* counts_dev, const int32_t* nn_rev_dev, void* stream);

/* Seam S-G — replaces cv2.estimateAffine3D(src, dst, force_rotation=True) and the quantities the gates need
 * (ref: src/openVO/stereo_odometer.py:204-221).  count_dev: i32 device scalar (number of points, e.g. counts[0]).
 * out: f64 [16] = rows of [R|t] (12), scale, rotation angle (= |Rodrigues(R)|), |t|, n used. */
int ovo_rigid_transform(ovo_ctx* ctx, const float* pts1_dev, const float* pts2_dev, const int32_t* count_dev, int cap,
                        double* out_dev, void* stream);

/* §8(f) n3 — the optional pre-filters of StereoOdometer.point_cloud_transform (off by default in the reference).
 * ovo_rigid_body_filter replaces rigid_body_filter + the boolean-mask indexing (ref: src/openVO/stereo_odometer.py:82-105,178-181):
 * greedy clique on the graph | ||p_i-p_j|| - ||q_i-q_j|| | < thr (float32, as numpy evaluates it), points compacted in place and
 * *count_dev updated.  ovo_outlier_filter replaces the single-pass outlier removal (ref: :189-197): relative residual under T
 * (f64 [12], rows of [R|t], e.g. the output of ovo_rigid_transform), keep residual < thr + median.  cap <= 16383. */
int ovo_rigid_body_filter(ovo_ctx* ctx, float* pts_prev_dev, float* pts_cur_dev, int32_t* count_dev, int cap, double thr, void* stream);
int ovo_outlier_filter(ovo_ctx* ctx, float* pts_prev_dev, float* pts_cur_dev, int32_t* count_dev, int cap, const double* T_dev,
                       double thr, void* stream);

/* §8(f) n4, OPT-IN, not the reference's behaviour (its pose is Umeyama, ovo_rigid_transform): batched P3P RANSAC over a fixed,
 * seed-determined hypothesis schedule + Levenberg-Marquardt refinement of the reprojection error on the inliers.  3-D points of
 * frame A (pts1, as produced by ovo_match_points) against the pixel positions of the matched keypoints of frame B (matches[i][1]
 * indexes kp2).  Intrinsics come from Q (f = Q[2][3], c = -Q[0..1][3]).  iters <= 4096.  Specification: oracle/pnp_restate.py.
 * out: f64 [16] = rows of [R|t] (12), inlier count, rotation angle, |t|, index of the winning hypothesis. */
int ovo_pnp_ransac(ovo_ctx* ctx, const float* pts1_dev, const int32_t* matches_dev, const float* kp2_dev, const int32_t* count_dev, int cap,
                   int iters, double reproj_px, unsigned long long seed, double* out_dev, void* stream);

/* Batched form of S-E + a8/S-F + S-G for n independent frame pairs (n <= max_batch): 2-NN, ratio test + fused 3-D lookup and
 * rigid alignment of every pair in four launches.  `items` is a HOST array; all pointers inside are device pointers.
 * out: f64 [18] per pair = the 16 values of ovo_rigid_transform followed by the two int32 counts of ovo_match_points packed in
 * slot 16.  `scratch` is filled in by the library. */
typedef struct ovo_pair_item {
    const uint8_t *q_desc, *t_desc;
    int nq, nt;
    const float *kp1, *kp2, *disp1_f32, *disp2_f32;
    int32_t *nn, *matches;
    float *pts1, *pts2;
    double* out;
    uint32_t* scratch;
    int32_t* nn_rev; /* NULL, or a [nt][4] buffer: enables the cross-check for this pair */
} ovo_pair_item;
int ovo_pair_batch(ovo_ctx* ctx, int n, const ovo_pair_item* items_host, double match_threshold, void* stream);

/* Instrumentation for bench.py: number of kernels launched by this library so far; optional per-kernel CUDA-event timing
 * (events recorded on the launching stream around every kernel while enabled).  ovo_profile_read synchronises the device
 * and returns, per kernel name ('\n'-separated in `names`), the summed milliseconds and the launch count since the last
 * read. */
long long ovo_launch_count(void);
/* bytes this context has moved between host and device itself (the keypoint-selection staging of ovo_orb_detect_compute) */
void ovo_transfer_bytes(ovo_ctx* ctx, long long* h2d, long long* d2h);
void ovo_profile_enable(int on);
int ovo_profile_read(char* names, int names_len, float* total_ms, int* counts, int max_entries);

#ifdef __cplusplus
}
#endif
#endif
