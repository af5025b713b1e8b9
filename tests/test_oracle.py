"""CPU tier: the oracle (oracle/*.cpp + oracle/openvo_port.py) pinned against (a) the installed cv2 binary, (b) the golden
fixtures generated from the unmodified reference, (c) the reference itself when /root/reference is present."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, SGBM_CASES, block_mask, occluded_pair, sgbm_params
from openvo_b200 import synth
from oracle import openvo_port as O

cv2 = pytest.importorskip("cv2")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("W,H,D,kw", SGBM_CASES)
def test_sgbm_restated_equals_cv2(W, H, D, kw):
    L, R = occluded_pair(W, H)
    p = sgbm_params(D, **kw)
    ref = cv2.StereoSGBM_create(*[p[k] for k in O.SGBM_KEYS]).compute(L, R)
    assert np.array_equal(O.sgbm_compute(L, R, p), ref)


def test_sgbm_rejects_unpinned_domain():
    L, R = occluded_pair(200, 60)
    with pytest.raises(ValueError):
        O.sgbm_compute(L, R, sgbm_params(32, blockSize=15, P1=1800, P2=7200))


@pytest.mark.parametrize("W,H,n,usemask", [(640, 200, 500, False), (415, 333, 300, True), (1240, 375, 2000, True)])
def test_orb_restated_equals_cv2(W, H, n, usemask):
    L, _ = synth.kat_pair(W, H)
    mask = block_mask(H, W) if usemask else None
    kps, desc = cv2.ORB_create(nfeatures=n).detectAndCompute(L, mask)
    ref = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)
    kp, d = O.orb_detect_compute(L, mask, n)
    assert np.array_equal(kp, ref) and np.array_equal(d, desc)


def test_orb_no_keypoints_on_noise_free_blocks():
    # without the tie-breaking noise equal adjacent FAST scores annihilate under strict NMS at level 0
    rng = np.random.default_rng(0)
    t = rng.integers(0, 256, (50, 80), dtype=np.uint8)
    img = np.ascontiguousarray(np.kron(t, np.ones((4, 4), np.uint8)))
    kp, _ = O.orb_detect_compute(img, None, 500)
    kps, _ = cv2.ORB_create(nfeatures=500).detectAndCompute(img, None)
    assert len(kp) == len(kps) and (kp[:, 5] == 0).sum() == sum(1 for k in kps if k.octave == 0)


def test_knn_restated_equals_cv2_with_ties():
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8) & 0xF0
    t = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    t[128:] &= 0xF0
    t[3] = t[1]
    mm = cv2.BFMatcher.create(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    ref = np.array([[m[0].trainIdx, int(m[0].distance), m[1].trainIdx, int(m[1].distance)] for m in mm], np.int32)
    got = O.knn2_hamming(q, t)
    assert np.array_equal(got, ref) and (ref[:, 1] == ref[:, 3]).sum() > 10


def test_ratio_integer_equivalence():
    # d0 < 0.8*d1 in double  <=>  5*d0 < 4*d1 for all 0..256
    for d0 in range(257):
        for d1 in range(257):
            assert (float(d0) < 0.8 * float(d1)) == (5 * d0 < 4 * d1)


def test_reproject_restated_equals_cv2():
    rng = np.random.default_rng(2)
    disp = (rng.integers(-16, 1600, (40, 60)).astype(np.float32)) / 16
    disp[0, :5] = 0
    Q = np.array([[1, 0, 0, -29.5], [0, 1, 0, -19.5], [0, 0, 0, 300.0], [0, 0, 1.862, 0]]) + rng.normal(0, 1e-3, (4, 4))
    a, b = O.reproject_to_3d(disp, Q), cv2.reprojectImageTo3D(disp, Q)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_remap_and_gray_restated_equal_cv2():
    rng = np.random.default_rng(6)
    args = synth.camera_args_distorted(320, 200, 32)
    cam = O.StereoCameraPort(**args, backend="cv2")
    img = rng.integers(0, 256, (200, 320), dtype=np.uint8)
    for m1, m2 in ((cam.map_left_1, cam.map_left_2), (cam.map_right_1, cam.map_right_2)):
        assert np.array_equal(O.remap_linear(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    # maps that leave the image on every side (BORDER_CONSTANT taps)
    m1 = np.stack([rng.integers(-3, 323, (200, 320)), rng.integers(-3, 203, (200, 320))], -1).astype(np.int16)
    m2 = rng.integers(0, 1024, (200, 320)).astype(np.uint16)
    assert np.array_equal(O.remap_linear(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    bgr = rng.integers(0, 256, (200, 320, 3), dtype=np.uint8)
    assert np.array_equal(O.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))


def test_umeyama_restated_equals_cv2():
    rng = np.random.default_rng(4)
    src = rng.normal(0, 5, (200, 3)).astype(np.float32)
    ang = 0.05
    Rm = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    dst = (src @ Rm.T * 1.01 + [0.1, -0.02, 0.5] + rng.normal(0, 0.01, (200, 3))).astype(np.float32)
    T, s = cv2.estimateAffine3D(src, dst, force_rotation=True)
    To, so = O.umeyama(src, dst)
    assert np.abs(To - T).max() < 1e-12 and abs(so - s) < 1e-12
    # B2: t uses the similarity scale although R is unscaled
    assert np.abs(To[:, 3] - (dst.astype(np.float64).mean(0) - so * To[:, :3] @ src.astype(np.float64).mean(0))).max() < 1e-12
    assert abs(O.rotation_angle(T[:, :3]) - np.linalg.norm(cv2.Rodrigues(T[:, :3])[0])) < 1e-12


def test_kat_hashes_match_survey_appendix_b():
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["K"]
    L, R = synth.kat_pair(kat["W"], kat["H"])
    assert sha(L) == kat["left"] and sha(R) == kat["right"]
    kpL, dL = O.orb_detect_compute(L, None, kat["n"])
    kpR, dR = O.orb_detect_compute(R, None, kat["n"])
    assert sha(kpL) == kat["kpL"] and sha(dL) == kat["descL"] and sha(kpR) == kat["kpR"] and sha(dR) == kat["descR"]
    nn = O.knn2_hamming(dL, dR)
    assert sha(nn) == kat["knn"]
    assert int((5 * nn[:, 1] < 4 * nn[:, 3]).sum()) == kat["ratio_pass"]
    assert sha(O.sgbm_compute(L, R, sgbm_params(kat["D"]))) == kat["sgbm"]


def _replay(g, backend, **kw):
    W, H, D, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    args = (synth.camera_args_distorted if bool(g["distorted"]) else synth.camera_args)(W, H, D)
    cam = O.StereoCameraPort(**args, backend=backend)
    assert tuple(cam.valid_region_left) == tuple(int(v) for v in g["roi"])
    od = O.StereoOdometerPort(cam, nfeatures=n, preprocessed_frames=bool(g["preprocessed"]), **kw)
    for i in range(len(g["left"])):
        ok = od.update(g["left"][i], g["right"][i])
        assert ok == bool(g["ok_%d" % i]), i
        assert od.skip_cause == str(g["cause_%d" % i]), i
        assert od.skipped_frames == int(g["skipped_%d" % i]), i
        if od.cur is not None:
            assert np.array_equal(np.rint(od.cur[1] * 16).astype(np.int16), g["disp16_%d" % i]), i
            assert np.array_equal(od.cur[3], g["kp_%d" % i]) and np.array_equal(od.cur[4], g["desc_%d" % i]), i
            assert sha(od.cur[2]) == str(g["xyz_sha_%d" % i]), i
        assert np.abs(od.c_T_w - g["cTw_%d" % i]).max() < 1e-9, i
        assert np.abs(od.current_pose() - g["pose_%d" % i]).max() < 1e-9, i


@pytest.mark.parametrize("backend", ["cv2", "restated"])
@pytest.mark.parametrize("name,kw", [("seq_small", {}), ("seq_skip", {}), ("seq_rectify", {}),
                                     ("seq_filters", dict(rigidity_threshold=0.06, outlier_threshold=0.02))])
def test_port_reproduces_reference_fixtures(golden, backend, name, kw):
    _replay(golden(name), backend, **kw)


def test_seam_fixture(golden):
    g = golden("seams_small")
    p = sgbm_params(64)
    assert np.array_equal(O.sgbm_compute(g["left"], g["right"], p), g["sgbm"])
    kp1, d1 = O.orb_detect_compute(g["left"], None, 400)
    kp2, d2 = O.orb_detect_compute(g["right"], None, 400)
    assert np.array_equal(kp1, g["kp1"]) and np.array_equal(d1, g["desc1"])
    assert np.array_equal(kp2, g["kp2"]) and np.array_equal(d2, g["desc2"])
    assert np.array_equal(O.knn2_hamming(d1, d2), g["nn"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/openVO"), reason="reference checkout not present on this box")
def test_port_equals_live_reference():
    sys.path.insert(0, "/root/reference/src")
    import openVO
    W, H, D, n = 400, 140, 48, 300
    Ls, Rs, _ = synth.make_sequence(W, H, 3, seed=99)
    args = synth.camera_args(W, H, D)
    ref = openVO.StereoOdometer(openVO.StereoCamera(**args), nfeatures=n, preprocessed_frames=True)
    port = O.StereoOdometerPort(O.StereoCameraPort(**args, backend="restated"), nfeatures=n, preprocessed_frames=True)
    for i in range(3):
        assert ref.update(Ls[i], Rs[i]) == port.update(Ls[i], Rs[i])
        assert np.array_equal(ref.current_disparity, port.cur[1]) and np.array_equal(ref.current_desc, port.cur[4])
        assert np.array_equal(ref.current_3d.view(np.uint32), port.cur[2].view(np.uint32))
        assert np.abs(ref.c_T_w - port.c_T_w).max() < 1e-9


def test_sgbm_mode_hh_restated_equals_cv2():
    # extension row n4: 8-direction MODE_HH (not used by the reference) pinned against cv2's own implementation
    for W, H, D, kw in SGBM_CASES[:3]:
        L, R = occluded_pair(W, H)
        p = sgbm_params(D, **kw)
        ref = cv2.StereoSGBM_create(*[p[k] for k in O.SGBM_KEYS], mode=cv2.STEREO_SGBM_MODE_HH).compute(L, R)
        assert np.array_equal(O.sgbm_compute_mode(L, R, p, 1), ref)


def test_pnp_ransac_specification_vs_cv2_and_truth():
    # extension row n4 (parity unpinned: OpenCV's RNG schedule is not reproducible): the numpy specification recovers the pose
    # and coincides with cv2.solvePnPRansac whenever the inlier sets coincide
    from oracle import pnp_restate as P
    rng = np.random.default_rng(0)
    f, cx, cy = 718.856, 620.0, 187.5
    m = 600
    X = np.stack([rng.uniform(-8, 8, m), rng.uniform(-2, 2, m), rng.uniform(5, 25, m)], 1)
    R, t = P.rodrigues(np.array([0.01, -0.03, 0.005])), np.array([0.03, -0.01, -0.2])
    uv, _ = P.project(R, t, X, f, cx, cy)
    # exact P3P on three clean correspondences
    bear = np.stack([(uv[:3, 0] - cx) / f, (uv[:3, 1] - cy) / f, np.ones(3)], 1)
    bear /= np.linalg.norm(bear, axis=1, keepdims=True)
    sols = P.p3p(X[:3], bear)
    assert min(np.abs(s[0] - R).max() + np.abs(s[1] - t).max() for s in sols) < 1e-9
    uv = uv + rng.normal(0, 0.3, uv.shape)
    out = rng.choice(m, 150, replace=False)
    uv[out] += rng.normal(0, 40, (150, 2))
    r = P.pnp_ransac(X, uv, f, cx, cy, iters=256, thr=8.0, seed=3)
    dR = r["R"] @ R.T
    assert np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)) < 2e-4 and np.linalg.norm(r["t"] - t) < 2e-3
    ok, rv, tv, inl = cv2.solvePnPRansac(X, uv, np.array([[f, 0, cx], [0, f, cy], [0, 0, 1]]), None, iterationsCount=256, reprojectionError=8.0)
    if ok and len(inl) == r["n_inliers"]:
        assert np.abs(cv2.Rodrigues(rv)[0] - r["R"]).max() < 1e-6 and np.abs(tv.ravel() - r["t"]).max() < 1e-6
