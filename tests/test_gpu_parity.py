"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the oracle and the reference-generated fixtures.
Bit-exact for disparity / keypoints / descriptors / matches; poses within 1e-4 rad and 1e-3 relative translation
(BASELINE.json north_star) — in practice ~1e-14."""
import ctypes

import numpy as np
import pytest

from conftest import SGBM_CASES, block_mask, occluded_pair, sgbm_params
from openvo_b200 import StereoCamera, StereoOdometer, synth
from oracle import openvo_port as O

pytestmark = pytest.mark.gpu
ROT_TOL, TRANS_REL_TOL = 1e-4, 1e-3


def _cam(W, H, D, **kw):
    args = synth.camera_args(W, H, D)
    args["sgbm_params"].update(kw)
    return StereoCamera(**args), args


def _pose_close(T, Tref):
    dR = T[:3, :3] @ Tref[:3, :3].T
    ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
    dt = np.linalg.norm(T[:3, 3] - Tref[:3, 3])
    return ang <= ROT_TOL and dt <= TRANS_REL_TOL * max(np.linalg.norm(Tref[:3, 3]), 1e-9) + 1e-12


@pytest.mark.parametrize("W,H,D,kw", SGBM_CASES + [(400, 100, 256, {}), (1241, 376, 128, {})])
def test_sgbm_bit_exact(W, H, D, kw):
    cam, args = _cam(W, H, D, **kw)
    L, R = occluded_pair(W, H, d=min(24, D // 2))
    got = cam.stereoSGBM.compute(L, R)
    assert got.dtype == np.int16 and np.array_equal(got, O.sgbm_compute(L, R, args["sgbm_params"]))


def test_sgbm_batch_and_idempotence():
    W, H, D = 320, 96, 64
    cam, args = _cam(W, H, D)
    eng = cam.engine(max_batch=3)
    L, R = occluded_pair(W, H)
    Ls = np.stack([np.roll(L, 5 * i, 1) for i in range(3)])
    Rs = np.stack([np.roll(R, 5 * i, 1) for i in range(3)])
    a = eng.sgbm(eng.upload(Ls, "l"), eng.upload(Rs, "r")).cpu().numpy()
    b = eng.sgbm(eng.upload(Ls, "l"), eng.upload(Rs, "r")).cpu().numpy()
    assert np.array_equal(a, b)
    for f in range(3):
        assert np.array_equal(a[f], O.sgbm_compute(Ls[f], Rs[f], args["sgbm_params"]))


def test_sgbm_properties_full_size():
    # size-independent properties at the KITTI shape: first D columns invalid, constant-shift scene recovers its shift
    W, H, D = 1241, 376, 128
    cam, _ = _cam(W, H, D)
    L, R = synth.kat_pair(W, H, d=24)
    d = cam.stereoSGBM.compute(L, R)
    assert (d[:, :D] == -16).all()
    valid = d[d >= 0]
    assert valid.size > 0.85 * (W - D) * H and np.median(valid) == 24 * 16


def test_sgbm_rejects_bad_input():
    cam, _ = _cam(320, 96, 64)
    L, R = occluded_pair(320, 96)
    with pytest.raises(ValueError):
        cam.stereoSGBM.compute(L, R[:, :-1])
    with pytest.raises(ValueError):
        cam.stereoSGBM.compute(L.astype(np.float32), R)


@pytest.mark.parametrize("W,H,n,usemask", [(640, 200, 500, False), (415, 333, 300, True), (1241, 376, 2000, True),
                                           (1241, 376, 1000, False)])
def test_orb_bit_exact(W, H, n, usemask):
    cam, _ = _cam(W, H, 16)
    od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    eng = od._engine()
    L, _r = synth.kat_pair(W, H)
    img = np.ascontiguousarray(L[:eng.ch, :eng.cw])
    mask = block_mask(eng.ch, eng.cw) if usemask else None
    kps, desc = od.orb.detectAndCompute(img, mask)
    rk, rd = O.orb_detect_compute(img, mask, n)
    got = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)
    assert np.array_equal(got, rk) and np.array_equal(desc, rd)
    assert all(k.class_id == -1 for k in kps)


def test_orb_empty_and_ragged():
    cam, _ = _cam(320, 120, 16)
    od = StereoOdometer(cam, nfeatures=300, preprocessed_frames=True)
    eng = od._engine()
    flat = np.full((eng.ch, eng.cw), 77, np.uint8)
    assert od.orb.detectAndCompute(flat, None) == ((), None)
    L, _r = synth.kat_pair(320, 120)
    img = np.ascontiguousarray(L[:eng.ch, :eng.cw])
    zero_mask = np.zeros_like(img)
    assert od.orb.detectAndCompute(img, zero_mask) == ((), None)
    with pytest.raises(ValueError):
        od.orb.detectAndCompute(img[:, :-3], None)


@pytest.mark.parametrize("nq,nt", [(1, 2), (5, 2), (300, 257), (2000, 2000), (1798, 1801)])
def test_knn_bit_exact_with_ties(nq, nt):
    cam, _ = _cam(640, 200, 16)
    od = StereoOdometer(cam, nfeatures=max(nq, nt), preprocessed_frames=True)
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8) & 0xF0
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[nt // 2:] &= 0xF0
    ref = O.knn2_hamming(q, t)
    mm = od.matcher.knnMatch(q, t, k=2)
    got = np.array([[m[0].trainIdx, int(m[0].distance), m[1].trainIdx, int(m[1].distance)] for m in mm], np.int32)
    assert np.array_equal(got, ref)
    assert all(m[0].queryIdx == i and m[0].imgIdx == 0 for i, m in enumerate(mm))


def test_compute_3d_matches_reference_contract():
    W, H, D = 480, 160, 64
    cam, args = _cam(W, H, D)
    Ls, Rs, _ = synth.make_sequence(W, H, 1)
    xyz, disp, img = cam.compute_3d(Ls[0], Rs[0], preprocessed=True)
    port = O.StereoCameraPort(**args, backend="restated")
    rxyz, rdisp, rimg = port.compute_3d(Ls[0], Rs[0], preprocessed=True)
    assert xyz.dtype == np.float32 and disp.dtype == np.float32 and img.dtype == np.uint8
    assert np.array_equal(disp, rdisp) and np.array_equal(img, rimg)
    assert np.array_equal(xyz.view(np.uint32), rxyz.view(np.uint32))


def _replay(g, **kw):
    W, H, D, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    cam = StereoCamera(**(synth.camera_args_distorted if bool(g["distorted"]) else synth.camera_args)(W, H, D))
    assert tuple(cam.valid_region_left) == tuple(int(v) for v in g["roi"])
    od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=bool(g["preprocessed"]), **kw)
    for i in range(len(g["left"])):
        ok = od.update(g["left"][i], g["right"][i])
        assert ok == bool(g["ok_%d" % i]), i
        assert od.skip_cause == str(g["cause_%d" % i]), i
        assert od.skipped_frames == int(g["skipped_%d" % i]), i
        if od.current_disparity is not None:
            assert np.array_equal(np.rint(od.current_disparity * 16).astype(np.int16), g["disp16_%d" % i]), i
            got = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in od.current_kps], np.float32)
            assert np.array_equal(got.reshape(-1, 6), g["kp_%d" % i]) and np.array_equal(od.current_desc, g["desc_%d" % i]), i
        assert _pose_close(od.c_T_w, g["cTw_%d" % i]), i
        assert _pose_close(od.current_pose(), g["pose_%d" % i]), i
    return od


@pytest.mark.parametrize("name,kw", [("seq_small", {}), ("seq_skip", {}), ("seq_rectify", {}),
                                     ("seq_filters", dict(rigidity_threshold=0.06, outlier_threshold=0.02))])
def test_update_reproduces_reference_fixtures(golden, name, kw):
    _replay(golden(name), **kw)


def test_seam_fixture(golden):
    g = golden("seams_small")
    H, W = g["left"].shape
    cam, _ = _cam(W, H, 64)
    assert np.array_equal(cam.stereoSGBM.compute(g["left"], g["right"]), g["sgbm"])


def test_update_vs_oracle_kitti_shape():
    W, H, D, n = 1241, 376, 128, 2000
    Ls, Rs, _ = synth.make_sequence(W, H, 3)
    cam, args = _cam(W, H, D)
    od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    po = O.StereoOdometerPort(O.StereoCameraPort(**args, backend="cv2"), nfeatures=n, preprocessed_frames=True)
    for i in range(3):
        assert od.update(Ls[i], Rs[i]) == po.update(Ls[i], Rs[i])
        assert np.array_equal(od.current_disparity, po.cur[1])
        assert np.array_equal(od._host(od._cur, "kp_array"), po.cur[3]) and np.array_equal(od.current_desc, po.cur[4])
        if i:
            eng = od._engine()
            assert np.array_equal(eng.matches[0, :od.last_match_count].cpu().numpy(), po.last_matches)
        assert _pose_close(od.c_T_w, po.c_T_w)
    # lazily materialised reference-typed state
    assert od.current_3d.shape == (od._engine().ch, od._engine().cw, 3)
    assert np.array_equal(od.current_3d.view(np.uint32), po.cur[2].view(np.uint32))
    assert od.prev_img is not None and od.current_img.dtype == np.uint8


def test_batch_odometer_equals_single_stream(golden):
    from openvo_b200.batch import BatchOdometer
    g = golden("seq_skip")
    W, H, D, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    cam, _ = _cam(W, H, D)
    S = 3
    bo = BatchOdometer(cam, S, nfeatures=n, preprocessed_frames=True)
    nfr = len(g["left"])
    # sequence s replays the fixture delayed by s frames (so the batch holds frames in different states)
    for step in range(nfr + S - 1):
        idx = [min(max(step - s, 0), nfr - 1) for s in range(S)]
        bo.update(g["left"][idx], g["right"][idx])
    singles = []
    for s in range(S):
        od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
        for step in range(nfr + S - 1):
            i = min(max(step - s, 0), nfr - 1)
            od.update(g["left"][i], g["right"][i])
        singles.append(od)
    for s in range(S):
        assert np.array_equal(bo.odometers[s].c_T_w, singles[s].c_T_w)
        assert bo.odometers[s].skip_cause == singles[s].skip_cause
        assert bo.odometers[s].skipped_frames == singles[s].skipped_frames


def test_batches_in_flight_round_robin_equals_update(golden):
    """begin()/finish() with two batches in flight on two streams (the bench driver) == update() called step by step."""
    import torch
    from openvo_b200.batch import BatchOdometer
    g = golden("seq_skip")
    W, H, D, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    cam, _ = _cam(W, H, D)
    S, G = 2, 2
    nfr = len(g["left"])
    steps = nfr + S * G - 1

    def frames_of(step, grp):
        idx = [min(max(step - (grp * S + s), 0), nfr - 1) for s in range(S)]
        return g["left"][idx], g["right"][idx]
    bos = [BatchOdometer(cam, S, nfeatures=n, engine_tag=10 + grp, preprocessed_frames=True) for grp in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    got = [[] for _ in range(G)]
    for grp in range(G):
        with torch.cuda.stream(streams[grp]):
            bos[grp].begin(*frames_of(0, grp))
    for step in range(steps):
        for grp in range(G):
            with torch.cuda.stream(streams[grp]):
                got[grp].append(bos[grp].finish())
                if step + 1 < steps:
                    bos[grp].begin(*frames_of(step + 1, grp))
    torch.cuda.synchronize()
    for grp in range(G):
        ref = BatchOdometer(cam, S, nfeatures=n, engine_tag=20 + grp, preprocessed_frames=True)
        want = [ref.update(*frames_of(step, grp)) for step in range(steps)]
        assert got[grp] == want
        for s in range(S):
            assert np.array_equal(bos[grp].odometers[s].c_T_w, ref.odometers[s].c_T_w)
            assert bos[grp].odometers[s].skip_cause == ref.odometers[s].skip_cause


def test_rectify_and_gray_bit_exact():
    import cv2
    W, H = 480, 160
    args = synth.camera_args_distorted(W, H, 64)
    cam = StereoCamera(**args)
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    assert np.array_equal(cam.undistort_rectify_left(img), cv2.remap(img, cam.map_left_1, cam.map_left_2, cv2.INTER_LINEAR))
    assert np.array_equal(cam.undistort_rectify_right(img), cv2.remap(img, cam.map_right_1, cam.map_right_2, cv2.INTER_LINEAR))
    # colour + remap through compute_3d vs the cv2-backed port
    Ls, Rs, _ = synth.make_sequence(W, H, 1)
    Lc, Rc = synth.to_bgr(Ls[0], 1), synth.to_bgr(Rs[0], 2)
    port = O.StereoCameraPort(**args, backend="cv2")
    for pre in (False, True):
        xyz, disp, left = cam.compute_3d(Lc, Rc, preprocessed=pre)
        rxyz, rdisp, rleft = port.compute_3d(Lc, Rc, preprocessed=pre)
        assert np.array_equal(left, rleft) and np.array_equal(disp, rdisp)
        assert np.array_equal(xyz.view(np.uint32), rxyz.view(np.uint32))


def test_frame_chunk_sharding_equals_sequential(golden):
    # config 4 plumbing: two "ranks" run their frame chunks (with a one-frame halo) independently; the replayed chain equals
    # the sequential odometer's
    from openvo_b200 import dist as D
    g = golden("seq_small")
    W, H, Dn, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    cam, _ = _cam(W, H, Dn)
    nfr = len(g["left"])
    Tall = np.tile(np.eye(4), (nfr, 1, 1))
    sall = np.zeros(nfr, np.int32)
    for rank in range(2):
        od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
        start, T, st = D.run_frame_chunk(od, g["left"], g["right"], rank, 2)
        Tall[start:start + len(st)] = T
        sall[start:start + len(st)] = st
    chain = D.replay_chains(Tall[None], sall[None])[0]
    assert _pose_close(chain, g["cTw_%d" % (nfr - 1)])


@pytest.mark.parametrize("W,H,D", [(200, 60, 32), (320, 96, 64), (1241, 376, 128)])
def test_sgbm_mode_hh_opt_in(W, H, D):
    # extension row n4 (not the reference's behaviour): 8-direction MODE_HH, bit-exact vs the oracle restatement / cv2's MODE_HH
    args = synth.camera_args(W, H, D)
    args["sgbm_params"]["mode"] = 1
    cam = StereoCamera(**args)
    L, R = occluded_pair(W, H, d=min(24, D // 2))
    got = cam.stereoSGBM.compute(L, R)
    assert np.array_equal(got, O.sgbm_compute_mode(L, R, args["sgbm_params"], 1))
    cv2 = pytest.importorskip("cv2")
    p = args["sgbm_params"]
    ref = cv2.StereoSGBM_create(*[p[k] for k in O.SGBM_KEYS], mode=cv2.STEREO_SGBM_MODE_HH).compute(L, R)
    assert np.array_equal(got, ref)


def test_cross_check_opt_in(golden):
    # extension row n4: left-right cross-check of the matches (the reference's "TODO crosscheck"); default stays off
    g = golden("seq_small")
    W, H, D, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    cam, _ = _cam(W, H, D)
    od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True, cross_check=True)
    plain = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    for i in range(2):
        assert od.update(g["left"][i], g["right"][i])
    got = od._engine().matches[0, :od.last_match_count, 0].cpu().numpy()   # the two odometers share one engine's buffers
    for i in range(2):
        assert plain.update(g["left"][i], g["right"][i])
    fwd = O.knn2_hamming(g["desc_0"], g["desc_1"])
    rev = O.knn2_hamming(g["desc_1"], g["desc_0"])
    want = [i for i in range(len(fwd)) if float(fwd[i, 1]) < 0.8 * float(fwd[i, 3]) and rev[fwd[i, 0], 0] == i]
    assert od.last_match_count == len(want) < plain.last_match_count
    assert np.array_equal(got, want)


def test_sgbm_random_shapes_and_params():
    # seeded sweep over ragged widths / heights, every block size, padded and unpadded disparity ranges, odd penalties
    rng = np.random.default_rng(2024)
    for trial in range(24):
        D = int(rng.choice([16, 32, 48, 64, 80, 96, 128, 160, 256]))
        W = D + int(rng.integers(17, 160))
        H = int(rng.integers(16, 70))
        bs = int(rng.choice([3, 5, 7, 9, 11]))
        P1 = int(rng.integers(1, 60)) * bs
        P2 = P1 + int(rng.integers(1, 400))
        kw = dict(blockSize=bs, P1=P1, P2=P2, disp12MaxDiff=int(rng.choice([-1, 0, 1, 2, 5])), preFilterCap=int(rng.choice([1, 15, 31, 63])),
                  uniquenessRatio=int(rng.choice([0, 5, 10, 15, 40])), speckleWindowSize=int(rng.choice([0, 20, 100])),
                  speckleRange=int(rng.choice([1, 2, 4])))
        if bs * bs * (2 * (max(kw["preFilterCap"], 15) | 1) + 63) + max(P2, P1 + 1) > 32767:
            continue
        cam, args = _cam(W, H, D, **kw)
        L, R = occluded_pair(W, H, d=min(9, D // 2))
        got = cam.stereoSGBM.compute(L, R)
        assert np.array_equal(got, O.sgbm_compute(L, R, args["sgbm_params"])), (trial, W, H, D, kw)


def test_orb_random_shapes():
    # level widths decide where the blur's FMA body ends (w//32*32, w//4*4): sweep ragged sizes, with and without mask
    rng = np.random.default_rng(77)
    for trial in range(10):
        W, H = int(rng.integers(140, 700)), int(rng.integers(130, 400))
        n = int(rng.choice([50, 300, 1000]))
        cam, _ = _cam(W, H, 16)
        od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
        eng = od._engine()
        L, _r = synth.kat_pair(W, H, seed=trial)
        img = np.ascontiguousarray(L[:eng.ch, :eng.cw])
        mask = block_mask(eng.ch, eng.cw, seed=trial) if trial % 2 else None
        kps, desc = od.orb.detectAndCompute(img, mask)
        rk, rd = O.orb_detect_compute(img, mask, n)
        got = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)
        assert np.array_equal(got, rk) and (len(rk) == 0 or np.array_equal(desc, rd)), (trial, W, H, n)


def test_pnp_ransac_opt_in(golden):
    # extension row n4: batched P3P RANSAC + LM vs its numpy specification (oracle/pnp_restate.py), then through the odometer
    import ctypes
    import torch
    from oracle import pnp_restate as P
    from openvo_b200 import _native as N
    from test_emu_kernels import _pnp_case
    X, kp2, matches, Q, (f, cx, cy), _ = _pnp_case(seed=1, m=1500, n_out=400)
    cam, _ = _cam(300, 150, 32)
    eng = cam.engine(nfeatures=2000)
    eng.cfg.Q[:] = (ctypes.c_double * 16)(*Q.reshape(-1))      # synthetic intrinsics for this unit test
    ctx2 = eng.lib.ovo_create(ctypes.byref(eng.cfg), eng.workspace.data_ptr(), eng.workspace.numel())
    m = len(X)
    pts = torch.zeros((eng.kp_cap, 3), dtype=torch.float32, device="cuda"); pts[:m] = torch.from_numpy(X).cuda()
    mt = torch.zeros((eng.kp_cap, 3), dtype=torch.int32, device="cuda"); mt[:m] = torch.from_numpy(matches).cuda()
    kp = torch.zeros((eng.kp_cap, 6), dtype=torch.float32, device="cuda"); kp[:m] = torch.from_numpy(kp2).cuda()
    cnt = torch.tensor([m], dtype=torch.int32, device="cuda")
    out = torch.zeros(16, dtype=torch.float64, device="cuda")
    N.check(eng.lib, eng.lib.ovo_pnp_ransac(ctx2, pts.data_ptr(), mt.data_ptr(), kp.data_ptr(), cnt.data_ptr(), eng.kp_cap, 1024, 8.0, 7,
                                            out.data_ptr(), None))
    o = out.cpu().numpy()
    eng.lib.ovo_destroy(ctx2)
    ref = P.pnp_ransac(X.astype(np.float64), kp2[matches[:, 1], :2].astype(np.float64), f, cx, cy, iters=1024, thr=8.0, seed=7)
    assert int(o[15]) == ref["best"] and int(o[12]) == ref["n_inliers"] > 900
    T, Tr = np.eye(4), np.eye(4)
    T[:3, :4] = o[:12].reshape(3, 4)
    Tr[:3, :3], Tr[:3, 3] = ref["R"], ref["t"]
    assert _pose_close(T, Tr)
    # through the odometer: same frames, opt-in estimator; deterministic, and (unlike the reference's outlier-sensitive Umeyama fit)
    # close to the ground-truth motion of the synthetic sequence: 2 steps of (0.01, 0, 0.05) m with 0.002 rad yaw each
    g = golden("seq_small")
    W, H, D, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    cam2, _ = _cam(W, H, D)
    od = StereoOdometer(cam2, nfeatures=n, preprocessed_frames=True, pose_method="pnp_ransac", ransac_iters=512, ransac_seed=1)
    od2 = StereoOdometer(cam2, nfeatures=n, preprocessed_frames=True, pose_method="pnp_ransac", ransac_iters=512, ransac_seed=1)
    for i in range(3):
        assert od.update(g["left"][i], g["right"][i]) and od2.update(g["left"][i], g["right"][i])
    assert np.array_equal(od.c_T_w, od2.c_T_w)                      # deterministic for a fixed seed and schedule
    _, _, poses = synth.make_sequence(W, H, 3)
    assert np.linalg.norm(od.c_T_w[:3, 3] - poses[2][:3, 3]) < 0.03
    assert np.linalg.norm(g["cTw_2"][:3, 3] - poses[2][:3, 3]) > np.linalg.norm(od.c_T_w[:3, 3] - poses[2][:3, 3])


def test_update_vs_oracle_1080p_shape():
    # BASELINE config 3: 1920x1080, ORB 5000, 256 disparities — whole update() against the cv2-backed port
    W, H, D, n = 1920, 1080, 256, 5000
    Ls, Rs, _ = synth.make_sequence(W, H, 2)
    cam, args = _cam(W, H, D)
    od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    po = O.StereoOdometerPort(O.StereoCameraPort(**args, backend="cv2"), nfeatures=n, preprocessed_frames=True)
    for i in range(2):
        assert od.update(Ls[i], Rs[i]) == po.update(Ls[i], Rs[i])
        assert np.array_equal(od.current_disparity, po.cur[1])
        assert np.array_equal(od._host(od._cur, "kp_array"), po.cur[3]) and np.array_equal(od.current_desc, po.cur[4])
    assert np.array_equal(od._engine().matches[0, :od.last_match_count].cpu().numpy(), po.last_matches)
    assert od.skip_cause == po.skip_cause and _pose_close(od.c_T_w, po.c_T_w)


def test_reference_error_behaviour():
    """The reference's hard failures are reproduced, not swallowed: all four taps of a lookup unusable -> ZeroDivisionError
    (ref: src/openVO/stereo_odometer.py:79 with num = den = 0); a next frame with a single keypoint -> IndexError (ref: :164)."""
    import torch
    from openvo_b200.engine import Frame
    W, H, D, n = 480, 160, 64, 400
    Ls, Rs, _ = synth.make_sequence(W, H, 2)
    cam, _ = _cam(W, H, D)
    od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    assert od.update(Ls[0], Rs[0])
    eng = od._engine()
    good = od._cur
    # same features, but a disparity map that is 0 everywhere: reprojection gives +-inf at every tap
    zero = Frame(good.img, torch.zeros_like(good.disp), good.kp, good.desc, good.n_kp)
    with pytest.raises(ZeroDivisionError):
        od._relative(zero, zero)
    one = Frame(good.img, good.disp, good.kp, good.desc, 1)
    od.min_matches = 1
    with pytest.raises(IndexError):
        od._relative(good, one)


def test_reference_method_surface(golden):
    """Drop-in details of the reference's class API (ref: src/openVO/stereo_odometer.py:24-31,107-160): update() dispatches
    its steps through `self.` (subclass overrides take effect), save_frame_update takes the reference's five arguments,
    current_* / prev_* are assignable, and malformed input raises cv2.error."""
    import cv2
    g = golden("seq_small")
    W, H, D, n = int(g["W"]), int(g["H"]), int(g["D"]), int(g["nfeatures"])
    cam, _ = _cam(W, H, D)

    calls = []

    class Hooked(StereoOdometer):
        def point_clouds(self, *a):
            calls.append("point_clouds")
            return super().point_clouds(*a)

        def point_cloud_transform(self, p, q):
            calls.append("transform")
            return super().point_cloud_transform(p, q)

    # the hooked path (host glue, device seams) reproduces the reference's record like the fused path does
    od = Hooked(cam, nfeatures=n, preprocessed_frames=True)
    for i in range(len(g["left"])):
        assert od.update(g["left"][i], g["right"][i]) == bool(g["ok_%d" % i])
        assert od.skip_cause == str(g["cause_%d" % i])
        assert np.array_equal(np.rint(od.current_disparity * 16).astype(np.int16), g["disp16_%d" % i])
        assert np.array_equal(od.current_desc, g["desc_%d" % i])
        assert _pose_close(od.c_T_w, g["cTw_%d" % i]), i
    assert calls.count("point_clouds") == len(g["left"]) - 1 and "transform" in calls

    class Blind(StereoOdometer):
        def feature_mask(self, disparity):
            return np.zeros(disparity.shape, np.uint8)

    blind = Blind(cam, nfeatures=n, preprocessed_frames=True)
    assert blind.update(g["left"][0], g["right"][0]) is False and blind.skip_cause == "keypoints"

    # save_frame_update with the reference's signature + assignable state: seed an odometer with the products of frame 0 taken
    # from another odometer, then continue with frame 1 on the fused path
    src = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    assert src.update(g["left"][0], g["right"][0])
    dst = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    assert dst.current_img is None
    dst.save_frame_update(src.current_img, src.current_disparity, src.current_3d, src.current_kps, src.current_desc)
    assert dst.prev_img is None and np.array_equal(dst.current_desc, src.current_desc) and len(dst.current_kps) == len(src.current_kps)
    assert dst.update(g["left"][1], g["right"][1]) and src.update(g["left"][1], g["right"][1])
    assert np.array_equal(dst.c_T_w, src.c_T_w)
    dst.current_desc = dst.current_desc.copy()       # plain attribute semantics: assignable
    dst.prev_kps = dst.prev_kps
    assert dst.update(g["left"][2], g["right"][2]) and src.update(g["left"][2], g["right"][2])
    assert np.array_equal(dst.c_T_w, src.c_T_w)
    with pytest.raises(cv2.error):
        src.update(g["left"][0][:, :-1], g["right"][0])
    with pytest.raises(cv2.error):
        cam.stereoSGBM.compute(g["left"][0].astype(np.float32), g["right"][0])


def test_sgbm_one_volume_per_direction_path():
    """OVO_SGBM_FUSED=0 (the switch is read once per process -> child interpreter): the unfused vertical kernel, the A/B partner of
    k_sgbm_vsum and what MODE_HH builds on, stays bit-exact on the GPU too."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = ("import sys; sys.path[:0] = [%r, %r]; import numpy as np; from conftest import occluded_pair; "
            "from openvo_b200 import StereoCamera, synth; from oracle import openvo_port as O\n"
            "for W, H, D in ((320, 96, 64), (1241, 376, 128), (400, 100, 256)):\n"
            "    a = synth.camera_args(W, H, D); L, R = occluded_pair(W, H, d=min(24, D // 2))\n"
            "    assert np.array_equal(StereoCamera(**a).stereoSGBM.compute(L, R), O.sgbm_compute(L, R, a['sgbm_params'])), (W, H, D)\n"
            % (ROOT, os.path.join(ROOT, "tests")))
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, OVO_SGBM_FUSED="0"))
