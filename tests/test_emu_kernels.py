"""CPU tier: the CUDA sources compiled against tests/emu/cuda_emu.h (a fiber-based execution emulator) and driven through
the same C ABI, checked bit-exactly against the oracle.  This exercises the kernels' indexing / warp-collective / packed
arithmetic logic without a GPU; the real parity tests are the -m gpu ones."""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, block_mask, occluded_pair, sgbm_params
from openvo_b200 import _native as N
from openvo_b200 import synth
from oracle import openvo_port as O

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


@pytest.fixture(scope="module")
def emu():
    import build_emu
    return N.load(build_emu.build())


class Ctx:
    def __init__(self, lib, W, H, params, roi, Q, n, nb=1):
        self.lib = lib
        self.cfg = N.make_config(W, H, params, roi, Q, n, max_batch=nb)
        nbytes = lib.ovo_workspace_bytes(ctypes.byref(self.cfg))
        assert nbytes, lib.ovo_last_error()
        self.ws = np.zeros(nbytes + 256, np.uint8)
        off = (-self.ws.ctypes.data) % 256
        self.ctx = lib.ovo_create(ctypes.byref(self.cfg), self.ws.ctypes.data + off, nbytes)
        assert self.ctx, lib.ovo_last_error()
        self.cap = lib.ovo_kp_capacity(ctypes.byref(self.cfg))
        cw, ch = ctypes.c_int(), ctypes.c_int()
        lib.ovo_cropped_size(ctypes.byref(self.cfg), ctypes.byref(cw), ctypes.byref(ch))
        self.cw, self.ch = cw.value, ch.value

    def __del__(self):
        self.lib.ovo_destroy(self.ctx)


@pytest.mark.parametrize("W,H,D,nb,kw", [
    (160, 40, 32, 1, {}),
    (200, 36, 64, 1, dict(blockSize=3, P1=72, P2=288, uniquenessRatio=0, speckleWindowSize=0)),
    (150, 40, 48, 1, dict(blockSize=7, preFilterCap=15, uniquenessRatio=15, disp12MaxDiff=2, speckleWindowSize=50, speckleRange=1)),
    (200, 34, 128, 2, {}),
    (190, 21, 128, 1, dict(uniquenessRatio=0)),                                  # odd height: the two-rows-per-warp kernel's tail
    (170, 19, 96, 1, dict(blockSize=3, P1=72, P2=288, disp12MaxDiff=2)),        # padded 96 -> 128
    (180, 18, 112, 1, dict(mode=1)),                                            # opt-in MODE_HH through the same kernel
    (214, 70, 64, 1, dict(uniquenessRatio=5, _d=9)),                            # three 32-row bands x three 64-column blocks of the fused vertical kernel
    (300, 17, 256, 1, dict(uniquenessRatio=5, _d=31)),                          # 256 disparities; true disparity at a selection-lane edge
    (330, 16, 256, 1, dict(_d=62)),
    (282, 16, 208, 1, dict(blockSize=3, P1=72, P2=288)),                        # padded 208 -> 256
    (160, 20, 64, 1, dict(uniquenessRatio=100)),                                # uniquenessRatio >= 100: cell-by-cell selection
    (265, 33, 128, 1, dict(_d=20)),                                             # fused vertical kernel: 3 tiles x 3 bands, last band = 1 row
    (140, 18, 128, 1, dict(_d=3)),                                              # W1 = 12: narrower than the halo of a tile
    (353, 17, 256, 1, dict(_d=40)),                                             # 256 disparities: 4 tiles of 32 columns, bands of 8 rows
])
def test_sgbm_kernels(emu, W, H, D, nb, kw):
    _sgbm_case(emu, W, H, D, nb, kw)


def test_sgbm_kernels_one_volume_per_direction(emu, monkeypatch):
    """The unfused vertical kernel (OVO_SGBM_FUSED=0; what MODE_HH builds on) stays bit-exact too.  The switch is read once per
    process, so this runs in a child interpreter."""
    import subprocess
    code = ("import sys; sys.path[:0] = [%r, %r, %r]; import test_emu_kernels as T, build_emu; from openvo_b200 import _native as N; "
            "lib = N.load(build_emu.build()); T._sgbm_case(lib, 200, 34, 128, 2, {}); T._sgbm_case(lib, 300, 17, 256, 1, dict(_d=31)); "
            "T._sgbm_case(lib, 170, 19, 96, 1, dict(blockSize=3, P1=72, P2=288, disp12MaxDiff=2))"
            % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")))
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, OVO_SGBM_FUSED="0"))


@pytest.mark.parametrize("sched", ["1", "5", "11"])
def test_sgbm_kernels_other_warp_interleavings(emu, sched):
    """The fused vertical kernel hands rows from warp to warp through producer / consumer named barriers and copy-engine flags, with
    no CTA-wide barrier: its result must not depend on how the warps interleave.  OVO_EMU_SCHED runs the emulator's threads in
    reverse / pseudo-random order (reshuffled every sweep) and OVO_EMU_POISON fills stacks and shared memory with a byte pattern;
    both are read once per process, hence the child interpreter.  A thread arriving twice in one barrier phase or a wait that
    never completes aborts the child."""
    import subprocess
    code = ("import sys; sys.path[:0] = [%r, %r, %r]; import test_emu_kernels as T, build_emu; from openvo_b200 import _native as N; "
            "lib = N.load(build_emu.build()); T._sgbm_case(lib, 214, 70, 64, 1, dict(uniquenessRatio=5, _d=9)); "
            "T._sgbm_case(lib, 265, 33, 128, 1, dict(_d=20)); T._sgbm_case(lib, 353, 17, 256, 1, dict(_d=40)); "
            "T._sgbm_case(lib, 140, 18, 128, 1, dict(_d=3))"
            % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")))
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, OVO_EMU_SCHED=sched, OVO_EMU_POISON="0xA5"))


def _sgbm_case(emu, W, H, D, nb, kw):
    kw = dict(kw)
    shift = kw.pop("_d", 7)
    p = sgbm_params(D, **kw)
    L, R = occluded_pair(W, H, d=shift)
    Ls = np.stack([np.roll(L, 3 * i, 1) for i in range(nb)])
    Rs = np.stack([np.roll(R, 3 * i, 1) for i in range(nb)])
    c = Ctx(emu, W, H, p, (0, 0, W, H), np.eye(4), 100, nb)
    out = np.zeros((nb, H, W), np.int16)
    N.check(emu, emu.ovo_sgbm_compute(c.ctx, N.ptr(Ls), N.ptr(Rs), W, W * H, nb, N.ptr(out), None))
    for f in range(nb):
        assert np.array_equal(out[f], O.sgbm_compute_mode(Ls[f], Rs[f], p, p.get("mode", 0)))


@pytest.mark.parametrize("W,H,n,usemask,nb", [(320, 120, 300, False, 1), (300, 170, 300, True, 2), (161, 97, 150, True, 1),
                                              (257, 66, 200, False, 1), (96, 230, 120, True, 1)])
def test_orb_kernels(emu, W, H, n, usemask, nb):
    c = Ctx(emu, W, H, sgbm_params(16), (0, 0, W, H), np.eye(4), n, nb)
    L, R = synth.kat_pair(W, H)
    imgs = np.stack([L, R][:nb])
    mask = np.stack([block_mask(H, W)] * nb) if usemask else None
    kp = np.zeros((nb, c.cap, 6), np.float32)
    desc = np.zeros((nb, c.cap, 32), np.uint8)
    nk = (ctypes.c_int * nb)()
    N.check(emu, emu.ovo_orb_detect_compute(c.ctx, N.ptr(imgs), N.ptr(mask), nb, N.ptr(kp), N.ptr(desc), nk, None))
    for f in range(nb):
        rk, rd = O.orb_detect_compute(imgs[f], None if mask is None else mask[f], n)
        assert nk[f] == len(rk) and np.array_equal(kp[f, :nk[f]], rk) and np.array_equal(desc[f, :nk[f]], rd)


def test_match_and_pose_kernels(emu):
    rng = np.random.default_rng(3)
    W, H, n = 300, 150, 300
    Q = np.array([[1, 0, 0, -(W - 1) / 2], [0, 1, 0, -(H - 1) / 2], [0, 0, 0, 300.0], [0, 0, 1 / 0.537, 0.0]]) + rng.normal(0, 1e-3, (4, 4))
    roi = (2, 3, W - 1, H - 2)
    c = Ctx(emu, W, H, sgbm_params(32), roi, Q, n)
    # 2-NN with ties
    for nq, nt in ((1, 1), (5, 2), (300, 257)):
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8) & 0xF0
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        t[nt // 2:] &= 0xF0
        nn = np.zeros((nq, 4), np.int32)
        N.check(emu, emu.ovo_knn2_hamming(c.ctx, N.ptr(q), nq, N.ptr(t), nt, N.ptr(nn), None))
        assert np.array_equal(nn, O.knn2_hamming(q, t))
    # disparity post + reprojection
    d16 = rng.integers(-1, 1600, (H, W)).astype(np.int16)
    d16[rng.random((H, W)) < 0.1] = -16
    d16[rng.random((H, W)) < 0.05] = 0
    df = np.zeros((c.ch, c.cw), np.float32)
    mk = np.zeros((c.ch, c.cw), np.uint8)
    N.check(emu, emu.ovo_disparity_post(c.ctx, N.ptr(d16), 1, N.ptr(df), N.ptr(mk), None))
    full = d16.astype(np.float32) / 16
    crop = full[roi[1]:roi[3], roi[0]:roi[2]]
    assert np.array_equal(df, crop) and np.array_equal(mk, ((crop >= 4) * (crop <= 100)).astype(np.uint8) * 255)
    xyz = np.zeros((c.ch, c.cw, 3), np.float32)
    N.check(emu, emu.ovo_reproject_3d(c.ctx, N.ptr(df), N.ptr(xyz), None))
    ref3 = O.reproject_to_3d(full, Q)[roi[1]:roi[3], roi[0]:roi[2]]
    assert np.array_equal(xyz.view(np.uint32), ref3.view(np.uint32))
    # ratio + ordered compaction + fused lookup
    nq, nt = 200, 180
    kp1, kp2 = np.zeros((nq, 6), np.float32), np.zeros((nt, 6), np.float32)
    for kp, m in ((kp1, nq), (kp2, nt)):
        kp[:, 0] = rng.uniform(0, c.cw - 0.01, m)
        kp[:, 1] = rng.uniform(0, c.ch - 0.01, m)
    kp1[:30, :2] = np.floor(kp1[:30, :2])
    kp1[0, :2] = (c.cw - 1, c.ch - 1)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[:100] = q[:100] ^ (rng.random((100, 32)) < 0.05).astype(np.uint8)
    nn = np.zeros((nq, 4), np.int32)
    N.check(emu, emu.ovo_knn2_hamming(c.ctx, N.ptr(q), nq, N.ptr(t), nt, N.ptr(nn), None))
    matches, p1, p2, cnt = np.zeros((nq, 3), np.int32), np.zeros((nq, 3), np.float32), np.zeros((nq, 3), np.float32), np.zeros(2, np.int32)
    N.check(emu, emu.ovo_match_points(c.ctx, N.ptr(nn), nq, 0.8, N.ptr(kp1), N.ptr(kp2), N.ptr(df), N.ptr(df), N.ptr(matches),
                                     N.ptr(p1), N.ptr(p2), N.ptr(cnt), None, None))
    keep = [i for i in range(nq) if float(nn[i, 1]) < 0.8 * float(nn[i, 3])]
    assert cnt[0] == len(keep) >= 90 and np.array_equal(matches[:cnt[0], 0], keep)
    with np.errstate(all="ignore"):
        for j, i in enumerate(keep):
            a = O.bilinear_lookup(ref3, float(kp1[i, 0]), float(kp1[i, 1]))
            b = O.bilinear_lookup(ref3, float(kp2[nn[i, 0], 0]), float(kp2[nn[i, 0], 1]))
            assert np.array_equal(a.view(np.uint32), p1[j].view(np.uint32)) or np.isnan(a).any()
            assert np.array_equal(b.view(np.uint32), p2[j].view(np.uint32)) or np.isnan(b).any()
    # batched pair step (ovo_pair_batch) == the three single-pair seams
    items = (N.PairItem * 2)()
    outs, keepers = [], []
    for j, (qq, tt, ka, kb) in enumerate(((q, t, kp1, kp2), (t[:150], q, kp2[:150], kp1))):
        bufs = dict(q=np.ascontiguousarray(qq), t=np.ascontiguousarray(tt), ka=np.ascontiguousarray(ka), kb=np.ascontiguousarray(kb),
                    nn=np.zeros((len(qq), 4), np.int32), m=np.zeros((len(qq), 3), np.int32), p1=np.zeros((len(qq), 3), np.float32),
                    p2=np.zeros((len(qq), 3), np.float32), out=np.zeros(18))
        keepers.append(bufs)
        it = items[j]
        it.q_desc, it.t_desc, it.nq, it.nt = N.ptr(bufs["q"]), N.ptr(bufs["t"]), len(qq), len(tt)
        it.kp1, it.kp2, it.disp1, it.disp2 = N.ptr(bufs["ka"]), N.ptr(bufs["kb"]), N.ptr(df), N.ptr(df)
        it.nn, it.matches, it.pts1, it.pts2, it.out = N.ptr(bufs["nn"]), N.ptr(bufs["m"]), N.ptr(bufs["p1"]), N.ptr(bufs["p2"]), N.ptr(bufs["out"])
    c2 = Ctx(emu, W, H, sgbm_params(32), roi, Q, n, nb=2)
    N.check(emu, emu.ovo_pair_batch(c2.ctx, 2, items, 0.8, None))
    for bufs in keepers:
        qq, tt = bufs["q"], bufs["t"]
        nn1 = np.zeros((len(qq), 4), np.int32)
        N.check(emu, emu.ovo_knn2_hamming(c.ctx, N.ptr(qq), len(qq), N.ptr(tt), len(tt), N.ptr(nn1), None))
        assert np.array_equal(nn1, bufs["nn"])
        m1, a1, b1, c1 = np.zeros((len(qq), 3), np.int32), np.zeros((len(qq), 3), np.float32), np.zeros((len(qq), 3), np.float32), np.zeros(2, np.int32)
        N.check(emu, emu.ovo_match_points(c.ctx, N.ptr(nn1), len(qq), 0.8, N.ptr(bufs["ka"]), N.ptr(bufs["kb"]), N.ptr(df), N.ptr(df),
                                         N.ptr(m1), N.ptr(a1), N.ptr(b1), N.ptr(c1), None, None))
        cnt_b = bufs["out"][16:17].view(np.int32)
        assert cnt_b[0] == c1[0] and cnt_b[1] == c1[1]
        k = c1[0]
        assert np.array_equal(m1[:k], bufs["m"][:k]) and np.array_equal(a1[:k].view(np.uint32), bufs["p1"][:k].view(np.uint32))
        o1 = np.zeros(16)
        N.check(emu, emu.ovo_rigid_transform(c.ctx, N.ptr(a1), N.ptr(b1), N.ptr(c1), c.cap, N.ptr(o1), None))
        assert np.array_equal(o1, bufs["out"][:16], equal_nan=True)
    # opt-in cross-check: batched pair step with nn_rev == forward matches filtered by mutual nearest neighbours (cv2 tie rule)
    items2 = (N.PairItem * 1)()
    b0 = keepers[0]
    rev = np.zeros((len(b0["t"]), 4), np.int32)
    out2, m2 = np.zeros(18), np.zeros((len(b0["q"]), 3), np.int32)
    it = items2[0]
    it.q_desc, it.t_desc, it.nq, it.nt = N.ptr(b0["q"]), N.ptr(b0["t"]), len(b0["q"]), len(b0["t"])
    it.kp1, it.kp2, it.disp1, it.disp2 = N.ptr(b0["ka"]), N.ptr(b0["kb"]), N.ptr(df), N.ptr(df)
    it.nn, it.matches, it.pts1, it.pts2, it.out = N.ptr(b0["nn"]), N.ptr(m2), N.ptr(b0["p1"]), N.ptr(b0["p2"]), N.ptr(out2)
    it.nn_rev = N.ptr(rev)
    N.check(emu, emu.ovo_pair_batch(c2.ctx, 1, items2, 0.8, None))
    assert np.array_equal(rev, O.knn2_hamming(b0["t"], b0["q"]))
    fwd = O.knn2_hamming(b0["q"], b0["t"])
    want = [i for i in range(len(fwd)) if float(fwd[i, 1]) < 0.8 * float(fwd[i, 3]) and rev[fwd[i, 0], 0] == i]
    k2 = int(out2[16:17].view(np.int32)[0])
    assert k2 == len(want) and np.array_equal(m2[:k2, 0], want) and 0 < k2 <= cnt_b[0] + 1000
    # Umeyama
    src = rng.normal(0, 5, (150, 3)).astype(np.float32)
    ang = 0.04
    Rm = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    dst = (src @ Rm.T + [0.1, -0.02, 0.5] + rng.normal(0, 0.01, (150, 3))).astype(np.float32)
    out = np.zeros(16)
    m = np.array([150], np.int32)
    N.check(emu, emu.ovo_rigid_transform(c.ctx, N.ptr(src), N.ptr(dst), N.ptr(m), 150, N.ptr(out), None))
    T, s = O.umeyama(src, dst)
    assert np.abs(out[:12].reshape(3, 4) - T).max() < 1e-12 and abs(out[12] - s) < 1e-12
    assert abs(out[13] - O.rotation_angle(T[:, :3])) < 1e-10 and abs(out[14] - np.linalg.norm(T[:, 3])) < 1e-12


def test_rectify_kernel(emu):
    rng = np.random.default_rng(9)
    W, H = 200, 120
    c = Ctx(emu, W, H, sgbm_params(32), (0, 0, W, H), np.eye(4), 100, nb=2)
    gray = rng.integers(0, 256, (2, H, W), dtype=np.uint8)
    bgr = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    m1 = np.stack([rng.integers(-3, W + 3, (H, W)), rng.integers(-3, H + 3, (H, W))], -1).astype(np.int16)
    m2 = rng.integers(0, 1024, (H, W)).astype(np.uint16)
    out = np.zeros((2, H, W), np.uint8)
    N.check(emu, emu.ovo_rectify(c.ctx, N.ptr(gray), 1, W, W * H, 2, N.ptr(m1), N.ptr(m2), N.ptr(out), None))
    for f in range(2):
        assert np.array_equal(out[f], O.remap_linear(gray[f], m1, m2))
    N.check(emu, emu.ovo_rectify(c.ctx, N.ptr(bgr), 3, 3 * W, 3 * W * H, 2, N.ptr(m1), N.ptr(m2), N.ptr(out), None))
    for f in range(2):
        assert np.array_equal(out[f], O.remap_linear(O.bgr2gray(bgr[f]), m1, m2))
    N.check(emu, emu.ovo_rectify(c.ctx, N.ptr(bgr), 3, 3 * W, 3 * W * H, 2, None, None, N.ptr(out), None))
    for f in range(2):
        assert np.array_equal(out[f], O.bgr2gray(bgr[f]))


def test_pose_filter_kernels(emu):
    """ovo_rigid_body_filter / ovo_outlier_filter vs the reference's numpy formulation (oracle port)."""
    rng = np.random.default_rng(11)
    W, H, n = 300, 150, 300
    c = Ctx(emu, W, H, sgbm_params(32), (0, 0, W, H), np.eye(4), n)
    m = 180
    ang = 0.03
    Rm = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    prev = rng.uniform(-5, 5, (m, 3)).astype(np.float32) + np.float32([0, 0, 12])
    cur = (prev @ Rm.T + [0.05, 0.0, 0.2] + rng.normal(0, 0.01, (m, 3))).astype(np.float32)
    bad = rng.choice(m, 40, replace=False)
    cur[bad] += rng.normal(0, 1.0, (40, 3)).astype(np.float32)

    class Dummy:
        backend_name, sgbm_params = "restated", sgbm_params(32)
    port = O.StereoOdometerPort(Dummy(), nfeatures=0, rigidity_threshold=0.06, outlier_threshold=0.02)
    clique = port.rigid_body_filter(prev, cur)
    a, b = np.zeros((c.cap, 3), np.float32), np.zeros((c.cap, 3), np.float32)
    a[:m], b[:m] = prev, cur
    cnt = np.array([m], np.int32)
    N.check(emu, emu.ovo_rigid_body_filter(c.ctx, N.ptr(a), N.ptr(b), N.ptr(cnt), c.cap, float(np.float32(0.06)), None))
    k = int(clique.sum())
    assert cnt[0] == k and 100 < k < m
    assert np.array_equal(a[:k], prev[clique > 0]) and np.array_equal(b[:k], cur[clique > 0])
    # outlier filter on the survivors
    p, q = prev[clique > 0], cur[clique > 0]
    q = q.copy()
    q[::9] += np.float32(0.8)
    b[:k] = q
    T, _ = O.umeyama(p, q)
    T4 = np.vstack([T, [0, 0, 0, 1]])
    hn = np.hstack([q, np.ones((k, 1))])
    hp = np.hstack([p, np.ones((k, 1))])
    err = np.array([np.linalg.norm(hn[i] - T4 @ hp[i]) / np.linalg.norm(hn[i]) for i in range(k)])
    keep = err < 0.02 + np.median(err)
    Tdev = np.ascontiguousarray(T.reshape(-1))
    N.check(emu, emu.ovo_outlier_filter(c.ctx, N.ptr(a), N.ptr(b), N.ptr(cnt), c.cap, N.ptr(Tdev), 0.02, None))
    assert cnt[0] == keep.sum() and 0 < keep.sum() < k
    assert np.array_equal(a[:cnt[0]], p[keep]) and np.array_equal(b[:cnt[0]], q[keep])


def test_sgbm_mode_hh_kernels(emu):
    """opt-in 8-direction MODE_HH vs its oracle restatement (which is pinned to cv2's MODE_HH)."""
    W, H, D = 160, 40, 32
    p = sgbm_params(D)
    p["mode"] = 1
    L, R = occluded_pair(W, H)
    c = Ctx(emu, W, H, p, (0, 0, W, H), np.eye(4), 100)
    out = np.zeros((1, H, W), np.int16)
    N.check(emu, emu.ovo_sgbm_compute(c.ctx, N.ptr(L), N.ptr(R), W, W * H, 1, N.ptr(out), None))
    ref = O.sgbm_compute_mode(L, R, p, 1)
    assert np.array_equal(out[0], ref)
    assert (ref != O.sgbm_compute(L, R, p)).sum() > 0  # and it is a different result from the reference's MODE_SGBM


def test_sgbm_kernels_random_small(emu):
    rng = np.random.default_rng(5)
    for trial in range(4):
        D = int(rng.choice([16, 48, 80]))
        W, H = D + int(rng.integers(17, 60)), int(rng.integers(16, 30))
        bs = int(rng.choice([3, 5, 9]))
        kw = dict(blockSize=bs, P1=8 * bs, P2=8 * bs + int(rng.integers(1, 300)), disp12MaxDiff=int(rng.choice([-1, 1, 3])),
                  preFilterCap=int(rng.choice([1, 31, 63])), uniquenessRatio=int(rng.choice([0, 10, 40])),
                  speckleWindowSize=int(rng.choice([0, 30])), speckleRange=2)
        p = sgbm_params(D, **kw)
        L, R = occluded_pair(W, H, d=5)
        c = Ctx(emu, W, H, p, (0, 0, W, H), np.eye(4), 100)
        out = np.zeros((1, H, W), np.int16)
        N.check(emu, emu.ovo_sgbm_compute(c.ctx, N.ptr(L), N.ptr(R), W, W * H, 1, N.ptr(out), None))
        assert np.array_equal(out[0], O.sgbm_compute(L, R, p)), (trial, W, H, D, kw)


def test_sgbm_kernels_tiny_widths(emu):
    """W - D of 1..13 columns: empty phases, a lone partial segment, rendezvous at the row ends (both modes)."""
    for D, w1s in ((16, (8, 9, 12, 13)), (64, (1, 2, 3, 4, 5, 7, 8, 9, 13))):
        for w1 in w1s:
            for mode in (0, 1):
                W, H = D + w1, 17
                p = sgbm_params(D, uniquenessRatio=5)
                p["mode"] = mode
                L, R = occluded_pair(W, H, d=3)
                c = Ctx(emu, W, H, p, (0, 0, W, H), np.eye(4), 100)
                out = np.zeros((1, H, W), np.int16)
                N.check(emu, emu.ovo_sgbm_compute(c.ctx, N.ptr(L), N.ptr(R), W, W * H, 1, N.ptr(out), None))
                assert np.array_equal(out[0], O.sgbm_compute_mode(L, R, p, mode)), (D, w1, mode)


def _pnp_case(seed=0, m=500, n_out=120):
    from oracle import pnp_restate as P
    rng = np.random.default_rng(seed)
    f, cx, cy = 300.0, 149.5, 74.5
    X = np.stack([rng.uniform(-8, 8, m), rng.uniform(-2, 2, m), rng.uniform(5, 25, m)], 1).astype(np.float32)
    R = P.rodrigues(np.array([0.01, -0.03, 0.005]))
    t = np.array([0.03, -0.01, -0.2])
    uv, _ = P.project(R, t, X.astype(np.float64), f, cx, cy)
    uv += rng.normal(0, 0.3, uv.shape)
    out = rng.choice(m, n_out, replace=False)
    uv[out] += rng.normal(0, 40, (n_out, 2))
    kp2 = np.zeros((m, 6), np.float32)
    kp2[:, :2] = uv
    matches = np.stack([np.arange(m), rng.permutation(m), np.zeros(m)], 1).astype(np.int32)
    kp2p = np.zeros_like(kp2)
    kp2p[matches[:, 1]] = kp2           # keypoint of match i sits at row matches[i][1]
    Q = np.array([[1, 0, 0, -cx], [0, 1, 0, -cy], [0, 0, 0, f], [0, 0, 1 / 0.537, 0.0]])
    return X, kp2p, matches, Q, (f, cx, cy), (R, t)


def test_pnp_ransac_kernels(emu):
    """opt-in P3P-RANSAC + LM vs its numpy specification: same winning hypothesis, same inlier count,
    pose within the north-star tolerance (1e-4 rad, 1e-3 relative translation)."""
    from oracle import pnp_restate as P
    X, kp2, matches, Q, (f, cx, cy), _ = _pnp_case()
    W, H = 300, 150
    c = Ctx(emu, W, H, sgbm_params(32), (0, 0, W, H), Q, 600)
    m = len(X)
    pts = np.zeros((c.cap, 3), np.float32)
    pts[:m] = X
    mt = np.zeros((c.cap, 3), np.int32)
    mt[:m] = matches
    kp = np.zeros((c.cap, 6), np.float32)
    kp[:m] = kp2
    cnt = np.array([m], np.int32)
    out = np.zeros(16)
    iters, seed = 64, 3
    N.check(emu, emu.ovo_pnp_ransac(c.ctx, N.ptr(pts), N.ptr(mt), N.ptr(kp), N.ptr(cnt), c.cap, iters, 8.0, seed, N.ptr(out), None))
    uv = kp2[matches[:, 1], :2].astype(np.float64)
    ref = P.pnp_ransac(X.astype(np.float64), uv, f, cx, cy, iters=iters, thr=8.0, seed=seed)
    assert int(out[15]) == ref["best"] and int(out[12]) == ref["n_inliers"] > 300
    Rg, tg = out[:12].reshape(3, 4)[:, :3], out[:12].reshape(3, 4)[:, 3]
    dR = Rg @ ref["R"].T
    assert np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)) < 1e-4
    assert np.linalg.norm(tg - ref["t"]) < 1e-3 * np.linalg.norm(ref["t"])
