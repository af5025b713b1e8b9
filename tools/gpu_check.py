"""Development probe for a GPU box: every seam through the C ABI vs the oracle, with per-stage diagnostics and timings.
Writes gpurun_out/gpu_check.json.  (The formal parity tests live in tests/; this script prints more.)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openvo_b200 import StereoCamera, StereoOdometer, synth  # noqa: E402
from oracle import openvo_port as O  # noqa: E402

OUT = {}


def timed(fn, n=5):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / n


def check_sgbm(W, H, D, tag, occl=True, **kw):
    args = synth.camera_args(W, H, D)
    args["sgbm_params"].update(kw)
    cam = StereoCamera(**args)
    L, R = synth.kat_pair(W, H, d=min(24, D // 2))
    if occl:
        rng = np.random.default_rng(1)
        R = R.copy()
        R[H // 3:H // 2, W // 3:W // 2] = rng.integers(0, 256, (H // 2 - H // 3, W // 2 - W // 3))
    eng = cam.engine()
    l, r = eng.upload(L[None], "l"), eng.upload(R[None], "r")
    got = eng.sgbm(l, r)[0].cpu().numpy()
    t0 = time.time()
    ref = O.sgbm_compute(L, R, args["sgbm_params"])
    t_or = time.time() - t0
    bad = int((got != ref).sum())
    ms = timed(lambda: eng.sgbm(l, r))
    print("SGBM %-10s %dx%d D=%d %s mismatches=%d/%d valid=%.3f gpu=%.3f ms oracle=%.1fs" % (tag, W, H, D, kw, bad, ref.size, (ref >= 0).mean(), ms, t_or), flush=True)
    if bad:
        ys, xs = np.nonzero(got != ref)
        print("   first:", list(zip(ys[:6].tolist(), xs[:6].tolist())), got[ys[:6], xs[:6]], ref[ys[:6], xs[:6]])
    OUT["sgbm_" + tag] = dict(W=W, H=H, D=D, mismatches=bad, ms=ms)
    return bad == 0


def check_orb(W, H, n, tag, usemask=True):
    args = synth.camera_args(W, H, 16)
    cam = StereoCamera(**args)
    eng = cam.engine(nfeatures=n)
    L, _ = synth.kat_pair(W, H)
    img = np.ascontiguousarray(L[:eng.ch, :eng.cw])
    mask = None
    if usemask:
        rng = np.random.default_rng(5)
        m = rng.integers(0, 2, (eng.ch // 8 + 1, eng.cw // 8 + 1)).astype(np.uint8)
        mask = np.ascontiguousarray((np.kron(m, np.ones((8, 8), np.uint8))[:eng.ch, :eng.cw] * 255).astype(np.uint8))
    di = eng.upload(img[None], "i")
    dm = eng.upload(mask[None], "m") if usemask else None
    kp, desc, cnt = eng.orb(di, dm)
    k = cnt[0]
    kp, desc = kp[0, :k].cpu().numpy(), desc[0, :k].cpu().numpy()
    rk, rd = O.orb_detect_compute(img, mask, n)
    ok = k == len(rk) and np.array_equal(kp, rk) and np.array_equal(desc, rd)
    ms = timed(lambda: eng.orb(di, dm))
    print("ORB  %-10s %dx%d n=%d mask=%s got=%d ref=%d %s gpu=%.3f ms" % (tag, eng.cw, eng.ch, n, usemask, k, len(rk), "EXACT" if ok else "MISMATCH", ms), flush=True)
    if not ok and k == len(rk):
        print("   kp field mismatches", (kp != rk).sum(0), "desc rows", int((desc != rd).any(1).sum()))
        bad = np.nonzero((kp != rk).any(1))[0][:5]
        for b in bad:
            print("   ", b, kp[b], rk[b])
    OUT["orb_" + tag] = dict(W=eng.cw, H=eng.ch, n=n, ok=bool(ok), ms=ms)
    return ok, (kp, desc)


def check_knn(nq, nt, tag):
    cam = StereoCamera(**synth.camera_args(640, 200, 16))
    eng = cam.engine(nfeatures=max(nq, nt))
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8) & 0xF0
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[nt // 2:] &= 0xF0
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    nn = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    eng.knn2(dq, nq, dt, nt, out=nn)
    ref = O.knn2_hamming(q, t)
    ok = np.array_equal(nn.cpu().numpy(), ref)
    ms = timed(lambda: eng.knn2(dq, nq, dt, nt, out=nn))
    print("KNN  %-10s %dx%d %s ties=%d gpu=%.3f ms" % (tag, nq, nt, "EXACT" if ok else "MISMATCH", int((ref[:, 1] == ref[:, 3]).sum()), ms), flush=True)
    OUT["knn_" + tag] = dict(nq=nq, nt=nt, ok=bool(ok), ms=ms)
    return ok


def check_sequence(W, H, D, n, nframes, tag):
    Ls, Rs, _ = synth.make_sequence(W, H, nframes)
    args = synth.camera_args(W, H, D)
    cam = StereoCamera(**args)
    od = StereoOdometer(cam, nfeatures=n, preprocessed_frames=True)
    pc = O.StereoCameraPort(**args, backend="cv2")
    po = O.StereoOdometerPort(pc, nfeatures=n, preprocessed_frames=True)
    ok_all = True
    for i in range(nframes):
        t0 = time.time()
        a = od.update(Ls[i], Rs[i])
        torch.cuda.synchronize()
        t1 = time.time()
        b = po.update(Ls[i], Rs[i])
        t2 = time.time()
        dT = float(np.abs(od.c_T_w - po.c_T_w).max())
        same_kp = od._cur is not None and po.cur is not None and od._cur.n_kp == len(po.cur[3]) and \
            np.array_equal(od._host(od._cur, "kp_array"), po.cur[3]) and np.array_equal(od.current_desc, po.cur[4])
        same_disp = po.cur is not None and np.array_equal(od.current_disparity, po.cur[1])
        ok = a == b and od.skip_cause == po.skip_cause and dT < 1e-9 and same_kp and same_disp
        ok_all &= bool(ok)
        print("SEQ  %-8s frame %d ok=%s/%s cause=%r/%r |dT|=%.2e kp=%s disp=%s matches=%d gpu=%.1f ms cpu=%.0f ms" % (
            tag, i, a, b, od.skip_cause, po.skip_cause, dT, same_kp, same_disp, od.last_match_count, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
    OUT["seq_" + tag] = dict(ok=ok_all)
    return ok_all


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    quick = "--quick" in sys.argv
    if "--only4k" in sys.argv:
        check_sgbm(3840, 2160, 256, "U", occl=False)
        check_orb(3840, 2160, 10000, "U")
        check_knn(10000, 10000, "U")
        return
    check_sgbm(200, 60, 32, "small32")
    check_sgbm(320, 48, 64, "small64", blockSize=3, P1=72, P2=288, uniquenessRatio=0, speckleWindowSize=0)
    check_sgbm(240, 64, 48, "pad48", blockSize=7, preFilterCap=15, uniquenessRatio=15, disp12MaxDiff=2, speckleWindowSize=50, speckleRange=1)
    check_sgbm(400, 100, 256, "d256")
    check_orb(640, 200, 500, "small", usemask=False)
    check_orb(415, 333, 300, "mask")
    check_knn(300, 257, "small")
    check_knn(2000, 2000, "K")
    check_sequence(640, 200, 64, 500, 4, "small")
    if not quick:
        check_sgbm(1241, 376, 128, "K")
        check_orb(1241, 376, 2000, "K")
        check_sequence(1241, 376, 128, 2000, 4, "K")
        check_sgbm(1920, 1080, 256, "F", occl=False)
        check_orb(1920, 1080, 5000, "F")
        check_knn(5000, 5000, "F")
    if "--4k" in sys.argv:
        check_sgbm(3840, 2160, 256, "U", occl=False)
        check_orb(3840, 2160, 10000, "U")
        check_knn(10000, 10000, "U")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as fh:
        json.dump(OUT, fh, indent=1)


if __name__ == "__main__":
    main()
