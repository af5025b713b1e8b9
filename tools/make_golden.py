"""Generate tests/golden/* by running the UNMODIFIED reference (imported from /root/reference/src) in this container.

The reference ships no tests or golden vectors; these fixtures — outputs of the reference itself on
seeded synthetic inputs, with cv2 4.13.0 / numpy 2.3 — are what pins the oracle and the CUDA path.  /root/reference does
not exist on the GPU box, so the fixtures (small) are committed together with this script.

    python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
import cv2  # noqa: E402
import openVO  # noqa: E402  (the unmodified reference)
from openvo_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def kp_array(kps):
    return np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32).reshape(-1, 6)


def run_sequence(name, W, H, D, n, lefts, rights, distorted=False, **od_kw):
    args = (synth.camera_args_distorted if distorted else synth.camera_args)(W, H, D)
    cam = openVO.StereoCamera(**args)
    od_kw.setdefault("preprocessed_frames", True)
    od = openVO.StereoOdometer(cam, nfeatures=n, **od_kw)
    rec = dict(roi=np.array(cam.valid_region_left), Q=cam.Q, W=W, H=H, D=D, nfeatures=n, distorted=np.array(distorted),
               preprocessed=np.array(od_kw["preprocessed_frames"]))
    for i in range(len(lefts)):
        ok = od.update(lefts[i], rights[i])
        rec["ok_%d" % i] = np.array(ok)
        rec["cause_%d" % i] = np.array(od.skip_cause)
        rec["skipped_%d" % i] = np.array(od.skipped_frames)
        rec["cTw_%d" % i] = od.c_T_w.copy()
        rec["pose_%d" % i] = od.current_pose()
        if od.current_disparity is not None:
            rec["disp16_%d" % i] = np.rint(od.current_disparity * 16).astype(np.int16)
            rec["kp_%d" % i] = kp_array(od.current_kps)
            rec["desc_%d" % i] = od.current_desc.copy()
            rec["xyz_sha_%d" % i] = np.array(sha(od.current_3d))
    rec["left"], rec["right"] = lefts, rights
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(name, "frames", len(lefts), [str(rec["cause_%d" % i]) for i in range(len(lefts))],
          [bool(rec["ok_%d" % i]) for i in range(len(lefts))])


def update_kat(W_, H_, n_, D_):
    """Appendix B last row: two update() calls of the unmodified reference, frame 2 = both images rolled by 3 px."""
    L, R = synth.kat_pair(W_, H_)
    cam = openVO.StereoCamera(**synth.camera_args(W_, H_, D_))
    od = openVO.StereoOdometer(cam, nfeatures=n_, preprocessed_frames=True)
    ok = [bool(od.update(L, R)), bool(od.update(np.roll(L, 3, axis=1), np.roll(R, 3, axis=1)))]
    return dict(update_ok=ok, roi=[int(v) for v in cam.valid_region_left], cTw=sha(od.c_T_w),
                pose_t=[float(v) for v in od.current_pose()[:3, 3]], cTw_rows=[[float(v) for v in row] for row in od.c_T_w])


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--kat-only" not in sys.argv:
        fixtures()
    kats()


def fixtures():
    # 1. a small clean sequence
    W, H, D, n = 480, 160, 64, 400
    Ls, Rs, _ = synth.make_sequence(W, H, 4)
    run_sequence("seq_small", W, H, D, n, Ls, Rs)
    # 2. a sequence that exercises the skip / fall-back state machine (B4): a texture-less frame ("keypoints"), a jump
    #    ("bigdist"), then recovery against the last committed frame
    Ls2, Rs2, _ = synth.make_sequence(W, H, 6)
    far_l, far_r, _ = synth.make_sequence(W, H, 2, step=(0.0, 0.0, 2.5))       # 2.5 m jump -> "bigdist"
    oth_l, oth_r, _ = synth.make_sequence(W, H, 1, seed=777)                   # unrelated scene -> "matches"
    blank = np.full_like(Ls2[0], 90)                                           # no corners -> "keypoints"
    lefts = np.stack([Ls2[0], Ls2[1], blank, Ls2[2], oth_l[0], Ls2[3], far_l[1], Ls2[4], Ls2[5]])
    rights = np.stack([Rs2[0], Rs2[1], blank, Rs2[2], oth_r[0], Rs2[3], far_r[1], Rs2[4], Rs2[5]])
    run_sequence("seq_skip", W, H, D, n, lefts, rights)
    # 3. optional filters on
    run_sequence("seq_filters", W, H, D, n, Ls, Rs, rigidity_threshold=0.06, outlier_threshold=0.02)
    # 3b. the reference's DEFAULT path: colour input + cv2.remap rectification (preprocessed_frames=False), distorted rig
    Lc = np.stack([synth.to_bgr(Ls[i], seed=i) for i in range(3)])
    Rc = np.stack([synth.to_bgr(Rs[i], seed=100 + i) for i in range(3)])
    run_sequence("seq_rectify", W, H, D, n, Lc, Rc, distorted=True, preprocessed_frames=False)
    # 4. per-seam vectors on a KAT pair incl. matcher output
    L, R = synth.kat_pair(W, H, d=12)
    sg = cv2.StereoSGBM_create(0, D, 5, 200, 800, 1, 63, 10, 100, 2).compute(L, R)
    orb = cv2.ORB_create(nfeatures=n)
    k1, d1 = orb.detectAndCompute(L, None)
    k2, d2 = orb.detectAndCompute(R, None)
    mm = cv2.BFMatcher.create(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2)
    nn = np.array([[m[0].trainIdx, int(m[0].distance), m[1].trainIdx, int(m[1].distance)] for m in mm], np.int32)
    np.savez_compressed(os.path.join(GOLD, "seams_small.npz"), left=L, right=R, sgbm=sg, kp1=kp_array(k1), desc1=d1,
                        kp2=kp_array(k2), desc2=d2, nn=nn)


def kats():
    # 5. known-answer hashes at the BASELINE shapes — recomputed here, compared with the survey's
    kat = {}
    for tag, (W_, H_, n_, D_) in dict(K=(1241, 376, 2000, 128), F=(1920, 1080, 5000, 256), U=(3840, 2160, 10000, 256)).items():
        L, R = synth.kat_pair(W_, H_)
        sg = cv2.StereoSGBM_create(0, D_, 5, 200, 800, 1, 63, 10, 100, 2).compute(L, R)
        orb = cv2.ORB_create(nfeatures=n_)
        k1, d1 = orb.detectAndCompute(L, None)
        k2, d2 = orb.detectAndCompute(R, None)
        mm = cv2.BFMatcher.create(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2)
        nn = np.array([[m[0].trainIdx, int(m[0].distance), m[1].trainIdx, int(m[1].distance)] for m in mm], np.int32)
        kat[tag] = dict(W=W_, H=H_, n=n_, D=D_, left=sha(L), right=sha(R), sgbm=sha(sg), kpL=sha(kp_array(k1)), descL=sha(d1),
                        kpR=sha(kp_array(k2)), descR=sha(d2), knn=sha(nn),
                        ratio_pass=int(sum(1 for m in mm if m[0].distance < 0.8 * m[1].distance)))
        kat[tag].update(update_kat(W_, H_, n_, D_))
        print(tag, kat[tag])
    survey = dict(K=dict(left="aa22cdb96cabde5d", right="1f355f21111242db", sgbm="194bde87fbd3fa38", kpL="3b45536482e8160b",
                         descL="6db71f5fe13b6209", kpR="0ec841d23b440a4d", descR="7b4eb48af6bb80ac", knn="8245e9a332564a00",
                         ratio_pass=1420),
                  F=dict(left="244befc2e79231dd", right="8a528ddd997c7255", sgbm="87d752db476c997f", kpL="8721f81f69b9aeef",
                         descL="5b25a7e84dd19004", kpR="9f6ca90f4318da37", descR="09a93c9152a01c02", knn="75d50723b36f6021",
                         ratio_pass=3454),
                  U=dict(left="bb481aff9dd6d734", right="4dd8dd07d16800c9", sgbm="9b3955d3537c6f3c", kpL="a44f2f647e432d36",
                         descL="cafa72b97bebba6a", kpR="76801604156db0a1", descR="00a68eeae5e5de65", knn="7d728f84b553f18d",
                         ratio_pass=6782))
    survey["K"].update(update_ok=[True, True], roi=[0, 0, 1240, 375], cTw="c99217ea23eff2d6")
    survey["F"].update(update_ok=[True, True], roi=[1, 0, 1919, 1079], cTw="0d5ea186a2c4e514")
    survey["U"].update(update_ok=[True, True], roi=[0, 0, 3839, 2159], cTw="95ae0e9000a08e61")
    advisory = ("roi", "cTw")  # the survey itself calls the pose hashes advisory (f64 SVD / BLAS); the seam hashes are the contract
    for tag in survey:
        for k, v in survey[tag].items():
            if k in advisory and kat[tag][k] != v:
                print("note: %s.%s = %s here, SURVEY.md Appendix B has %s (advisory row)" % (tag, k, kat[tag][k], v))
                continue
            assert kat[tag][k] == v, ("KAT differs from SURVEY.md Appendix B", tag, k, kat[tag][k], v)
    with open(os.path.join(GOLD, "kat.json"), "w") as fh:
        json.dump(kat, fh, indent=1)
    print("KATs agree with SURVEY.md Appendix B")


if __name__ == "__main__":
    main()
