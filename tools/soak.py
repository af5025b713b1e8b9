"""Determinism soak on the GPU: the same batch through SGBM / ORB many times on two concurrent streams must give identical
bytes every time (races in the kernels would show up as run-to-run differences), and must equal the oracle on frame 0."""
import sys, hashlib
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from openvo_b200 import StereoCamera, synth
from openvo_b200.batch import BatchOdometer
from oracle import openvo_port as O

cfg = bench.CONFIGS["K"]
L, R = bench.make_frames(cfg)
cam_args = synth.camera_args(cfg["W"], cfg["H"], cfg["D"])
cam = StereoCamera(**cam_args)
NB, REP = 12, int(sys.argv[1]) if len(sys.argv) > 1 else 20
streams = [torch.cuda.Stream() for _ in range(2)]
bos = [BatchOdometer(cam, NB, nfeatures=cfg["n"], engine_tag=50 + g, preprocessed_frames=True) for g in range(2)]
idx = [i % len(L) for i in range(NB)]
dl, dr = torch.from_numpy(L[idx]).cuda(), torch.from_numpy(R[idx]).cuda()
ref = None
for rep in range(REP):
    outs = []
    for g in range(2):
        with torch.cuda.stream(streams[g]):
            eng = bos[g].engine
            fr = eng.frames(dl, dr)
            outs.append(fr)
    torch.cuda.synchronize()
    for g in range(2):
        h = hashlib.sha1()
        for f in outs[g]:
            h.update(f.disp.cpu().numpy().tobytes())
            h.update(f.kp[:f.n_kp].cpu().numpy().tobytes())
            h.update(f.desc[:f.n_kp].cpu().numpy().tobytes())
        if ref is None:
            ref = h.hexdigest()
        assert h.hexdigest() == ref, ("nondeterministic", rep, g)
print("identical over", REP, "x 2 streams:", ref)
# frame 0 vs the oracle (cv2 back end)
port = O.StereoCameraPort(**cam_args, backend="cv2")
xyz, disp, left = port.compute_3d(L[idx[0]], R[idx[0]], preprocessed=True)
f0 = outs[0][0]
assert np.array_equal(f0.disp.cpu().numpy(), disp), "disparity differs from the oracle"
print("frame 0 disparity == oracle")
