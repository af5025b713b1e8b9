"""Host-side profile of one bench thread (8 KITTI sequences, one stream): cProfile over N steps, so that the time spent in Python
(per-sequence state machine, tensor slicing, ctypes marshalling) can be told from the time blocked in the library."""
import cProfile, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from openvo_b200 import StereoCamera, synth
from openvo_b200.batch import BatchOdometer

cfg = bench.CONFIGS["K"]
L, R = bench.make_frames(cfg)
cam = StereoCamera(**synth.camera_args(cfg["W"], cfg["H"], cfg["D"]))
SP = 24
bo = BatchOdometer(cam, SP, nfeatures=cfg["n"], engine_tag=0, preprocessed_frames=True)
dev_L, dev_R = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()


def step(s):
    idx = [bench.frame_index(s, q) for q in range(SP)]
    ti = torch.tensor(idx, device="cuda")
    return bo.update_device(dev_L[ti], dev_R[ti])


for s in range(4):
    step(s)
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 30
pr = cProfile.Profile()
pr.enable()
for s in range(4, 4 + N):
    step(s)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print("ms per step (8 frames, 1 thread): %.3f" % (1e3 * dt / N))
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
