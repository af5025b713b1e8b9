"""Probe: the single_sequence record of bench.py, repeated (run under gpurun)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from openvo_b200 import StereoCamera, synth
cfg = bench.CONFIGS["K"]
L, R = bench.make_frames(cfg)
cam = StereoCamera(**synth.camera_args(cfg["W"], cfg["H"], cfg["D"]))
pin_L = [torch.from_numpy(L[i]).pin_memory() for i in range(len(L))]
pin_R = [torch.from_numpy(R[i]).pin_memory() for i in range(len(R))]
for rep in range(4):
    r = bench.single_sequence(cam, cfg, pin_L, pin_R, L, R)
    print(rep, round(r["value"], 1), round(r["streaming_update_frames_per_s"], 1), r["identical_to_streaming"], flush=True)
