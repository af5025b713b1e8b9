#!/usr/bin/env python
"""bench.py — stereo frames/sec of the openVO per-frame hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path on the host cores

A "step" advances S = 72 independent KITTI-shaped synthetic stereo sequences by one frame each on every GPU (three BatchOdometers
of 24 sequences in flight on three streams, driven by one host thread:
SGBM 128 disp + ORB 2000 kp + Hamming 2-NN/ratio + fused 3-D lookup + Umeyama + the reference's skip state machine), so a
step is S frames per GPU (weak scaling: S per GPU is fixed).  `value` is measured with the frames already resident in
HBM; `e2e` is the same loop through the public host-buffer API (numpy frames in pinned memory -> H2D every step, poses
read back every step).  Timing: CUDA events around the K steps, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "K": dict(W=1241, H=376, D=128, n=2000, seqs=72, name="KITTI-shaped 1241x376 stereo, ORB 2000 kp, StereoSGBM 128 disp"),
    "F": dict(W=1920, H=1080, D=256, n=5000, seqs=36, name="1920x1080 stereo, ORB 5000 kp, StereoSGBM 256 disp"),
    "U": dict(W=3840, H=2160, D=256, n=10000, seqs=9, name="3840x2160 stereo, ORB 10000 kp, StereoSGBM 256 disp"),
    "S": dict(W=640, H=200, D=64, n=500, seqs=72, name="small 640x200 stereo, ORB 500 kp, StereoSGBM 64 disp (dev only)"),
}
N_DISTINCT = 6  # distinct rendered frames; sequences ping-pong through them with different phases


def frame_index(step, seq):
    period = 2 * (N_DISTINCT - 1)
    k = (step + seq) % period
    return k if k < N_DISTINCT else period - k


def make_frames(cfg):
    from openvo_b200 import synth
    cache = os.path.join(tempfile.gettempdir(), "ovo_bench_%dx%d_%d.npz" % (cfg["W"], cfg["H"], N_DISTINCT))
    if os.path.exists(cache):
        z = np.load(cache)
        return z["L"], z["R"]
    L, R, _ = synth.make_sequence(cfg["W"], cfg["H"], N_DISTINCT)
    try:
        np.savez(cache, L=L, R=R)
    except Exception:
        pass
    return L, R


# ---------------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port with the cv2 back end (what a user of the reference runs), one process per host core
# ---------------------------------------------------------------------------------------------------------------------------
def cpu_worker(cfg_key, frames_file, seq, warmup, steps):
    import cv2
    from openvo_b200 import synth
    from oracle import openvo_port as O
    cv2.setNumThreads(1)
    cfg = CONFIGS[cfg_key]
    z = np.load(frames_file)
    L, R = z["L"], z["R"]
    args = synth.camera_args(cfg["W"], cfg["H"], cfg["D"])
    od = O.StereoOdometerPort(O.StereoCameraPort(**args, backend="cv2"), nfeatures=cfg["n"], preprocessed_frames=True)
    for s in range(warmup):
        od.update(L[frame_index(s, seq)], R[frame_index(s, seq)])
    sys.stdout.write("READY\n")
    sys.stdout.flush()
    sys.stdin.readline()  # start gun
    t0 = time.time()
    for s in range(warmup, warmup + steps):
        od.update(L[frame_index(s, seq)], R[frame_index(s, seq)])
    sys.stdout.write("DONE %.6f\n" % (time.time() - t0))
    sys.stdout.flush()


def run_cpu(cfg_key, L, R, warmup, steps, nprocs):
    """-> (frames/s aggregate, wall seconds).  All processes start their timed frames together."""
    with tempfile.NamedTemporaryFile(suffix=".npz", delete=False) as fh:
        np.savez(fh, L=L, R=R)
        frames_file = fh.name
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--_cpu_worker", cfg_key, frames_file, str(i), str(warmup),
                               str(steps)], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env) for i in range(nprocs)]
    for p in procs:
        assert p.stdout.readline().strip() == "READY"
    t0 = time.time()
    for p in procs:
        p.stdin.write("go\n")
        p.stdin.flush()
    for p in procs:
        line = p.stdout.readline()
        assert line.startswith("DONE"), line
    wall = time.time() - t0
    for p in procs:
        p.wait()
    os.unlink(frames_file)
    return nprocs * steps / wall, wall


def cpu_profile_worker(cfg_key, frames_file, nframes):
    """Single-stream figure (what a user of the reference gets: one process, cv2's default thread pool) and per-stage medians
    of the reference's CPU path: each cv2 call of update() timed in place.  Prints one JSON line."""
    import cv2
    from openvo_b200 import synth
    from oracle import openvo_port as O
    cfg = CONFIGS[cfg_key]
    z = np.load(frames_file)
    L, R = z["L"], z["R"]
    args = synth.camera_args(cfg["W"], cfg["H"], cfg["D"])
    od = O.StereoOdometerPort(O.StereoCameraPort(**args, backend="cv2"), nfeatures=cfg["n"], preprocessed_frames=True)
    stages = {}

    def timed(obj, name, tag):
        fn = getattr(obj, name)

        def wrap(*a, **k):
            t0 = time.perf_counter()
            out = fn(*a, **k)
            stages.setdefault(tag, []).append(time.perf_counter() - t0)
            return out
        setattr(obj, name, wrap)
    timed(od.stereo.be, "disparity", "sgbm")          # ref: stereo_camera.py:51
    timed(od.stereo.be, "reproject", "reproject")     # ref: stereo_camera.py:52
    timed(od.be, "features", "orb")                   # ref: stereo_odometer.py:117
    timed(od.be, "knn2", "knn")                       # ref: stereo_odometer.py:163
    timed(od.be, "rigid", "umeyama")                  # ref: stereo_odometer.py:190,204
    pc = od.point_clouds

    def point_clouds(*a, **k):                        # ref: stereo_odometer.py:162-175 (knn + ratio test + the bilinear loop)
        t0 = time.perf_counter()
        out = pc(*a, **k)
        stages.setdefault("point_clouds_total", []).append(time.perf_counter() - t0)
        return out
    od.point_clouds = point_clouds
    for s in range(2):
        od.update(L[frame_index(s, 0)], R[frame_index(s, 0)])
    for v in stages.values():
        v.clear()
    t0 = time.perf_counter()
    for s in range(2, 2 + nframes):
        od.update(L[frame_index(s, 0)], R[frame_index(s, 0)])
    wall = time.perf_counter() - t0
    med = {k: 1e3 * float(np.median(v)) for k, v in stages.items() if v}
    if "point_clouds_total" in med and "knn" in med:
        med["bilinear_loop"] = med.pop("point_clouds_total") - med["knn"]   # ref: stereo_odometer.py:172-174
    print(json.dumps({"single_stream_fps": nframes / wall, "cv2_threads": cv2.getNumThreads(), "frames": nframes, "stages_ms": med}))


def run_cpu_profile(cfg_key, L, R, nframes):
    with tempfile.NamedTemporaryFile(suffix=".npz", delete=False) as fh:
        np.savez(fh, L=L, R=R)
        frames_file = fh.name
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--_cpu_profile", cfg_key, frames_file, str(nframes)],
                             capture_output=True, text=True, timeout=600)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}
    finally:
        os.unlink(frames_file)


def single_sequence(cam, cfg, pin_L, pin_R, L, R, n_frames=200):
    """BASELINE configs 1/2: ONE 200-frame KITTI-shaped sequence on one GPU.  `value`: frame-parallel chunks
    (openvo_b200.sequence.SequenceOdometer.run: host frames in pinned memory -> poses), identical results to the streaming
    update() loop, which is timed next to it (the latency-bound way a live camera would drive the odometer)."""
    import torch
    from openvo_b200 import StereoOdometer
    from openvo_b200.sequence import SequenceOdometer
    idx = [frame_index(s, 0) for s in range(n_frames)]
    lefts, rights = [pin_L[i] for i in idx], [pin_R[i] for i in idx]
    od = SequenceOdometer(cam, chunk=24, nfeatures=cfg["n"], preprocessed_frames=True)
    od.run(lefts, rights)  # warm-up: the whole sequence once (allocator, pinned staging, first launches)
    runs = []
    for _ in range(3):
        od = SequenceOdometer(cam, chunk=24, nfeatures=cfg["n"], preprocessed_frames=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        oks = od.run(lefts, rights)
        torch.cuda.synchronize()
        runs.append(n_frames / (time.perf_counter() - t0))
    dt = n_frames / float(np.median(runs))
    trace = getattr(od, "trace", None)
    st = StereoOdometer(cam, nfeatures=cfg["n"], preprocessed_frames=True, _engine_tag=998)
    for k in range(3):
        st.update(L[idx[k]], R[idx[k]])
    st = StereoOdometer(cam, nfeatures=cfg["n"], preprocessed_frames=True, _engine_tag=998)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    want = [st.update(L[i], R[i]) for i in idx]
    torch.cuda.synchronize()
    dts = time.perf_counter() - t0
    return {"workload": "one synthetic %s sequence of %d frames, host frames in, poses out" % (cfg["name"], n_frames),
            "value": n_frames / dt, "unit": "frames/s", "runs": runs, "chunk": 24, "frames_committed": int(sum(oks)),
            **({"host_ms_per_chunk_last_run(begin, extract+select, pair, replay)": [[round(1e3 * v, 2) for v in row] for row in trace]} if trace else {}),
            "streaming_update_frames_per_s": n_frames / dts, "streaming_ms_per_frame": 1e3 * dts / n_frames,
            "identical_to_streaming": bool(oks == want and np.array_equal(od.c_T_w, st.c_T_w) and od.skip_cause == st.skip_cause)}


def frame_sharded_4k(rank, world, n_frames=16, n_render=4):
    """BASELINE config 4: ONE 3840x2160 / ORB 10000 / 256-disparity sequence sharded by contiguous frame chunk over the ranks
    (openvo_b200.dist.run_frame_chunk: one-frame halo, no data-path collective), per-frame transforms all-gathered once over
    NCCL, chain replayed; rank 0 also runs the sequence sequentially and the two chains are compared."""
    import torch
    import torch.distributed as dist
    from openvo_b200 import StereoCamera, StereoOdometer, synth
    from openvo_b200 import dist as odist
    cfg = CONFIGS["U"]
    Lr, Rr, _ = synth.make_sequence(cfg["W"], cfg["H"], n_render)   # every rank renders the same deterministic frames
    period = 2 * (n_render - 1)
    idx = [(k % period) if (k % period) < n_render else period - (k % period) for k in range(n_frames)]
    lefts, rights = [Lr[i] for i in idx], [Rr[i] for i in idx]
    cam = StereoCamera(**synth.camera_args(cfg["W"], cfg["H"], cfg["D"]))
    od = StereoOdometer(cam, nfeatures=cfg["n"], preprocessed_frames=True)
    od.update(lefts[0], rights[0])          # warm-up (allocations, first launches), then a fresh state machine
    od = StereoOdometer(cam, nfeatures=cfg["n"], preprocessed_frames=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    start, T, st = odist.run_frame_chunk(od, lefts, rights, rank, world)
    Tall, sall = odist.gather_frame_chunks(start, T, st, n_frames)
    chain = odist.replay_chains(Tall[None], sall[None])[0]
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    rec = None
    if rank == 0:
        seq = StereoOdometer(cam, nfeatures=cfg["n"], preprocessed_frames=True)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        oks = [seq.update(lefts[k], rights[k]) for k in range(n_frames)]
        s1.record()
        torch.cuda.synchronize()
        dR = chain[:3, :3] @ seq.c_T_w[:3, :3].T
        ang = float(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
        dt = float(np.linalg.norm(chain[:3, 3] - seq.c_T_w[:3, 3]))
        rec = {"workload": "one synthetic %s sequence of %d frames, sharded by contiguous frame chunk (one-frame halo)" % (cfg["name"], n_frames),
               "value": n_frames / (ms * 1e-3), "unit": "frames/s", "ms": ms, "n_gpus": world,
               "sequential_1gpu_frames_per_s": n_frames / (s0.elapsed_time(s1) * 1e-3),
               "frames_committed": int(sum(1 for v in sall[1:] if v)) + 1, "frames": n_frames, "sequential_committed": int(sum(oks)),
               "chain_equals_sequential": bool(np.array_equal(chain, seq.c_T_w)),
               "chain_vs_sequential": {"rotation_rad": ang, "translation": dt}}
    dist.barrier()
    return rec


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 128))  # one single-threaded reference process per core; capped to bound memory on very wide hosts


# ---------------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def algorithmic_bytes(tag, eng, nb, fused=True):
    """ALGORITHMIC HBM bytes of one launch of the named kernel (DESIGN.md 'Kernels'); V = one frame's cost volume."""
    D = eng.cfg.sgbm.numDisparities
    Dp = 64 if D <= 64 else (128 if D <= 128 else 256)
    W, H = eng.W, eng.H
    V = H * (W - D) * Dp * 2
    bands = -(-H // (8 if Dp == 256 else 16))  # launches of the fused vertical kernel per batch (one per band of rows)
    table = {
        "k_sgbm_prep": 2 * W * H + 16 * W * H,
        "k_sgbm_cost_t": 16 * W * H + V,
        "k_sgbm_vert_t": V + 3 * V,                                          # unfused (OVO_SGBM_FUSED=0 / MODE_HH): C in, three volumes out
        "k_sgbm_vsum_t": 2.0 * V / bands,                                    # fused: C in, Sv out, one band per launch
        "k_sgbm_horiz_t": (2 * V if fused else 4 * V) + 2 * W * H,           # C + the vertical sums in, disparity out
    }
    return table.get(tag, 0) * nb


def source_sha1():
    """sha1 of the SGBM kernel source: stamps profiles/traffic.json, so a captured DRAM figure is only quoted for the code it
    was captured from (the GPU box has no .git)."""
    import hashlib
    with open(os.path.join(ROOT, "openvo_b200", "csrc", "sgbm.cu"), "rb") as fh:
        return hashlib.sha1(fh.read()).hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="K", choices=list(CONFIGS))
    ap.add_argument("--seqs", type=int, default=0, help="independent sequences (= frames per step) per GPU; 0 = the config's default "
                                                        "(72 at the KITTI shape: three batches of 24 in flight)")
    ap.add_argument("--select-threads", type=int, default=0,
                    help="host threads per batch for the keypoint-selection step (OVO_SELECT_THREADS); 0 = min(8, cores / ranks)")
    ap.add_argument("--threads", type=int, default=1, help="host threads (each drives --groups batches round-robin)")
    ap.add_argument("--groups", type=int, default=3,
                    help="batches in flight per host thread: each has its own CUDA stream / workspace and a share of the sequences; the "
                         "thread finishes batch g of step s, queues batch g of step s+1 and moves on to g+1, so the host-side "
                         "keypoint selection of one batch overlaps device work of the others")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra records of the default run (other_configs: short F and U runs; frame-sharded 4K at N > 1)")
    ap.add_argument("--_cpu_worker", nargs=5, default=None)
    ap.add_argument("--_cpu_profile", nargs=3, default=None)
    a = ap.parse_args()
    if a._cpu_worker:
        k, f, seq, wu, st = a._cpu_worker
        return cpu_worker(k, f, int(seq), int(wu), int(st))
    if a._cpu_profile:
        k, f, n = a._cpu_profile
        return cpu_profile_worker(k, f, int(n))
    cfg = CONFIGS[a.config]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    workload = "synthetic %s, independent sequences advanced 1 frame each per step" % cfg["name"]
    metric, unit = "stereo frames/sec", "frames/s"
    # `config` is identical in both arms (the driver compares it); per-arm run details go to `run`
    config = {"workload": workload, "width": cfg["W"], "height": cfg["H"], "nfeatures": cfg["n"], "num_disparities": cfg["D"],
              "distinct_frames": N_DISTINCT}

    L, R = make_frames(cfg)

    if a.impl == "reference":
        if rank != 0:
            return 0
        cores = host_cores()
        fps, wall = run_cpu(a.config, L, R, max(a.warmup, 1), a.steps, cores)
        sample = "%d processes x %d frames each (cv2.setNumThreads(1)), wall %.1f s" % (cores, a.steps, wall)
        config["l2_policy"] = "no flush: the working set of a step (>= 0.47 GB of SGBM volumes per frame) is far larger than the 126 MB L2"
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16",
                "data": "synthetic", "config": config,
                "run": {"frames_per_step": cores, "sequences": cores, "note": "one single-threaded reference process per host core"},
                "cpu_baseline": {"value": fps, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": fps, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---- CPU baseline first (before CUDA is initialised in this process), rank 0 at N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = host_cores()
        nfr = 24 if a.config != "F" else 4
        fps, wall = run_cpu(a.config, L, R, 2, nfr, cores)
        cpu_baseline = {"value": fps, "unit": unit, "cores": cores, "kind": "port",
                        "sample": "oracle port on cv2 (the reference's CPU path): %d processes x %d frames (cv2 threads=1 each), "
                                  "wall %.1f s" % (cores, nfr, wall)}
        # what a user of the reference gets from one update() stream, and where its time goes
        prof = run_cpu_profile(a.config, L, R, 12 if a.config == "K" else 3)
        cpu_baseline["single_stream"] = {"value": prof.get("single_stream_fps"), "unit": unit, "cv2_threads": prof.get("cv2_threads"),
                                         "frames": prof.get("frames")}
        cpu_baseline["stages_ms_per_frame"] = prof.get("stages_ms")

    if "OVO_SELECT_THREADS" not in os.environ:
        # host threads for the per-frame retainBest ordering: the rank's share of the cores minus one for the driving thread
        os.environ["OVO_SELECT_THREADS"] = str(a.select_threads or max(2, min(8, host_cores() // max(1, world) - 1)))
    import torch
    import torch.distributed as dist
    from openvo_b200 import StereoCamera, synth, _native
    from openvo_b200.batch import BatchOdometer
    from openvo_b200 import dist as odist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    S = a.seqs or cfg["seqs"]
    cam = StereoCamera(**synth.camera_args(cfg["W"], cfg["H"], cfg["D"]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    NT = max(1, min(a.threads, S))
    NG = max(1, min(a.groups, S // NT))
    assert S % (NT * NG) == 0, "--seqs must be divisible by --threads x --groups"
    SP = S // (NT * NG)  # sequences per batch
    streams = [[torch.cuda.Stream() for _ in range(NG)] for _ in range(NT)]

    def fresh():
        return [[BatchOdometer(cam, SP, nfeatures=cfg["n"], engine_tag=t * NG + g, preprocessed_frames=True) for g in range(NG)]
                for t in range(NT)]

    dev_L, dev_R = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    pin_L = [torch.from_numpy(L[i]).pin_memory() for i in range(len(L))]   # e2e inputs: host frames in pinned memory
    pin_R = [torch.from_numpy(R[i]).pin_memory() for i in range(len(R))]
    lib = _native.load()

    def begin(bo, t, g, s, host):
        idx = [frame_index(s, rank * S + (t * NG + g) * SP + q) for q in range(SP)]
        if host:
            bo.begin([pin_L[i] for i in idx], [pin_R[i] for i in idx])  # one pinned host frame per sequence, H2D every step
        else:
            ti = torch.tensor(idx, device="cuda")
            bo.begin_device(dev_L[ti], dev_R[ti])

    def run_part(bos_t, t, steps, first_step, host, hist=None):
        """One host thread, NG batches round-robin: finish batch g of step s, queue batch g of step s+1, go on to batch g+1.
        hist (optional): [T [S, steps, 4, 4], status [S, steps]] filled with every frame's relative transform / commit mode."""
        ok = 0
        for g in range(NG):
            with torch.cuda.stream(streams[t][g]):
                begin(bos_t[g], t, g, first_step, host)
        for s in range(first_step, first_step + steps):
            for g in range(NG):
                with torch.cuda.stream(streams[t][g]):
                    res = bos_t[g].finish()
                    ok += sum(res)
                    if hist is not None:
                        base = (t * NG + g) * SP
                        for q, od in enumerate(bos_t[g].odometers):
                            if res[q] and od.last_mode:
                                hist[0][base + q, s - first_step] = od.last_T
                                hist[1][base + q, s - first_step] = od.last_mode
                    if s + 1 < first_step + steps:
                        begin(bos_t[g], t, g, s + 1, host)
        return ok

    def run(bos, steps, first_step, host, hist=None):
        oks, errs = [0] * NT, []
        torch.cuda.synchronize()

        def work(t):
            try:
                torch.cuda.set_device(local_rank)
                oks[t] = run_part(bos[t], t, steps, first_step, host, hist)
            except Exception as e:  # pragma: no cover
                errs.append(e)
        if NT == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(t,)) for t in range(NT)]
            for x in th:
                x.start()
            for x in th:
                x.join()
        if errs:
            raise errs[0]
        for sl in streams:
            for st in sl:
                torch.cuda.current_stream().wait_stream(st)
        return sum(oks)

    def timed(host):
        bos = fresh()
        run(bos, warmup, 0, host)
        engs = [b.engine for bt in bos for b in bt]
        h2d0, d2h0, l0 = sum(e.h2d_bytes for e in engs), sum(e.d2h_bytes for e in engs), lib.ovo_launch_count()
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hist = [np.tile(np.eye(4), (S, a.steps, 1, 1)), np.zeros((S, a.steps), np.int32)]
        ok = run(bos, a.steps, warmup, host, hist)
        chains = None
        if world > 1:  # the only exchange: every frame's relative transform + commit mode of the chunk, gathered once
            Tall, sall, _ = odist.gather_poses(hist[0], hist[1], odist.shard_sequences(S * world, rank, world), S * world)
            chains = odist.replay_chains(Tall, sall)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.result()
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        causes = {}
        for bt in bos:
            for b in bt:
                for od in b.odometers:
                    key = od.skip_cause or "none"
                    causes[key] = causes.get(key, 0) + 1
        return dict(ms=ms, ok=ok, causes=causes, chains=chains, h2d=(sum(e.h2d_bytes for e in engs) - h2d0) / a.steps, d2h=(sum(e.d2h_bytes for e in engs) - d2h0) / a.steps,
                    launches=lib.ovo_launch_count() - l0, clocks=clocks, bos=bos)

    dev = timed(host=False)
    e2e = timed(host=True)
    frames = S * world * a.steps
    value = frames / (dev["ms"] * 1e-3)
    e2e_value = frames / (e2e["ms"] * 1e-3)

    # ---- verification: sequence 0 of this rank replayed frame by frame through the single-stream StereoOdometer.update()
    # must land on exactly the pose the batched, two-streams-in-flight run produced (same kernels, different driver)
    from openvo_b200 import StereoOdometer
    import hashlib
    single = StereoOdometer(cam, nfeatures=cfg["n"], preprocessed_frames=True, _engine_tag=999)
    for s_ in range(warmup + a.steps):
        i = frame_index(s_, rank * S)
        single.update(L[i], R[i])
    od0 = dev["bos"][0][0].odometers[0]
    e2e0 = e2e["bos"][0][0].odometers[0]
    verified = bool(np.array_equal(single.c_T_w, od0.c_T_w) and np.array_equal(single.c_T_w, e2e0.c_T_w) and
                    single.skip_cause == od0.skip_cause and single.skipped_frames == od0.skipped_frames)
    pose_hash = hashlib.sha1(np.ascontiguousarray(np.stack([od.c_T_w for bt in dev["bos"] for b in bt for od in b.odometers])).tobytes()).hexdigest()[:16]
    del single

    # ---- roofline leg: per-kernel CUDA-event durations over an identical region (events on the launching stream)
    roofline, per_kernel = None, {}
    if rank == 0:
        bos = dev["bos"]
        lib.ovo_profile_enable(1)
        b0 = bos[0][0]
        with torch.cuda.stream(streams[0][0]):  # one stream only: event pairs around each launch must not see another stream's kernels
            for s_ in range(warmup + a.steps, warmup + a.steps + min(a.steps, 5)):
                begin(b0, 0, 0, s_, False)
                b0.finish()
        torch.cuda.synchronize()
        prof = _native.profile_read(lib)
        lib.ovo_profile_enable(0)
        tot = sum(v[0] for v in prof.values()) or 1.0
        per_kernel = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share": v[0] / tot} for k, v in prof.items()}
        top = max(prof, key=lambda k: prof[k][0])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, which = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        abytes = algorithmic_bytes(top, b0.engine, SP, fused="k_sgbm_vsum_t" in prof)
        dur_s = prof[top][0] / prof[top][1] * 1e-3
        achieved = abytes / dur_s / 1e9 if dur_s > 0 else 0.0
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("source_sha1") == source_sha1():   # a capture of other code says nothing about this build: report null
                traffic = tj.get(a.config, {}).get(top)
        except Exception:
            pass
        roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": abytes,
                    "ms_per_launch": dur_s * 1e3, "share_of_step": prof[top][0] / tot,
                    "note": "kernel with the largest share of the step.  Since round 2 the path kernels write ONE vertical volume "
                            "(Sv) instead of three, which halves their algorithmic bytes; they are no longer HBM-bound (ncu: DRAM 29 % "
                            "for k_sgbm_vsum, 55 % for k_sgbm_horiz) but limited by the ALU / DPX pipe and shared-memory latency (46 % / 69 % pipe utilisation, "
                            "profiles/r02_ncu_full_summary.txt), so this HBM fraction is low by construction"}
    stages = None
    if rank == 0 and per_kernel:
        # per-stage view (SURVEY.md §8(d) contract numerators); times are per frame from the single-stream event profile
        def t_of(prefixes):
            return sum(v["ms_per_launch"] * v["launches"] for k, v in per_kernel.items() if k.startswith(prefixes)) / (min(a.steps, 5) * SP)
        W_, H_, D_ = cfg["W"], cfg["H"], cfg["D"]
        eng0 = dev["bos"][0][0].engine
        nkp = float(np.mean([od._cur.n_kp for bt in dev["bos"] for b in bt for od in b.odometers if od._cur is not None] or [0]))
        ipk = {}
        try:
            ipk = json.load(open(os.path.join(ROOT, "profiles", "int_peaks.json")))
        except Exception:
            pass
        t4 = t_of(("k_sgbm", "k_median3", "k_ccl"))
        t12 = t_of(("k_orb",))
        t3 = t_of(("k_knn2",))
        t5 = t_of(("k_match_gather", "k_umeyama"))
        ops4 = 76.0 * (W_ - D_) * H_ * D_
        bytes12 = 19.5 * eng0.cw * eng0.ch + 1400.0 * nkp
        ops3 = 27.0 * nkp * nkp
        int_peak = ipk.get("int32_add_logic_gops")
        popc_peak = ipk.get("popc_xor_add_gops")
        stages = {
            "stage4_sgbm": {"ms_per_frame": t4, "int_ops_per_frame": ops4, "achieved_gops": ops4 / (t4 * 1e-3) / 1e9 if t4 else None,
                            "peak_gops": int_peak, "frac": (ops4 / (t4 * 1e-3) / 1e9 / int_peak) if (t4 and int_peak) else None,
                            "peak_source": "profiles/int_peaks.json int32_add_logic (measured, tools/int_peak.cu)",
                            "frac_of_dpx_int16_peak": (ops4 / (t4 * 1e-3) / 1e9 / ipk["dpx_viaddmnmx_u16x2_int16_gops"])
                            if (t4 and ipk.get("dpx_viaddmnmx_u16x2_int16_gops")) else None,
                            "note": "the path kernels run packed int16 on the half-rate DPX pipe (VIADDMNMX.U16x2): the DPX figure counts two "
                                    "int16 operations per lane instruction"},
            "stage12_orb": {"ms_per_frame": t12, "bytes_per_frame": bytes12, "achieved_gbs": bytes12 / (t12 * 1e-3) / 1e9 if t12 else None,
                            "peak_gbs": peak, "frac": (bytes12 / (t12 * 1e-3) / 1e9 / peak) if t12 else None, "keypoints": nkp,
                            "note": "device kernels only; the retainBest host step is outside"},
            "stage3_match": {"ms_per_frame": t3, "int_ops_per_frame": ops3, "achieved_gops": ops3 / (t3 * 1e-3) / 1e9 if t3 else None,
                             "peak_gops": popc_peak, "frac": (ops3 / (t3 * 1e-3) / 1e9 / popc_peak) if (t3 and popc_peak) else None,
                             "peak_source": "profiles/int_peaks.json popc_xor_add (measured)"},
            "stage5_pose": {"ms_per_frame": t5, "note": "latency-bound; no roofline (SURVEY.md §8(d))"},
        }
    # ---- extra records (default run only) -------------------------------------------------------------------------------
    run_info = {"frames_per_step": S * world, "sequences_per_gpu": S, "host_threads": NT, "batches_in_flight_per_thread": NG,
                "sequences_per_batch": SP, "workspace_gb_per_frame": dev["bos"][0][0].engine.workspace.numel() / max(1, SP) / 1e9,
                "frames_committed": dev["ok"], "frames": S * a.steps, "last_skip_cause_per_sequence": dev["causes"],
                "pose_hash": pose_hash}
    if dev["chains"] is not None:  # chains of the timed chunk, replayed on every rank from the all-gathered per-frame transforms
        run_info["gathered_chains"] = int(dev["chains"].shape[0])
        run_info["gathered_chains_hash"] = hashlib.sha1(np.ascontiguousarray(dev["chains"]).tobytes()).hexdigest()[:16]
    extras = {}
    if not a.no_extras:
        # free this configuration's workspaces first: 4K needs 16 GB per frame in flight
        del dev["bos"], e2e["bos"], od0, e2e0
        bos = b0 = eng0 = None
        cam._engines.clear()
        gc.collect()
        torch.cuda.empty_cache()
        if world == 1 and a.config == "K":
            extras["single_sequence"] = single_sequence(cam, cfg, pin_L, pin_R, L, R)
            cam._engines.clear()
            gc.collect()
            torch.cuda.empty_cache()
        if world > 1:
            extras["frame_sharded_4k"] = frame_sharded_4k(rank, world)
        elif a.config == "K":
            other = {}
            for key, steps in (("F", 5), ("U", 4)):
                try:
                    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", key, "--steps", str(steps), "--warmup", "3",
                                          "--no-cpu-baseline", "--no-extras"], capture_output=True, text=True, timeout=900)
                    d_ = json.loads(out.stdout.strip().splitlines()[-1])
                    other[key] = {k: d_[k] for k in ("value", "unit", "ms_per_step", "steps", "config", "run", "verified", "clocks", "gpu_launches")}
                    other[key]["e2e"] = d_["e2e"]
                    other[key]["stages"] = {k: {"ms_per_frame": v.get("ms_per_frame"), "frac": v.get("frac")} for k, v in (d_.get("stages") or {}).items()}
                except Exception as e:  # pragma: no cover
                    other[key] = {"error": repr(e)}
            extras["other_configs"] = other
    if rank == 0:
        config["l2_policy"] = "no flush: the working set of a step (>= 0.47 GB of SGBM volumes per frame) is far larger than the 126 MB L2"
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": warmup,
                "ms_per_step": dev["ms"] / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16",
                "data": "synthetic", "config": config, "run": run_info, "verified": verified,
                "clocks": dev["clocks"], "gpu_launches": dev["launches"],
                "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                        "ms_per_step": e2e["ms"] / a.steps},
                "roofline": roofline, "stages": stages, "kernels": per_kernel}
        line.update(extras)
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
