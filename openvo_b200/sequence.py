"""SequenceOdometer — ONE stereo sequence processed frame-parallel on one GPU (BASELINE.json configs 1/2: "200-frame
KITTI-shaped sequence on 1xB200").

Extraction and disparity are independent per stereo pair; only the pose chain is ordered
(ref: src/openVO/stereo_odometer.py:136-160).  ``run`` therefore
  1. extracts features + disparity of a chunk of consecutive frames in one batched launch set (grid.z = frame),
  2. aligns every ADJACENT pair of the chunk speculatively in one batched pair step (2-NN, ratio test, 3-D lookup, Umeyama),
  3. replays the reference's skip / fall-back state machine (B4) serially on the host: a speculative result is used only if its
     first frame is in fact the odometer's current frame; after a failed frame the next one must be aligned to the last
     COMMITTED frame (and, on failure, to the one before it), so that pair is re-matched on demand — rare.
Three chunks are in flight on three streams / workspaces so that the host-side keypoint selection and state-machine replay of
one overlap device work of the others.  Results are identical to calling ``update`` frame by frame (same kernels, same state machine object).
"""
import numpy as np

from .stereo_odometer import StereoOdometer


class SequenceOdometer(StereoOdometer):
    def __init__(self, stereo_camera, chunk=24, in_flight=3, **kw):
        super().__init__(stereo_camera, _max_batch=int(chunk), **kw)
        self.chunk = int(chunk)
        self.in_flight = max(1, int(in_flight))  # chunks queued on the device at any time (one engine + stream each)
        self._active = None     # engine whose pair buffers hold the result being replayed
        self._streams = None

    def _engine(self):
        if self._active is not None:
            return self._active
        return super()._engine()

    def _chunk_engine(self, k):
        return self.stereo.engine(self._nfeatures, self.chunk, float(self.MIN_VALID_DISPARITY), float(self.MAX_VALID_DISPARITY),
                                  tag=("sequence", self._engine_tag, k))

    def run(self, lefts, rights):
        """lefts / rights: sequences of host frames ([n,H,W] / [n,H,W,3] arrays or lists) -> list of n bools, exactly what n
        calls of ``update`` would have returned; all odometer state (c_T_w, skip_cause, skipped_frames, current_* / prev_*) ends
        up as after those calls."""
        import torch
        n = len(lefts)
        if n == 0:
            return []
        NF = self.in_flight
        engines = [self._chunk_engine(k) for k in range(NF)]
        for e in engines:
            e.async_finish = True   # one driver thread: let the library's worker start each chunk's host half as soon as it can
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=engines[0].device) for _ in range(NF)]
        streams = self._streams
        bounds = [(i, min(i + self.chunk, n)) for i in range(0, n, self.chunk)]
        tokens = [None] * NF

        def begin(c):
            i0, i1 = bounds[c]
            eng = engines[c % NF]
            with torch.cuda.stream(streams[c % NF]):
                l, r = self.stereo._prepare_device(eng, lefts[i0:i1], rights[i0:i1], self.preprocessed_frames, key="seq%d" % (c % NF))
                tokens[c % NF] = eng.frames_begin(l, r)

        import os
        import time
        trace = [] if os.environ.get("OVO_SEQ_TRACE") else None   # per-chunk host timings (begin, extract wait + select, pair, replay)
        self.trace = trace
        out = []
        last = None  # the frame before the chunk's first one (committed or not: the speculation is checked at replay time)
        for c in range(min(NF - 1, len(bounds))):
            begin(c)
        try:
            for c in range(len(bounds)):
                t0 = time.perf_counter()
                if c + NF - 1 < len(bounds):
                    begin(c + NF - 1)
                t1 = time.perf_counter()
                eng = engines[c % NF]
                with torch.cuda.stream(streams[c % NF]):
                    frames = eng.frames_finish(tokens[c % NF])
                    t2 = time.perf_counter()
                    tokens[c % NF] = None
                    firsts = [last] + frames[:-1]
                    jobs, queued = [], {}
                    for k, (a, b) in enumerate(zip(firsts, frames)):
                        if a is not None and a.n_kp >= self.min_matches and b.n_kp >= self.min_matches and b.n_kp >= 2:
                            jobs.append((a, b, k))
                            queued[k] = a
                    eng.pair_batch_async(jobs, self.match_threshold, self.cross_check)
                    res = eng.pair_collect(len(frames)) if jobs else []
                    t3 = time.perf_counter()
                    self._active = eng
                    for k, fr in enumerate(frames):
                        spec = (k, res[k]) if (k in queued and queued[k] is self._cur) else None
                        out.append(self._advance(fr, first=spec))
                    last = frames[-1]
                    if trace is not None:
                        trace.append((t1 - t0, t2 - t1, t3 - t2, time.perf_counter() - t3))
        finally:
            self._active = None
        return out
