"""BatchOdometer — S independent stereo sequences advanced one frame each per call (BASELINE.json config 5, and the
throughput path of bench.py).  Per-sequence semantics are exactly StereoOdometer.update's (the same state machine object
is used); only the device work is batched: one SGBM / ORB launch set over all S frames (grid.z = frame), S pair steps
queued back to back, a single 144*S-byte read-back and one stream synchronisation per call.
"""
import numpy as np

from .stereo_odometer import StereoOdometer


class BatchOdometer:
    def __init__(self, stereo_camera, n_sequences, nfeatures=500, engine_tag=0, **kw):
        self.n = int(n_sequences)
        self.stereo = stereo_camera
        self.odometers = [StereoOdometer(stereo_camera, nfeatures=nfeatures, _max_batch=self.n, _engine_tag=engine_tag, **kw) for _ in range(self.n)]
        self.engine = self.odometers[0]._engine()
        self._pending = None

    def update(self, lefts, rights):
        """lefts/rights: numpy uint8 host frames [S,H,W] (gray) or [S,H,W,3] (BGR); rectified on the device unless the
        odometers were built with preprocessed_frames=True -> list of S bools."""
        eng = self.engine
        l, r = self.stereo._prepare_device(eng, lefts, rights, self.odometers[0].preprocessed_frames, key="b")
        return self.update_device(l, r)

    def update_device(self, lefts, rights):
        """Same, with the frames already resident on the device (torch uint8 [S,H,W])."""
        self.begin_device(lefts, rights)
        return self.finish()

    # update() in two halves, so that one host thread can keep several BatchOdometers (each on its own CUDA stream) busy:
    # begin() only queues device work for the next frames; finish() waits for it, selects keypoints on the host, runs the
    # pair step and advances the S state machines.  begin + finish == update.
    def begin(self, lefts, rights):
        eng = self.engine
        l, r = self.stereo._prepare_device(eng, lefts, rights, self.odometers[0].preprocessed_frames, key="b")
        self.begin_device(l, r)

    def begin_device(self, lefts, rights):
        assert self._pending is None, "finish() the previous batch first"
        self._pending = self.engine.frames_begin(lefts, rights)

    def finish(self):
        eng = self.engine
        token, self._pending = self._pending, None
        frames = eng.frames_finish(token)
        queued, jobs = [], []
        for i, (od, fr) in enumerate(zip(self.odometers, frames)):
            if od._cur is not None and fr.n_kp >= od.min_matches and fr.n_kp >= 2:
                jobs.append((od._cur, fr, i))
                queued.append(i)
        eng.pair_batch_async(jobs, self.odometers[0].match_threshold, self.odometers[0].cross_check)
        res = eng.pair_collect(self.n) if queued else []
        out, errors = [], []
        for i, (od, fr) in enumerate(zip(self.odometers, frames)):
            # a hard failure of one sequence (the reference's ZeroDivisionError / IndexError paths) must not leave the state
            # machines of the others un-advanced: collect, advance everyone, then re-raise the first
            try:
                out.append(od._advance(fr, first=(i, res[i]) if i in queued else None))
            except (ZeroDivisionError, IndexError) as e:
                out.append(False)
                errors.append((i, e))
        if errors:
            self.failed = errors
            raise errors[0][1]
        return out

    def poses(self):
        return np.stack([od.current_pose() for od in self.odometers])

    def relative_transforms(self):
        """Last committed relative transform per sequence (identity when the last frame was skipped) and status words."""
        T = np.stack([od.last_T if od.last_T is not None else np.eye(4) for od in self.odometers])
        return T
