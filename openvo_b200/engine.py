"""Device engine: owns the C-ABI context, its torch-allocated workspace and the per-frame device buffers.

PyTorch is used for device memory and streams only; every arithmetic step is a kernel behind include/openvo_b200.h.
"""
import ctypes
import os

import numpy as np
import torch

from . import _native as N

# Engine.async_finish: the host half of the ORB seam runs on a worker thread of the library (it starts the moment the detection
# phase lands) instead of on the calling thread inside frames_finish.  It makes a single driver thread less sensitive to its own
# latency (SequenceOdometer turns it on: 2400 -> 2950 frames/s) but costs host threads: with 8 ranks on 32 cores the batched
# throughput drops (0.93 -> 0.85 of linear), so it is off by default.  OVO_ASYNC_FINISH=0/1 forces it for every engine.
ASYNC_FINISH = {"0": False, "1": True}.get(os.environ.get("OVO_ASYNC_FINISH", ""), None)


class Frame:
    """Device-resident products of one stereo pair (what the reference keeps in current_*/prev_*,
    ref: src/openVO/stereo_odometer.py:107-113)."""
    __slots__ = ("img", "disp", "kp", "desc", "n_kp", "_host", "_dirty")

    def __init__(self, img, disp, kp, desc, n_kp):
        self.img, self.disp, self.kp, self.desc, self.n_kp = img, disp, kp, desc, n_kp
        self._host = {}      # host-side copies in the reference's types, materialised on first read (or assigned by the caller)
        self._dirty = False  # host-side products were assigned: the device copy must be rebuilt before the next pair step


class Engine:
    """One engine = one C-ABI context + workspace + pinned staging.  An engine is NOT re-entrant: its staging buffers and its
    workspace are reused by every call, so it must be driven from one host thread on one CUDA stream at a time (StereoCamera
    hands out one engine per (nfeatures, batch, tag); drivers that want concurrency use distinct tags, as bench.py does)."""

    def __init__(self, width, height, sgbm_params, roi, Q, nfeatures, max_batch=1, min_valid=4.0, max_valid=100.0,
                 device=None, lib_path=None):
        if not torch.cuda.is_available():
            raise N.NativeError("openvo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = N.load(lib_path)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.cfg = N.make_config(width, height, sgbm_params, roi, Q, nfeatures, max_batch, min_valid, max_valid)
        self.W, self.H, self.max_batch, self.nfeatures = int(width), int(height), int(max_batch), int(nfeatures)
        self.Q = np.asarray(Q, np.float64).copy()
        nbytes = self.lib.ovo_workspace_bytes(ctypes.byref(self.cfg))
        if nbytes == 0:
            raise N.NativeError(self.lib.ovo_last_error().decode())
        cw, ch = ctypes.c_int(), ctypes.c_int()
        N.check(self.lib, self.lib.ovo_cropped_size(ctypes.byref(self.cfg), ctypes.byref(cw), ctypes.byref(ch)))
        self.cw, self.ch = cw.value, ch.value
        self.kp_cap = self.lib.ovo_kp_capacity(ctypes.byref(self.cfg))
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.ctx = self.lib.ovo_create(ctypes.byref(self.cfg), self.workspace.data_ptr(), nbytes)
        if not self.ctx:
            raise N.NativeError(self.lib.ovo_last_error().decode())
        self._pin = {}
        self.async_finish = False
        # persistent buffers of the pair step, one slot per frame of the batch
        nb = self.max_batch
        self.nn = torch.empty((nb, self.kp_cap, 4), dtype=torch.int32, device=self.device)
        self.nn_rev = torch.empty((nb, self.kp_cap, 4), dtype=torch.int32, device=self.device)  # cross-check (opt-in)
        self.matches = torch.empty((nb, self.kp_cap, 3), dtype=torch.int32, device=self.device)
        self.pts1 = torch.empty((nb, self.kp_cap, 3), dtype=torch.float32, device=self.device)
        self.pts2 = torch.empty((nb, self.kp_cap, 3), dtype=torch.float32, device=self.device)
        self.pair_out = torch.zeros((nb, 18), dtype=torch.float64, device=self.device)  # [0:16] rigid, [16] two i32 counts
        self.pair_host = torch.zeros((nb, 18), dtype=torch.float64).pin_memory()
        self._d2h = 0
        self._h2d = 0

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.ovo_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def _lib_bytes(self):
        h, d = ctypes.c_longlong(), ctypes.c_longlong()
        self.lib.ovo_transfer_bytes(self.ctx, ctypes.byref(h), ctypes.byref(d))
        return h.value, d.value

    @property
    def h2d_bytes(self):
        """host->device bytes moved so far (frame uploads + the library's own staging)."""
        return self._h2d + self._lib_bytes()[0]

    @property
    def d2h_bytes(self):
        return self._d2h + self._lib_bytes()[1]

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _pinned(self, key, shape, dtype):
        buf = self._pin.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype).pin_memory()
            self._pin[key] = buf
        return buf

    def upload(self, arr, key):
        """numpy uint8 [nb,H,W(,3)] (or a list of nb frames) -> device tensor via a pinned staging buffer (one host copy)."""
        if isinstance(arr, (list, tuple)) and isinstance(arr[0], torch.Tensor):
            # frames that already live in pinned host memory (a loader decoding into pinned buffers): straight H2D, no staging copy
            dst = torch.empty((len(arr),) + tuple(arr[0].shape), dtype=torch.uint8, device=self.device)
            for i, fr in enumerate(arr):
                dst[i].copy_(fr, non_blocking=True)
                self._h2d += fr.numel()
            return dst
        if isinstance(arr, (list, tuple)):
            first = np.asarray(arr[0])
            stage = self._pinned(key, (len(arr),) + first.shape, torch.uint8)
            view = stage.numpy()
            for i, fr in enumerate(arr):
                view[i] = fr
            self._h2d += view.nbytes
            return stage.to(self.device, non_blocking=True)
        a = np.ascontiguousarray(arr)
        stage = self._pinned(key, a.shape, torch.uint8)
        stage.numpy()[...] = a
        self._h2d += a.nbytes
        return stage.to(self.device, non_blocking=True)

    # ---- seams -------------------------------------------------------------------------------------------------------
    def sgbm(self, left, right):
        """left/right: device u8 [nb,H,W] -> device i16 [nb,H,W] (seam S-A)."""
        nb = left.shape[0]
        disp = torch.empty((nb, self.H, self.W), dtype=torch.int16, device=self.device)
        N.check(self.lib, self.lib.ovo_sgbm_compute(self.ctx, left.data_ptr(), right.data_ptr(), self.W, self.W * self.H, nb,
                                                    disp.data_ptr(), self._stream()))
        return disp

    def disparity_post(self, disp16):
        nb = disp16.shape[0]
        d = torch.empty((nb, self.ch, self.cw), dtype=torch.float32, device=self.device)
        m = torch.empty((nb, self.ch, self.cw), dtype=torch.uint8, device=self.device)
        N.check(self.lib, self.lib.ovo_disparity_post(self.ctx, disp16.data_ptr(), nb, d.data_ptr(), m.data_ptr(), self._stream()))
        return d, m

    def crop(self, img):
        nb = img.shape[0]
        out = torch.empty((nb, self.ch, self.cw), dtype=torch.uint8, device=self.device)
        N.check(self.lib, self.lib.ovo_crop_left(self.ctx, img.data_ptr(), self.W, self.W * self.H, nb, out.data_ptr(), self._stream()))
        return out

    def rectify(self, img, maps=None):
        """img: device u8 [nb,H,W] (gray) or [nb,H,W,3] (BGR); maps: (map1 i16 [H,W,2], map2 u16 [H,W]) device tensors or None
        (colour conversion only) -> device u8 [nb,H,W] (cv2.cvtColor + cv2.remap seams)."""
        nb = img.shape[0]
        ch = 3 if img.dim() == 4 else 1
        out = torch.empty((nb, self.H, self.W), dtype=torch.uint8, device=self.device)
        m1, m2 = (maps[0].data_ptr(), maps[1].data_ptr()) if maps is not None else (None, None)
        N.check(self.lib, self.lib.ovo_rectify(self.ctx, img.data_ptr(), ch, self.W * ch, self.W * self.H * ch, nb, m1, m2,
                                               out.data_ptr(), self._stream()))
        return out

    def device_maps(self, key, map1, map2):
        """Upload (once) a CV_16SC2 rectification map pair."""
        cache = self.__dict__.setdefault("_maps", {})
        if key not in cache:
            cache[key] = (torch.from_numpy(np.ascontiguousarray(map1, np.int16)).to(self.device),
                          torch.from_numpy(np.ascontiguousarray(map2).view(np.int16).astype(np.int16)).to(self.device))
        return cache[key]

    def reproject(self, disp_f32):
        xyz = torch.empty((self.ch, self.cw, 3), dtype=torch.float32, device=self.device)
        N.check(self.lib, self.lib.ovo_reproject_3d(self.ctx, disp_f32.data_ptr(), xyz.data_ptr(), self._stream()))
        return xyz

    def orb(self, img, mask):
        """img/mask: device u8 [nb,ch,cw] -> kp f32 [nb,cap,6], desc u8 [nb,cap,32], list of counts (seam S-D)."""
        nb = img.shape[0]
        kp = torch.empty((nb, self.kp_cap, N.KP_FIELDS), dtype=torch.float32, device=self.device)
        desc = torch.empty((nb, self.kp_cap, 32), dtype=torch.uint8, device=self.device)
        n = (ctypes.c_int * nb)()
        N.check(self.lib, self.lib.ovo_orb_detect_compute(self.ctx, img.data_ptr(), None if mask is None else mask.data_ptr(), nb,
                                                          kp.data_ptr(), desc.data_ptr(), n, self._stream()))
        return kp, desc, list(n)

    def orb_begin(self, img, mask):
        """First half of orb(): queue the detection phase (asynchronous).  Keep img / mask alive until orb_finish."""
        N.check(self.lib, self.lib.ovo_orb_detect_begin(self.ctx, img.data_ptr(), None if mask is None else mask.data_ptr(),
                                                        img.shape[0], self._stream()))

    def orb_finish(self, nb):
        """Second half of orb(): wait for the detection phase, select on the host, queue the descriptor phase."""
        kp = torch.empty((nb, self.kp_cap, N.KP_FIELDS), dtype=torch.float32, device=self.device)
        desc = torch.empty((nb, self.kp_cap, 32), dtype=torch.uint8, device=self.device)
        n = (ctypes.c_int * nb)()
        N.check(self.lib, self.lib.ovo_orb_detect_finish(self.ctx, nb, kp.data_ptr(), desc.data_ptr(), n, self._stream()))
        return kp, desc, list(n)

    def knn2(self, desc_q, nq, desc_t, nt, out=None):
        nn = self.nn[0] if out is None else out
        N.check(self.lib, self.lib.ovo_knn2_hamming(self.ctx, desc_q.data_ptr(), nq, desc_t.data_ptr(), nt, nn.data_ptr(), self._stream()))
        return nn

    def pair_async(self, a, b, match_threshold, slot=0, cross_check=False):
        return self.pair_batch_async([(a, b, slot)], match_threshold, cross_check)

    def pair_batch_async(self, jobs, match_threshold, cross_check=False):
        """jobs: list of (frame a, frame b, slot).  All pairs in four launches (ovo_pair_batch); results land in pair_out[slot]."""
        if not jobs:
            return
        items = (N.PairItem * len(jobs))()
        for it, (a, b, slot) in zip(items, jobs):
            it.q_desc, it.t_desc, it.nq, it.nt = a.desc.data_ptr(), b.desc.data_ptr(), a.n_kp, b.n_kp
            it.kp1, it.kp2, it.disp1, it.disp2 = a.kp.data_ptr(), b.kp.data_ptr(), a.disp.data_ptr(), b.disp.data_ptr()
            it.nn, it.matches = self.nn[slot].data_ptr(), self.matches[slot].data_ptr()
            it.pts1, it.pts2, it.out = self.pts1[slot].data_ptr(), self.pts2[slot].data_ptr(), self.pair_out[slot].data_ptr()
            it.nn_rev = self.nn_rev[slot].data_ptr() if cross_check else None
        N.check(self.lib, self.lib.ovo_pair_batch(self.ctx, len(jobs), items, float(match_threshold), self._stream()))

    def pair_collect(self, nslots=1):
        """One D2H (144 bytes per slot) + one stream sync -> list of (n_matches, n_bad_lookups, out16 numpy)."""
        self.pair_host[:nslots].copy_(self.pair_out[:nslots], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self._d2h += nslots * 144
        host = self.pair_host.numpy()
        res = []
        for s in range(nslots):
            counts = host[s, 16:17].view(np.int32)
            res.append((int(counts[0]), int(counts[1]), host[s, :16].copy()))
        return res

    def pair(self, a, b, match_threshold, cross_check=False):
        self.pair_async(a, b, match_threshold, 0, cross_check)
        return self.pair_collect(1)[0]

    def filtered_transform(self, pts1, pts2, count_dev, rigidity_threshold, outlier_threshold, min_matches):
        """point_cloud_transform with the optional filters on (ref: stereo_odometer.py:177-205), on device buffers
        pts1/pts2 [cap,3] and a device count (all modified in place).  Returns (n_after_rigidity, n_final, out16 or None)."""
        st = self._stream()
        cap = pts1.shape[0]
        host = torch.empty(1, dtype=torch.int32).pin_memory()

        def read_count():
            host.copy_(count_dev, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            self._d2h += 4
            return int(host[0])
        out = torch.empty(16, dtype=torch.float64, device=self.device)
        if rigidity_threshold > 0:
            N.check(self.lib, self.lib.ovo_rigid_body_filter(self.ctx, pts1.data_ptr(), pts2.data_ptr(), count_dev.data_ptr(), cap,
                                                             float(np.float32(rigidity_threshold)), st))
        n1 = read_count()
        n2 = n1
        if outlier_threshold > 0 and n1 >= 10:
            N.check(self.lib, self.lib.ovo_rigid_transform(self.ctx, pts1.data_ptr(), pts2.data_ptr(), count_dev.data_ptr(), cap, out.data_ptr(), st))
            N.check(self.lib, self.lib.ovo_outlier_filter(self.ctx, pts1.data_ptr(), pts2.data_ptr(), count_dev.data_ptr(), cap, out.data_ptr(),
                                                          float(outlier_threshold), st))
            n2 = read_count()
        if n2 < min_matches:
            return n1, n2, None
        N.check(self.lib, self.lib.ovo_rigid_transform(self.ctx, pts1.data_ptr(), pts2.data_ptr(), count_dev.data_ptr(), cap, out.data_ptr(), st))
        self._d2h += 128
        return n1, n2, out.cpu().numpy()

    def pnp_ransac(self, b, slot, n, iters, reproj_px, seed):
        """Opt-in P3P-RANSAC + LM pose (ovo_pnp_ransac) from the buffers the pair step left in `slot`: 3-D points of the query
        frame and the matched keypoints of frame b.  -> out16 numpy."""
        cnt = torch.tensor([n], dtype=torch.int32, device=self.device)
        out = torch.empty(16, dtype=torch.float64, device=self.device)
        N.check(self.lib, self.lib.ovo_pnp_ransac(self.ctx, self.pts1[slot].data_ptr(), self.matches[slot].data_ptr(), b.kp.data_ptr(),
                                                  cnt.data_ptr(), self.kp_cap, int(iters), float(reproj_px), int(seed), out.data_ptr(),
                                                  self._stream()))
        self._d2h += 128
        return out.cpu().numpy()

    def rigid(self, pts1, pts2):
        """numpy float32 [m,3] x2 -> out16 (estimateAffine3D seam for the optional filter paths)."""
        m = len(pts1)
        p1 = torch.from_numpy(np.ascontiguousarray(pts1, np.float32)).to(self.device)
        p2 = torch.from_numpy(np.ascontiguousarray(pts2, np.float32)).to(self.device)
        cnt = torch.tensor([m], dtype=torch.int32, device=self.device)
        out = torch.empty(16, dtype=torch.float64, device=self.device)
        N.check(self.lib, self.lib.ovo_rigid_transform(self.ctx, p1.data_ptr(), p2.data_ptr(), cnt.data_ptr(), m, out.data_ptr(),
                                                       self._stream()))
        return out.cpu().numpy()

    # ---- whole-frame feature extraction ---------------------------------------------------------------------------------
    def frames(self, left, right):
        """left/right: device u8 [nb,H,W] (rectified, gray) -> list of Frame."""
        return self.frames_finish(self.frames_begin(left, right))

    def frames_begin(self, left, right):
        """Queue disparity + detection for a batch without waiting (one batch in flight per engine) -> token for frames_finish."""
        nb = left.shape[0]
        # persistent input / output buffers: ovo_extract_begin replays one CUDA graph per (buffers, batch size); the products a
        # frame keeps (disparity, cropped image) are copied out of them
        px = self.__dict__.get("_xbuf")
        if px is None:
            mb = self.max_batch
            px = self._xbuf = dict(
                left=torch.empty((mb, self.H, self.W), dtype=torch.uint8, device=self.device),
                right=torch.empty((mb, self.H, self.W), dtype=torch.uint8, device=self.device),
                disp16=torch.empty((mb, self.H, self.W), dtype=torch.int16, device=self.device),
                disp=torch.empty((mb, self.ch, self.cw), dtype=torch.float32, device=self.device),
                mask=torch.empty((mb, self.ch, self.cw), dtype=torch.uint8, device=self.device),
                img=torch.empty((mb, self.ch, self.cw), dtype=torch.uint8, device=self.device))
        px["left"][:nb].copy_(left, non_blocking=True)
        px["right"][:nb].copy_(right, non_blocking=True)
        N.check(self.lib, self.lib.ovo_extract_begin(self.ctx, px["left"].data_ptr(), px["right"].data_ptr(), self.W, self.W * self.H, nb,
                                                     px["disp16"].data_ptr(), px["disp"].data_ptr(), px["mask"].data_ptr(),
                                                     px["img"].data_ptr(), self._stream()))
        disp, img = px["disp"][:nb].clone(), px["img"][:nb].clone()
        # the second half of the ORB seam starts on a worker thread of the library the moment the detection phase lands
        # (host-side retainBest, then the descriptor launch); frames_finish only joins it
        kp = torch.empty((nb, self.kp_cap, N.KP_FIELDS), dtype=torch.float32, device=self.device)
        desc = torch.empty((nb, self.kp_cap, 32), dtype=torch.uint8, device=self.device)
        n = (ctypes.c_int * nb)()
        started = self.async_finish if ASYNC_FINISH is None else ASYNC_FINISH
        if started:
            N.check(self.lib, self.lib.ovo_orb_detect_finish_async(self.ctx, nb, kp.data_ptr(), desc.data_ptr(), n, self._stream()))
        return (left, right, kp, disp, desc, img, n, started)

    def frames_finish(self, token):
        _, _, kp, disp, desc, img, n, started = token
        if started:
            N.check(self.lib, self.lib.ovo_orb_detect_wait(self.ctx))
        else:
            N.check(self.lib, self.lib.ovo_orb_detect_finish(self.ctx, img.shape[0], kp.data_ptr(), desc.data_ptr(), n, self._stream()))
        n = list(n)
        return [Frame(img[i], disp[i], kp[i], desc[i], n[i]) for i in range(img.shape[0])]
