"""Device engine: owns the C-ABI context, its torch-allocated workspace and the per-frame device buffers.

PyTorch is used for device memory and streams only; every arithmetic step is a kernel behind include/openvo_b200.h.
"""
import ctypes

import numpy as np
import torch

from . import _native as N


class Frame:
    """Device-resident products of one stereo pair (what the reference keeps in current_*/prev_*,
    ref: src/openVO/stereo_odometer.py:107-113)."""
    __slots__ = ("img", "disp", "kp", "desc", "n_kp", "_host")

    def __init__(self, img, disp, kp, desc, n_kp):
        self.img, self.disp, self.kp, self.desc, self.n_kp = img, disp, kp, desc, n_kp
        self._host = {}


class Engine:
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
        # small persistent buffers of the pair step
        self.nn = torch.empty((self.kp_cap, 4), dtype=torch.int32, device=self.device)
        self.matches = torch.empty((self.kp_cap, 3), dtype=torch.int32, device=self.device)
        self.pts1 = torch.empty((self.kp_cap, 3), dtype=torch.float32, device=self.device)
        self.pts2 = torch.empty((self.kp_cap, 3), dtype=torch.float32, device=self.device)
        self.pair_out = torch.empty(18, dtype=torch.float64, device=self.device)  # [0:16] rigid, [16:18] counts (as i32 view)
        self.pair_host = torch.empty(18, dtype=torch.float64).pin_memory()

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.ovo_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

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
        """numpy uint8 [H,W] or [nb,H,W] -> device tensor via a pinned staging buffer."""
        a = np.ascontiguousarray(arr)
        stage = self._pinned(key, a.shape, torch.uint8)
        stage.numpy()[...] = a
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

    def knn2(self, desc_q, nq, desc_t, nt, out=None):
        nn = self.nn if out is None else out
        N.check(self.lib, self.lib.ovo_knn2_hamming(self.ctx, desc_q.data_ptr(), nq, desc_t.data_ptr(), nt, nn.data_ptr(), self._stream()))
        return nn

    def pair(self, a, b, match_threshold):
        """Frames a (query) and b (train): 2-NN + ratio + fused 3-D lookup + rigid alignment, all on the device, one
        D2H of 144 bytes.  Returns (n_matches, n_bad_lookups, out16 numpy)."""
        self.knn2(a.desc, a.n_kp, b.desc, b.n_kp)
        counts_ptr = self.pair_out.data_ptr() + 16 * 8
        N.check(self.lib, self.lib.ovo_match_points(self.ctx, self.nn.data_ptr(), a.n_kp, float(match_threshold), a.kp.data_ptr(),
                                                    b.kp.data_ptr(), a.disp.data_ptr(), b.disp.data_ptr(), self.matches.data_ptr(),
                                                    self.pts1.data_ptr(), self.pts2.data_ptr(), counts_ptr, self._stream()))
        N.check(self.lib, self.lib.ovo_rigid_transform(self.ctx, self.pts1.data_ptr(), self.pts2.data_ptr(), counts_ptr, self.kp_cap,
                                                       self.pair_out.data_ptr(), self._stream()))
        self.pair_host.copy_(self.pair_out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        host = self.pair_host.numpy()
        counts = host[16:18].view(np.int32)
        return int(counts[0]), int(counts[1]), host[:16].copy()

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
        disp16 = self.sgbm(left, right)
        disp, mask = self.disparity_post(disp16)
        img = self.crop(left)
        kp, desc, n = self.orb(img, mask)
        return [Frame(img[i], disp[i], kp[i], desc[i], n[i]) for i in range(left.shape[0])]
