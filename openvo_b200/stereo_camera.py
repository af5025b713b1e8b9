"""StereoCamera — same constructor, attributes and methods as the reference class
(ref: src/openVO/stereo_camera.py:6-55), with the per-frame arithmetic on the B200.

Calibration (stereoRectify / initUndistortRectifyMap) is init-time only and stays a cv2 call, exactly as in the
reference (out of the hot path).  ``compute_3d`` keeps the reference's contract — it returns numpy
arrays cropped with the (x, y, w, h)-as-(x1, y1, x2, y2) slices (bug-compatible B1) — while StereoOdometer uses the
device-resident path (``frames_device``) and never materialises the 3-D image.
"""
import pickle

import numpy as np

from . import _native as N


class _SgbmHandle:
    """Stands in for the cv2.StereoSGBM object the reference keeps in ``stereoSGBM``
    (ref: src/openVO/stereo_camera.py:23-27): ``compute(left, right)`` -> int16 disparity * 16."""

    def __init__(self, cam):
        self._cam = cam

    def compute(self, left, right):
        cam = self._cam
        left, right = np.asarray(left), np.asarray(right)
        if left.shape != right.shape or left.dtype != np.uint8 or right.dtype != np.uint8 or left.ndim != 2:
            raise N.cv2_error("StereoSGBM.compute: left and right must be 2-D uint8 images of equal size "
                              "(OpenCV: left.size() == right.size() && left.type() == right.type())")
        if left.shape != (cam.img_size[1], cam.img_size[0]):
            raise N.cv2_error("StereoSGBM.compute: image size differs from the camera's img_size")
        eng = cam.engine()
        l, r = eng.upload(left[None], "sgbm_l"), eng.upload(right[None], "sgbm_r")
        return eng.sgbm(l, r)[0].cpu().numpy()

    def __getattr__(self, name):
        # getMinDisparity(), getNumDisparities(), ... like the cv2 object
        if name.startswith("get"):
            key = name[3].lower() + name[4:]
            key = {"p1": "P1", "p2": "P2"}.get(key, key)
            p = self._cam.sgbm_params
            if key in p:
                return lambda: p[key]
            if key == "mode":
                return lambda: p.get("mode", 0)
        raise AttributeError(name)


class StereoCamera:
    @classmethod
    def from_pfiles(cls, left_cam_file, right_cam_file, rect_file, sgbm_file, img_size):
        # ref: src/openVO/stereo_camera.py:7-14
        loaded = []
        for path in (left_cam_file, right_cam_file, rect_file, sgbm_file):
            with open(path, "rb") as fh:
                loaded.append(pickle.load(fh))
        cam_l, cam_r, rect, sgbm = loaded
        return cls(cam_l["K"], cam_l["dist"], cam_r["K"], cam_r["dist"], rect, sgbm, img_size)

    def __init__(self, K_left, dist_left, K_right, dist_right, rect_params, sgbm_params, img_size):
        import cv2  # init-time calibration only (ref: src/openVO/stereo_camera.py:17-22)
        R1, R2, P1, P2, self.Q, self.valid_region_left, self.valid_region_right = cv2.stereoRectify(
            K_left, dist_left, K_right, dist_right, img_size, rect_params["R"], rect_params["T"])
        self.map_left_1, self.map_left_2 = cv2.initUndistortRectifyMap(K_left, dist_left, R1, P1, img_size, cv2.CV_16SC2)
        self.map_right_1, self.map_right_2 = cv2.initUndistortRectifyMap(K_right, dist_right, R2, P2, img_size, cv2.CV_16SC2)
        self.img_size = (int(img_size[0]), int(img_size[1]))
        self.sgbm_params = {k: int(sgbm_params[k]) for k in N.SGBM_KEYS}
        # opt-in extension (not in the reference, whose `mode=1` is commented out at stereo_camera.py:27): 'mode': 1 -> MODE_HH
        self.sgbm_params["mode"] = int(sgbm_params.get("mode", 0))
        self.stereoSGBM = _SgbmHandle(self)
        self._engines = {}

    # ---- device engine (one per (nfeatures, batch, mask range)) -------------------------------------------------------
    def engine(self, nfeatures=500, max_batch=1, min_valid=4.0, max_valid=100.0, lib_path=None, tag=0):
        from .engine import Engine
        key = (int(nfeatures), int(max_batch), float(min_valid), float(max_valid), tag)
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(self.img_size[0], self.img_size[1], self.sgbm_params, self.valid_region_left, self.Q, nfeatures,
                         max_batch, min_valid, max_valid, lib_path=lib_path)
            self._engines[key] = eng
        return eng

    # ---- reference API ---------------------------------------------------------------------------------------------------
    def undistort_rectify_left(self, img):
        # ref: src/openVO/stereo_camera.py:29-30 (cv2.remap on the device)
        return self._rectify_host(img, "left")

    def undistort_rectify_right(self, img):
        # ref: src/openVO/stereo_camera.py:32-33
        return self._rectify_host(img, "right")

    def _maps(self, eng, side):
        m = (self.map_left_1, self.map_left_2) if side == "left" else (self.map_right_1, self.map_right_2)
        return eng.device_maps(side, m[0], m[1])

    def _rectify_host(self, img, side):
        img = np.asarray(img)
        self._check(img)
        if img.ndim == 3:
            raise N.cv2_error("cv2.remap keeps the channel count; the hot path only rectifies single-channel images "
                              "(compute_3d converts colour input to gray first, like the reference)")
        eng = self.engine()
        return eng.rectify(eng.upload(img[None], "rect_" + side), self._maps(eng, side))[0].cpu().numpy()

    def crop_to_valid_region_left(self, img):
        r = self.valid_region_left
        return img[r[1]:r[3], r[0]:r[2]]

    def crop_to_valid_region_right(self, img):
        r = self.valid_region_right
        return img[r[1]:r[3], r[0]:r[2]]

    def _check(self, img):
        if hasattr(img, "data_ptr"):  # torch CPU tensor
            import torch
            if img.dtype != torch.uint8 or img.device.type != "cpu" or not img.is_contiguous():
                raise N.cv2_error("torch frames must be contiguous uint8 CPU (ideally pinned) tensors")
        elif img.dtype != np.uint8:
            raise N.cv2_error("images must be uint8")
        if len(img.shape) not in (2, 3) or (len(img.shape) == 3 and img.shape[2] != 3):
            raise N.cv2_error("images must be uint8, HxW (gray) or HxWx3 (BGR)")
        if tuple(img.shape[:2]) != (self.img_size[1], self.img_size[0]):
            raise N.cv2_error("image size differs from the camera's img_size")

    def _prepare_device(self, eng, img_left, img_right, preprocessed, key="prep"):
        """Host frames ([H,W] / [H,W,3], or batches [S,H,W] / [S,H,W,3]) -> rectified gray device tensors [S,H,W]
        (ref: src/openVO/stereo_camera.py:44-50: cvtColor if colour, remap unless preprocessed — both on the device)."""
        out = []
        for side, img in (("left", img_left), ("right", img_right)):
            hw = (self.img_size[1], self.img_size[0])
            if isinstance(img, (list, tuple)):   # one array (or pinned torch tensor) per sequence
                for fr in img:
                    self._check(fr if hasattr(fr, "data_ptr") else np.asarray(fr))
                colour = len(img[0].shape) == 3
            else:
                img = np.asarray(img)
                if img.shape[:2] == hw and img.ndim in (2, 3):  # a single frame
                    img = img[None]
                self._check(img[0])
                colour = img.ndim == 4
            dev = eng.upload(img, key + "_" + side)
            if colour or not preprocessed:
                dev = eng.rectify(dev, None if preprocessed else self._maps(eng, side))
            out.append(dev)
        if out[0].shape != out[1].shape:
            raise N.cv2_error("left and right must be of equal size")
        return out[0], out[1]

    def compute_3d(self, img_left, img_right, preprocessed=False):
        """ref: src/openVO/stereo_camera.py:43-55 -> (img_3d f32 H'xW'x3, disparity f32 H'xW', img_left u8 H'xW')."""
        eng = self.engine()
        l, r = self._prepare_device(eng, img_left, img_right, preprocessed, key="c3d")
        disp, _ = eng.disparity_post(eng.sgbm(l, r))
        xyz = eng.reproject(disp[0])
        return xyz.cpu().numpy(), disp[0].cpu().numpy(), eng.crop(l)[0].cpu().numpy()
