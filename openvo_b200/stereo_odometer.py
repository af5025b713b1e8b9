"""StereoOdometer — same constructor, constants, attributes and per-frame ``update`` as the reference class
(ref: src/openVO/stereo_odometer.py:4-226), with every per-frame stage on the B200.

``update`` keeps all frame products on the device (disparity, keypoints, descriptors; the 3-D image is never
materialised: the point lookup reprojects on the fly) and reads back 144 bytes per frame pair.  The reference's
public state (``current_kps`` as cv2.KeyPoint tuples, ``current_3d`` as an HxWx3 float image, ...) is materialised
lazily when read.  Control flow — the skip / fall-back state machine B4 — is host Python exactly as in the reference.
"""
import numpy as np

from . import _native as N


def _keypoints(kp):
    import cv2
    return tuple(cv2.KeyPoint(float(r[0]), float(r[1]), float(r[2]), float(r[3]), float(r[4]), int(r[5]), -1) for r in kp)


class _OrbHandle:
    """Stands in for the cv2.ORB object in ``StereoOdometer.orb`` (ref: src/openVO/stereo_odometer.py:22,117)."""

    def __init__(self, od):
        self._od = od

    def detectAndCompute(self, image, mask=None):
        eng = self._od._engine()
        image = np.asarray(image)
        if image.dtype != np.uint8 or image.shape != (eng.ch, eng.cw):
            raise N.cv2_error("ORB: expected a uint8 image of the camera's cropped size %dx%d" % (eng.cw, eng.ch))
        img = eng.upload(image[None], "orb_img")
        m = None
        if mask is not None:
            mask = np.asarray(mask)
            if mask.dtype != np.uint8 or mask.shape != image.shape:
                raise N.cv2_error("ORB: mask must be uint8 and of the image's size")
            m = eng.upload(mask[None], "orb_mask")
        kp, desc, n = eng.orb(img, m)
        if n[0] == 0:
            return (), None
        return _keypoints(kp[0, :n[0]].cpu().numpy()), desc[0, :n[0]].cpu().numpy()

    def getMaxFeatures(self):
        return self._od._nfeatures


class _MatcherHandle:
    """Stands in for cv2.BFMatcher(NORM_HAMMING) in ``StereoOdometer.matcher`` (ref: stereo_odometer.py:22,163)."""

    def __init__(self, od):
        self._od = od

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2):
        import cv2
        import torch
        if k != 2:
            raise ValueError("only k=2 is on the hot path")
        eng = self._od._engine()
        q = np.ascontiguousarray(queryDescriptors, np.uint8)
        t = np.ascontiguousarray(trainDescriptors, np.uint8)
        nn = torch.empty((len(q), 4), dtype=torch.int32, device=eng.device)
        eng.knn2(torch.from_numpy(q).to(eng.device), len(q), torch.from_numpy(t).to(eng.device), len(t), out=nn)
        out = []
        for i, (i0, d0, i1, d1) in enumerate(nn.cpu().numpy().tolist()):
            row = []
            if i0 >= 0:
                row.append(cv2.DMatch(i, i0, 0, float(d0)))
            if i1 >= 0:
                row.append(cv2.DMatch(i, i1, 0, float(d1)))
            out.append(tuple(row))
        return tuple(out)


class StereoOdometer:
    # ref: src/openVO/stereo_odometer.py:6-12 (read through ``self.`` so subclass / instance overrides work)
    MIN_VALID_DISPARITY = 4
    MAX_VALID_DISPARITY = 100
    MAX_DISTANCE_CHANGE = 1
    MAX_ROTATION_CHANGE = np.pi / 3

    def __init__(self, stereo_camera, nfeatures=500, match_threshold=0.8, rigidity_threshold=0, outlier_threshold=0,
                 preprocessed_frames=False, min_matches=10, cross_check=False, pose_method="umeyama", ransac_iters=512,
                 ransac_reproj_px=8.0, ransac_seed=0, _max_batch=1, _engine_tag=0):
        self.stereo = stereo_camera
        self._nfeatures = nfeatures
        self._max_batch = _max_batch  # >1 only when driven by openvo_b200.batch.BatchOdometer
        self._engine_tag = _engine_tag
        self._cur = None   # last committed frame (device resident)
        self._prev = None  # the one before
        self.orb, self.matcher = _OrbHandle(self), _MatcherHandle(self)
        self.match_threshold, self.rigidity_threshold = match_threshold, rigidity_threshold
        self.outlier_threshold, self.preprocessed_frames = outlier_threshold, preprocessed_frames
        self.min_matches = min_matches
        # opt-in extension: left-right cross-check of the matches (the reference's "TODO crosscheck", stereo_odometer.py:21)
        self.cross_check = cross_check
        # opt-in extension (north-star stage 5): "pnp_ransac" = batched P3P RANSAC + LM instead of the reference's Umeyama alignment
        if pose_method not in ("umeyama", "pnp_ransac"):
            raise ValueError("pose_method must be 'umeyama' (the reference's behaviour) or 'pnp_ransac'")
        if pose_method == "pnp_ransac" and (rigidity_threshold > 0 or outlier_threshold > 0):
            import warnings
            warnings.warn("pose_method='pnp_ransac' is ignored while rigidity_threshold / outlier_threshold are set: the filtered "
                          "path ends in the reference's Umeyama alignment (ref: src/openVO/stereo_odometer.py:177-205)")
        self.pose_method, self.ransac_iters = pose_method, ransac_iters
        self.ransac_reproj_px, self.ransac_seed = ransac_reproj_px, ransac_seed
        self.skipped_frames = 0
        self.c_T_w = np.eye(4)
        self.c_T_w_prev = np.eye(4)
        self.skip_cause = ""
        self.last_match_count = 0
        self.last_T = None
        self.last_mode = 0  # 0: frame not committed / first frame, 1: aligned to the current frame, 2: fall-back to the previous one

    def _engine(self):
        return self.stereo.engine(self._nfeatures, self._max_batch, float(self.MIN_VALID_DISPARITY), float(self.MAX_VALID_DISPARITY), tag=self._engine_tag)

    # ---- lazily materialised public state (reference types) --------------------------------------------------------------
    def _host(self, frame, what):
        if frame is None:
            return None
        h = frame._host
        if what not in h:
            if frame.disp is None and what != "kps":   # a frame assembled from caller-assigned products only
                return None
            if what == "img":
                h[what] = frame.img.cpu().numpy()
            elif what == "disparity":
                h[what] = frame.disp.cpu().numpy()
            elif what == "3d":
                h[what] = self._engine().reproject(frame.disp).cpu().numpy()
            elif what == "kp_array":
                h[what] = frame.kp[:frame.n_kp].cpu().numpy()
            elif what == "kps":
                h[what] = _keypoints(self._host(frame, "kp_array"))
            elif what == "desc":
                h[what] = frame.desc[:frame.n_kp].cpu().numpy()
        return h[what]

    # The reference keeps these as plain attributes (ref: src/openVO/stereo_odometer.py:24-31,107-113): readable AND assignable.
    # Reading materialises the device-resident product on the host (cached); assigning replaces the product, and the device
    # copy is rebuilt from the host value before the frame is next used for matching.
    def _state_property(which, what):
        def get(self):
            return self._host(getattr(self, which), what)

        def put(self, value):
            frame = getattr(self, which)
            if value is None:
                if what == "img":            # the reference tests `current_img is None` / `prev_img is None` for "no frame yet"
                    setattr(self, which, None)
                return
            if frame is None:
                from .engine import Frame
                frame = Frame(None, None, None, None, 0)
                setattr(self, which, frame)
            frame._host[what] = value
            if what == "kps":
                frame._host.pop("kp_array", None)
            frame._dirty = True
        return property(get, put)

    current_img = _state_property("_cur", "img")
    current_disparity = _state_property("_cur", "disparity")
    current_3d = _state_property("_cur", "3d")
    current_kps = _state_property("_cur", "kps")
    current_desc = _state_property("_cur", "desc")
    prev_img = _state_property("_prev", "img")
    prev_disparity = _state_property("_prev", "disparity")
    prev_3d = _state_property("_prev", "3d")
    prev_kps = _state_property("_prev", "kps")
    prev_desc = _state_property("_prev", "desc")
    del _state_property

    def _device_frame(self, frame):
        """Rebuild the device copy of a frame whose host-side products were assigned by the caller."""
        if not getattr(frame, "_dirty", False):
            return frame
        import torch
        eng = self._engine()
        h = frame._host
        dev = eng.device
        if "img" in h:
            frame.img = torch.from_numpy(np.ascontiguousarray(h["img"], np.uint8)).to(dev)
        if "disparity" in h:
            frame.disp = torch.from_numpy(np.ascontiguousarray(h["disparity"], np.float32)).to(dev)
        if "kps" in h or "kp_array" in h:
            arr = h.get("kp_array")
            if arr is None:
                arr = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in h["kps"]], np.float32).reshape(-1, 6)
            if len(arr) > eng.kp_cap:
                raise N.cv2_error("more keypoints than the odometer's capacity")
            kp = torch.zeros((eng.kp_cap, N.KP_FIELDS), dtype=torch.float32, device=dev)
            kp[:len(arr)] = torch.from_numpy(np.ascontiguousarray(arr, np.float32)).to(dev)
            frame.kp, frame.n_kp = kp, len(arr)
        if "desc" in h and h["desc"] is not None:
            dsc = np.ascontiguousarray(h["desc"], np.uint8)
            desc = torch.zeros((eng.kp_cap, 32), dtype=torch.uint8, device=dev)
            desc[:len(dsc)] = torch.from_numpy(dsc).to(dev)
            frame.desc = desc
        frame._dirty = False
        return frame

    # ---- helpers that are part of the reference's method surface --------------------------------------------------------
    def feature_mask(self, disparity):
        # ref: src/openVO/stereo_odometer.py:38-41 (the hot path fuses this into ovo_disparity_post)
        ok = (disparity >= self.MIN_VALID_DISPARITY) * (disparity <= self.MAX_VALID_DISPARITY)
        return ok.astype(np.uint8) * 255

    def valid_distance_change(self, prev_kp_idx, current_kp_idx):
        # ref: src/openVO/stereo_odometer.py:43-48 (dead code in the reference: guarded by ``if (False)``)
        px, py = self.prev_kps[prev_kp_idx].pt
        cx, cy = self.current_kps[current_kp_idx].pt
        gap = np.linalg.norm(self.prev_3d[int(py)][int(px)]) - np.linalg.norm(self.current_3d[int(cy)][int(cx)])
        return gap <= self.MAX_DISTANCE_CHANGE * (self.skipped_frames + 1)

    def bilinear_interpolate_pixels(self, img, x, y):
        # ref: src/openVO/stereo_odometer.py:50-79, host helper for callers that hold a numpy 3-D image; ``update`` uses the
        # fused device lookup (ovo_match_points) instead.
        fx, fy = int(x), int(y)
        h, w = img.shape[0:2]
        rx, ry = x - fx, y - fy
        num, den = 0, 0
        for px, py, wt in ((fx, fy, (1 - rx) * (1 - ry)), (fx, fy + 1, (1 - rx) * ry), (fx + 1, fy, rx * (1 - ry)),
                           (fx + 1, fy + 1, rx * ry)):
            if px < w and py < h and not np.isinf(img[py, px]).any():
                num = num + wt * img[py, px]
                den = den + wt
        return num / den

    def save_frame_update(self, next_img, next_disp=None, next_3d=None, next_kps=None, next_desc=None):
        """ref: src/openVO/stereo_odometer.py:107-113 — same five arguments (host products of the next frame): the current
        frame becomes the previous one, the given products the current one.  (The per-frame path commits its device-resident
        frame through the same rotation, without materialising anything on the host.)"""
        from .engine import Frame
        if isinstance(next_img, Frame):
            frame = next_img
        else:
            frame = Frame(None, None, None, None, 0)
            frame._host.update({"img": next_img, "disparity": next_disp, "3d": next_3d, "kps": tuple(next_kps), "desc": next_desc})
            frame._dirty = True
        self._prev, self._cur = self._cur, frame

    def _hooks_overridden(self):
        # the reference dispatches these through `self.`: a subclass (or instance) override must take effect in update()
        base = StereoOdometer
        return any(getattr(type(self), n) is not getattr(base, n) or n in self.__dict__
                   for n in ("feature_mask", "point_clouds", "point_cloud_transform", "bilinear_interpolate_pixels", "rigid_body_filter",
                             "save_frame_update"))

    def _update_through_hooks(self, img_left, img_right):
        """update() exactly as the reference writes it (ref: src/openVO/stereo_odometer.py:115-160), every step dispatched
        through `self.` on host-side products: used when a subclass overrides one of the steps.  Each step still runs on the
        device (compute_3d, ORB, knnMatch, estimateAffine3D are the C-ABI seams); only the glue is host Python."""
        next_3d, next_disp, next_img = self.stereo.compute_3d(img_left, img_right, preprocessed=self.preprocessed_frames)
        next_kps, next_desc = self.orb.detectAndCompute(next_img, self.feature_mask(next_disp))
        if len(next_kps) < self.min_matches:
            self.skipped_frames += 1
            self.skip_cause = "keypoints"
            return False
        if self.current_img is None:
            self.save_frame_update(next_img, next_disp, next_3d, next_kps, next_desc)
            return True
        T = None
        current_pts, next_pts = self.point_clouds(self.current_kps, next_kps, self.current_desc, next_desc, self.current_3d, next_3d)
        if current_pts is None:
            self.skip_cause = "matches"
        else:
            T = self.point_cloud_transform(current_pts, next_pts)
            if T is not None:
                self.c_T_w_prev = self.c_T_w
                self.c_T_w = T @ self.c_T_w
        if T is None and self.prev_img is not None:
            prev_pts, next_pts = self.point_clouds(self.prev_kps, next_kps, self.prev_desc, next_desc, self.prev_3d, next_3d)
            if prev_pts is None:
                self.skip_cause = "matches"
            else:
                T = self.point_cloud_transform(prev_pts, next_pts)
                if T is not None:
                    T_prev = self.c_T_w_prev
                    self.c_T_w_prev = self.c_T_w
                    self.c_T_w = T @ T_prev
                    self.skipped_frames = 0
        if T is None:
            self.skipped_frames += 1
            return False
        self.skipped_frames = 0
        self.save_frame_update(next_img, next_disp, next_3d, next_kps, next_desc)
        return True

    # ---- the per-frame call --------------------------------------------------------------------------------------------------
    def update(self, img_left, img_right):
        """ref: src/openVO/stereo_odometer.py:115-160."""
        if self._hooks_overridden():
            return self._update_through_hooks(img_left, img_right)
        eng = self._engine()
        left, right = self.stereo._prepare_device(eng, img_left, img_right, self.preprocessed_frames, key="upd")
        frame = eng.frames(left, right)[0]
        return self._advance(frame)

    def _advance(self, frame, first=None):
        """The state machine of ``update`` for an already extracted frame.  ``first`` optionally carries the device result
        of the (current -> frame) pair step when a batch driver has already launched it."""
        self.last_mode = 0
        if frame.n_kp < self.min_matches:
            self.skipped_frames += 1
            self.skip_cause = "keypoints"
            return False
        if self._cur is None:
            self.save_frame_update(frame)
            return True
        T = self._relative(self._cur, frame, first)
        if T is not None:
            self.c_T_w_prev = self.c_T_w
            self.c_T_w = T @ self.c_T_w
            self.last_mode = 1
        if T is None and self._prev is not None:
            T = self._relative(self._prev, frame)
            if T is not None:
                older = self.c_T_w_prev
                self.c_T_w_prev = self.c_T_w
                self.c_T_w = T @ older
                self.skipped_frames = 0
                self.last_mode = 2
        if T is None:
            self.skipped_frames += 1
            return False
        self.skipped_frames = 0
        self.last_T = T
        self.save_frame_update(frame)
        return True

    def _relative(self, a, b, result=None):
        """point_clouds + point_cloud_transform for device frames a -> b."""
        eng = self._engine()
        a, b = self._device_frame(a), self._device_frame(b)
        if b.n_kp < 2:
            raise IndexError("tuple index out of range")  # the reference indexes m[1] (ref: stereo_odometer.py:164)
        slot = 0
        if result is None:
            n, bad, out = eng.pair(a, b, self.match_threshold, self.cross_check)
        else:
            slot, (n, bad, out) = result
        self.last_match_count = n
        if n < self.min_matches:
            self.skip_cause = "matches"
            return None
        if bad:
            raise ZeroDivisionError("division by zero")  # ref: stereo_odometer.py:79 with every tap skipped
        if self.rigidity_threshold > 0 or self.outlier_threshold > 0:
            import torch
            cnt = torch.tensor([n], dtype=torch.int32, device=eng.device)
            return self._filtered(eng.pts1[slot], eng.pts2[slot], cnt)
        if n < 10:
            self.skip_cause = "rigidity"
        if self.pose_method == "pnp_ransac":
            out = eng.pnp_ransac(b, slot, n, self.ransac_iters, self.ransac_reproj_px, self.ransac_seed)
            if not (out[12] >= self.min_matches):   # too few inliers (or no valid hypothesis)
                self.skip_cause = "ransac"
                return None
        return self._gate(out)

    def _gate(self, out):
        # ref: src/openVO/stereo_odometer.py:204-223
        T = np.eye(4)
        T[:3, :4] = out[:12].reshape(3, 4)
        if np.isnan(T).any():
            self.skip_cause = "nan"
            return None
        k = self.skipped_frames + 1
        far = np.linalg.norm(T[0:3, 3]) > self.MAX_DISTANCE_CHANGE * k
        spun = out[13] > self.MAX_ROTATION_CHANGE * k
        if far:
            self.skip_cause = "bigdist"
        if spun:
            self.skip_cause = "bigrot"
        return None if (far or spun) else T

    def point_clouds(self, kps1, kps2, desc1, desc2, im3d1, im3d2):
        """ref: src/openVO/stereo_odometer.py:162-175 for callers that hold host-side products."""
        matches = self.matcher.knnMatch(desc1, desc2, k=2)
        matches = [m[0] for m in matches if m[0].distance < self.match_threshold * m[1].distance]
        if len(matches) < self.min_matches:
            return None, None
        pts1 = [self.bilinear_interpolate_pixels(im3d1, *kps1[m.queryIdx].pt) for m in matches]
        pts2 = [self.bilinear_interpolate_pixels(im3d2, *kps2[m.trainIdx].pt) for m in matches]
        return np.array(pts1), np.array(pts2)

    def rigid_body_filter(self, prev_pts, pts):
        """ref: src/openVO/stereo_odometer.py:82-105 for callers that hold host arrays; ``update`` runs ovo_rigid_body_filter."""
        d_now = np.linalg.norm(pts[:, None, :] - pts[None, :, :], axis=2)
        d_old = np.linalg.norm(prev_pts[:, None, :] - prev_pts[None, :, :], axis=2)
        consistency = (np.abs(d_now - d_old) < self.rigidity_threshold).astype(int)
        clique = np.zeros(len(pts), int)
        degree = consistency.sum(0)
        first = int(np.argmax(degree))
        clique[first] = 1
        compatible = consistency[first]
        for _ in range(len(pts)):
            cand = (compatible - clique).astype(int)
            if cand.sum() == 0:
                break
            clique[int(np.argmax(degree * cand))] = 1
            compatible = (consistency @ clique >= clique.sum()).astype(int)
        return clique

    def _filtered(self, pts1_dev, pts2_dev, count_dev):
        """ref: src/openVO/stereo_odometer.py:177-223 with the optional filters, on device buffers."""
        n1, n2, out = self._engine().filtered_transform(pts1_dev, pts2_dev, count_dev, self.rigidity_threshold, self.outlier_threshold,
                                                        self.min_matches)
        rigidity_cause = n1 < 10
        if rigidity_cause:
            self.skip_cause = "rigidity"
        if out is None:
            if not rigidity_cause:
                self.skip_cause = "outlier"
            return None
        return self._gate(out)

    def point_cloud_transform(self, current_pts, next_pts):
        """ref: src/openVO/stereo_odometer.py:177-223 for callers that hold host-side point sets (numpy float32 [m,3])."""
        import torch
        eng = self._engine()
        m = len(current_pts)
        if m > eng.kp_cap:
            raise ValueError("more points than the keypoint capacity")
        p1 = torch.zeros((eng.kp_cap, 3), dtype=torch.float32, device=eng.device)
        p2 = torch.zeros((eng.kp_cap, 3), dtype=torch.float32, device=eng.device)
        if m:
            p1[:m] = torch.from_numpy(np.ascontiguousarray(current_pts, np.float32)).to(eng.device)
            p2[:m] = torch.from_numpy(np.ascontiguousarray(next_pts, np.float32)).to(eng.device)
        cnt = torch.tensor([m], dtype=torch.int32, device=eng.device)
        return self._filtered(p1, p2, cnt)

    def current_pose(self):
        # ref: src/openVO/stereo_odometer.py:225-226
        return np.linalg.inv(self.c_T_w)
