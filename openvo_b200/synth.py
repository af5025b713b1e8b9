"""Deterministic synthetic rectified stereo (numpy only) for tests and bench.py — SURVEY.md §8(d) / Appendix B.

Not part of the hot path.  Two generators:
  * ``kat_pair``      — the Appendix-B known-answer pair (constant disparity, corner-dense blocky texture);
  * ``make_sequence`` — a multi-plane scene (background + occluding foreground rectangles, so the uniqueness / LR /
                        speckle paths fire) seen by a stereo rig moving along a known SE(3) trajectory, rendered by
                        ray/plane intersection with a painter's algorithm.
"""
import numpy as np

BASELINE_M = 0.537


def kat_pair(W, H, d=24, seed=12345, blk=4):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (H // blk + 1, (W + 64) // blk + 1), dtype=np.uint8)
    tex = np.kron(t, np.ones((blk, blk), np.uint8))[:H, :W + 64]
    tex = (tex.astype(np.int32) * 3 // 4 + rng.integers(0, 64, tex.shape)).astype(np.uint8)
    return np.ascontiguousarray(tex[:, 32:32 + W]), np.ascontiguousarray(tex[:, 32 + d:32 + d + W])


def camera_args(W, H, num_disparities):
    """Constructor arguments of StereoCamera for the synthetic zero-distortion rig."""
    f = 718.856 * (W / 1241.0)
    K = np.array([[f, 0, (W - 1) / 2.0], [0, f, (H - 1) / 2.0], [0, 0, 1]], np.float64)
    dist = np.zeros(5)
    rect = {"R": np.eye(3), "T": np.array([-BASELINE_M, 0.0, 0.0])}
    sgbm = dict(minDisparity=0, numDisparities=int(num_disparities), blockSize=5, P1=200, P2=800, disp12MaxDiff=1,
                preFilterCap=63, uniquenessRatio=10, speckleWindowSize=100, speckleRange=2)
    return dict(K_left=K, dist_left=dist, K_right=K.copy(), dist_right=dist.copy(), rect_params=rect, sgbm_params=sgbm,
                img_size=(W, H))


def camera_args_distorted(W, H, num_disparities):
    """A rig with lens distortion, slightly different intrinsics and a small relative rotation: exercises cv2.remap
    rectification (preprocessed_frames=False, the reference's default) and a non-trivial valid-pixel ROI (B1)."""
    a = camera_args(W, H, num_disparities)
    a["dist_left"] = np.array([-0.12, 0.03, 0.0008, -0.0006, 0.0])
    a["dist_right"] = np.array([-0.10, 0.02, -0.0005, 0.0007, 0.0])
    a["K_right"] = a["K_right"] + np.array([[2.0, 0, 1.5], [0, 2.0, -1.0], [0, 0, 0]])
    ax = np.array([0.004, -0.006, 0.002])
    th = np.linalg.norm(ax)
    k = ax / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    a["rect_params"] = {"R": np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx, "T": np.array([-BASELINE_M, 0.002, 0.001])}
    return a


def to_bgr(gray, seed=0):
    """A colour image whose cv2 BGR2GRAY conversion is close to `gray` (channels differ by small seeded offsets)."""
    rng = np.random.default_rng(seed)
    g = gray.astype(np.int16)
    off = rng.integers(-12, 13, gray.shape + (3,))
    return np.clip(g[..., None] + off, 0, 255).astype(np.uint8)


def _texture(rng, n=2048, blk=4):
    t = rng.integers(0, 256, (n // blk, n // blk)).astype(np.float32)
    tex = np.kron(t, np.ones((blk, blk), np.float32)) * 0.75 + rng.integers(0, 64, (n, n)).astype(np.float32)
    return tex


def _sample(tex, u, v):
    n = tex.shape[0]
    u0, v0 = np.floor(u), np.floor(v)
    fu, fv = (u - u0).astype(np.float32), (v - v0).astype(np.float32)
    iu, iv = u0.astype(np.int64) % n, v0.astype(np.int64) % n
    iu1, iv1 = (iu + 1) % n, (iv + 1) % n
    return (tex[iv, iu] * (1 - fu) * (1 - fv) + tex[iv, iu1] * fu * (1 - fv) + tex[iv1, iu] * (1 - fu) * fv +
            tex[iv1, iu1] * fu * fv)


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def make_scene(W, seed=12345):
    rng = np.random.default_rng(seed)
    f = 718.856 * (W / 1241.0)
    # one texel per pixel on every plane, so the (wrapping) texture must be wider than the image: a 2048-texel texture repeats
    # inside a 3840-pixel frame, and the duplicated patches turn half of the 4K matches into 2048-pixel outliers
    n = 2048 if W <= 2040 else 4096
    planes = [dict(Z=20.0, ext=None, tex=_texture(rng, n), cell=20.0 / f)]
    for Z, cx, cy, hw, hh in ((14.0, -6.0, -0.5, 3.0, 1.6), (10.0, 4.0, 0.8, 1.8, 1.2), (7.5, -2.0, 0.6, 1.1, 0.8),
                              (6.0, 1.4, -0.4, 0.7, 0.5), (5.0, -0.9, 0.7, 0.4, 0.3), (16.0, 7.0, -2.0, 3.0, 2.0)):
        planes.append(dict(Z=Z, ext=(cx - hw, cx + hw, cy - hh, cy + hh), tex=_texture(rng, n if 2 * hw * f / Z > 2040 else 2048), cell=Z / f))
    planes.sort(key=lambda p: -p["Z"])
    return planes


def render(planes, W, H, R, t, noise_rng=None):
    """Image of the scene from a camera with X_c = R X_w + t."""
    f = 718.856 * (W / 1241.0)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dc = np.stack([(u - cx) / f, (v - cy) / f, np.ones_like(u)], -1)
    dw = dc @ R  # = R^T dc per pixel
    o = -R.T @ t
    img = np.zeros((H, W), np.float32)
    for p in planes:
        ys, xs = slice(0, H), slice(0, W)
        if p["ext"] is not None:
            # a bounded plane is only evaluated inside the (padded) bounding box of its projected corners: same values, less work
            x0, x1, y0, y1 = p["ext"]
            c = np.array([[x0, y0, p["Z"]], [x1, y0, p["Z"]], [x0, y1, p["Z"]], [x1, y1, p["Z"]]]) @ R.T + t
            if (c[:, 2] > 1e-6).all():
                uu, vv = f * c[:, 0] / c[:, 2] + cx, f * c[:, 1] / c[:, 2] + cy
                xs = slice(int(np.clip(np.floor(uu.min()) - 3, 0, W)), int(np.clip(np.ceil(uu.max()) + 4, 0, W)))
                ys = slice(int(np.clip(np.floor(vv.min()) - 3, 0, H)), int(np.clip(np.ceil(vv.max()) + 4, 0, H)))
                if xs.start >= xs.stop or ys.start >= ys.stop:
                    continue
        d = dw[ys, xs]
        lam = (p["Z"] - o[2]) / d[..., 2]
        X, Y = o[0] + lam * d[..., 0], o[1] + lam * d[..., 1]
        hit = lam > 0
        if p["ext"] is not None:
            x0, x1, y0, y1 = p["ext"]
            hit &= (X >= x0) & (X <= x1) & (Y >= y0) & (Y <= y1)
        val = _sample(p["tex"], X / p["cell"], Y / p["cell"])
        img[ys, xs] = np.where(hit, val, img[ys, xs])
    if noise_rng is not None:
        img = img + noise_rng.normal(0, 1.0, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_sequence(W, H, nframes, seed=12345, step=(0.01, 0.0, 0.05), yaw_step=0.002):
    """-> lefts u8 [n,H,W], rights u8 [n,H,W], list of ground-truth c_T_w (4x4) per frame."""
    planes = make_scene(W, seed)
    rng = np.random.default_rng(seed + 1)
    lefts, rights, poses = [], [], []
    for i in range(nframes):
        Rwc = _rot_y(yaw_step * i)            # camera orientation in the world
        pos = np.array(step) * i              # camera position in the world
        R = Rwc.T
        t = -R @ pos
        lefts.append(render(planes, W, H, R, t, rng))
        rights.append(render(planes, W, H, R, t - np.array([BASELINE_M, 0, 0]), rng))
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        poses.append(T)
    return np.stack(lefts), np.stack(rights), poses
