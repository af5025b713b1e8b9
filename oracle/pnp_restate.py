"""ORACLE (test infrastructure only — never imported by the product path).

Specification-by-restatement of the OPT-IN robust pose stage (north-star stage 5): batched P3P RANSAC
over a fixed, seed-determined hypothesis schedule followed by a small Levenberg-Marquardt refinement of the reprojection error
on the inliers.  The reference does NOT do this (its pose is cv2.estimateAffine3D / Umeyama, ref: src/openVO/stereo_odometer.py:204;
SURVEY.md §0 D1), and OpenCV's own solvePnPRansac draws its samples from cv::RNG, whose schedule is not reproducible from outside —
so there is no reference output to be bit-compatible with: **parity unpinned**.  What is pinned is (a) this numpy restatement vs
the CUDA kernels (same schedule, poses within 1e-4 rad / 1e-3 relative translation, identical inlier sets on non-borderline data),
and (b) sanity against cv2.solvePnPRansac and ground truth on synthetic data (tests/test_oracle.py).

Schedule: hypothesis h uses correspondences idx_k = splitmix64(seed, h, counter) mod M, k = 0..3, skipping repeats; the first three
feed P3P, the fourth disambiguates.  P3P: law-of-cosines system in depth ratios u = s2/s1, v = s3/s1 reduced to a quartic in v
(u is linear in v after subtracting two of the equations); real roots in [1/16, 16]; rigid alignment of the three points by triads.
Score: number of correspondences with squared reprojection error < thr^2; best = most inliers, ties to the lowest h.
Refinement: 20 LM iterations (lambda0 = 1e-3, x0.1 on success, x10 on failure) on the inliers of the best hypothesis, pose
parametrised as (rotation vector increment applied on the left, translation).
"""
import numpy as np

MASK64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def sample4(seed, h, m):
    out, ctr = [], 0
    while len(out) < 4:
        r = splitmix64((seed * 0x100000001B3 + h * 0x9E3779B1 + ctr) & MASK64) % m
        ctr += 1
        if r not in out:
            out.append(int(r))
        if ctr > 64:
            return None
    return out


def _polymul(a, b):
    return np.convolve(a, b)  # ascending coefficients


def p3p(P, j):
    """P: 3x3 world points, j: 3x3 unit bearings (rows).  -> list of (R, t) with X_cam = R X_world + t."""
    a2 = np.sum((P[1] - P[2]) ** 2)
    b2 = np.sum((P[0] - P[2]) ** 2)
    c2 = np.sum((P[0] - P[1]) ** 2)
    ca, cb, cg = j[1] @ j[2], j[0] @ j[2], j[0] @ j[1]
    q = np.array([1.0, -2 * cb, 1.0])                       # 1 - 2 cb v + v^2 (ascending)
    N = np.array([b2, 0.0, -b2]) + (a2 - c2) * q            # numerator of u
    Dn = np.array([2 * b2 * cg, -2 * b2 * ca])              # denominator of u
    poly = b2 * _polymul(N, N) - 2 * b2 * cg * np.append(_polymul(N, Dn), 0.0) + _polymul(np.array([b2, 0, 0]) - c2 * q, _polymul(Dn, Dn))
    sols = []
    if abs(poly[-1]) < 1e-300:
        return sols
    for root in np.roots(poly[::-1]):
        if abs(root.imag) > 1e-9 * max(1.0, abs(root.real)):
            continue
        v = root.real
        if not (1.0 / 16 <= v <= 16.0):
            continue
        den = Dn[0] + Dn[1] * v
        if abs(den) < 1e-12 * b2:
            continue
        u = (N[0] + N[1] * v + N[2] * v * v) / den
        if u <= 0:
            continue
        qv = 1 - 2 * cb * v + v * v
        if qv <= 0:
            continue
        s1 = np.sqrt(b2 / qv)
        Q = np.stack([s1 * j[0], u * s1 * j[1], v * s1 * j[2]])

        def triad(X):
            e1 = X[1] - X[0]
            e1 = e1 / np.linalg.norm(e1)
            e3 = np.cross(e1, X[2] - X[0])
            e3 = e3 / np.linalg.norm(e3)
            return np.stack([e1, np.cross(e3, e1), e3], 1)
        R = triad(Q) @ triad(P).T
        sols.append((R, Q[0] - R @ P[0]))
    return sols


def project(R, t, X, f, cx, cy):
    Y = X @ R.T + t
    z = Y[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.stack([f * Y[:, 0] / z + cx, f * Y[:, 1] / z + cy], 1), z


def rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3) + np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def pnp_ransac(X, uv, f, cx, cy, iters=512, thr=8.0, seed=0, lm_iters=20):
    """X: [m,3] points in frame A, uv: [m,2] pixels in frame B.  -> dict(R, t, inliers mask, best hypothesis) or None."""
    X = np.asarray(X, np.float64)
    uv = np.asarray(uv, np.float64)
    m = len(X)
    if m < 4:
        return None
    bear = np.stack([(uv[:, 0] - cx) / f, (uv[:, 1] - cy) / f, np.ones(m)], 1)
    bear /= np.linalg.norm(bear, axis=1, keepdims=True)
    best = (-1, None, None, None)
    for h in range(iters):
        idx = sample4(seed, h, m)
        if idx is None:
            continue
        cand = p3p(X[idx[:3]], bear[idx[:3]])
        pick, pick_err = None, np.inf
        for R, t in cand:
            p, z = project(R, t, X[idx[3:4]], f, cx, cy)
            if z[0] <= 0 or not np.isfinite(p).all():
                continue
            e = float(np.sum((p[0] - uv[idx[3]]) ** 2))
            if e < pick_err:
                pick, pick_err = (R, t), e
        if pick is None:
            continue
        p, z = project(pick[0], pick[1], X, f, cx, cy)
        inl = (z > 0) & (np.sum((p - uv) ** 2, 1) < thr * thr)
        n = int(inl.sum())
        if n > best[0]:
            best = (n, pick, inl, h)
    if best[1] is None or best[0] < 4:
        return None
    R, t = best[1]
    inl = best[2]
    Xi, ui = X[inl], uv[inl]

    def cost(R, t):
        p, z = project(R, t, Xi, f, cx, cy)
        return float(np.sum((p - ui) ** 2))
    lam, c0 = 1e-3, cost(R, t)
    for _ in range(lm_iters):
        Y = Xi @ R.T + t
        x, y, z = Y[:, 0], Y[:, 1], Y[:, 2]
        r = np.stack([f * x / z + cx - ui[:, 0], f * y / z + cy - ui[:, 1]], 1).reshape(-1)
        J = np.zeros((len(Xi), 2, 6))
        # d(proj)/dY
        dY = np.zeros((len(Xi), 2, 3))
        dY[:, 0, 0] = f / z
        dY[:, 0, 2] = -f * x / (z * z)
        dY[:, 1, 1] = f / z
        dY[:, 1, 2] = -f * y / (z * z)
        # Y' = exp(w) Y + dt  ->  dY/dw = -[Y]x, dY/dt = I
        for k in range(len(Xi)):
            Yx = np.array([[0, -z[k], y[k]], [z[k], 0, -x[k]], [-y[k], x[k], 0]])
            J[k, :, :3] = dY[k] @ (-Yx)
            J[k, :, 3:] = dY[k]
        J = J.reshape(-1, 6)
        A, g = J.T @ J, J.T @ r
        step = np.linalg.solve(A + lam * np.diag(np.diag(A)), -g)
        Rn = rodrigues(step[:3]) @ R
        tn = rodrigues(step[:3]) @ t + step[3:]
        c1 = cost(Rn, tn)
        if np.isfinite(c1) and c1 < c0:
            R, t, c0, lam = Rn, tn, c1, lam * 0.1
        else:
            lam *= 10.0
    return dict(R=R, t=t, inliers=inl, best=best[3], n_inliers=best[0], cost=c0)
