import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


def occluded_pair(W, H, d=7):
    """KAT pair with a block of the right image replaced by noise (fires uniqueness / LR / speckle paths)."""
    from openvo_b200 import synth
    L, R = synth.kat_pair(W, H, d=d)
    rng = np.random.default_rng(1)
    R = R.copy()
    R[H // 3:H // 2, W // 3:W // 2] = rng.integers(0, 256, (H // 2 - H // 3, W // 2 - W // 3))
    return L, R


SGBM_CASES = [
    (200, 60, 32, {}),
    (320, 48, 64, dict(blockSize=3, P1=72, P2=288, uniquenessRatio=0, speckleWindowSize=0)),
    (240, 64, 48, dict(blockSize=7, preFilterCap=15, uniquenessRatio=15, disp12MaxDiff=2, speckleWindowSize=50, speckleRange=1)),
    (240, 64, 16, dict(blockSize=11, P1=100, P2=1000, disp12MaxDiff=-1)),
    # cv2.StereoSGBM_create's own defaults: P1 = P2 = 0 (-> 2 / 5 inside OpenCV) and uniquenessRatio < 0 (-> 10)
    (200, 60, 32, dict(P1=0, P2=0)),
    (200, 60, 32, dict(P1=7, P2=0, uniquenessRatio=-1)),
]


def sgbm_params(D, **kw):
    p = dict(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=63, uniquenessRatio=10,
             speckleWindowSize=100, speckleRange=2)
    p.update(kw)
    return p


def block_mask(h, w, seed=5):
    rng = np.random.default_rng(seed)
    m = rng.integers(0, 2, (h // 8 + 1, w // 8 + 1)).astype(np.uint8)
    return np.ascontiguousarray((np.kron(m, np.ones((8, 8), np.uint8))[:h, :w] * 255).astype(np.uint8))
