"""CPU tier: the nvcc-built library loads and exports every symbol include/openvo_b200.h declares (no compute calls),
and the product refuses to run without it / without a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from openvo_b200 import _native as N


def _declared():
    hdr = open(os.path.join(ROOT, "include", "openvo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ovo_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from openvo_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(N.EXPORTS)
    assert N.load().ovo_abi_version() == 1


def test_config_validation_without_gpu():
    lib = N.load()
    from conftest import sgbm_params
    import numpy as np
    good = N.make_config(640, 200, sgbm_params(64), (0, 0, 639, 199), np.eye(4), 500)
    assert lib.ovo_workspace_bytes(ctypes.byref(good)) > 0
    cw, ch = ctypes.c_int(), ctypes.c_int()
    assert lib.ovo_cropped_size(ctypes.byref(good), ctypes.byref(cw), ctypes.byref(ch)) == 0
    assert (cw.value, ch.value) == (639, 199)   # reference slice semantics (B1)
    for bad in (sgbm_params(64, minDisparity=1), sgbm_params(40), sgbm_params(64, blockSize=15, P1=1800, P2=7200),
                sgbm_params(64, blockSize=4)):
        cfg = N.make_config(640, 200, bad, (0, 0, 639, 199), np.eye(4), 500)
        assert lib.ovo_workspace_bytes(ctypes.byref(cfg)) == 0
        assert lib.ovo_last_error()


def test_missing_extension_fails_loudly(tmp_path):
    with pytest.raises(N.NativeError):
        N.load(str(tmp_path / "nope.so"))


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from openvo_b200 import StereoCamera, synth
    cam = StereoCamera(**synth.camera_args(320, 120, 32))
    with pytest.raises(N.NativeError):
        cam.engine()


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under openvo_b200/ may import, link or execute it."""
    import re
    pkg = os.path.join(ROOT, "openvo_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#include\s+\".*oracle)", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not pat.search(src), f
                assert "oracle/_build" not in src and "liboracle" not in src, f
