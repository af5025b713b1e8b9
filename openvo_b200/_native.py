"""ctypes binding of the C ABI in include/openvo_b200.h.

The product path loads exactly one library: the nvcc-built ``openvo_b200/lib/libopenvo_b200.so``.  If it is missing the
import of any compute entry point raises — there is no CPU fallback.  (The test-suite's CPU tier builds the same sources
against an execution emulator and passes that library's path explicitly to ``load()``; nothing in the package does.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OVO_B200_LIB") or os.path.join(_HERE, "lib", "libopenvo_b200.so")  # env override: kernel-variant experiments

SGBM_KEYS = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
             "uniquenessRatio", "speckleWindowSize", "speckleRange")
KP_FIELDS = 6


class SgbmParams(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in SGBM_KEYS]


class Config(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("sgbm", SgbmParams), ("roi", ctypes.c_int * 4),
                ("Q", ctypes.c_double * 16), ("nfeatures", ctypes.c_int), ("max_batch", ctypes.c_int),
                ("min_valid_disparity", ctypes.c_float), ("max_valid_disparity", ctypes.c_float), ("sgbm_mode", ctypes.c_int)]


class PairItem(ctypes.Structure):
    _fields_ = [("q_desc", ctypes.c_void_p), ("t_desc", ctypes.c_void_p), ("nq", ctypes.c_int), ("nt", ctypes.c_int),
                ("kp1", ctypes.c_void_p), ("kp2", ctypes.c_void_p), ("disp1", ctypes.c_void_p), ("disp2", ctypes.c_void_p),
                ("nn", ctypes.c_void_p), ("matches", ctypes.c_void_p), ("pts1", ctypes.c_void_p), ("pts2", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("scratch", ctypes.c_void_p), ("nn_rev", ctypes.c_void_p)]


class NativeError(RuntimeError):
    pass


def cv2_error(msg):
    """The reference's hard input errors surface as cv2.error (an OpenCV C++ assertion); ours are raised as a
    class that IS a cv2.error (so callers written against the reference keep working) and also a ValueError."""
    try:
        import cv2
        cls = cv2_error.__dict__.get("cls")
        if cls is None:
            cls = type("InputError", (cv2.error, ValueError), {})
            cv2_error.cls = cls
        return cls(msg)
    except Exception:
        return ValueError(msg)


_vp, _i, _sz, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_double
_SIGNATURES = {
    "ovo_last_error": (ctypes.c_char_p, []),
    "ovo_abi_version": (_i, []),
    "ovo_cropped_size": (_i, [ctypes.POINTER(Config), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "ovo_kp_capacity": (_i, [ctypes.POINTER(Config)]),
    "ovo_workspace_bytes": (_sz, [ctypes.POINTER(Config)]),
    "ovo_create": (_vp, [ctypes.POINTER(Config), _vp, _sz]),
    "ovo_destroy": (None, [_vp]),
    "ovo_sgbm_compute": (_i, [_vp, _vp, _vp, _i, _sz, _i, _vp, _vp]),
    "ovo_disparity_post": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "ovo_crop_left": (_i, [_vp, _vp, _i, _sz, _i, _vp, _vp]),
    "ovo_reproject_3d": (_i, [_vp, _vp, _vp, _vp]),
    "ovo_rectify": (_i, [_vp, _vp, _i, _i, _sz, _i, _vp, _vp, _vp, _vp]),
    "ovo_orb_detect_compute": (_i, [_vp, _vp, _vp, _i, _vp, _vp, ctypes.POINTER(_i), _vp]),
    "ovo_orb_detect_begin": (_i, [_vp, _vp, _vp, _i, _vp]),
    "ovo_orb_detect_finish": (_i, [_vp, _i, _vp, _vp, ctypes.POINTER(_i), _vp]),
    "ovo_orb_detect_finish_async": (_i, [_vp, _i, _vp, _vp, ctypes.POINTER(_i), _vp]),
    "ovo_orb_detect_wait": (_i, [_vp]),
    "ovo_extract_begin": (_i, [_vp, _vp, _vp, _i, _sz, _i, _vp, _vp, _vp, _vp, _vp]),
    "ovo_extract_finish": (_i, [_vp, _i, _vp, _vp, ctypes.POINTER(_i), _vp]),
    "ovo_knn2_hamming": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "ovo_match_points": (_i, [_vp, _vp, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ovo_rigid_transform": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "ovo_pnp_ransac": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _d, ctypes.c_ulonglong, _vp, _vp]),
    "ovo_rigid_body_filter": (_i, [_vp, _vp, _vp, _vp, _i, _d, _vp]),
    "ovo_outlier_filter": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _d, _vp]),
    "ovo_pair_batch": (_i, [_vp, _i, ctypes.POINTER(PairItem), _d, _vp]),
    "ovo_launch_count": (ctypes.c_longlong, []),
    "ovo_transfer_bytes": (None, [_vp, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]),
    "ovo_profile_enable": (None, [_i]),
    "ovo_profile_read": (_i, [ctypes.c_char_p, _i, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_i), _i]),
}
EXPORTS = tuple(_SIGNATURES)

_LIBS = {}


def load(path=None):
    """Load (once) and type the shared library.  ``path=None`` -> the in-tree nvcc build; raises if absent."""
    path = os.path.abspath(path or LIB_PATH)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise NativeError(
            "openvo_b200: CUDA extension %s is missing — build it with `python -m openvo_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.ovo_abi_version() != 1:
        raise NativeError("openvo_b200: ABI version mismatch")
    _LIBS[path] = lib
    return lib


def check(lib, rc):
    if rc != 0:
        raise NativeError(lib.ovo_last_error().decode("utf-8", "replace"))


def make_config(width, height, sgbm_params, roi, Q, nfeatures, max_batch=1, min_valid=4.0, max_valid=100.0, sgbm_mode=None):
    cfg = Config()
    cfg.width, cfg.height = int(width), int(height)
    for k in SGBM_KEYS:
        setattr(cfg.sgbm, k, int(sgbm_params[k]))
    for i in range(4):
        cfg.roi[i] = int(roi[i])
    flat = [float(v) for row in Q for v in row]
    for i in range(16):
        cfg.Q[i] = flat[i]
    cfg.nfeatures, cfg.max_batch = int(nfeatures), int(max_batch)
    cfg.min_valid_disparity, cfg.max_valid_disparity = float(min_valid), float(max_valid)
    cfg.sgbm_mode = int(sgbm_params.get("mode", 0) if sgbm_mode is None else sgbm_mode)
    return cfg


def ptr(x):
    """Raw address of a torch tensor / numpy array / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x.ctypes.data


def profile_read(lib):
    """-> {kernel name: (total ms, launches)} since the last read."""
    names = ctypes.create_string_buffer(8192)
    ms = (ctypes.c_float * 128)()
    cnt = (ctypes.c_int * 128)()
    n = lib.ovo_profile_read(names, 8192, ms, cnt, 128)
    tags = names.value.decode().split("\n")[:n]
    return {t: (float(ms[i]), int(cnt[i])) for i, t in enumerate(tags)}
