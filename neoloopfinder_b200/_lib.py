"""ctypes binding of libnlf_b200.so (include/nlf_b200.h).

This is the whole Python <-> CUDA boundary: plain pointers and sizes, numpy arrays owned by the
caller.  There is NO CPU fallback: if the library is missing or no B200 is visible, every entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnlf_b200.so")

NLF_ERR_EMPTY_FIT = -3
NLF_ERR_LOWER_TRI = -4

_lib = None
_lock = threading.Lock()

dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
llp = C.POINTER(C.c_longlong)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_ubyte)
vp = C.c_void_p

_SIGNATURES = {
    "nlf_version": (C.c_int, []),
    "nlf_last_error": (C.c_char_p, []),
    "nlf_device_count": (C.c_int, [ip]),
    "nlf_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "nlf_ctx_destroy": (C.c_int, [vp]),
    "nlf_ctx_sync": (C.c_int, [vp]),
    "nlf_timer_start": (C.c_int, [vp]),
    "nlf_timer_stop": (C.c_int, [vp, fp]),
    "nlf_launch_count": (C.c_int, [vp, llp]),
    "nlf_profile_enable": (C.c_int, [vp, C.c_int]),
    "nlf_profile_read": (C.c_int, [vp, C.c_int, fp, llp]),
    "nlf_profile_reset": (C.c_int, [vp]),
    "nlf_l2_flush": (C.c_int, [vp]),
    "nlf_fp64_peak": (C.c_int, [vp, dp]),
    "nlf_forest_create": (C.c_int, [vp, C.c_int, i64p, ip, ip, ip, dp, u8p, dp, C.c_int, C.POINTER(vp)]),
    "nlf_forest_compact_host": (C.c_int, [C.c_int, i64p, ip, ip, ip, dp, dp, C.c_longlong, C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_uint32), C.c_longlong, dp, C.POINTER(C.c_uint32), llp, llp]),
    "nlf_forest_destroy": (C.c_int, [vp]),
    "nlf_forest_predict": (C.c_int, [vp, fp, C.c_longlong, dp]),
    "nlf_pixels_create": (C.c_int, [vp, C.c_longlong, i64p, ip, ip, dp, C.POINTER(vp)]),
    "nlf_pixels_destroy": (C.c_int, [vp]),
    "nlf_assembly_create_from_pixels": (C.c_int, [vp, vp, C.c_int, i64p, i64p, u8p, C.c_int, C.POINTER(vp)]),
    "nlf_assembly_size": (C.c_int, [vp, ip]),
    "nlf_assembly_set_bias": (C.c_int, [vp, dp]),
    "nlf_assembly_fetch_matrix": (C.c_int, [vp, dp, dp]),
    "nlf_expected_from_pixels": (C.c_int, [vp, vp, C.c_int, C.c_longlong, C.c_longlong, C.c_int, dp, llp]),
    "nlf_batch_add_pixels": (C.c_int, [vp, vp, C.c_int, C.c_longlong, C.c_longlong, ip, ip]),
    "nlf_assembly_create": (C.c_int, [vp, dp, C.c_int, dp, ip, ip, C.c_int, C.POINTER(vp)]),
    "nlf_assembly_diag_means": (C.c_int, [vp, C.c_int, C.c_int, dp, ip]),
    "nlf_assembly_rect_means": (C.c_int, [vp, C.c_int, ip, ip, ip, ip, C.c_int, C.c_int, dp, ip]),
    "nlf_assembly_set_scales": (C.c_int, [vp, ip, dp]),
    "nlf_assembly_remove_gaps": (C.c_int, [vp, ip, ip]),
    "nlf_assembly_fetch_hcm": (C.c_int, [vp, dp]),
    "nlf_assembly_fetch_gap_free": (C.c_int, [vp, dp]),
    "nlf_assembly_destroy": (C.c_int, [vp]),
    "nlf_expected_diagonals": (C.c_int, [vp, dp, C.c_int, C.c_int, u8p, dp, llp]),
    "nlf_batch_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "nlf_batch_add_dense": (C.c_int, [vp, dp, C.c_int]),
    "nlf_batch_add_assembly": (C.c_int, [vp, vp]),
    "nlf_batch_add_band": (C.c_int, [vp, dp, C.c_int, C.c_int]),
    "nlf_batch_keep_features": (C.c_int, [vp, C.c_int]),
    "nlf_batch_score": (C.c_int, [vp, vp, dp, C.c_int, C.c_int, dp, C.c_int, C.c_double]),
    "nlf_batch_counts": (C.c_int, [vp, llp, llp, llp]),
    "nlf_batch_fetch_donuts": (C.c_int, [vp, C.c_int, ip, ip, dp, dp]),
    "nlf_batch_fetch_donuts_all": (C.c_int, [vp, C.c_longlong, ip, ip, ip, dp, dp, llp]),
    "nlf_batch_fetch_scored": (C.c_int, [vp, C.c_int, C.c_longlong, ip, ip, dp]),
    "nlf_batch_fetch_expected": (C.c_int, [vp, C.c_int, dp, dp]),
    "nlf_batch_fetch_candidates": (C.c_int, [vp, C.c_int, C.c_longlong, ip, ip]),
    "nlf_batch_fetch_features32": (C.c_int, [vp, C.c_int, C.c_longlong, fp]),
    "nlf_batch_getwindow": (C.c_int, [vp, C.c_int, ip, ip, C.c_longlong, C.c_int, dp, C.c_int, dp, C.c_int, dp, ip, llp]),
    "nlf_distance_normalize": (C.c_int, [vp, dp, C.c_longlong, ip, ip, C.c_int, dp, C.c_int, dp, u8p]),
    "nlf_batch_reset": (C.c_int, [vp]),
    "nlf_batch_njobs": (C.c_int, [vp]),
    "nlf_batch_destroy": (C.c_int, [vp]),
    "nlf_pool": (C.c_int, [vp, C.c_int, i64p, ip, ip, dp, C.c_int, C.c_int, C.c_int, ip, i64p, ip, i64p, u8p]),
    "nlf_pool_find_anchors": (C.c_int, [vp, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, C.c_int, ip]),
    "nlf_pool_find_anchors_wlen": (C.c_int, [vp, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, C.c_int, ip]),
    "nlf_batch_rethreshold": (C.c_int, [vp, C.c_double]),
    "nlf_pool_cluster_core": (C.c_int, [vp, ip, ip, C.c_int, C.c_int, u8p, u8p]),
    "nlf_pool_peak_heights": (C.c_int, [vp, C.c_int, i64p, ip, C.c_int, C.c_int, C.c_int, dp, i64p, ip]),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)


class NLFError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnlf_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `python -m neoloopfinder_b200.build` "
                    "(nvcc, sm_100a). neoloopfinder_b200 has no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc):
    if rc < 0:
        msg = load().nlf_last_error()
        raise NLFError(rc, msg.decode("utf-8", "replace") if msg else "unknown")
    return rc


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    n = C.c_int(0)
    check(load().nlf_device_count(C.byref(n)))
    return n.value


class Context:
    """One per GPU and per worker process; lazily created."""

    _per_device = {}

    def __init__(self, device=0):
        self.h = vp()
        check(load().nlf_ctx_create(int(device), C.byref(self.h)))
        self.device = int(device)

    @staticmethod
    def default_device() -> int:
        """NLF_DEVICE, else LOCAL_RANK (torchrun), else -- inside a pool worker such as the loky processes the
        reference fans assemblies out to (scripts/neoloop-caller:155-159) -- worker number modulo the visible
        GPUs, else 0 (device = worker_id mod n_gpus)."""
        for key in ("NLF_DEVICE", "LOCAL_RANK"):
            if key in os.environ:
                return int(os.environ[key])
        import multiprocessing
        import re
        m = re.search(r"(\d+)$", multiprocessing.current_process().name)     # LokyProcess-3, SpawnPoolWorker-2, ...
        if m is None:
            return 0
        n = device_count()
        return (int(m.group(1)) - 1) % n if n > 1 else 0

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = cls.default_device()
        key = (os.getpid(), device)
        if key not in cls._per_device:
            cls._per_device[key] = Context(device)
        return cls._per_device[key]

    def sync(self):
        check(load().nlf_ctx_sync(self.h))

    def timer_start(self):
        check(load().nlf_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(load().nlf_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_longlong(0)
        check(load().nlf_launch_count(self.h, C.byref(n)))
        return n.value

    def l2_flush(self):
        check(load().nlf_l2_flush(self.h))

    def fp64_peak_gops(self) -> float:
        g = C.c_double(0)
        check(load().nlf_fp64_peak(self.h, C.byref(g)))
        return g.value

    def profile(self, on=True):
        check(load().nlf_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        check(load().nlf_profile_reset(self.h))

    def profile_read(self, which):
        ms = C.c_float(0)
        n = C.c_longlong(0)
        check(load().nlf_profile_read(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value


PROF = {"bandify": 0, "diag": 1, "feature": 2, "forest": 3, "threshold": 4, "norm": 5, "pool": 6, "score": 7, "isotonic": 8, "count": 9}
