"""ctypes binding of the C ABI declared in include/gpyreg_b200.h.

There is no CPU fallback: if the shared library is missing or no B200 is
visible, every compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# GPYREG_B200_LIB: another build of the library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("GPYREG_B200_LIB") or os.path.join(HERE, "libgpyreg_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/gpyreg_b200.h declares
SYMBOLS = {
    "gpb_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "gpb_destroy": (None, [_vp]),
    "gpb_last_error": (C.c_char_p, [_vp]),
    "gpb_version": (C.c_int, []),
    "gpb_set_stream": (C.c_int, [_vp, C.c_uint64]),
    "gpb_set_workspace_limit": (C.c_int, [_vp, C.c_uint64]),
    "gpb_set_model": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "gpb_set_data": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int]),
    "gpb_nlz_batch": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, _vp]),
    "gpb_nlz_batch_dev": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, _vp]),
    "gpb_posterior_batch": (C.c_int, [_vp, _vp, C.c_int64, C.POINTER(_vp)]),
    "gpb_posterior_count": (C.c_int64, [_vp]),
    "gpb_posterior_fetch": (C.c_int, [_vp, C.c_int64, C.c_int, _vp]),
    "gpb_posterior_free": (None, [_vp]),
    "gpb_posterior_size": (C.c_int64, [_vp]),
    "gpb_posterior_append": (C.c_int, [_vp, _vp, _vp, C.c_double, _vp]),
    "gpb_posterior_rebuild": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "gpb_predict": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int,
                              _vp, _vp, _vp]),
    "gpb_predict_dev": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]),
    "gpb_predict_full": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp]),
    "gpb_quad": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]),
    "gpb_cov": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp,
                          C.c_int64, C.c_int, _vp, _vp]),
    "gpb_mean": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int64, C.c_int, _vp, _vp]),
    "gpb_noise": (C.c_int, [_vp, C.POINTER(C.c_int), _vp, _vp, _vp, C.c_int64, _vp, _vp]),
    "gpb_debug_gemm_nt": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double,
                                    C.c_double]),
    "gpb_debug_potrf": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "gpb_debug_diag_bench": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "gpb_debug_gemm_bench": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "gpb_last_timings": (C.c_int, [_vp, _vp]),
    "gpb_launch_count": (C.c_int64, [_vp]),
    "gpb_cache_stats": (C.c_int, [_vp, _vp, _vp]),
}

_lib = None


class GpbError(RuntimeError):
    """A call into libgpyreg_b200.so failed."""

    def __init__(self, code, msg):
        super().__init__(f"gpyreg_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load the shared library and bind every symbol (no GPU needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m gpyreg_b200._build` "
            "(nvcc, sm_100a). gpyreg_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Raw pointer of a C-contiguous float64/int32 array, or None."""
    if a is None:
        return None
    return a.ctypes.data


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a
