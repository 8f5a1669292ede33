"""Measurement hooks of libb200mnn that are not part of the reference-facing surface (used by bench.py)."""
from __future__ import annotations

import ctypes as C

from . import _lib


def gemm_profile(enable: bool):
    """enable=True: start recording CUDA events around every split-fp16 GEMM launch.  enable=False: stop and return
    (summed milliseconds, launches, executed tensor flops)."""
    if enable:
        _lib.call("b200mnn_gemm_profile_enable", 1)
        return None
    ms, n, fl = C.c_double(0), C.c_int64(0), C.c_double(0)
    _lib.call("b200mnn_gemm_profile_collect", C.byref(ms), C.byref(n), C.byref(fl))
    _lib.call("b200mnn_gemm_profile_enable", 0)
    return ms.value, n.value, fl.value
