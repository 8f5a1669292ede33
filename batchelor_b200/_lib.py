"""ctypes binding of libb200mnn.so (the C ABI declared in include/b200mnn.h).

The library is the product: there is no Python/NumPy/torch fallback for any compute entry point.  Loading fails
loudly when the shared object has not been built (``python -c "import __graft_entry__ as g; g.build()"``), and every
compute call fails with :class:`B200Error` when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200MNN_LIB: load a differently built copy of the same library (kernel A/B measurements on one box)
LIB_PATH = os.environ.get("B200MNN_LIB") or os.path.join(HERE, "lib", "libb200mnn.so")

i64 = C.c_int64
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
vp = C.c_void_p


class B200Error(RuntimeError):
    """Raised for every non-zero status of the C ABI; ``code`` is the B200MNN_E* value."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


# name -> argtypes.  Every symbol include/b200mnn.h declares is listed here (tests check the two agree).
SIGNATURES = {
    "b200mnn_last_error": [],
    "b200mnn_version": [],
    "b200mnn_launch_count": [],
    "b200mnn_device_count": [],
    "b200mnn_set_device": [C.c_int],
    "b200mnn_query_knn": [f64p, i64, f64p, i64, C.c_int, C.c_int, C.c_int, i32p, f64p],
    "b200mnn_find_mutual_nn": [f64p, i64, f64p, i64, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p, i64, i64p],
    "b200mnn_find_mutual_nns": [i32p, i64, C.c_int, i32p, i64, C.c_int, i32p, i32p, i64p],
    "b200mnn_smooth_gaussian_kernel": [f64p, i64, i64, i32p, i64, f64p, i64, i64, C.c_double, f64p],
    "b200mnn_adjust_shift_variance": [f64p, i64, i64, f64p, i64, i64, f64p, i64, i64, C.c_double, i32p, i64, i32p, i64, f64p],
    "b200mnn_cosine_norm": [f64p, i64, i64, f64p, f64p],
    "b200mnn_average_correction": [f64p, i64, f64p, i64, C.c_int, i32p, i32p, i64, f64p, i32p, i64p],
    "b200mnn_center_along_batch_vector": [f64p, i64, C.c_int, f64p, i32p, i64, f64p],
    "b200mnn_tricube_weighted_correction": [f64p, i64, C.c_int, f64p, i32p, i64, C.c_int, C.c_double, f64p],
    "b200mnn_reduced_mnn": [C.POINTER(f64p), i64p, C.c_int, C.c_int, C.c_int, i32p, i32p, C.c_int, C.c_double, C.c_double, C.c_double,
                            C.POINTER(i32p), i64p, C.c_int, C.POINTER(vp)],
    "b200mnn_result_ncells": [vp],
    "b200mnn_result_npairs": [vp, C.c_int],
    "b200mnn_result_pairs": [vp, C.c_int, i32p, i32p],
    "b200mnn_result_corrected": [vp, f64p, C.c_int],
    "b200mnn_result_info": [vp, i32p, i64p, f64p, i32p, f64p],
    "b200mnn_result_merges": [vp, i32p, i32p],
    "b200mnn_result_free": [vp],
    # device-pointer entry points: raw addresses (void*) so torch tensors' data_ptr() can be passed directly
    "b200mnn_dev_query_knn": [vp, i64, vp, i64, C.c_int, C.c_int, vp, vp, vp, vp],
    "b200mnn_dev_find_mutual_nns": [vp, i64, C.c_int, vp, i64, C.c_int, vp, vp, i64, vp, vp],
    "b200mnn_dev_average_correction": [vp, i64, vp, i64, C.c_int, vp, vp, i64, vp, vp, vp, vp],
    "b200mnn_dev_center_along_batch_vector": [vp, i64, C.c_int, vp, vp, i64, vp],
    "b200mnn_dev_tricube_apply": [vp, i64, C.c_int, vp, i64, vp, vp, C.c_int, C.c_double, vp, vp],
    "b200mnn_dev_smooth_gaussian_kernel": [vp, i64, i64, vp, vp, i64, i64, C.c_double, vp, vp],
    "b200mnn_smooth_last_check": [C.POINTER(C.c_int), f64p, f64p],
    "b200mnn_dev_adjust_shift_variance": [vp, i64, vp, i64, i64, vp, C.c_double, vp, i64, vp, i64, vp, vp],
    "b200mnn_dev_smooth_gaussian_from_centroids": [vp, i64, C.c_int, vp, vp, C.c_int, C.c_double, vp, vp],
    "b200mnn_dev_cosine_norm": [vp, i64, i64, vp, vp, vp],
    "b200mnn_dev_transpose_f64": [vp, i64, i64, vp, vp],
    "b200mnn_dev_debug_candidates": [vp, i64, vp, i64, C.c_int, C.c_int, vp, vp, vp, i64, i64p, vp],
    "b200mnn_dev_debug_gemm": [vp, i64, vp, i64, i64, C.c_int, C.c_int, vp, i64, vp],
    "b200mnn_gemm_profile_enable": [C.c_int],
    "b200mnn_gemm_profile_collect": [f64p, i64p, f64p],
    "b200mnn_profile_enable": [C.c_int],
    "b200mnn_profile_collect": [f64p, i64p, f64p],
    "b200mnn_profile_collect_executed": [f64p],
}

_lib = None


def load():
    """Load (once) and return the ctypes handle.  Raises ImportError if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m batchelor_b200._build` (needs nvcc). "
                "batchelor_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.argtypes = argtypes
            fn.restype = (C.c_char_p if name == "b200mnn_last_error" else
                          C.c_int64 if name in ("b200mnn_launch_count", "b200mnn_result_ncells", "b200mnn_result_npairs") else
                          None if name == "b200mnn_result_free" else C.c_int)
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200mnn_last_error()
        raise B200Error(rc, (msg or b"unknown error").decode("utf-8", "replace"))


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
