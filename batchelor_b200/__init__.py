"""batchelor_b200 -- B200-native (sm_100a) MNN hot path of LTLA/batchelor behind the reference's own interface.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of ``include/b200mnn.h``), ``_lib``
(ctypes binding), ``device`` (device-pointer wrappers; torch supplies memory/streams/process groups) and ``api``
(the host-side mirror of batchelor's R functions).  There is no CPU fallback anywhere in this package.
"""
from ._lib import B200Error, LIB_PATH, load  # noqa: F401
from .api import (  # noqa: F401
    B200Param,
    MNNResult,
    SerialParam,
    adjust_shift_variance,
    cosineNorm,
    fastMNN,
    findMutualNN,
    find_mutual_nns,
    mnnCorrect,
    multiBatchPCA,
    propagate_to_cells,
    queryKNN,
    reducedMNN,
    smooth_gaussian_kernel,
)

__all__ = [
    "B200Error", "B200Param", "MNNResult", "SerialParam", "adjust_shift_variance", "cosineNorm", "fastMNN", "findMutualNN",
    "find_mutual_nns", "mnnCorrect", "multiBatchPCA", "propagate_to_cells", "queryKNN", "reducedMNN", "smooth_gaussian_kernel", "load", "LIB_PATH",
]
