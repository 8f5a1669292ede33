"""TEST INFRASTRUCTURE ONLY -- ctypes bindings onto the two CPU checkers.

* ``_build/liboracle.so``  : this repo's C restatement (``mnn_oracle.c``) + KMKNN port (``kmknn_port.cpp``)
* ``_ref/libbatchelor_ref.so`` : the reference's own kernels compiled unmodified from /root/reference/src

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module.  The product package ``batchelor_b200`` never does.

All wrappers take/return numpy arrays in *R conventions* (cells x dims matrices, 1-based neighbour ids)
unless stated otherwise, so the parity tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libbatchelor_ref.so")

_i64, _i32p, _f64p = C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile both checkers (the reference one only when /root/reference is present)."""
    if force and os.path.exists(_ORACLE_SO):
        os.remove(_ORACLE_SO)
    subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)   # no-op when up to date
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["ALIGNED"] + (["F_CONTIGUOUS"] if order == "F" else ["C_CONTIGUOUS"]))


def _i32(a, order="F"):
    return np.require(np.asarray(a, dtype=np.int32), requirements=["ALIGNED"] + (["F_CONTIGUOUS"] if order == "F" else ["C_CONTIGUOUS"]))


def _p(a, t):
    return a.ctypes.data_as(t)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_ORACLE_SO)
        L.oracle_query_knn.argtypes = [_f64p, _i64, _f64p, _i64, C.c_int, C.c_int, C.c_int, _i32p, _f64p, C.c_int]
        L.oracle_find_mutual_nns.argtypes = [_i32p, _i64, C.c_int, _i32p, _i64, C.c_int, _i32p, _i32p, C.POINTER(_i64)]
        L.oracle_smooth_gaussian_kernel.argtypes = [_f64p, _i64, _i64, _i32p, _i64, _f64p, _i64, _i64, C.c_double, _f64p, C.c_int]
        L.oracle_adjust_shift_variance.argtypes = [_f64p, _i64, _i64, _f64p, _i64, _i64, _f64p, _i64, _i64, C.c_double,
                                                   _i32p, _i64, _i32p, _i64, _f64p, C.c_int]
        L.oracle_cosine_norm.argtypes = [_f64p, _i64, _i64, _f64p, _f64p]
        L.kmknn_build.restype = C.c_void_p
        L.kmknn_build.argtypes = [_f64p, _i64, C.c_int, C.c_int, C.c_uint64]
        L.kmknn_free.argtypes = [C.c_void_p]
        L.kmknn_query.argtypes = [C.c_void_p, _f64p, _i64, C.c_int, _i32p, _f64p, C.c_int]
        _lib = L
    return _lib


def have_ref() -> bool:
    build()
    return os.path.exists(_REF_SO)


def ref():
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(_REF_SO)
        R.ref_last_error.restype = C.c_char_p
        R.ref_find_mutual_nns.argtypes = [_i32p, _i64, C.c_int, _i32p, _i64, C.c_int, _i32p, _i32p, C.POINTER(_i64)]
        R.ref_smooth_gaussian_kernel.argtypes = [_f64p, _i64, _i64, _i32p, _i64, _f64p, _i64, _i64, C.c_double, _f64p]
        R.ref_adjust_shift_variance.argtypes = [_f64p, _i64, _i64, _f64p, _i64, _i64, _f64p, _i64, _i64, C.c_double,
                                                _i32p, _i64, _i32p, _i64, _f64p]
        _ref = R
    return _ref


class OracleError(RuntimeError):
    pass


# --------------------------------------------------------------------------------------------
# a1 exact kNN
# --------------------------------------------------------------------------------------------
def query_knn(X, query, k, nthreads=0):
    """Brute-force fp64 queryKNN(X, query, k): returns (index [nq x k] 1-based int32, distance [nq x k])."""
    X = _f64(X); Q = _f64(query)
    n, d = X.shape
    nq = Q.shape[0]
    assert Q.shape[1] == d
    idx = np.zeros((nq, k), dtype=np.int32, order="F")
    dist = np.zeros((nq, k), dtype=np.float64, order="F")
    rc = lib().oracle_query_knn(_p(X, _f64p), n, _p(Q, _f64p), nq, d, k, 1, _p(idx, _i32p), _p(dist, _f64p), nthreads)
    if rc:
        raise OracleError(f"oracle_query_knn rc={rc}")
    return idx, dist


class Kmknn:
    """Exact KMKNN port (multi-threaded); same results as :func:`query_knn`."""

    def __init__(self, X, nthreads=0, seed=42):
        self.X = _f64(X, "C")
        self.n, self.d = self.X.shape
        self.h = lib().kmknn_build(_p(self.X, _f64p), self.n, self.d, nthreads, seed)

    def query(self, query, k, nthreads=0, want_dist=True):
        Q = _f64(query, "C")
        nq = Q.shape[0]
        idx = np.zeros((nq, k), dtype=np.int32)
        dist = np.zeros((nq, k), dtype=np.float64) if want_dist else None
        rc = lib().kmknn_query(self.h, _p(Q, _f64p), nq, k, _p(idx, _i32p), _p(dist, _f64p) if want_dist else None, nthreads)
        if rc:
            raise OracleError(f"kmknn_query rc={rc}")
        return idx + 1, dist

    def __del__(self):
        try:
            if self.h:
                lib().kmknn_free(self.h)
                self.h = None
        except Exception:
            pass


# --------------------------------------------------------------------------------------------
# a3 mutual pairs
# --------------------------------------------------------------------------------------------
def _mutual(fn, left, right):
    left = _i32(left); right = _i32(right)
    n1, k2 = left.shape
    n2, k1 = right.shape
    first = np.zeros(max(1, n1 * k2), dtype=np.int32)
    second = np.zeros(max(1, n1 * k2), dtype=np.int32)
    npairs = _i64(0)
    rc = fn(_p(left, _i32p), n1, k2, _p(right, _i32p), n2, k1, _p(first, _i32p), _p(second, _i32p), C.byref(npairs))
    if rc:
        raise OracleError(f"find_mutual_nns rc={rc}")
    return first[: npairs.value].copy(), second[: npairs.value].copy()


def find_mutual_nns(left, right):
    return _mutual(lib().oracle_find_mutual_nns, left, right)


def ref_find_mutual_nns(left, right):
    return _mutual(ref().ref_find_mutual_nns, left, right)


def find_mutual_nn(data1, data2, k1, k2, nthreads=0):
    """findMutualNN(data1, data2, k1, k2) as used at R/MNN_tree.R:129: two exact searches + pair extraction."""
    w21, _ = query_knn(data2, data1, k2, nthreads)  # neighbours of batch-1 cells in batch 2
    w12, _ = query_knn(data1, data2, k1, nthreads)  # neighbours of batch-2 cells in batch 1
    return find_mutual_nns(w21, w12)


# --------------------------------------------------------------------------------------------
# a5 Gaussian smoothing
# --------------------------------------------------------------------------------------------
def smooth_gaussian_kernel(averaged, index0, mat, sigma2, nthreads=0):
    A = _f64(averaged); M = _f64(mat); I = _i32(index0)
    G, nmnn = A.shape
    Gd, nc = M.shape
    out = np.zeros((G, nc), dtype=np.float64, order="F")
    rc = lib().oracle_smooth_gaussian_kernel(_p(A, _f64p), G, nmnn, _p(I, _i32p), I.size, _p(M, _f64p), Gd, nc, float(sigma2),
                                             _p(out, _f64p), nthreads)
    if rc == 1:
        raise OracleError("'index' must have length equal to number of rows in 'averaged'")
    if rc:
        raise OracleError(f"smooth_gaussian_kernel rc={rc}")
    return out


def ref_smooth_gaussian_kernel(averaged, index0, mat, sigma2):
    A = _f64(averaged); M = _f64(mat); I = _i32(index0)
    G, nmnn = A.shape
    Gd, nc = M.shape
    out = np.zeros((G, nc), dtype=np.float64, order="F")
    rc = ref().ref_smooth_gaussian_kernel(_p(A, _f64p), G, nmnn, _p(I, _i32p), I.size, _p(M, _f64p), Gd, nc, float(sigma2), _p(out, _f64p))
    if rc:
        raise OracleError(ref().ref_last_error().decode())
    return out


# --------------------------------------------------------------------------------------------
# a7 shift variance
# --------------------------------------------------------------------------------------------
def adjust_shift_variance(data1, data2, vect, sigma2, restrict1, restrict2, nthreads=0):
    D1 = _f64(data1); D2 = _f64(data2); V = _f64(vect); r1 = _i32(restrict1); r2 = _i32(restrict2)
    out = np.zeros(D2.shape[1], dtype=np.float64)
    rc = lib().oracle_adjust_shift_variance(_p(D1, _f64p), D1.shape[0], D1.shape[1], _p(D2, _f64p), D2.shape[0], D2.shape[1],
                                            _p(V, _f64p), V.shape[0], V.shape[1], float(sigma2),
                                            _p(r1, _i32p), r1.size, _p(r2, _i32p), r2.size, _p(out, _f64p), nthreads)
    if rc == 1:
        raise OracleError("number of genes do not match up between matrices")
    if rc == 4:
        raise OracleError("number of cells do not match up between matrices")
    if rc == 3:
        raise OracleError("subset indices out of range")
    if rc:
        raise OracleError(f"adjust_shift_variance rc={rc}")
    return out


def adjust_shift_variance_cells(data1, data2, vect_rows, cells, sigma2, restrict1, restrict2, nthreads=0):
    """Scaling of the requested batch-2 cells only (same per-cell loop): data1 [G x n1], data2 [G x n2] column-major,
    vect_rows [len(cells) x G] = the correction rows of those cells."""
    D1 = _f64(data1); D2 = _f64(data2); r1 = _i32(restrict1); r2 = _i32(restrict2)
    V = np.ascontiguousarray(np.asarray(vect_rows, dtype=np.float64))
    cells = np.ascontiguousarray(np.asarray(cells, dtype=np.int64))
    out = np.zeros(cells.size, dtype=np.float64)
    fn = lib().oracle_adjust_shift_variance_cells
    fn.restype = C.c_int
    fn.argtypes = [_f64p, C.c_int64, C.c_int64, _f64p, C.c_int64, C.c_int64, _f64p, C.c_int64, C.c_int64, C.c_double, _i32p, C.c_int64,
                   _i32p, C.c_int64, C.POINTER(C.c_int64), C.c_int64, C.c_int, _f64p, C.c_int]
    rc = fn(_p(D1, _f64p), D1.shape[0], D1.shape[1], _p(D2, _f64p), D2.shape[0], D2.shape[1], V.ctypes.data_as(_f64p), D2.shape[1], D1.shape[0],
            float(sigma2), _p(r1, _i32p), r1.size, _p(r2, _i32p), r2.size, cells.ctypes.data_as(C.POINTER(C.c_int64)), cells.size, 1,
            out.ctypes.data_as(_f64p), nthreads)
    if rc:
        raise OracleError(f"adjust_shift_variance_cells rc={rc}")
    return out


def ref_adjust_shift_variance(data1, data2, vect, sigma2, restrict1, restrict2):
    D1 = _f64(data1); D2 = _f64(data2); V = _f64(vect); r1 = _i32(restrict1); r2 = _i32(restrict2)
    out = np.zeros(D2.shape[1], dtype=np.float64)
    rc = ref().ref_adjust_shift_variance(_p(D1, _f64p), D1.shape[0], D1.shape[1], _p(D2, _f64p), D2.shape[0], D2.shape[1],
                                         _p(V, _f64p), V.shape[0], V.shape[1], float(sigma2),
                                         _p(r1, _i32p), r1.size, _p(r2, _i32p), r2.size, _p(out, _f64p))
    if rc:
        raise OracleError(ref().ref_last_error().decode())
    return out


def cosine_norm(x):
    """x [G x n] (cells in columns) -> (normalised, l2norm)."""
    X = _f64(x)
    out = np.zeros_like(X, order="F")
    l2 = np.zeros(X.shape[1])
    lib().oracle_cosine_norm(_p(X, _f64p), X.shape[0], X.shape[1], _p(out, _f64p), _p(l2, _f64p))
    return out, l2
