"""Host-side mirror of batchelor's R interface for the MNN hot path, driving libb200mnn through its C ABI.

Same names, argument meaning and error behaviour as the reference (LTLA/batchelor v1.23.1), with R's dotted argument
names spelt with underscores: ``fastMNN`` / ``reducedMNN`` (R/fastMNN.R:283, R/reducedMNN.R:61), ``mnnCorrect``
(R/mnnCorrect.R:125), ``findMutualNN`` (R/findMutualNN.R:1-3), ``queryKNN`` (BiocNeighbors, call site
R/fastMNN.R:605), ``cosineNorm`` (R/cosineNorm.R:53) and the three ``.Call`` wrappers of R/RcppExports.R:4-14.
``BNPARAM=B200Param()`` plays the role of BiocNeighbors' backend selector; ``BPPARAM`` must be serial (multi-GPU
parallelism comes from ``torch.distributed``: one process per GPU, query rows sharded, see device.query_knn_sharded).

Matrices follow R's orientation: ``reducedMNN``/``queryKNN`` take [cells x dims]; ``mnnCorrect``/``cosineNorm`` take
[genes x cells].  Neighbour ids, pair ids, ``restrict`` and ``merge_order`` are 1-based like R's.  Host control flow
(merge tree, bookkeeping of pair positions) is Python; every numeric step runs in a CUDA kernel of the library.
"""
from __future__ import annotations

import ctypes as C
import math
import warnings
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _lib

B200Error = _lib.B200Error


# ----------------------------------------------------------------------------------------------------------
# Parameter objects (stand-ins for BiocNeighbors::KmknnParam / BiocParallel::SerialParam in the signatures)
# ----------------------------------------------------------------------------------------------------------
@dataclass
class B200Param:
    """BNPARAM-style backend selector: exact Euclidean search on a B200 (same results as KmknnParam())."""
    distance: str = "Euclidean"
    device: Optional[int] = None

    def __post_init__(self):
        if self.distance != "Euclidean":
            raise ValueError("B200Param only supports distance='Euclidean' (what batchelor uses)")


@dataclass
class SerialParam:
    """BPPARAM placeholder.  The GPU backend does its own parallelism; forked workers cannot share a CUDA context."""
    workers: int = 1


def _check_bpparam(BPPARAM) -> None:
    if BPPARAM is not None and getattr(BPPARAM, "workers", 1) != 1:
        raise ValueError("the B200 backend requires a serial BPPARAM (it shards across GPUs itself)")


def _select_device(BNPARAM) -> None:
    if BNPARAM is not None and getattr(BNPARAM, "device", None) is not None:
        _lib.call("b200mnn_set_device", int(BNPARAM.device))


def _f64(a, order=None):
    a = np.asarray(a, dtype=np.float64)
    if order == "F":
        return np.asfortranarray(a)
    if order == "C":
        return np.ascontiguousarray(a)
    return a if (a.flags.f_contiguous or a.flags.c_contiguous) else np.ascontiguousarray(a)


def _i32(a, order="C"):
    a = np.asarray(a)
    if a.dtype != np.int32:
        if a.size and (np.any(a != np.floor(a)) if a.dtype.kind == "f" else False):
            raise ValueError("indices must be integers")
        a = a.astype(np.int32)
    return np.asfortranarray(a) if order == "F" else np.ascontiguousarray(a)


def _fp(a):
    return a.ctypes.data_as(_lib.f64p)


def _ip(a):
    return a.ctypes.data_as(_lib.i32p)


# ----------------------------------------------------------------------------------------------------------
# kNN and mutual pairs (host buffers, R layout)
# ----------------------------------------------------------------------------------------------------------
def queryKNN(X, query, k, get_index=True, get_distance=True, BNPARAM=None, BPPARAM=None) -> Dict[str, Any]:
    """BiocNeighbors::queryKNN(X, query, k): ``index`` [nq x k] (1-based) and ``distance`` [nq x k]."""
    _check_bpparam(BPPARAM)
    _select_device(BNPARAM)
    X = _f64(X); Q = _f64(query)
    if X.ndim != 2 or Q.ndim != 2 or X.shape[1] != Q.shape[1]:
        raise ValueError("'X' and 'query' must be matrices with the same number of columns")
    n, d = X.shape
    nq = Q.shape[0]
    k = int(k)
    if k > n:
        warnings.warn("'k' capped at the number of observations")  # BiocNeighbors' behaviour
        k = n
    col_major = X.flags.f_contiguous and not X.flags.c_contiguous
    if col_major:
        Q = np.asfortranarray(Q)
    else:
        X = np.ascontiguousarray(X); Q = np.ascontiguousarray(Q)
    order = "F" if col_major else "C"
    idx = np.zeros((nq, k), dtype=np.int32, order=order)
    dist = np.zeros((nq, k), dtype=np.float64, order=order) if get_distance else None
    _lib.call("b200mnn_query_knn", _fp(X), n, _fp(Q), nq, d, k, 1 if col_major else 0, _ip(idx), _fp(dist) if get_distance else None)
    out: Dict[str, Any] = {}
    if get_index:
        out["index"] = idx
    if get_distance:
        out["distance"] = dist
    return out


def findMutualNN(data1, data2, k1, k2=None, BNPARAM=None, BPPARAM=None) -> Dict[str, np.ndarray]:
    """BiocNeighbors::findMutualNN (re-exported at R/findMutualNN.R:1-3): ``first``/``second`` 1-based pair ids."""
    _check_bpparam(BPPARAM)
    _select_device(BNPARAM)
    if k2 is None:
        k2 = k1
    d1 = np.ascontiguousarray(_f64(data1)); d2 = np.ascontiguousarray(_f64(data2))
    if d1.ndim != 2 or d2.ndim != 2 or d1.shape[1] != d2.shape[1]:
        raise ValueError("'data1' and 'data2' must be matrices with the same number of columns")
    n1, d = d1.shape
    n2 = d2.shape[0]
    cap = max(1, n1 * min(int(k2), n2))
    first = np.empty(cap, dtype=np.int32)    # upper bound n1 * k2; only the pages the library writes are ever touched
    second = np.empty(cap, dtype=np.int32)
    npairs = C.c_int64(0)
    _lib.call("b200mnn_find_mutual_nn", _fp(d1), n1, _fp(d2), n2, d, int(k1), int(k2), 0, _ip(first), _ip(second), cap, C.byref(npairs))
    m = npairs.value
    return {"first": first[:m], "second": second[:m]}   # views: copying 2 x 4 m bytes would cost as much as the D2H transfer


# ----------------------------------------------------------------------------------------------------------
# The three .Call wrappers (R/RcppExports.R:4-14)
# ----------------------------------------------------------------------------------------------------------
def find_mutual_nns(left, right):
    """left [n1 x k2], right [n2 x k1] integer matrices of 1-based neighbour ids -> [first, second] (1-based)."""
    L = _i32(left, "F"); R = _i32(right, "F")
    if L.ndim != 2 or R.ndim != 2:
        raise ValueError("'left' and 'right' must be integer matrices")
    n1, k2 = L.shape
    n2, k1 = R.shape
    first = np.zeros(max(1, n1 * k2), dtype=np.int32)
    second = np.zeros(max(1, n1 * k2), dtype=np.int32)
    npairs = C.c_int64(0)
    _lib.call("b200mnn_find_mutual_nns", _ip(L), n1, k2, _ip(R), n2, k1, _ip(first), _ip(second), C.byref(npairs))
    m = npairs.value
    return [first[:m].copy(), second[:m].copy()]


def smooth_gaussian_kernel(averaged, index, mat, sigma2):
    """averaged [G x nmnn], index (0-based columns of mat), mat [Gdist x ncells] -> [G x ncells]."""
    A = _f64(averaged, "F"); M = _f64(mat, "F"); I = _i32(index)
    G, nmnn = A.shape
    Gd, nc = M.shape
    out = np.zeros((G, nc), dtype=np.float64, order="F")
    _lib.call("b200mnn_smooth_gaussian_kernel", _fp(A), G, nmnn, _ip(I), I.size, _fp(M), Gd, nc, float(sigma2), _fp(out))
    return out


def adjust_shift_variance(data1, data2, vect, sigma2, restrict1, restrict2):
    """data1 [G x n1], data2 [G x n2], vect [n2 x G], restricts 0-based -> scaling factor per cell of batch 2."""
    D1 = _f64(data1, "F"); D2 = _f64(data2, "F"); V = _f64(vect, "F")
    r1 = _i32(restrict1); r2 = _i32(restrict2)
    out = np.zeros(D2.shape[1], dtype=np.float64)
    _lib.call("b200mnn_adjust_shift_variance", _fp(D1), D1.shape[0], D1.shape[1], _fp(D2), D2.shape[0], D2.shape[1], _fp(V), V.shape[0],
              V.shape[1], float(sigma2), _ip(r1), r1.size, _ip(r2), r2.size, _fp(out))
    return out


# ----------------------------------------------------------------------------------------------------------
# cosineNorm (R/cosineNorm.R:53-82)
# ----------------------------------------------------------------------------------------------------------
def cosineNorm(x, mode="matrix", subset_row=None, BPPARAM=None):
    """x [genes x cells].  mode: 'matrix' | 'all' | 'l2norm' (as the reference's match.arg)."""
    if mode not in ("matrix", "all", "l2norm"):
        raise ValueError("'arg' should be one of 'matrix', 'all', 'l2norm'")
    x = np.asarray(x, dtype=np.float64)
    if subset_row is not None:
        x = x[_subset_to_index(subset_row, x.shape[0]) - 1, :]
    X = np.asfortranarray(x)
    G, n = X.shape
    l2 = np.zeros(n, dtype=np.float64)
    if mode == "l2norm":
        _lib.call("b200mnn_cosine_norm", _fp(X), G, n, None, _fp(l2))
        return l2
    out = np.zeros((G, n), dtype=np.float64, order="F")
    _lib.call("b200mnn_cosine_norm", _fp(X), G, n, _fp(out), _fp(l2))
    return out if mode == "matrix" else {"matrix": out, "l2norm": l2}


def _subset_to_index(subset, n) -> np.ndarray:
    """R/utils_subset.R: logical or 1-based integer subset -> 1-based integer index."""
    s = np.asarray(subset)
    if s.dtype == bool:
        if s.size != n:
            raise ValueError("logical subset has the wrong length")
        return np.nonzero(s)[0].astype(np.int64) + 1
    s = s.astype(np.int64)
    if s.size and (s.min() < 1 or s.max() > n):
        raise IndexError("subset indices out of range")
    return s


# ----------------------------------------------------------------------------------------------------------
# Merge tree (R/MNN_tree.R:2-109) -- host control flow
# ----------------------------------------------------------------------------------------------------------
class _Node:
    __slots__ = ("index", "data", "restrict", "origin", "extras")

    def __init__(self, index, data, restrict, origin=None, extras=None):
        self.index = list(index)
        self.data = data            # torch [n x d] float64 cuda
        self.restrict = restrict    # torch int64 0-based rows, or None
        self.origin = np.repeat(self.index[0], data.shape[0]) if origin is None else origin
        self.extras = [] if extras is None else extras


def _binarize(tree):
    if not isinstance(tree, (list, tuple)):
        return tree
    n = len(tree)
    if n == 0:
        raise ValueError("merge tree contains a node with no children")
    if n == 1:
        return _binarize(tree[0])
    cur = [_binarize(tree[0]), _binarize(tree[1])]
    for i in range(2, n):
        cur = [cur, _binarize(tree[i])]
    return cur


def _leaves(tree):
    return [tree] if not isinstance(tree, list) else _leaves(tree[0]) + _leaves(tree[1])


def _predefined_tree(nb: int, merge_order):
    if merge_order is None:
        merge_order = list(range(1, nb + 1))
    merge_order = list(merge_order) if isinstance(merge_order, (tuple, np.ndarray)) else merge_order
    if not any(isinstance(m, (list, tuple)) for m in merge_order) and len(merge_order) > 1:
        tree = [merge_order[0], merge_order[1]]
        for i in merge_order[2:]:
            tree = [tree, i]
    else:
        tree = merge_order
    tree = _binarize(tree)
    leaves = _leaves(tree)
    ok = all(isinstance(l, (int, np.integer)) for l in leaves)
    if not ok or len(set(leaves)) != len(leaves) or any(l < 1 or l > nb for l in leaves) or len(leaves) != nb:
        raise ValueError("invalid leaf nodes specified in 'merge.order'")
    return tree


def _fill(tree, make_leaf):
    if not isinstance(tree, list):
        return make_leaf(int(tree))
    return [_fill(tree[0], make_leaf), _fill(tree[1], make_leaf)]


def _next_merge(tree, path=()):
    """R/MNN_tree.R:61-69: descend into the second child while it is a list -> right subtree first."""
    if not isinstance(tree[0], list) and not isinstance(tree[1], list):
        return tree[0], tree[1], path
    if isinstance(tree[1], list):
        return _next_merge(tree[1], path + (1,))
    return _next_merge(tree[0], path + (0,))


def _update(tree, path, node):
    if len(path) == 0:
        return node
    tree[path[0]] = _update(tree[path[0]], path[1:], node)
    return tree


def _choose_k(k, prop_k, N):
    """R/MNN_tree.R:140-146."""
    if prop_k is None:
        return int(k)
    return int(min(N, max(k, int(round(prop_k * N)))))


def _restore_original_order(batch_ordering, ncells_per_batch):
    """R/utils_reorder.R:1-21 (1-based permutation)."""
    reorder: List[Optional[np.ndarray]] = [None] * len(batch_ordering)
    last = 0
    for idx in batch_ordering:
        n = int(ncells_per_batch[idx - 1])
        reorder[idx - 1] = last + np.arange(1, n + 1)
        last += n
    return np.concatenate(reorder) if reorder else np.zeros(0, dtype=np.int64)


def _reindex_pairings(pairings, new_order):
    """R/utils_reorder.R:23-36."""
    new_order = np.asarray(new_order, dtype=np.int64)
    rev = np.zeros(new_order.size, dtype=np.int64)
    rev[new_order - 1] = np.arange(1, new_order.size + 1)
    return [{"left": rev[p["left"] - 1], "right": rev[p["right"] - 1]} for p in pairings]


@dataclass
class MNNResult:
    """``corrected`` is [cells x dims] for reducedMNN/fastMNN and [genes x cells] for mnnCorrect, like the reference."""
    corrected: np.ndarray
    batch: np.ndarray
    merge_info: Dict[str, Any] = field(default_factory=dict)


def _check_restrict(batches_ncells: Sequence[int], restrict):
    """checkRestrictions (R/checkInputs.R:96-125): list of 1-based / logical subsets or None -> list of 1-based or None."""
    if restrict is None:
        return None
    if len(restrict) != len(batches_ncells):
        raise ValueError("'restrictions' must of length equal to the number of batches")
    out = []
    for n, r in zip(batches_ncells, restrict):
        if r is None:
            out.append(None)
            continue
        idx = _subset_to_index(r, n)
        if idx.size == 0:
            raise ValueError("no cells remaining in a batch after restriction")
        out.append(idx)
    return out


def _divide_into_batches(x, batch, byrow, restrict):
    """divideIntoBatches (R/divideIntoBatches.R): split by sorted factor levels, remember how to undo it."""
    ncell = x.shape[0] if byrow else x.shape[1]
    if batch is None:
        raise ValueError("'batch' must be specified if '...' has only one object")
    batch = np.asarray(batch)
    if batch.size != ncell:
        raise ValueError("'length(batch)' should be equal to number of cells in '...'")
    levels = np.unique(batch)
    rmask = None
    if restrict is not None:
        rmask = np.zeros(ncell, dtype=bool)
        rmask[_subset_to_index(restrict, ncell) - 1] = True
    batches, restricted = [], ([] if restrict is not None else None)
    reorder = np.zeros(ncell, dtype=np.int64)
    last = 0
    for b in levels:
        keep = batch == b
        cur = x[keep, :] if byrow else x[:, keep]
        n = int(keep.sum())
        if rmask is not None:
            cr = np.nonzero(rmask[keep])[0] + 1
            if cr.size == 0:
                raise ValueError("no cells remaining in a batch after restriction")
            restricted.append(cr)
        batches.append(cur)
        reorder[keep] = last + np.arange(1, n + 1)
        last += n
    return batches, reorder, restricted, levels


# ----------------------------------------------------------------------------------------------------------
# reducedMNN / fastMNN core (R/fastMNN.R:436-562) -- device-resident merge loop
# ----------------------------------------------------------------------------------------------------------
def _perbatch_var(data, index, origin):
    """.compute_perbatch_var (R/fastMNN.R:651-658): diagnostic only (merge.info$lost.var); torch reduction."""
    import torch

    out = np.zeros(len(index))
    for i, b in enumerate(index):
        rows = torch.from_numpy(np.nonzero(origin == b)[0]).to(data.device)
        sub = data.index_select(0, rows)
        out[i] = float(sub.var(dim=0, unbiased=True).sum().item()) if sub.shape[0] > 1 else float("nan")
    return out


def _restricted_mnn(dev, ld, lres, rd, rres, k, prop_k):
    """.restricted_mnn (R/MNN_tree.R:113-133) on device; returns 0-based pair ids into ld / rd (torch int64)."""
    L = ld if lres is None else ld.index_select(0, lres)
    R = rd if rres is None else rd.index_select(0, rres)
    k1 = _choose_k(k, prop_k, L.shape[0])
    k2 = _choose_k(k, prop_k, R.shape[0])
    first, second, _, _ = dev.find_mutual_nn(L, R, k1, k2)
    first = first.long(); second = second.long()
    if lres is not None:
        first = lres.index_select(0, first)
    if rres is not None:
        second = rres.index_select(0, second)
    return first, second


def _combine_restrict(nl, lres, nr, rres, device):
    import torch

    if lres is None and rres is None:
        return None
    if lres is None:
        lres = torch.arange(nl, device=device)
    if rres is None:
        rres = torch.arange(nr, device=device)
    return torch.cat([lres, rres + nl])


def _finish(tree, full, pairings, left_set, right_set, extra) -> MNNResult:
    full_order = tree.index
    full_origin = np.asarray(tree.origin)
    pairs = []
    for p, ls, rs in zip(pairings, left_set, right_set):
        bonus1 = int(np.nonzero(full_origin == ls[0])[0][0])
        bonus2 = int(np.nonzero(full_origin == rs[0])[0][0])
        pairs.append({"left": p[0] + 1 + bonus1, "right": p[1] + 1 + bonus2})
    if any(full_order[i] > full_order[i + 1] for i in range(len(full_order) - 1)):
        ncells = np.bincount(full_origin, minlength=max(full_order) + 1)[1:]
        ordering = _restore_original_order(full_order, ncells)
        full = full[ordering - 1]
        full_origin = full_origin[ordering - 1]
        pairs = _reindex_pairings(pairs, ordering)
    info = {"left": left_set, "right": right_set, "pairs": pairs}
    info.update(extra)
    return MNNResult(corrected=full, batch=full_origin, merge_info=info)


class _Final:
    """What :func:`_finish` needs of the final tree node."""

    def __init__(self, index, origin):
        self.index = index
        self.origin = origin


def _merge_sequence(nb: int, merge_order):
    """Flattens the merge tree into the order .get_next_merge walks it (R/MNN_tree.R:61-69): node ids 0..nb-1 are the
    batches, merge m creates node nb + m.  Returns (left ids, right ids, batches of each left node, of each right node)."""
    tree = _predefined_tree(nb, merge_order)
    members = {i: [i + 1] for i in range(nb)}

    def to_ids(t):
        return [to_ids(t[0]), to_ids(t[1])] if isinstance(t, list) else int(t) - 1

    tree = to_ids(tree)
    lefts, rights, lset, rset = [], [], [], []
    for m in range(nb - 1):
        left, right, path = _next_merge(tree)
        lefts.append(left); rights.append(right)
        lset.append(list(members[left])); rset.append(list(members[right]))
        members[nb + m] = members[left] + members[right]
        tree = _update(tree, path, nb + m)
    return lefts, rights, lset, rset


def _fast_mnn_core_c(batches, k, prop_k, restrict, ndist, merge_order, min_batch_skip, get_variance=True, auto_merge=False) -> MNNResult:
    """The merge loop as ONE call of the C ABI (b200mnn_reduced_mnn, csrc/merge.cu): what the R shim does.  With
    ``auto_merge`` the merge order is searched on the device (R/MNN_tree.R:154-226) and read back afterwards."""
    nb = len(batches)
    d = batches[0].shape[1]
    mats = []
    for b in batches:
        if b.ndim != 2 or b.shape[1] != d:
            raise ValueError("number of columns is not the same across batches")
        mats.append(np.ascontiguousarray(b, dtype=np.float64))
    if auto_merge:
        lefts, rights, left_set, right_set = [], [], [], []
    else:
        lefts, rights, left_set, right_set = _merge_sequence(nb, merge_order)
    ptrs = (_lib.f64p * nb)(*[_fp(m) for m in mats])
    ncells = (C.c_int64 * nb)(*[m.shape[0] for m in mats])
    ml = np.asarray(lefts, dtype=np.int32); mr = np.asarray(rights, dtype=np.int32)
    rptr, rn, keep = None, None, []
    if restrict is not None:
        keep = [None if r is None else np.ascontiguousarray(r, dtype=np.int32) for r in restrict]
        rptr = (_lib.i32p * nb)(*[C.cast(None, _lib.i32p) if r is None else _ip(r) for r in keep])
        rn = (C.c_int64 * nb)(*[0 if r is None else r.size for r in keep])
    skip = float("nan") if (min_batch_skip is None or (isinstance(min_batch_skip, float) and math.isnan(min_batch_skip))) else float(min_batch_skip)
    # Fresh pageable memory is faulted in at ~3 GB/s when a device-to-host copy first touches it, so a helper thread
    # allocates and touches the output buffer while the GPUs work; the final copy then runs at PCIe speed.
    import threading
    ntot = int(sum(m.shape[0] for m in mats))
    host_out = {}

    def _prefault():
        buf = np.empty((ntot, d), dtype=np.float64)
        buf.fill(0.0)
        host_out["buf"] = buf

    prefault = threading.Thread(target=_prefault, daemon=True)
    prefault.start()
    handle = C.c_void_p(None)
    _lib.call("b200mnn_reduced_mnn", ptrs, ncells, nb, d, 0, _ip(ml) if (nb > 1 and not auto_merge) else None,
              _ip(mr) if (nb > 1 and not auto_merge) else None, int(k),
              -1.0 if prop_k is None else float(prop_k), float(ndist), skip, rptr, rn, 1 if get_variance else 0, C.byref(handle))
    try:
        ntotal = int(_lib.load().b200mnn_result_ncells(handle))
        prefault.join()
        corrected = host_out["buf"] if host_out.get("buf") is not None and host_out["buf"].shape[0] == ntotal else np.empty((ntotal, d), dtype=np.float64)
        _lib.call("b200mnn_result_corrected", handle, _fp(corrected), 0)
        order = np.zeros(nb, dtype=np.int32); counts = np.zeros(nb, dtype=np.int64)
        nm = nb - 1
        batch_size = np.full(max(nm, 1), np.nan); skipped = np.zeros(max(nm, 1), dtype=np.int32); lost = np.zeros((max(nm, 1), nb))
        _lib.call("b200mnn_result_info", handle, _ip(order), counts.ctypes.data_as(_lib.i64p), _fp(batch_size), _ip(skipped), _fp(lost))
        if auto_merge and nm > 0:
            gl = np.zeros(nm, dtype=np.int32); gr = np.zeros(nm, dtype=np.int32)
            _lib.call("b200mnn_result_merges", handle, _ip(gl), _ip(gr))
            members = {i: [i + 1] for i in range(nb)}
            for m in range(nm):
                left_set.append(list(members[int(gl[m])])); right_set.append(list(members[int(gr[m])]))
                members[nb + m] = members[int(gl[m])] + members[int(gr[m])]
        pairings = []
        for m in range(nm):
            npairs = int(_lib.load().b200mnn_result_npairs(handle, m))
            first = np.zeros(max(npairs, 1), dtype=np.int32); second = np.zeros(max(npairs, 1), dtype=np.int32)
            _lib.call("b200mnn_result_pairs", handle, m, _ip(first), _ip(second))
            pairings.append((first[:npairs].astype(np.int64) - 1, second[:npairs].astype(np.int64) - 1))
    finally:
        _lib.load().b200mnn_result_free(handle)
    final = _Final([int(x) for x in order], np.repeat(order.astype(np.int64), counts))
    if nm == 0:
        return _finish(final, corrected, [], [], [], dict(batch_size=np.zeros(0), skipped=np.zeros(0, bool), lost_var=np.zeros((0, 1))))
    return _finish(final, corrected, pairings, left_set, right_set,
                   dict(batch_size=batch_size[:nm], skipped=skipped[:nm].astype(bool), lost_var=lost[:nm]))


def _fast_mnn_core(batches, k, prop_k, restrict, ndist, merge_order, min_batch_skip, get_variance=True, auto_merge=False) -> MNNResult:
    """One call of the C ABI (device-resident merge loop, all visible GPUs driven from this process) -- unless this process
    is one rank of a torch.distributed job (one GPU per process: the loop below shards the searches over the ranks) or
    B200MNN_PYLOOP=1 asks for the Python-driven loop."""
    import os

    import torch

    from . import device as dev

    dev.require_cuda()
    in_group = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    if auto_merge or (not in_group and os.environ.get("B200MNN_PYLOOP") != "1"):
        return _fast_mnn_core_c(batches, k, prop_k, restrict, ndist, merge_order, min_batch_skip, get_variance, auto_merge)
    cuda = torch.device("cuda", torch.cuda.current_device())
    nb = len(batches)
    d = batches[0].shape[1]
    for b in batches:
        if b.ndim != 2 or b.shape[1] != d:
            raise ValueError("number of columns is not the same across batches")
    tree = _predefined_tree(nb, merge_order)

    def leaf(i):
        data = torch.from_numpy(np.ascontiguousarray(batches[i - 1], dtype=np.float64)).to(cuda)
        r = None
        if restrict is not None and restrict[i - 1] is not None:
            r = torch.from_numpy(np.asarray(restrict[i - 1], dtype=np.int64) - 1).to(cuda)
        return _Node([i], data, r)

    # The corrected matrix goes back into a host array of the total size.  Fresh pageable memory is faulted in at ~3 GB/s
    # when the device-to-host copy first touches it (0.25 s for 2M x 50 doubles), so a helper thread allocates and
    # touches the buffer while the GPU works; the final copy then runs at PCIe speed.
    import threading

    ntotal = int(sum(b.shape[0] for b in batches))
    host_out = {}

    def _prefault():
        buf = np.empty((ntotal, d), dtype=np.float64)
        buf.fill(0.0)
        host_out["buf"] = buf

    prefault = threading.Thread(target=_prefault, daemon=True)
    prefault.start()

    def to_host(t):
        prefault.join()
        buf = host_out.get("buf")
        if buf is None or buf.shape != tuple(t.shape):
            return t.cpu().numpy()
        torch.from_numpy(buf).copy_(t)
        return buf

    tree = _fill(tree, leaf)
    if nb == 1:
        node = tree
        return _finish(node, to_host(node.data), [], [], [], dict(batch_size=np.zeros(0), skipped=np.zeros(0, bool), lost_var=np.zeros((0, 1))))
    nmerges = nb - 1
    pairings, left_set, right_set = [], [], []
    batch_size = np.full(nmerges, np.nan)
    skipped = np.zeros(nmerges, dtype=bool)
    var_kept = np.ones((nmerges, nb))
    for mdx in range(nmerges):
        left, right, path = _next_merge(tree)
        ld, rd = left.data.clone(), right.data.clone()
        if get_variance:
            left_old = _perbatch_var(ld, left.index, left.origin)
            right_old = _perbatch_var(rd, right.index, right.origin)
        left_set.append(list(left.index)); right_set.append(list(right.index))
        # orthogonalise each side along the other side's earlier batch vectors (R/fastMNN.R:473-474)
        for vec in left.extras:
            dev.center_along_batch_vector(rd, vec, right.restrict)
        for vec in right.extras:
            dev.center_along_batch_vector(ld, vec, left.restrict)
        first, second = _restricted_mnn(dev, ld, left.restrict, rd, right.restrict, k, prop_k)
        if first.numel() == 0:
            raise B200Error(1, "no MNN pairs found between the batches being merged")
        averaged, _ = dev.average_correction(ld, rd, first, second)
        overall = averaged.mean(dim=0)
        do_correct = True
        if min_batch_skip is not None and not (isinstance(min_batch_skip, float) and math.isnan(min_batch_skip)):
            ave_l2sq = float((averaged ** 2).mean(dim=0).sum().item())
            mag = 0.0 if ave_l2sq == 0 else math.sqrt(float((overall ** 2).sum().item()) / ave_l2sq)
            batch_size[mdx] = mag
            if mag < min_batch_skip:
                do_correct = False
                skipped[mdx] = True
        if do_correct:
            dev.center_along_batch_vector(ld, overall, left.restrict)
            dev.center_along_batch_vector(rd, overall, right.restrict)
            if get_variance:   # recorded straight after the centring, before the tricube smoothing (R/fastMNN.R:500-501)
                left_new = _perbatch_var(ld, left.index, left.origin)
                right_new = _perbatch_var(rd, right.index, right.origin)
            to_add = [overall]
            re_avg, re_second = dev.average_correction(ld, rd, first, second)
            kk = min(_choose_k(k, prop_k, rd.shape[0]), re_second.shape[0])
            sub = rd.index_select(0, re_second.long())
            idx, dist = dev.query_knn_sharded(sub, rd, kk, want_dist=True)
            rd = dev.tricube_apply(rd, re_avg, idx, dist, ndist)
        else:
            to_add = []
            if get_variance:   # R/fastMNN.R:510-512
                left_new = _perbatch_var(ld, left.index, left.origin)
                right_new = _perbatch_var(rd, right.index, right.origin)
        if get_variance:
            var_kept[mdx, np.asarray(left.index) - 1] = left_new / left_old
            var_kept[mdx, np.asarray(right.index) - 1] = right_new / right_old
        pairings.append((first.cpu().numpy(), second.cpu().numpy()))
        node = _Node(left.index + right.index, torch.cat([ld, rd], dim=0),
                     _combine_restrict(ld.shape[0], left.restrict, rd.shape[0], right.restrict, cuda),
                     origin=np.concatenate([left.origin, right.origin]), extras=left.extras + right.extras + to_add)
        tree = _update(tree, path, node)
    return _finish(tree, to_host(tree.data), pairings, left_set, right_set,
                   dict(batch_size=batch_size, skipped=skipped, lost_var=1 - var_kept))


def reducedMNN(*batches, batch=None, k=20, prop_k=None, restrict=None, ndist=3, merge_order=None, auto_merge=False,
               min_batch_skip=0, BNPARAM=None, BPPARAM=None) -> MNNResult:
    """reducedMNN (R/reducedMNN.R:61-95): MNN correction of precomputed low-dimensional coordinates [cells x dims]."""
    _check_bpparam(BPPARAM)
    _select_device(BNPARAM)
    mats = [np.asarray(b, dtype=np.float64) for b in batches]
    if len(mats) == 0:
        raise ValueError("at least one batch must be supplied")
    if len(mats) == 1:
        parts, reorder, restricted, _ = _divide_into_batches(mats[0], batch, True, None if restrict is None else restrict[0])
        out = _fast_mnn_core(parts, k, prop_k, restricted, ndist, merge_order, min_batch_skip, auto_merge=auto_merge)
        out.corrected = out.corrected[reorder - 1]
        out.batch = out.batch[reorder - 1]
        out.merge_info["pairs"] = _reindex_pairings(out.merge_info["pairs"], reorder)
        return out
    restrict = _check_restrict([m.shape[0] for m in mats], restrict)
    return _fast_mnn_core(mats, k, prop_k, restrict, ndist, merge_order, min_batch_skip, auto_merge=auto_merge)


def multiBatchPCA(*batches, d=50, subset_row=None, weights=None, get_all_genes=False, get_variance=False, BSPARAM=None, BPPARAM=None):
    """multiBatchPCA (R/multiBatchPCA.R:211-322, .multi_pca_list) for [genes x cells] matrices: every batch centred on the
    weighted grand mean of the batch means, scaled by 1/sqrt(ncells/weight) so that each batch contributes equally, exact
    SVD of the scaled matrix, UNscaled centred batches projected on its left singular vectors.

    Returns ``(pcs, meta)``: ``pcs[b]`` is [cells x d]; ``meta`` has ``rotation`` [genes x d], ``centers`` and, with
    ``get_variance``, ``var_explained`` (d^2 / nbatches, :420) and ``var_total``.  Rotation columns are sign-normalised
    (largest |component| positive); an SVD leaves the sign free, the reference's LAPACK call included.

    B200 form (a front end of the hot path, not part of it): the G x G Gram matrix of the scaled
    data and the projections are fp64 GEMMs and the eigen-decomposition is cuSOLVER's, all through torch -- library code
    by design here; exact (ExactParam) semantics, no randomised SVD.  ``weights``: None/True (equal), False (by cell
    count) or one number per batch; tree-like weights are not supported."""
    import torch

    from . import device as dev

    _check_bpparam(BPPARAM)
    dev.require_cuda()
    cuda = torch.device("cuda", torch.cuda.current_device())
    mats = [np.asarray(b, dtype=np.float64) for b in batches]
    if len(mats) == 0:
        raise ValueError("at least one batch must be supplied")
    G = mats[0].shape[0]
    for m in mats:
        if m.ndim != 2 or m.shape[0] != G:
            raise ValueError("number of rows is not the same across batches")
    ncells = np.array([m.shape[1] for m in mats], dtype=np.float64)
    if weights is None or weights is True:
        w = np.ones(len(mats))
    elif weights is False:
        w = ncells.copy()
    elif isinstance(weights, (list, tuple, np.ndarray)) and not any(isinstance(x, (list, tuple)) for x in weights):
        w = np.asarray(weights, dtype=np.float64)
        if w.size != len(mats):
            raise ValueError("'length(weights)' should be the same as number of entries in '...'")
    else:
        raise NotImplementedError("tree-like 'weights' are not supported")
    keep = None if subset_row is None else _subset_to_index(subset_row, G) - 1
    dev_mats = [torch.from_numpy(np.ascontiguousarray(m.T)).to(cuda) for m in mats]           # [cells x genes]

    def process(cols):
        sub = [x if cols is None else x.index_select(1, torch.from_numpy(cols).to(cuda)) for x in dev_mats]
        centers = sum(x.mean(dim=0) * wi for x, wi in zip(sub, w)) / float(w.sum())        # grand average of batch centres (:268-281)
        centred = [x - centers for x in sub]
        scaled = [c / math.sqrt(n / wi) for c, n, wi in zip(centred, ncells, w)]             # (:307-313)
        return centers, centred, scaled

    centers, centred, scaled = process(keep)
    nsel = centred[0].shape[1]
    d_eff = int(min(d, nsel, int(ncells.sum())))
    gram = sum(sc.T @ sc for sc in scaled)                                                   # [genes x genes]
    evals, evecs = torch.linalg.eigh(gram)
    order = torch.argsort(evals, descending=True)[:d_eff]
    u = evecs.index_select(1, order)
    sv2 = torch.clamp(evals.index_select(0, order), min=0.0)
    piv = torch.argmax(u.abs(), dim=0)
    u = u * torch.sign(u[piv, torch.arange(d_eff, device=cuda)])[None, :]
    pcs = [(c @ u).cpu().numpy() for c in centred]                                           # crossprod(centered, u) (:238-240)
    rotation, all_centers = u, centers
    if get_all_genes and keep is not None:                                                   # .make_pca_metadata (:393-412)
        rest = np.setdiff1d(np.arange(G), keep)
        lc, _, lscaled = process(rest)
        # leftover.u = left.scaled %*% v / d with v = t(scaled) u / d  ->  (sum_b lscaled_b^T scaled_b) u / d^2
        cross = sum(ls.T @ sc for ls, sc in zip(lscaled, scaled))
        left_u = (cross @ u) / sv2[None, :]
        rotation = torch.zeros((G, d_eff), dtype=torch.float64, device=cuda)
        rotation[torch.from_numpy(keep).to(cuda)] = u
        rotation[torch.from_numpy(rest).to(cuda)] = left_u
        all_centers = torch.zeros(G, dtype=torch.float64, device=cuda)
        all_centers[torch.from_numpy(keep).to(cuda)] = centers
        all_centers[torch.from_numpy(rest).to(cuda)] = lc
    meta = {"rotation": rotation.cpu().numpy(), "centers": all_centers.cpu().numpy()}
    if get_variance:
        meta["var_explained"] = (sv2 / len(mats)).cpu().numpy()
        meta["var_total"] = float(sum((sc ** 2).sum() for sc in scaled).item()) / len(mats)
    return pcs, meta


def fastMNN(*batches, batch=None, k=20, prop_k=None, restrict=None, cos_norm=True, ndist=3, d=50, merge_order=None,
            auto_merge=False, min_batch_skip=0, subset_row=None, weights=None, get_variance=False, BNPARAM=None, BPPARAM=None) -> MNNResult:
    """fastMNN (R/fastMNN.R:283-331) for [genes x cells] matrices: cosineNorm -> multi-batch PCA -> reducedMNN.

    The PCA front end is :func:`multiBatchPCA` (R/multiBatchPCA.R:211-322: batch weights, ``subset_row``, variance
    explained; fp64 Gram matrix + ``eigh`` on the device).  ``corrected`` is [cells x d]."""
    import torch

    mats = [np.asarray(b, dtype=np.float64) for b in batches]
    if len(mats) == 1:
        parts, reorder, restricted, _ = _divide_into_batches(mats[0], batch, False, None if restrict is None else restrict[0])
        out = fastMNN(*parts, k=k, prop_k=prop_k, restrict=restricted, cos_norm=cos_norm, ndist=ndist, d=d, merge_order=merge_order,
                      auto_merge=auto_merge, min_batch_skip=min_batch_skip, subset_row=subset_row, weights=weights,
                      get_variance=get_variance, BNPARAM=BNPARAM, BPPARAM=BPPARAM)
        out.corrected = out.corrected[reorder - 1]
        out.batch = out.batch[reorder - 1]
        out.merge_info["pairs"] = _reindex_pairings(out.merge_info["pairs"], reorder)
        return out
    if cos_norm:   # R/fastMNN.R:349-350: cosine normalisation over the selected genes
        mats = [cosineNorm(m, subset_row=subset_row) for m in mats]
        subset_row = None
    pcs, meta = multiBatchPCA(*mats, d=d, subset_row=subset_row, weights=weights, get_variance=get_variance, BPPARAM=BPPARAM)   # :353
    out = reducedMNN(*pcs, k=k, prop_k=prop_k, restrict=restrict, ndist=ndist, merge_order=merge_order, auto_merge=auto_merge,
                     min_batch_skip=min_batch_skip, BNPARAM=BNPARAM, BPPARAM=BPPARAM)
    out.merge_info["rotation"] = meta["rotation"]
    out.merge_info["pca"] = meta
    return out


# ----------------------------------------------------------------------------------------------------------
# mnnCorrect (R/mnnCorrect.R:125-393) -- device-resident merge loop
# ----------------------------------------------------------------------------------------------------------
def mnnCorrect(*batches, batch=None, restrict=None, k=20, prop_k=None, sigma=0.1, cos_norm_in=True, cos_norm_out=True,
               svd_dim=0, var_adj=True, subset_row=None, correct_all=False, merge_order=None, auto_merge=False,
               BNPARAM=None, BPPARAM=None, _timings=None) -> MNNResult:
    """mnnCorrect for [genes x cells] matrices; ``corrected`` is [genes x cells] in the input batch order.

    ``_timings`` (measurement aid, not part of the reference's signature): a dict that receives the seconds spent per
    stage (upload, cosine norm, MNN search, averaging, smoothing, shift variance, download), each bracketed by a
    device synchronisation."""
    import time

    import torch

    def _tick(name, t0):
        if _timings is not None:
            torch.cuda.synchronize()
            _timings[name] = _timings.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    from . import device as dev

    _check_bpparam(BPPARAM)
    _select_device(BNPARAM)
    if svd_dim:
        raise NotImplementedError("svd.dim > 0 (biological-subspace removal) is outside the accelerated path (default is 0)")
    mats = [np.asarray(b, dtype=np.float64) for b in batches]
    do_split = len(mats) == 1
    if do_split:
        mats, reorder, restrict, _ = _divide_into_batches(mats[0], batch, False, None if restrict is None else restrict[0])
    else:
        restrict = _check_restrict([m.shape[1] for m in mats], restrict)
    if len(mats) < 2:
        raise ValueError("at least two batches must be specified")
    G = mats[0].shape[0]
    for m in mats:
        if m.shape[0] != G:
            raise ValueError("number of rows is not the same across batches")
    dev.require_cuda()
    cuda = torch.device("cuda", torch.cuda.current_device())

    # .prepare_input_data (R/mnnCorrect.R:398-442); device layout is [cells x genes]: R's column-major [genes x cells]
    # IS that layout, so a Fortran-ordered input uploads without any host-side transpose
    t0 = time.perf_counter()
    raw = [torch.from_numpy(m.T if m.flags.f_contiguous else np.ascontiguousarray(m.T)).to(cuda) for m in mats]
    t0 = _tick("upload", t0)
    sub = None
    if subset_row is not None:
        sidx = _subset_to_index(subset_row, G)
        if not np.array_equal(sidx, np.arange(1, G + 1)):
            sub = torch.from_numpy(sidx - 1).to(cuda)
    in_b = [r if sub is None else r.index_select(1, sub).contiguous() for r in raw]
    same_set = True
    if sub is not None and correct_all:
        same_set = False
        out_b = list(raw)
    else:
        out_b = list(in_b)
    norms = None
    if cos_norm_in:
        normed = [dev.cosine_norm(b) for b in in_b]
        in_b = [n[0] for n in normed]
        norms = [n[1] for n in normed]
        if same_set and cos_norm_out:
            out_b = list(in_b)
    if cos_norm_out:
        if norms is None:
            norms = [dev.cosine_norm(b, want_matrix=False)[1] for b in in_b]
        if not (same_set and cos_norm_in):
            out_b = [b / torch.clamp(l2, min=1e-8)[:, None] for b, l2 in zip(out_b, norms)]
    if cos_norm_out != cos_norm_in:
        same_set = False

    t0 = _tick("cosine_norm", t0)
    nb = len(mats)
    tree = _predefined_tree(nb, merge_order)

    def leaf(i):
        r = None
        if restrict is not None and restrict[i - 1] is not None:
            r = torch.from_numpy(np.asarray(restrict[i - 1], dtype=np.int64) - 1).to(cuda)
        return _Node([i], in_b[i - 1], r, extras=[None if same_set else out_b[i - 1]])

    def count_pairs(a, b):
        """.count_mnn_pairs with orthogonalize=FALSE (R/mnnCorrect.R:212): MNN pairs between two nodes."""
        f, _ = _restricted_mnn(dev, a.data, a.restrict, b.data, b.restrict, k, prop_k)
        return int(f.numel())

    if auto_merge:   # .initialize_auto_search (R/MNN_tree.R:154-168)
        remainders = [leaf(i) for i in range(1, nb + 1)]
        pairwise = np.zeros((nb, nb), dtype=np.int64)
        for i in range(nb):
            for j in range(i):
                pairwise[i, j] = count_pairs(remainders[i], remainders[j])
    else:
        tree = _fill(tree, leaf)
    pairings, left_set, right_set = [], [], []
    for _ in range(nb - 1):
        if auto_merge:   # .pick_best_merge: first maximum in column-major order; row = left, column = right
            cols, rows = np.nonzero(pairwise.T == pairwise.max())
            chosen = (int(rows[0]), int(cols[0]))
            left, right = remainders[chosen[0]], remainders[chosen[1]]
        else:
            left, right, path = _next_merge(tree)
        ld, rd = left.data, right.data
        lx, rx = left.extras[0], right.extras[0]
        t0 = time.perf_counter()
        s1, s2 = _restricted_mnn(dev, ld, left.restrict, rd, right.restrict, k, prop_k)
        if s1.numel() == 0:
            raise B200Error(1, "no MNN pairs found between the batches being merged")
        pairings.append((s1.cpu().numpy(), s2.cpu().numpy()))
        t0 = _tick("mnn_search", t0)
        left_set.append(list(left.index)); right_set.append(list(right.index))

        def correction(d1, d2, adjust_sub):
            # .compute_correction_vectors (R/mnnCorrect.R:451-460): distances always in the *input* space (rd)
            tc = time.perf_counter()
            averaged, uniq = dev.average_correction(d1, d2, s1, s2)
            tc = _tick("average_correction", tc)
            if _timings is not None:
                _timings["mnn_cells"] = int(uniq.shape[0])
            cor = dev.smooth_gaussian_kernel(averaged, uniq, rd, sigma)
            tc = _tick("smooth_gaussian_kernel", tc)
            if _timings is not None:
                _timings.setdefault("_smooth_io", []).append((averaged, uniq, rd, cor))
            if var_adj:  # .adjust_shift_variance (R/mnnCorrect.R:462-481)
                r1 = left.restrict if left.restrict is not None else torch.arange(d1.shape[0], device=cuda)
                r2 = right.restrict if right.restrict is not None else torch.arange(d2.shape[0], device=cuda)
                if adjust_sub is not None:
                    scaling = dev.adjust_shift_variance(d1.index_select(1, adjust_sub).contiguous(), d2.index_select(1, adjust_sub).contiguous(),
                                                        cor.index_select(1, adjust_sub).contiguous(), sigma, r1, r2)
                else:
                    scaling = dev.adjust_shift_variance(d1, d2, cor, sigma, r1, r2)
                if _timings is not None:   # what a sampled parity check of this stage needs: its inputs and its output
                    _timings.setdefault("_shiftvar_io", []).append((d1, d2, cor, r1, r2, scaling))
                cor = torch.clamp(scaling, min=1.0)[:, None] * cor  # pmax(scaling, 1) * correction
                tc = _tick("adjust_shift_variance", tc)
            return cor

        cor_in = correction(ld, rd, None)
        new_rd = rd + cor_in
        new_rx = None
        if not same_set:
            new_rx = rx + correction(lx, rx, sub if (sub is not None and correct_all) else None)
        node = _Node(left.index + right.index, torch.cat([ld, new_rd], dim=0),
                     _combine_restrict(ld.shape[0], left.restrict, rd.shape[0], right.restrict, cuda),
                     origin=np.concatenate([left.origin, right.origin]),
                     extras=[None if same_set else torch.cat([lx, new_rx], dim=0)])
        if auto_merge:   # .update_remainders (R/MNN_tree.R:205-226)
            keep = [i for i in range(len(remainders)) if i not in chosen]
            remainders = [remainders[i] for i in keep]
            if remainders:
                old_meta = pairwise[np.ix_(keep, keep)]
                new_stats = np.array([count_pairs(node, r) for r in remainders], dtype=np.int64)
                pairwise = np.hstack([np.vstack([old_meta, new_stats[None, :]]), np.zeros((len(keep) + 1, 1), dtype=np.int64)])
                remainders.append(node)
            else:
                tree = node
        else:
            tree = _update(tree, path, node)
    t0 = time.perf_counter()
    full = (tree.data if same_set else tree.extras[0]).cpu().numpy()
    res = _finish(tree, full, pairings, left_set, right_set, {})
    res.corrected = res.corrected.T   # [genes x cells], Fortran-ordered like an R matrix: no host-side transpose
    t0 = _tick("download", t0)
    if do_split:
        res.corrected = res.corrected[:, reorder - 1]
        res.batch = res.batch[reorder - 1]
        res.merge_info["pairs"] = _reindex_pairings(res.merge_info["pairs"], reorder)
    return res


# ----------------------------------------------------------------------------------------------------------
# clusterMNN's propagation step (R/clusterMNN.R:244-312) -- SURVEY.md section 8f, N4
# ----------------------------------------------------------------------------------------------------------
def propagate_to_cells(cells, centroids, corrected_centroids, restrict=None, BNPARAM=None, BPPARAM=None):
    """.propagate_to_cells for one batch in PC space (R/clusterMNN.R:267-283): `cells` [ncells x d] (already projected),
    `centroids` [nc x d] and their MNN-corrected positions.  sigma = median distance of the (restricted) cells to their
    nearest centroid (exact search, k = 1); returns cells + Gaussian-weighted centroid corrections."""
    import torch

    from . import device as dev

    _check_bpparam(BPPARAM)
    _select_device(BNPARAM)
    dev.require_cuda()
    cuda = torch.device("cuda", torch.cuda.current_device())
    x = torch.from_numpy(np.ascontiguousarray(cells, dtype=np.float64)).to(cuda)
    cen = torch.from_numpy(np.ascontiguousarray(centroids, dtype=np.float64)).to(cuda)
    delta = torch.from_numpy(np.ascontiguousarray(corrected_centroids, dtype=np.float64)).to(cuda) - cen
    q = x if restrict is None else x.index_select(0, torch.from_numpy(_subset_to_index(restrict, x.shape[0]) - 1).to(cuda))
    _, dist = dev.query_knn(cen, q, 1, want_dist=True)
    sigma = float(np.median(dist[:, 0].cpu().numpy()))     # stats::median on the host, as the reference does
    return dev.smooth_gaussian_from_centroids(x, cen, delta, sigma).cpu().numpy()
