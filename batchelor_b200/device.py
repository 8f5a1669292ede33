"""Device-resident wrappers over the ``b200mnn_dev_*`` C ABI.

torch supplies device memory, the current CUDA stream and (optionally) the ``torch.distributed`` process group;
every numeric step is a kernel of libb200mnn.  Tensors are fp64, row-major ``[cells x dims]``; ids are int32,
0-based.  Nothing here falls back to torch math when the library or the GPU is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

_launches = 0  # kernels of libb200mnn launched through this module (bench.py reports it as gpu_launches)


def launches() -> int:
    """Kernels of libb200mnn launched by this process so far (counted inside the library at every launch site)."""
    return int(_lib.load().b200mnn_launch_count())


def _count(n: int) -> None:
    global _launches
    _launches += n


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float64 or not t.is_cuda:
        raise TypeError("expected a CUDA float64 tensor")
    return t.contiguous()


def _i32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise TypeError("expected a CUDA tensor")
    return t.to(torch.int32).contiguous()


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise _lib.B200Error(2, "no usable CUDA device: batchelor_b200 has no CPU fallback")
    _lib.load()


# ----------------------------------------------------------------------------------------------------------
# a1: exact kNN
# ----------------------------------------------------------------------------------------------------------
def query_knn(X: torch.Tensor, Q: torch.Tensor, k: int, want_dist: bool = True, stats: Optional[torch.Tensor] = None
              ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """k nearest rows of X for every row of Q: (idx int32 [nq,k] 0-based, dist float64 [nq,k] or None)."""
    X = _f64(X); Q = _f64(Q)
    n, d = X.shape
    nq = Q.shape[0]
    if Q.shape[1] != d:
        raise ValueError("X and query must have the same number of dimensions")
    idx = torch.empty((nq, k), dtype=torch.int32, device=X.device)
    dist = torch.empty((nq, k), dtype=torch.float64, device=X.device) if want_dist else None
    _lib.call("b200mnn_dev_query_knn", _p(X), n, _p(Q), nq, d, k, _p(idx), _p(dist), _p(stats), _stream())
    _count(8)
    return idx, dist


def query_knn_sharded(X: torch.Tensor, Q: torch.Tensor, k: int, want_dist: bool = True, min_rows_per_rank: int = 2048
                      ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Query-sharded kNN: the reference set X is replicated, rank r searches a contiguous block of query rows and the
    per-shard (index, distance) blocks are all-gathered over NCCL.  Falls through to :func:`query_knn` when there is
    no process group (or the problem is tiny)."""
    import torch.distributed as dist_

    if not (dist_.is_available() and dist_.is_initialized()) or dist_.get_world_size() == 1:
        return query_knn(X, Q, k, want_dist)
    if Q.shape[0] < dist_.get_world_size() * min_rows_per_rank:
        return query_knn(X, Q, k, want_dist)
    return shard_rows_and_gather(Q.shape[0], k, lambda lo, hi: query_knn(X, Q[lo:hi], k, want_dist), X.device, want_dist)


def shard_bounds(nq: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous block of query rows owned by `rank`: (lo, hi, rows_per_rank)."""
    per = (nq + world - 1) // world
    return min(nq, rank * per), min(nq, (rank + 1) * per), per


def all_gather_rows(local: torch.Tensor, n: int, per: int) -> torch.Tensor:
    """Exchange step of the query-sharded search (backend-agnostic: NCCL on GPUs, gloo in the CPU tests): pads this
    rank's block of rows to `per` rows, all-gathers the blocks of every rank in rank order and returns the first n rows."""
    import torch.distributed as dist_

    ws = dist_.get_world_size()
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = torch.empty((ws * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist_.all_gather_into_tensor(full, pad)
    return full[:n]


def upload_sharded(host: torch.Tensor, device) -> torch.Tensor:
    """Host matrix -> replicated device matrix, paying the PCIe transfer once per row instead of once per rank: every rank
    copies only its contiguous block of rows (:func:`shard_bounds`) and the blocks are all-gathered over NVLink.  `host`
    must hold the same data on every rank (pinned memory makes the copy asynchronous).  Without a process group this is a
    plain upload."""
    import torch.distributed as dist_

    if not (dist_.is_available() and dist_.is_initialized()) or dist_.get_world_size() == 1:
        return host.to(device, non_blocking=True)
    ws, rank = dist_.get_world_size(), dist_.get_rank()
    n = host.shape[0]
    lo, hi, per = shard_bounds(n, ws, rank)
    local = host[lo:hi].to(device, non_blocking=True)
    return all_gather_rows(local, n, per)


def shard_rows_and_gather(nq: int, k: int, compute_local, device, want_dist: bool = True):
    """Host-side sharding logic: every rank computes `compute_local(lo, hi) -> (idx[hi-lo, k], dist or None)` for its
    contiguous block of rows (:func:`shard_bounds`); the blocks are exchanged with :func:`all_gather_rows`, so every rank
    ends with the full [nq, k] result in row order."""
    import torch.distributed as dist_

    ws, rank = dist_.get_world_size(), dist_.get_rank()
    lo, hi, per = shard_bounds(nq, ws, rank)
    idx_l, dist_l = compute_local(lo, hi)
    idx_all = all_gather_rows(idx_l.to(torch.int32), nq, per)
    dist_all = all_gather_rows(dist_l.to(torch.float64), nq, per) if want_dist else None
    return idx_all, dist_all


def debug_candidates(X: torch.Tensor, Q: torch.Tensor, k: int):
    """Tensor-core scoring stage only: (cand_idx [nq,C], approx_d2 [nq,C], thr [nq])."""
    X = _f64(X); Q = _f64(Q)
    n, d = X.shape
    nq = Q.shape[0]
    cap = 8 * 64
    cidx = torch.empty((nq, cap), dtype=torch.int32, device=X.device)
    cd2 = torch.empty((nq, cap), dtype=torch.float64, device=X.device)
    thr = torch.empty((nq,), dtype=torch.float64, device=X.device)
    nc = C.c_int64(0)
    _lib.call("b200mnn_dev_debug_candidates", _p(X), n, _p(Q), nq, d, k, _p(cidx), _p(cd2), _p(thr), cap, C.byref(nc), _stream())
    c = nc.value
    return cidx.view(-1)[: nq * c].view(nq, c), cd2.view(-1)[: nq * c].view(nq, c), thr


# ----------------------------------------------------------------------------------------------------------
# a3: mutual pairs
# ----------------------------------------------------------------------------------------------------------
def find_mutual_nns(left: torch.Tensor, right: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """left [n1,k2] ids into batch 2, right [n2,k1] ids into batch 1 (0-based) -> (first, second) int32, 0-based."""
    left = _i32(left); right = _i32(right)
    n1, k2 = left.shape
    n2, k1 = right.shape
    cap = max(1, n1 * k2)
    first = torch.empty((cap,), dtype=torch.int32, device=left.device)
    second = torch.empty((cap,), dtype=torch.int32, device=left.device)
    npairs = torch.zeros((1,), dtype=torch.int64, device=left.device)
    _lib.call("b200mnn_dev_find_mutual_nns", _p(left), n1, k2, _p(right), n2, k1, _p(first), _p(second), cap, _p(npairs), _stream())
    _count(5)
    m = int(npairs.item())
    return first[:m], second[:m]


def gather_pair_blocks(first_l: torch.Tensor, second_l: torch.Tensor, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exchange step of the row-sharded pair extraction (backend-agnostic): every rank holds the pairs of its block of batch-1
    rows (already in the reference's order, `first` in global numbering); the blocks have different lengths, so the counts are
    all-gathered first, the blocks padded to the longest, all-gathered, and the valid parts concatenated in rank order -- which
    IS the reference's order, because the row blocks are contiguous and ascending."""
    import torch.distributed as dist_

    cnt = torch.tensor([first_l.shape[0]], dtype=torch.int64, device=first_l.device)
    counts = torch.empty((world,), dtype=torch.int64, device=first_l.device)
    dist_.all_gather_into_tensor(counts, cnt)
    counts_h = counts.tolist()
    m = max(1, max(counts_h))
    pad = torch.zeros((2, m), dtype=torch.int32, device=first_l.device)
    pad[0, : first_l.shape[0]] = first_l
    pad[1, : second_l.shape[0]] = second_l
    full = torch.empty((world, 2, m), dtype=torch.int32, device=first_l.device)
    dist_.all_gather_into_tensor(full.view(world * 2, m), pad)
    first = torch.cat([full[r, 0, : counts_h[r]] for r in range(world)])
    second = torch.cat([full[r, 1, : counts_h[r]] for r in range(world)])
    return first, second


def find_mutual_nns_sharded(w21: torch.Tensor, w12: torch.Tensor, world: int, rank: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pair extraction over ranks: rank r probes its contiguous block of batch-1 rows (the probes into w12 are the work) and
    the per-block pair lists are exchanged (:func:`gather_pair_blocks`).  Every rank ends with the full lists."""
    n1 = w21.shape[0]
    lo, hi, _ = shard_bounds(n1, world, rank)
    if hi > lo:
        # the extraction identifies a batch-1 row by its position in `left`: shift the ids in w12 so that the block's rows
        # are 0 .. hi-lo-1 (ids outside the block then match nothing)
        f, s2 = find_mutual_nns(w21[lo:hi], w12 - lo if lo else w12)
        f = f + lo
    else:
        f = torch.zeros((0,), dtype=torch.int32, device=w21.device)
        s2 = torch.zeros((0,), dtype=torch.int32, device=w21.device)
    return gather_pair_blocks(f.to(torch.int32), s2.to(torch.int32), world)


_side_streams = {}


def _two_side_streams(device):
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _side_streams[key]


def direction_split(n1: int, n2: int, world: int, rank: int) -> Tuple[int, int, int, int, int]:
    """Work split of findMutualNN over `world` >= 2 ranks: the first h ranks search (contiguous blocks of) the batch-1
    rows in batch 2, the other world - h ranks the batch-2 rows in batch 1, h proportional to n1 : n2.  Every rank then
    builds the reference-side plan of ONE direction only (with both directions on every rank that plan -- k-means,
    grouping, operand preparation: work that does not shrink with the shard -- is paid twice per rank).
    Returns (direction, lo, hi, rows_per_rank_of_that_direction, h)."""
    h = min(world - 1, max(1, int(round(world * n1 / max(1, n1 + n2)))))
    if rank < h:
        lo, hi, per = shard_bounds(n1, h, rank)
        return 0, lo, hi, per, h
    lo, hi, per = shard_bounds(n2, world - h, rank - h)
    return 1, lo, hi, per, h


def direction_split_gather(local: torch.Tensor, n1: int, k2: int, n2: int, k1: int, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exchange step of :func:`direction_split` (backend-agnostic like :func:`all_gather_rows`): `local` is this rank's
    block of neighbour indices ([rows, k2] on the first h ranks, [rows, k1] on the others).  One all-gather of equally
    padded blocks; returns (w21 [n1, k2], w12 [n2, k1]) on every rank."""
    import torch.distributed as dist_

    h = direction_split(n1, n2, world, 0)[4]
    per1 = (n1 + h - 1) // h
    per2 = (n2 + (world - h) - 1) // (world - h)
    per, kmax = max(per1, per2), max(k1, k2)
    pad = torch.zeros((per, kmax), dtype=torch.int32, device=local.device)
    pad[: local.shape[0], : local.shape[1]] = local
    full = torch.empty((world, per, kmax), dtype=torch.int32, device=local.device)
    dist_.all_gather_into_tensor(full.view(world * per, kmax), pad)
    w21 = full[:h, :per1, :k2].reshape(h * per1, k2)[:n1].contiguous()
    w12 = full[h:, :per2, :k1].reshape((world - h) * per2, k1)[:n2].contiguous()
    return w21, w12


def find_mutual_nn(data1: torch.Tensor, data2: torch.Tensor, k1: int, k2: int, sharded: bool = True):
    """findMutualNN(data1, data2, k1, k2): two exact searches + mutual pairs.  Returns (first, second, w21, w12).

    The two searches are independent, so their kernels are enqueued on two side streams: the serial stretches of one
    direction's cluster plan (a one-block seeding kernel, small scans) overlap with the other direction's work.  With a
    process group the ranks split the DIRECTIONS first and the query rows second (:func:`direction_split`), and one
    all-gather of the index blocks follows on the caller's stream."""
    import torch.distributed as dist_

    k1 = min(k1, data1.shape[0]); k2 = min(k2, data2.shape[0])
    n1, n2 = data1.shape[0], data2.shape[0]
    ws = dist_.get_world_size() if (sharded and dist_.is_available() and dist_.is_initialized()) else 1
    if ws > 1 and min(n1, n2) >= ws * 2048:
        direction, lo, hi, _, _ = direction_split(n1, n2, ws, dist_.get_rank())
        if direction == 0:
            mine, _ = query_knn(data2, data1[lo:hi], k2, want_dist=False)   # neighbours of batch-1 cells in batch 2
        else:
            mine, _ = query_knn(data1, data2[lo:hi], k1, want_dist=False)   # neighbours of batch-2 cells in batch 1
        w21, w12 = direction_split_gather(mine, n1, k2, n2, k1, ws)
        first, second = find_mutual_nns_sharded(w21, w12, ws, dist_.get_rank())
        return first, second, w21, w12
    split1 = ws > 1 and n1 >= ws * 2048     # batch-1 cells are the queries of the first search
    split2 = ws > 1 and n2 >= ws * 2048
    rank = dist_.get_rank() if ws > 1 else 0
    lo1, hi1, per1 = shard_bounds(n1, ws, rank) if split1 else (0, n1, n1)
    lo2, hi2, per2 = shard_bounds(n2, ws, rank) if split2 else (0, n2, n2)
    cur = torch.cuda.current_stream()
    s1, s2 = _two_side_streams(data1.device)
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        a, _ = query_knn(data2, data1[lo1:hi1], k2, want_dist=False)   # neighbours of batch-1 cells in batch 2
    with torch.cuda.stream(s2):
        b, _ = query_knn(data1, data2[lo2:hi2], k1, want_dist=False)   # neighbours of batch-2 cells in batch 1
    cur.wait_stream(s1); cur.wait_stream(s2)
    a.record_stream(cur); b.record_stream(cur)

    w21 = all_gather_rows(a, n1, per1) if split1 else a
    w12 = all_gather_rows(b, n2, per2) if split2 else b
    first, second = find_mutual_nns(w21, w12)
    return first, second, w21, w12


# ----------------------------------------------------------------------------------------------------------
# a4 / a8 / a6 / a9
# ----------------------------------------------------------------------------------------------------------
def average_correction(ref: torch.Tensor, cur: torch.Tensor, first: torch.Tensor, second: torch.Tensor):
    ref = _f64(ref); cur = _f64(cur); first = _i32(first); second = _i32(second)
    n1, d = ref.shape
    n2 = cur.shape[0]
    npairs = first.shape[0]
    cap = max(1, min(npairs, n2))
    averaged = torch.empty((cap, d), dtype=torch.float64, device=ref.device)
    uniq = torch.empty((cap,), dtype=torch.int32, device=ref.device)
    nmnn = torch.zeros((1,), dtype=torch.int64, device=ref.device)
    _lib.call("b200mnn_dev_average_correction", _p(ref), n1, _p(cur), n2, d, _p(first), _p(second), npairs, _p(averaged), _p(uniq),
              _p(nmnn), _stream())
    _count(10)
    m = int(nmnn.item())
    return averaged[:m], uniq[:m]


def center_along_batch_vector(mat: torch.Tensor, batch_vec: torch.Tensor, restrict: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place on a contiguous fp64 CUDA matrix; returns it."""
    if not mat.is_contiguous():
        raise ValueError("mat must be contiguous (the kernel works in place)")
    mat = _f64(mat); batch_vec = _f64(batch_vec)
    r = _i32(restrict) if restrict is not None else None
    _lib.call("b200mnn_dev_center_along_batch_vector", _p(mat), mat.shape[0], mat.shape[1], _p(batch_vec), _p(r),
              0 if r is None else r.shape[0], _stream())
    _count(5)
    return mat


def tricube_apply(cur: torch.Tensor, correction: torch.Tensor, idx: torch.Tensor, dist: torch.Tensor, ndist: float) -> torch.Tensor:
    cur = _f64(cur); correction = _f64(correction); idx = _i32(idx); dist = _f64(dist)
    out = torch.empty_like(cur)
    _lib.call("b200mnn_dev_tricube_apply", _p(cur), cur.shape[0], cur.shape[1], _p(correction), correction.shape[0], _p(idx), _p(dist),
              idx.shape[1], float(ndist), _p(out), _stream())
    _count(1)
    return out


def cosine_norm(x: torch.Tensor, want_matrix: bool = True):
    """x [cells, genes] -> (normalised or None, l2norm [cells])."""
    x = _f64(x)
    out = torch.empty_like(x) if want_matrix else None
    l2 = torch.empty((x.shape[0],), dtype=torch.float64, device=x.device)
    _lib.call("b200mnn_dev_cosine_norm", _p(x), x.shape[0], x.shape[1], _p(out), _p(l2), _stream())
    _count(1)
    return out, l2


# ----------------------------------------------------------------------------------------------------------
# a5 / a7
# ----------------------------------------------------------------------------------------------------------
def smooth_gaussian_kernel(averaged: torch.Tensor, index0: torch.Tensor, mat: torch.Tensor, sigma2: float) -> torch.Tensor:
    """averaged [nmnn, G], index0 int [nmnn] rows of mat, mat [ncells, Gdist] -> [ncells, G]."""
    averaged = _f64(averaged); mat = _f64(mat); index0 = _i32(index0)
    nmnn, G = averaged.shape
    if index0.shape[0] != nmnn:
        raise _lib.B200Error(1, "'index' must have length equal to number of rows in 'averaged'")
    out = torch.empty((mat.shape[0], G), dtype=torch.float64, device=mat.device)
    _lib.call("b200mnn_dev_smooth_gaussian_kernel", _p(averaged), G, nmnn, _p(index0), _p(mat), mat.shape[1], mat.shape[0], float(sigma2),
              _p(out), _stream())
    _count(6)
    return out


def adjust_shift_variance(data1: torch.Tensor, data2: torch.Tensor, vect: torch.Tensor, sigma2: float, r1: torch.Tensor,
                          r2: torch.Tensor) -> torch.Tensor:
    """data1 [n1,G], data2 [n2,G], vect [n2,G]; r1/r2 0-based int -> scaling [n2]."""
    data1 = _f64(data1); data2 = _f64(data2); vect = _f64(vect); r1 = _i32(r1); r2 = _i32(r2)
    if data1.shape[1] != data2.shape[1] or data1.shape[1] != vect.shape[1]:
        raise _lib.B200Error(1, "number of genes do not match up between matrices")
    if vect.shape[0] != data2.shape[0]:
        raise _lib.B200Error(1, "number of cells do not match up between matrices")
    out = torch.empty((data2.shape[0],), dtype=torch.float64, device=data2.device)
    _lib.call("b200mnn_dev_adjust_shift_variance", _p(data1), data1.shape[0], _p(data2), data2.shape[0], data1.shape[1], _p(vect),
              float(sigma2), _p(r1), r1.shape[0], _p(r2), r2.shape[0], _p(out), _stream())
    _count(3)
    return out


def smooth_gaussian_from_centroids(x: torch.Tensor, centers: torch.Tensor, delta: torch.Tensor, sigma: float) -> torch.Tensor:
    """x [n,d] + soft-max-weighted centroid corrections (R/clusterMNN.R:289-312): centers, delta [nc,d]."""
    x = _f64(x); centers = _f64(centers); delta = _f64(delta)
    out = torch.empty_like(x)
    _lib.call("b200mnn_dev_smooth_gaussian_from_centroids", _p(x), x.shape[0], x.shape[1], _p(centers), _p(delta), centers.shape[0], float(sigma),
              _p(out), _stream())
    _count(1)
    return out
