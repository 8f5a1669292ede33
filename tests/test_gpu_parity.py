"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI (host-buffer entry points via
batchelor_b200.api, device-pointer entry points via batchelor_b200.device) and is compared with the CPU oracle on the
same seeded inputs, and with the committed golden vectors produced by the reference's own object code.

Bars: neighbour indices, distances' order and MNN pair lists are BIT-EXACT (0 mismatching slots, ties by index);
floating-point outputs agree within 1e-5 relative (north_star), most far tighter.
"""
import numpy as np
import pytest

import batchelor_b200 as bb
from batchelor_b200 import synth
from oracle import capi, host_oracle as ho

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance for floating-point outputs


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ---------------------------------------------------------------------------------------------------------------
# a1: exact kNN -- bit-exact indices and distances
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nq,d,k", [
    (1000, 700, 50, 20),      # the configured shape, small
    (5000, 3000, 50, 20),
    (300, 1000, 50, 20),      # fewer references than one tile
    (4097, 129, 50, 1),       # k = 1, ragged tiles
    (2000, 500, 50, 30),      # k > 24 -> 64-entry candidate lists
    (2000, 500, 50, 56),
    (2000, 300, 50, 60),      # k beyond the tensor path -> generic exact path
    (1500, 400, 1, 5), (1500, 400, 2, 5), (1500, 400, 5, 7), (1500, 400, 6, 7), (1500, 400, 16, 7), (1500, 400, 17, 7),
    (1500, 400, 32, 10), (1500, 400, 64, 10), (1500, 400, 100, 10), (800, 200, 192, 10),
    (800, 200, 250, 10),      # d beyond the resident-operand kernel -> generic exact path
    (40, 100, 50, 40),        # k = n
    # wide rows / large k: K-streamed tensor-core scoring + radix selection (csrc/knn_wide.cu)
    (3000, 1000, 512, 20), (2500, 700, 2000, 20), (3000, 500, 50, 100), (3000, 300, 50, 1000), (1300, 300, 333, 57),
    (700, 200, 6000, 10),     # a row wider than the rescue kernel's former shared-memory staging
    (3000, 64, 30, 1400),     # k beyond the candidate capacity -> generic exact scan
])
def test_query_knn_matches_oracle_bit_exact(n, nq, d, k):
    X, Q = synth.pc_batches(2, [n, nq], d=d, ncomp=8)
    got = bb.queryKNN(X, Q, k)
    idx, dist = capi.query_knn(X, Q, k)
    assert np.array_equal(got["index"], idx), f"{int((got['index'] != idx).sum())} mismatching neighbour slots"
    assert np.array_equal(got["distance"], dist)  # same arithmetic, same summation order -> identical doubles


def test_query_knn_column_major_inputs_like_r():
    X, Q = synth.pc_batches(2, [1200, 900], d=50, ncomp=8)
    got = bb.queryKNN(np.asfortranarray(X), np.asfortranarray(Q), 20)
    idx, dist = capi.query_knn(X, Q, 20)
    assert got["index"].flags.f_contiguous and np.array_equal(got["index"], idx) and np.array_equal(got["distance"], dist)


def test_query_knn_not_fp32_representable_inputs():
    rng = np.random.default_rng(3)
    X = rng.normal(size=(3000, 50)) * 7.3 + 100.0   # arbitrary doubles, far from the origin (hard for the split operands)
    Q = rng.normal(size=(1000, 50)) * 7.3 + 100.0
    got = bb.queryKNN(X, Q, 20)
    idx, dist = capi.query_knn(X, Q, 20)
    assert np.array_equal(got["index"], idx) and np.array_equal(got["distance"], dist)


def test_query_knn_tiny_and_huge_scales():
    X, Q = synth.pc_batches(2, [2000, 600], d=50, ncomp=8)
    for s in (1e-6, 3e4):
        got = bb.queryKNN(X * s, Q * s, 20)
        idx, dist = capi.query_knn(X * s, Q * s, 20)
        assert np.array_equal(got["index"], idx) and np.array_equal(got["distance"], dist)


def test_query_knn_exact_ties_broken_by_index():
    g = np.stack(np.meshgrid(np.arange(12.0), np.arange(12.0), indexing="ij"), -1).reshape(-1, 2)
    X = np.vstack([g, g, g])        # every point three times: ties everywhere, incl. > 32 equidistant points
    got = bb.queryKNN(X, g, 9)
    idx, dist = capi.query_knn(X, g, 9)
    assert np.array_equal(got["index"], idx) and np.array_equal(got["distance"], dist)
    X = np.zeros((500, 50)); X[:, 0] = np.repeat(np.arange(10.0), 50)   # 50 exact duplicates per location
    Q = X[::7] + 0.0
    got = bb.queryKNN(X, Q, 20)
    idx, dist = capi.query_knn(X, Q, 20)
    assert np.array_equal(got["index"], idx) and np.array_equal(got["distance"], dist)


def test_query_knn_k_capped_with_warning():
    X, Q = synth.pc_batches(2, [10, 30], d=50, ncomp=2)
    with pytest.warns(UserWarning, match="capped"):
        got = bb.queryKNN(X, Q, 20)
    idx, _ = capi.query_knn(X, Q, 10)
    assert got["index"].shape == (30, 10) and np.array_equal(got["index"], idx)


def test_query_knn_medium_against_kmknn_port():
    """100k x 100k: brute-force oracle is too slow; the KMKNN port (itself pinned to brute force in test_oracle.py) checks."""
    X, Q = synth.pc_batches(2, [100_000, 100_000], d=50)
    got = bb.queryKNN(X, Q, 20)
    idx, dist = capi.Kmknn(X).query(Q, 20)
    assert int((got["index"] != idx).sum()) == 0
    assert np.array_equal(got["distance"], dist)


# Pruned search (csrc/knn_cluster.cu): forced on small inputs with few clusters; the answer must not change at all.
@pytest.mark.parametrize("n,nq,d,k,ncomp,nclus", [
    (6000, 5000, 50, 20, 8, 16),
    (4097, 129, 50, 1, 8, 16),       # ragged tiles, k = 1
    (3000, 3000, 50, 30, 8, 16),     # 64-entry candidate lists
    (5000, 2000, 17, 7, 4, 32),      # more clusters than mixture components
    (3000, 1000, 100, 10, 8, 16),
    (2500, 4000, 3, 5, 2, 64),       # low dimension: bounds bite inside components too
    (700, 300, 50, 20, 8, 16),       # barely enough rows for the clustering
])
def test_query_knn_pruned_matches_oracle_bit_exact(n, nq, d, k, ncomp, nclus, monkeypatch):
    monkeypatch.setenv("B200MNN_PRUNE", "1")
    monkeypatch.setenv("B200MNN_CLUSTERS", str(nclus))
    X, Q = synth.pc_batches(2, [n, nq], d=d, ncomp=ncomp)
    got = bb.queryKNN(X, Q, k)
    idx, dist = capi.query_knn(X, Q, k)
    assert np.array_equal(got["index"], idx), f"{int((got['index'] != idx).sum())} mismatching neighbour slots"
    assert np.array_equal(got["distance"], dist)


def test_query_knn_pruned_degenerate_inputs(monkeypatch):
    """Duplicates (coincident centroids), exact ties, unclustered data and far-apart blobs with the pruning forced on."""
    monkeypatch.setenv("B200MNN_PRUNE", "1")
    monkeypatch.setenv("B200MNN_CLUSTERS", "16")
    rng = np.random.default_rng(11)
    cases = []
    X = np.zeros((600, 50)); X[:, 0] = np.repeat(np.arange(12.0), 50)          # 50 exact duplicates per location
    cases.append((X, X[::7] + 0.0, 20))
    cases.append((np.ones((400, 7)), np.ones((150, 7)), 9))                     # every point identical
    cases.append((rng.normal(size=(3000, 50)), rng.normal(size=(1000, 50)), 20))  # one blob: nothing can be skipped
    far = np.concatenate([rng.normal(size=(1500, 10)), rng.normal(size=(1500, 10)) + 1e3])
    cases.append((far, np.concatenate([far[::5] + 0.25, rng.normal(size=(100, 10)) + 500.0]), 20))   # queries between blobs
    cases.append((rng.normal(size=(2000, 50)) * 7.3 + 100.0, rng.normal(size=(700, 50)) * 7.3 + 100.0, 20))
    for X, Q, k in cases:
        got = bb.queryKNN(X, Q, k)
        idx, dist = capi.query_knn(X, Q, k)
        assert np.array_equal(got["index"], idx) and np.array_equal(got["distance"], dist)


def test_query_knn_pruned_skips_tiles_and_stays_exact(monkeypatch):
    """On clustered data most score tiles must be skipped (stats), with the answer still equal to the exact search."""
    import torch
    from batchelor_b200 import device as dev

    monkeypatch.setenv("B200MNN_PRUNE", "1")
    monkeypatch.setenv("B200MNN_CLUSTERS", "32")
    X, Q = synth.pc_batches(2, [60_000, 30_000], d=50, ncomp=16)
    Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    idx, dist = dev.query_knn(Xd, Qd, 20, stats=stats)
    s = stats.cpu().numpy()
    print(f"stats {s}")
    assert s[2] == 2 and s[0] == 0
    assert 0 < s[4] < 0.4 * s[6], f"pruning skipped too little: {s[4]} of {s[6]} tiles scored"
    want_idx, want_dist = capi.Kmknn(X).query(Q, 20)
    assert int((idx.cpu().numpy() + 1 != want_idx).sum()) == 0
    assert np.array_equal(dist.cpu().numpy(), want_dist)


@pytest.mark.parametrize("n,nq,d,k,ncomp,nclus", [
    (40_000, 25_000, 32, 30, 12, 32),     # 64-entry lists, two operand boxes in the first tier
    (50_000, 20_000, 64, 50, 6, 16),
    (30_000, 30_000, 20, 5, 40, 64),      # many small components
    (20_000, 70_000, 50, 20, 3, 128),     # far more clusters than components
])
def test_query_knn_pruned_medium_against_kmknn_port(n, nq, d, k, ncomp, nclus, monkeypatch):
    monkeypatch.setenv("B200MNN_PRUNE", "1")
    monkeypatch.setenv("B200MNN_CLUSTERS", str(nclus))
    X, Q = synth.pc_batches(2, [n, nq], d=d, ncomp=ncomp)
    got = bb.queryKNN(X, Q, k)
    idx, dist = capi.Kmknn(X).query(Q, k)
    assert int((got["index"] != idx).sum()) == 0
    assert np.array_equal(got["distance"], dist)


def test_query_knn_many_exact_duplicates_go_through_the_bounded_rescue():
    """40 copies of every location: every query has > 32 equidistant candidates at its k-th distance, so no scoring tier
    can certify it; the exact rescue (one bounded pass per query) must still return the (distance, index) order."""
    import torch
    from batchelor_b200 import device as dev

    base, _ = synth.pc_batches(2, [2_000, 10], d=50, ncomp=8)
    X = np.repeat(base, 40, axis=0)                      # 80,000 references: pruned path
    Q = base[::2] + 0.0                                  # 1,000 queries sitting on duplicated locations
    Q = np.concatenate([Q] * 17)                         # >= 16,384 queries
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    idx, dist = dev.query_knn(torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda(), 20, stats=stats)
    s = stats.cpu().numpy()
    assert s[0] > 0, "expected uncertifiable queries"
    want_idx, want_dist = capi.Kmknn(X).query(Q, 20)
    assert int((idx.cpu().numpy() + 1 != want_idx).sum()) == 0
    assert np.array_equal(dist.cpu().numpy(), want_dist)


@pytest.mark.parametrize("serial", [False, True])
def test_query_knn_a_handful_of_rescued_queries(monkeypatch, serial):
    """A few queries sitting on a 40-fold duplicated location in otherwise ordinary data: they alone need the exact rescue,
    which answers the first flagged queries with a scan sliced over many blocks (one block per query would stream every
    reference on its own); B200MNN_RESCUE_SERIAL keeps the one-block-per-query kernel.  Same exact answer either way."""
    import torch
    from batchelor_b200 import device as dev

    if serial:
        monkeypatch.setenv("B200MNN_RESCUE_SERIAL", "1")
    X, Q = synth.pc_batches(2, [70_000, 17_000], d=50, ncomp=8)
    X[100:140] = X[100]
    Q[:5] = X[100]
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    idx, dist = dev.query_knn(torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda(), 20, stats=stats)
    assert 5 <= int(stats[0]) < 200, "expected a handful of uncertifiable queries"
    want_idx, want_dist = capi.Kmknn(X).query(Q, 20)
    assert int((idx.cpu().numpy() + 1 != want_idx).sum()) == 0
    assert np.array_equal(dist.cpu().numpy(), want_dist)


def test_query_knn_dense_path_still_exact_at_medium_size(monkeypatch):
    monkeypatch.setenv("B200MNN_PRUNE", "0")
    X, Q = synth.pc_batches(2, [70_000, 20_000], d=50)
    got = bb.queryKNN(X, Q, 20)
    idx, dist = capi.Kmknn(X).query(Q, 20)
    assert int((got["index"] != idx).sum()) == 0 and np.array_equal(got["distance"], dist)


def test_candidate_scoring_error_is_far_inside_the_certificate_bound():
    """The tensor-core scores must approximate the exact squared distances much better than the eps the certificate
    assumes (2^-16 (|q| M + M^2)); otherwise the rescue path would carry the correctness."""
    import torch
    from batchelor_b200 import device as dev

    X, Q = synth.pc_batches(2, [20_000, 4_000], d=50)
    Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
    cidx, cd2, thr = dev.debug_candidates(Xd, Qd, 20)
    cidx, cd2 = cidx.cpu().numpy(), cd2.cpu().numpy()
    valid = cidx >= 0
    exact = ((Q[:, None, :] - X[np.where(valid, cidx, 0)]) ** 2).sum(-1)
    M2 = (X ** 2).sum(1).max(); qn = (Q ** 2).sum(1)
    eps = 2.0 ** -16 * (np.sqrt(qn * M2) + M2)
    err = np.abs(cd2 - exact)[valid]
    ratio = (np.abs(cd2 - exact) / eps[:, None])[valid].max()
    print(f"max |approx-exact| = {err.max():.3e}, max err/eps = {ratio:.3f}")
    assert ratio < 0.25
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    dev.query_knn(Xd, Qd, 20, stats=stats)
    s = stats.cpu().numpy()
    assert s[2] >= 1 and s[0] == 0, f"tensor path not taken or queries rescued: {s}"


# ---------------------------------------------------------------------------------------------------------------
# a2 / a3: mutual pairs -- identical lists in the reference's order
# ---------------------------------------------------------------------------------------------------------------
def test_find_mutual_nns_matches_golden(golden):
    g = golden["find_mutual_nns"]
    first, second = bb.find_mutual_nns(g["w21"], g["w12"])
    assert np.array_equal(first, g["first"]) and np.array_equal(second, g["second"])
    res = bb.findMutualNN(g["X1"], g["X2"], k1=10, k2=15)
    assert np.array_equal(res["first"], g["first"]) and np.array_equal(res["second"], g["second"])


@pytest.mark.parametrize("n1,n2,k1,k2", [(3000, 2500, 20, 20), (500, 4000, 5, 30), (64, 64, 64, 64), (1000, 10, 20, 20)])
def test_find_mutual_nn_matches_oracle(n1, n2, k1, k2):
    b1, b2 = synth.pc_batches(2, [n1, n2], d=50, ncomp=8)
    res = bb.findMutualNN(b1, b2, k1=k1, k2=k2)
    f, s = capi.find_mutual_nn(b1, b2, min(k1, n1), min(k2, n2))
    assert np.array_equal(res["first"], f) and np.array_equal(res["second"], s)
    assert np.all(np.diff(res["first"]) >= 0)  # order contract: first ascending


@pytest.mark.parametrize("col_major", [0, 1])
@pytest.mark.parametrize("prune", ["0", "1"])
def test_find_mutual_nn_pipelined_upload(monkeypatch, col_major, prune):
    """Large inputs take the pipelined host entry: batch 2 is uploaded in row chunks and searched chunk by chunk against a
    cached reference side of batch 1 (knn::RefCache; scale from the reference rows alone).  Forced here on small inputs,
    in both host layouts, with and without the cluster plan; ragged last chunk; pairs identical to the oracle."""
    import ctypes as C
    from batchelor_b200 import _lib
    monkeypatch.setenv("B200MNN_PIPELINE_MIN_ROWS", "256")
    monkeypatch.setenv("B200MNN_PRUNE", prune)
    n1, n2, k1, k2 = 2300, 3001, 20, 12
    b1, b2 = synth.pc_batches(2, [n1, n2], d=50, ncomp=8)
    b2 = b2 * 1.7   # queries of the cached search somewhat larger than its references
    h1 = np.asfortranarray(b1) if col_major else np.ascontiguousarray(b1)
    h2 = np.asfortranarray(b2) if col_major else np.ascontiguousarray(b2)
    cap = n1 * k2
    first = np.zeros(cap, np.int32); second = np.zeros(cap, np.int32)
    npairs = C.c_int64(0)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    _lib.call("b200mnn_find_mutual_nn", fp(h1), n1, fp(h2), n2, 50, k1, k2, col_major, ip(first), ip(second), cap, C.byref(npairs))
    f, s = capi.find_mutual_nn(b1, b2, k1, k2)
    m = npairs.value
    assert m == len(f) and np.array_equal(first[:m], f) and np.array_equal(second[:m], s)


@pytest.mark.parametrize("epi", ["0", "1"])
@pytest.mark.parametrize("n,nq,d,k", [(70000, 20000, 50, 20), (5000, 3000, 17, 24), (3000, 1000, 120, 5), (70000, 20000, 50, 30),
                                      (4000, 2500, 33, 26)])
def test_query_knn_both_epilogues(monkeypatch, epi, n, nq, d, k):
    """The TS kernel's two epilogues (B200MNN_EPI=1: per-thread register lists, the default; 0: replace-the-maximum lists)
    give the same exact answer, with the cluster plan on (70000-row shapes) and off; k = 26 / 30 use lists of 64 candidates per
    row (32 per thread in the register-list epilogue)."""
    monkeypatch.setenv("B200MNN_EPI", epi)
    X, Q = synth.pc_batches(2, [n, nq], d=d, ncomp=6)
    res = bb.queryKNN(X, Q, k)
    want_i, want_d = capi.query_knn(X, Q, k)
    assert np.array_equal(res["index"], want_i) and np.array_equal(res["distance"], want_d)


def test_find_mutual_nns_empty_and_no_pairs():
    first, second = bb.find_mutual_nns(np.zeros((0, 3), np.int32), np.ones((4, 2), np.int32))
    assert first.size == 0 and second.size == 0
    left = np.array([[1], [1]], np.int32); right = np.array([[2]], np.int32)  # only cell 2 <-> cell 1 is mutual
    first, second = bb.find_mutual_nns(left, right)
    assert np.array_equal(first, [2]) and np.array_equal(second, [1])
    with pytest.raises(bb.B200Error, match="out of range"):
        bb.find_mutual_nns(np.array([[5]], np.int32), np.array([[1]], np.int32))


# ---------------------------------------------------------------------------------------------------------------
# a5: Gaussian smoothing; a7: shift variance -- golden vectors from the reference object code + oracle
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["vanilla", "repeated", "many", "wide"])
def test_smooth_gaussian_kernel_matches_golden(golden, case):
    g = golden["smooth_gaussian_kernel"]
    out = bb.smooth_gaussian_kernel(g[f"{case}_averaged"], g[f"{case}_index0"], g["data2"].T, float(g[f"{case}_sigma"]))
    assert _relerr(out, g[f"{case}_out"]) < 1e-10


def _smooth_path():
    import ctypes as C
    from batchelor_b200 import _lib
    path, e0, e1 = C.c_int(0), C.c_double(0), C.c_double(0)
    _lib.call("b200mnn_smooth_last_check", C.byref(path), C.byref(e0), C.byref(e1))
    return path.value, e0.value, e1.value


@pytest.mark.parametrize("case", ["vanilla", "repeated", "many", "wide"])
def test_smooth_gaussian_kernel_tensor_path_matches_golden(golden, case, monkeypatch):
    """The tcgen05 path (forced here; by default only large problems take it) against the reference object code: 1e-5."""
    monkeypatch.setenv("B200MNN_SMOOTH", "tensor")
    g = golden["smooth_gaussian_kernel"]
    out = bb.smooth_gaussian_kernel(g[f"{case}_averaged"], g[f"{case}_index0"], g["data2"].T, float(g[f"{case}_sigma"]))
    path, e0, e1 = _smooth_path()
    err = _relerr(out, g[f"{case}_out"])
    print(f"{case}: path {path}, rel err {err:.2e}, in-call sample check rows {e0:.2e} log-density {e1:.2e}")
    assert path == 1 and err < RTOL


@pytest.mark.parametrize("ncells,nmnn,Gd,G,sigma,normalise", [(6000, 1500, 300, 300, 0.1, True), (3000, 700, 2000, 2100, 0.1, True),
                                                              (2500, 900, 130, 64, 1.0, False)])
def test_smooth_gaussian_kernel_tensor_path_matches_oracle(ncells, nmnn, Gd, G, sigma, normalise, monkeypatch):
    """Mid-size shapes (several tiles, ragged edges, G != Gdist, K-chunked accumulation) against the fp64 C restatement."""
    monkeypatch.setenv("B200MNN_SMOOTH", "tensor")
    rng = np.random.default_rng(ncells + G)
    centres = rng.normal(size=(5, Gd))
    mat = (centres[rng.integers(0, 5, size=ncells)] * 0.15 + rng.normal(size=(ncells, Gd))).T
    if normalise:
        mat = mat / np.linalg.norm(mat, axis=0, keepdims=True)
    else:
        mat *= 0.1
    idx = np.sort(rng.permutation(ncells)[:nmnn]).astype(np.int32)
    avg = rng.normal(size=(G, nmnn)) * 0.05 + 0.02
    out = bb.smooth_gaussian_kernel(avg, idx, mat, sigma)
    path, e0, e1 = _smooth_path()
    ref = capi.smooth_gaussian_kernel(avg, idx, mat, sigma)
    err = _relerr(out, ref)
    print(f"path {path}, rel err {err:.2e}, in-call sample check rows {e0:.2e} log-density {e1:.2e}")
    assert path == 1 and err < RTOL


def test_smooth_gaussian_kernel_more_genes_out_than_in_and_errors():
    rng = np.random.default_rng(8)
    mat = rng.normal(scale=0.1, size=(30, 700)); avg = rng.normal(size=(90, 64)); idx = rng.permutation(700)[:64].astype(np.int32)
    out = bb.smooth_gaussian_kernel(avg, idx, mat, 0.1)   # 'averaged' may cover more genes than 'mat' (:23)
    assert _relerr(out, capi.smooth_gaussian_kernel(avg, idx, mat, 0.1)) < 1e-10
    assert np.all(np.isnan(bb.smooth_gaussian_kernel(np.zeros((5, 0)), np.zeros(0, np.int32), mat, 0.1)))
    with pytest.raises(bb.B200Error, match="out of range"):
        bb.smooth_gaussian_kernel(avg, idx + 700, mat, 0.1)


@pytest.mark.parametrize("mode", ["fast", "fast_simt", "exact", "cell"])
def test_adjust_shift_variance_matches_golden(golden, mode, monkeypatch):
    """Bit-equal to the reference object code (tests/golden/make_golden.py) in every mode: FMA-order bulk + certificate
    (default), reference-order tiles, and the reference's loop per cell.  Identical doubles == identical quantile picks."""
    monkeypatch.setenv("B200MNN_SHIFTVAR", mode)
    g = golden["adjust_shift_variance"]
    for s in (1.0, 0.1):
        out = bb.adjust_shift_variance(g["data1"], g["data2"], g["vect"], s, np.arange(400), np.arange(1000))
        ref = g[f"out_sigma_{s}"]
        assert np.array_equal(out, ref), f"sigma={s}: {int((out != ref).sum())} of {ref.size} cells differ, max rel err {_relerr(out, ref):.2e}"
    out = bb.adjust_shift_variance(g["data1"], g["data2"], g["vect"], 1.0, g["r1"], g["r2"])
    assert np.array_equal(out, g["out_restricted"])


@pytest.mark.parametrize("mode", ["fast", "fast_simt", "exact"])
@pytest.mark.parametrize("n1,n2,G,sigma", [(3000, 700, 37, 0.5), (5000, 300, 130, 1.0), (700, 300, 2100, 0.1)])
def test_adjust_shift_variance_matches_oracle_bit_exact(n1, n2, G, sigma, mode, monkeypatch):
    """Larger shapes than the goldens: several column splits, bins of > 1 024 cells (second selection level), ragged
    tiles, G beyond one staged chunk and beyond the old shared-memory limit -- against the C restatement (itself pinned
    bit-identical to the reference object code in tests/test_oracle.py)."""
    monkeypatch.setenv("B200MNN_SHIFTVAR", mode)
    rng = np.random.default_rng(n1 + G)
    scale = 1.0 / np.sqrt(G)
    d1 = rng.normal(scale=scale, size=(G, n1)); d2 = rng.normal(scale=scale, size=(G, n2)) + 0.3 * scale
    if n1 == 5000:   # a tight clump of 4 000 reference cells + a wide halo: the crossing bin overflows -> second selection level
        d1[:, :4000] = d1[:, [0]] + 1e-4 * scale * rng.normal(size=(G, 4000))
    cv = rng.normal(size=(n2, G))
    r1 = rng.permutation(n1)[: n1 - 7].astype(np.int32); r2 = rng.permutation(n2)[: n2 - 5].astype(np.int32)
    out = bb.adjust_shift_variance(d1, d2, cv, sigma, r1, r2)
    ref = capi.adjust_shift_variance(d1, d2, cv, sigma, r1, r2)
    assert np.array_equal(out, ref), f"{int((out != ref).sum())} of {ref.size} cells differ, max rel err {_relerr(out, ref):.2e}"


def test_adjust_shift_variance_restrict_identity_and_zero_vector():  # test-mnn-correct.R:162-173
    rng = np.random.default_rng(100032)
    d1 = rng.normal(scale=0.1, size=(25, 120)); d2 = rng.normal(scale=0.1, size=(25, 200)); cv = rng.uniform(size=(200, 25))
    i1 = np.arange(10, 21); i2 = np.arange(20, 9, -1)
    A1 = np.hstack([d1, d1[:, i1 - 1]]); A2 = np.hstack([d2, d2[:, i2 - 1]])
    a = bb.adjust_shift_variance(d1, d2, cv, 1.0, np.arange(120), np.arange(200))
    b = bb.adjust_shift_variance(A1, A2, np.vstack([cv, cv[i2 - 1]]), 1.0, np.arange(120), np.arange(200))
    assert np.array_equal(a, b[:200]) and np.array_equal(a[i2 - 1], b[200:])   # expect_identical in the reference
    cv0 = cv.copy(); cv0[3] = 0.0
    out = bb.adjust_shift_variance(d1, d2, cv0, 1.0, np.arange(120), np.arange(200))
    ref = capi.adjust_shift_variance(d1, d2, cv0, 1.0, np.arange(120), np.arange(200))
    assert not np.isfinite(out[3]) and not np.isfinite(ref[3])   # 0/0 or x/0 kept (:64, :160)


# ---------------------------------------------------------------------------------------------------------------
# a4 / a6 / a8 / a9 through the host-buffer ABI
# ---------------------------------------------------------------------------------------------------------------
def _call_average(t1, m1, t2, m2):
    import ctypes as C
    from batchelor_b200 import _lib
    t1f, t2f = np.asfortranarray(t1), np.asfortranarray(t2)
    m1 = np.ascontiguousarray(m1, np.int32); m2 = np.ascontiguousarray(m2, np.int32)
    cap = max(1, min(m1.size, t2.shape[0]))
    avg = np.zeros(cap * t1.shape[1]); sec = np.zeros(cap, np.int32); n = C.c_int64(0)
    _lib.call("b200mnn_average_correction", t1f.ctypes.data_as(_lib.f64p), t1.shape[0], t2f.ctypes.data_as(_lib.f64p), t2.shape[0],
              t1.shape[1], m1.ctypes.data_as(_lib.i32p), m2.ctypes.data_as(_lib.i32p), m1.size, avg.ctypes.data_as(_lib.f64p),
              sec.ctypes.data_as(_lib.i32p), C.byref(n))
    return avg[: n.value * t1.shape[1]].reshape((n.value, t1.shape[1]), order="F"), sec[: n.value]


def test_average_correction_abi():  # test-fast-mnn.R:6-32
    rng = np.random.default_rng(1200001)
    t1 = rng.normal(size=(100, 10)); t2 = rng.normal(size=(200, 10))
    m1 = rng.integers(1, 101, size=250); m2 = rng.integers(1, 101, size=250)
    avg, sec = _call_average(t1, m1, t2, m2)
    ravg, rsec = ho.average_correction(t1, m1, t2, m2)
    assert np.array_equal(sec, rsec) and _relerr(avg, ravg) < 1e-12
    avg, sec = _call_average(t1, np.zeros(0, np.int32), t2, np.zeros(0, np.int32))
    assert avg.shape == (0, 10) and sec.size == 0


def test_center_and_tricube_and_cosine_abi():
    import ctypes as C
    from batchelor_b200 import _lib
    rng = np.random.default_rng(1200002)
    t = np.asfortranarray(rng.normal(size=(100, 10))); b = rng.normal(size=10)
    out = np.zeros((100, 10), order="F")
    _lib.call("b200mnn_center_along_batch_vector", t.ctypes.data_as(_lib.f64p), 100, 10, b.ctypes.data_as(_lib.f64p), None, 0, out.ctypes.data_as(_lib.f64p))
    assert _relerr(out, ho.center_along_batch_vector(t, b)) < 1e-12 and np.std(out @ b) < 1e-8
    t2 = np.asfortranarray(np.vstack([t, t[:10]])); keep = np.arange(1, 101, dtype=np.int32)
    out2 = np.zeros((110, 10), order="F")
    _lib.call("b200mnn_center_along_batch_vector", t2.ctypes.data_as(_lib.f64p), 110, 10, b.ctypes.data_as(_lib.f64p), keep.ctypes.data_as(_lib.i32p), 100, out2.ctypes.data_as(_lib.f64p))
    assert np.array_equal(out, out2[:100]) and np.array_equal(out2[:10], out2[100:])   # expect_identical (test-fast-mnn.R:43-50)

    corr = np.asfortranarray(rng.normal(size=(50, 10))); involved = (rng.permutation(100)[:50] + 1).astype(np.int32)
    for k, nd in [(20, 3.0), (11, 3.0), (11, 1.0), (80, 3.0)]:
        o = np.zeros((100, 10), order="F")
        _lib.call("b200mnn_tricube_weighted_correction", t.ctypes.data_as(_lib.f64p), 100, 10, corr.ctypes.data_as(_lib.f64p),
                  involved.ctypes.data_as(_lib.i32p), 50, k, nd, o.ctypes.data_as(_lib.f64p))
        assert _relerr(o, ho.tricube_weighted_correction(t, corr, involved, k=k, ndist=nd)) < 1e-12

    X = rng.normal(size=(20, 30)); X[:, 4] = 0
    res = bb.cosineNorm(X, mode="all")
    m, l2 = capi.cosine_norm(X)
    assert _relerr(res["matrix"], m) < 1e-14 and _relerr(res["l2norm"], l2) < 1e-14 and np.all(res["matrix"][:, 4] == 0)
    assert np.allclose(bb.cosineNorm(X, mode="l2norm"), l2)
    assert np.allclose(bb.cosineNorm(X, subset_row=np.arange(1, 11)), capi.cosine_norm(X[:10])[0])


# ---------------------------------------------------------------------------------------------------------------
# whole merges: reducedMNN / mnnCorrect against the oracle's restatement of the R loops
# ---------------------------------------------------------------------------------------------------------------
def test_reduced_mnn_toy_known_answers():  # tests/testthat/test-reduced-mnn.R:80-105
    core = np.stack([np.repeat(np.arange(1.0, 11), 10), np.tile(np.arange(1.0, 11), 10)], 1)
    b1 = core.copy(); b1[:, 0] += 20
    b2 = core.copy(); b2[:, 1] += 20
    o1 = bb.reducedMNN(core, b1, k=1).corrected
    assert np.allclose(o1[:, 0], 5.5) and np.allclose(o1[:, 1], np.r_[core[:, 1], b1[:, 1]])
    assert np.allclose(bb.reducedMNN(core, b1, b2, k=1).corrected, 5.5)
    oy = bb.reducedMNN(core + 10, b2 + 10, k=1).corrected
    assert np.allclose(oy[:, 0], np.r_[core[:, 0], b2[:, 0]] + 10) and np.allclose(oy[:, 1], 15.5)
    oz = bb.reducedMNN(core, b1, core + 10, b2 + 10, merge_order=[[1, 2], [3, 4]], k=1).corrected
    assert np.allclose(oz, 5.5)


def test_reduced_mnn_matches_oracle_three_batches_and_hierarchy():
    bs = synth.pc_batches(4, [1500, 1200, 900, 1100], d=50, ncomp=8)
    for order in (None, [3, 1, 2, 4], [[1, 2], [3, 4]]):
        got = bb.reducedMNN(*bs, k=20, merge_order=order)
        ref = ho.reduced_mnn(bs, k=20, merge_order=order)
        assert np.array_equal(got.batch, ref["batch"])
        assert got.merge_info["left"] == ref["merge_info"]["left"] and got.merge_info["right"] == ref["merge_info"]["right"]
        # first merge: identical inputs -> identical pairs; later merges see 1e-16-level input differences
        assert np.array_equal(got.merge_info["pairs"][0]["left"], ref["merge_info"]["pairs"][0][0])
        assert np.array_equal(got.merge_info["pairs"][0]["right"], ref["merge_info"]["pairs"][0][1])
        assert _relerr(got.corrected, ref["corrected"]) < RTOL
        assert np.allclose(got.merge_info["batch_size"], ref["merge_info"]["batch_size"], rtol=1e-6)
        assert np.allclose(got.merge_info["lost_var"], ref["merge_info"]["lost_var"], rtol=1e-5, atol=1e-8)


def test_reduced_mnn_restrict_propk_batch_and_skip():
    rng = np.random.default_rng(12000053)
    B1 = rng.normal(0, size=(300, 10)); B2 = rng.normal(1, size=(400, 10)); B3 = rng.normal(2, size=(200, 10))
    ref = bb.reducedMNN(B1, B2, B3)
    i1 = np.arange(100, 49, -1); i2 = np.arange(1, 21); i3 = np.arange(50, 101)
    C1 = np.vstack([B1, B1[i1 - 1]]); C2 = np.vstack([B2, B2[i2 - 1]]); C3 = np.vstack([B3, B3[i3 - 1]])
    out = bb.reducedMNN(C1, C2, C3, restrict=[np.arange(1, 301), np.arange(1, 401), np.arange(1, 201)])
    for b, keep, dup in [(1, 300, i1), (2, 400, i2), (3, 200, i3)]:   # expect_identical in test-reduced-mnn.R:127-135
        r = ref.corrected[ref.batch == b]; o = out.corrected[out.batch == b]
        assert np.array_equal(r, o[:keep]) and np.array_equal(r[dup - 1], o[keep:])
    p1 = bb.reducedMNN(B1, B1 + 1.0, k=10, prop_k=20 / 300); p2 = bb.reducedMNN(B1, B1 + 1.0, k=20)
    assert np.array_equal(p1.corrected, p2.corrected)   # test-reduced-mnn.R:39-58
    oracle = ho.reduced_mnn([B1, B2, B3])
    assert _relerr(ref.corrected, oracle["corrected"]) < RTOL
    # single matrix + batch factor (R/reducedMNN.R:81-88)
    allc = np.vstack([B1, B2, B3]); lab = np.r_[np.full(300, 1), np.full(400, 2), np.full(200, 3)]
    sh = rng.permutation(900)
    one = bb.reducedMNN(allc[sh], batch=lab[sh])
    assert _relerr(one.corrected, ref.corrected[sh]) < 1e-12 and np.array_equal(one.batch, lab[sh])
    # min.batch.skip (test-fast-mnn.R:409-457): no batch effect -> skipped, coordinates unchanged
    nob = rng.normal(0, size=(350, 10))
    same = bb.reducedMNN(B1, nob, min_batch_skip=0.5)
    want = ho.reduced_mnn([B1, nob], min_batch_skip=0.5)
    assert same.merge_info["skipped"][0] and want["merge_info"]["skipped"][0]
    assert np.isclose(same.merge_info["batch_size"][0], want["merge_info"]["batch_size"][0], rtol=1e-9) and same.merge_info["batch_size"][0] < 0.5
    assert np.array_equal(same.corrected[:300], B1) and np.array_equal(same.corrected[300:], nob)
    assert bb.reducedMNN(B1, B2, min_batch_skip=0.5).merge_info["batch_size"][0] > 0.5   # a real offset is not skipped


def test_mnn_correct_matches_oracle():
    A, B, Cc = synth.gene_batches(3, [260, 300, 220], G=120, latent=6, ncomp=4)
    # sigma is matched to the data scale: on raw (un-normalised) data squared distances are ~2*G, and with the default
    # sigma = 0.1 every Gaussian weight is one-hot, which makes the reference's discrete quantile pick
    # (src/adjust_shift_variance.cpp:145-156) a coin flip on last-bit rounding -- the reference's own tests skip
    # platforms over exactly this (tests/testthat/test-mnn-correct.R:140-141, :396-399).
    for kw in (dict(), dict(var_adj=False), dict(cos_norm_in=False, cos_norm_out=False, sigma=60.0),
               dict(cos_norm_in=False, cos_norm_out=False, var_adj=False), dict(cos_norm_out=False, var_adj=False),
               dict(cos_norm_out=False, sigma=60.0),
               dict(merge_order=[3, 1, 2]), dict(sigma=1.0)):
        got = bb.mnnCorrect(A, B, Cc, k=15, **kw)
        ref = ho.mnn_correct([A, B, Cc], k=15, **kw)
        assert np.array_equal(got.batch, ref["batch"])
        assert np.array_equal(got.merge_info["pairs"][0]["left"], ref["merge_info"]["pairs"][0][0])
        assert np.array_equal(got.merge_info["pairs"][0]["right"], ref["merge_info"]["pairs"][0][1])
        err = _relerr(got.corrected, ref["corrected"])
        print(kw, "rel err", err)
        assert err < RTOL


def test_mnn_correct_restrict_identity():  # test-mnn-correct.R:379-441
    rng = np.random.default_rng(10004)
    A = rng.normal(size=(15, 60)); B = rng.normal(size=(15, 80)) + 1
    ref = bb.mnnCorrect(A, B, k=10)
    i1 = np.arange(5, 16); i2 = np.arange(20, 9, -1)
    out = bb.mnnCorrect(np.hstack([A, A[:, i1 - 1]]), np.hstack([B, B[:, i2 - 1]]), k=10, restrict=[np.arange(1, 61), np.arange(1, 81)])
    r, o = ref.corrected, out.corrected
    assert np.array_equal(r[:, :60], o[:, :60]) and np.array_equal(r[:, 60:], o[:, 71:151]) and np.array_equal(r[:, 60:][:, i2 - 1], o[:, 151:])


def test_results_are_deterministic_run_to_run():
    b1, b2 = synth.pc_batches(2, [4000, 3500], d=50, ncomp=8)
    a = bb.reducedMNN(b1, b2, k=20); b = bb.reducedMNN(b1, b2, k=20)
    assert np.array_equal(a.corrected, b.corrected)
    assert np.array_equal(a.merge_info["pairs"][0]["left"], b.merge_info["pairs"][0]["left"])


# ---------------------------------------------------------------------------------------------------------------
# split-fp16 tcgen05 GEMM (engine of the gene-space kernels) against fp64 matmul
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,terms,cb", [(128, 256, 64, 3, 0), (300, 700, 200, 3, 2), (129, 257, 1000, 3, 8), (1500, 3000, 2000, 3, 8),
                                           (300, 700, 200, 1, 0), (2100, 1030, 130, 1, 1)])
def test_split_gemm_matches_fp64(M, N, K, terms, cb):
    import ctypes as C
    import torch
    from batchelor_b200 import _lib
    rng = np.random.default_rng(M + N + K)
    A = rng.normal(size=(M, K)) * 3.0 + 0.5; B = rng.normal(size=(N, K)) * 0.02
    dA = torch.from_numpy(A).cuda(); dB = torch.from_numpy(B).cuda()
    ldo = (N + 255) // 256 * 256
    out = torch.full((M, ldo), float("nan"), dtype=torch.float32, device="cuda")
    _lib.call("b200mnn_dev_debug_gemm", C.c_void_p(dA.data_ptr()), M, C.c_void_p(dB.data_ptr()), N, K, terms, cb, C.c_void_p(out.data_ptr()), ldo,
              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    got = out[:, :N].double().cpu().numpy()
    ref = A @ B.T
    scale = np.linalg.norm(A, axis=1)[:, None] * np.linalg.norm(B, axis=1)[None, :]
    err = np.max(np.abs(got - ref) / scale)
    print(f"max |err| / (|a||b|) = {err:.3e}")
    assert err < (2e-6 if terms == 3 else 2e-3)


def test_query_knn_wide_path_ties_and_gene_space():
    """Wide path: integer grid (massive exact ties across the selection threshold -> rescue) and cosine-normalised gene-space
    data as mnnCorrect searches it (R/mnnCorrect.R:288-289)."""
    rng = np.random.default_rng(5)
    X = rng.integers(0, 2, size=(1500, 300)).astype(np.float64); Q = rng.integers(0, 2, size=(400, 300)).astype(np.float64)
    X[100:140] = X[7]   # 41 identical references: more ties than any candidate margin absorbs
    got = bb.queryKNN(X, Q, 20)
    idx, dist = capi.query_knn(X, Q, 20)
    assert np.array_equal(got["index"], idx) and np.array_equal(got["distance"], dist)
    A, B = synth.gene_batches(2, [2600, 2200], G=2000)
    A = A / np.linalg.norm(A, axis=0); B = B / np.linalg.norm(B, axis=0)
    got = bb.findMutualNN(A.T, B.T, k1=20, k2=20)
    f, s = capi.find_mutual_nn(np.ascontiguousarray(A.T), np.ascontiguousarray(B.T), 20, 20)
    assert np.array_equal(got["first"], f) and np.array_equal(got["second"], s)


def test_reduced_mnn_single_c_abi_call_equals_python_driven_loop(monkeypatch):
    """b200mnn_reduced_mnn (the device-resident merge loop as one C-ABI call) against the Python-driven loop
    over the per-step entry points and against the numpy oracle: 4 batches, hierarchical merge order, restrict, prop.k."""
    rng = np.random.default_rng(77)
    Bs = [rng.normal(size=(n, 12)) + 0.8 * i for i, n in enumerate([500, 420, 380, 610])]
    restrict = [np.arange(1, 451), None, np.arange(20, 381), None]
    kw = dict(k=15, merge_order=[[1, 2], [4, 3]], restrict=restrict, prop_k=0.05)
    one = bb.reducedMNN(*Bs, **kw)
    monkeypatch.setenv("B200MNN_PYLOOP", "1")
    two = bb.reducedMNN(*Bs, **kw)
    assert np.array_equal(one.batch, two.batch)
    assert _relerr(one.corrected, two.corrected) < 1e-12
    for a, b in zip(one.merge_info["pairs"], two.merge_info["pairs"]):
        assert np.array_equal(a["left"], b["left"]) and np.array_equal(a["right"], b["right"])
    assert np.allclose(one.merge_info["lost_var"], two.merge_info["lost_var"], rtol=1e-9, atol=1e-12)
    assert one.merge_info["left"] == two.merge_info["left"] and one.merge_info["right"] == two.merge_info["right"]
    ref = ho.reduced_mnn(Bs, **kw)
    assert _relerr(one.corrected, ref["corrected"]) < RTOL
    assert np.allclose(one.merge_info["lost_var"], ref["merge_info"]["lost_var"], rtol=1e-5, atol=1e-8)


def test_reduced_mnn_auto_merge_matches_oracle():
    """auto.merge = TRUE (R/MNN_tree.R:154-226) searched on the device inside b200mnn_reduced_mnn: same merge order, same
    pairs, same corrected values as the numpy restatement; 4 batches so that a merged node is re-counted with orthogonalisation."""
    rng = np.random.default_rng(9)
    A = rng.normal(size=(400, 8)); B = rng.normal(size=(350, 8)) + 0.4; Cc = rng.normal(size=(300, 8)); Cc[:, 0] += 5.0
    D = rng.normal(size=(320, 8)); D[:, 0] += 5.3
    got = bb.reducedMNN(A, Cc, B, D, k=12, auto_merge=True)
    ref = ho.reduced_mnn([A, Cc, B, D], k=12, auto_merge=True)
    assert got.merge_info["left"] == ref["merge_info"]["left"] and got.merge_info["right"] == ref["merge_info"]["right"]
    for a, b in zip(got.merge_info["pairs"], ref["merge_info"]["pairs"]):
        assert np.array_equal(a["left"], b[0]) and np.array_equal(a["right"], b[1])
    assert np.array_equal(got.batch, ref["batch"]) and _relerr(got.corrected, ref["corrected"]) < RTOL


def test_sharded_search_under_nccl():
    """Hardware parity at N > 1: two ranks over NCCL (query rows sharded, reference replicated, all-gather of the top-k,
    sharded upload) must reproduce the single-GPU result bit for bit.  Needs two visible GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the multi-GPU tier and tools/ run it)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(root, "tests", "_dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_propagate_to_cells_matches_oracle():
    """clusterMNN's propagation (R/clusterMNN.R:267-312): sigma from the exact k = 1 search, Gaussian-weighted centroid deltas."""
    rng = np.random.default_rng(21)
    cen = rng.normal(scale=3.0, size=(37, 20))
    cells = cen[rng.integers(0, 37, size=5000)] + rng.normal(size=(5000, 20))
    corrected = cen + rng.normal(scale=0.5, size=cen.shape)
    restrict = np.arange(1, 4001)
    got = bb.propagate_to_cells(cells, cen, corrected, restrict=restrict)
    ref = ho.propagate_to_cells(cells, cen, corrected, restrict=restrict)
    assert _relerr(got, ref) < 1e-12


def test_multi_batch_pca_matches_exact_svd():
    """multiBatchPCA (R/multiBatchPCA.R:211-322) against numpy's exact SVD on data with a clear spectrum: rotation, PCs,
    centres and variance explained; unequal batch sizes exercise the per-batch scaling and the weights."""
    rng = np.random.default_rng(4)
    load = rng.normal(size=(8, 300)) * np.linspace(6.0, 1.5, 8)[:, None]
    mats = [(rng.normal(size=(n, 8)) @ load + 0.3 * rng.normal(size=(n, 300)) + off).T for n, off in [(400, 0.0), (250, 0.5), (900, -0.2)]]
    for weights in (None, [1.0, 2.0, 0.5], False):
        pcs, meta = bb.multiBatchPCA(*mats, d=6, weights=weights, get_variance=True)
        ref = ho.multi_batch_pca(mats, d=6, weights=weights, get_variance=True)
        assert _relerr(meta["centers"], ref["centers"]) < 1e-12
        assert _relerr(meta["rotation"], ref["rotation"]) < 1e-7
        for a, b in zip(pcs, ref["pcs"]):
            assert _relerr(a, b) < 1e-7
        assert np.allclose(meta["var_explained"], ref["var_explained"], rtol=1e-9) and np.isclose(meta["var_total"], ref["var_total"], rtol=1e-9)
    # get_all_genes: genes outside subset_row get rotation vectors by projection (:393-405); check against the definition
    keep = np.arange(0, 300, 2) + 1
    pcs, meta = bb.multiBatchPCA(*mats, d=5, subset_row=keep, get_all_genes=True)
    sub = ho.multi_batch_pca([m[keep - 1] for m in mats], d=5)
    assert _relerr(meta["rotation"][keep - 1], sub["rotation"]) < 1e-7 and meta["rotation"].shape == (300, 5)
    res = bb.fastMNN(*mats, d=6, k=15)
    assert res.corrected.shape == (1550, 6) and res.merge_info["rotation"].shape == (300, 6)


def test_mnn_correct_auto_merge_matches_oracle():
    """mnnCorrect(auto.merge=TRUE): R/mnnCorrect.R:211-223 (pair counts without orthogonalisation) + R/MNN_tree.R:196-226."""
    A, B, Cc = synth.gene_batches(3, [220, 260, 200], G=100, latent=5, ncomp=3)
    got = bb.mnnCorrect(A, B, Cc, k=12, auto_merge=True, sigma=1.0)
    ref = ho.mnn_correct([A, B, Cc], k=12, auto_merge=True, sigma=1.0)
    assert got.merge_info["left"] == ref["merge_info"]["left"] and got.merge_info["right"] == ref["merge_info"]["right"]
    assert np.array_equal(got.batch, ref["batch"]) and _relerr(got.corrected, ref["corrected"]) < RTOL
