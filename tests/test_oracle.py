"""CPU tests of the oracle itself: it must reproduce (1) the committed golden vectors, which were produced by the
reference's OWN kernels compiled unmodified (tests/golden/make_golden.py), (2) that object code live when
/root/reference is present, and (3) the reference's test expectations restated from tests/testthat/*.R."""
import math

import numpy as np
import pytest

from oracle import capi, host_oracle as ho


# ---------------------------------------------------------------------------------------------------------------
# golden vectors (reference object code outputs)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["vanilla", "repeated", "many", "wide"])
def test_smooth_matches_golden(golden, case):
    g = golden["smooth_gaussian_kernel"]
    out = capi.smooth_gaussian_kernel(g[f"{case}_averaged"], g[f"{case}_index0"], g["data2"].T, float(g[f"{case}_sigma"]))
    ref = g[f"{case}_out"]
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_shift_variance_matches_golden(golden):
    g = golden["adjust_shift_variance"]
    for s in (1.0, 0.1):
        out = capi.adjust_shift_variance(g["data1"], g["data2"], g["vect"], s, np.arange(400), np.arange(1000))
        assert np.array_equal(out, g[f"out_sigma_{s}"])  # same operation order -> bit-identical
    out = capi.adjust_shift_variance(g["data1"], g["data2"], g["vect"], 1.0, g["r1"], g["r2"])
    assert np.array_equal(out, g["out_restricted"])


def test_mutual_pairs_match_golden(golden):
    g = golden["find_mutual_nns"]
    first, second = capi.find_mutual_nns(g["w21"], g["w12"])
    assert np.array_equal(first, g["first"]) and np.array_equal(second, g["second"])
    # and the kNN restatement regenerates the neighbour matrices that fed the reference kernel
    w21, d21 = capi.query_knn(g["X2"], g["X1"], 15)
    assert np.array_equal(w21, g["w21"]) and np.array_equal(d21, g["dist21"])


@pytest.mark.skipif(not capi.have_ref(), reason="reference tree absent (GPU box): golden vectors cover this")
def test_oracle_matches_reference_object_code_live():
    rng = np.random.default_rng(5)
    X1 = rng.normal(size=(250, 7)); X2 = rng.normal(size=(300, 7)) + 0.3
    w21, _ = capi.query_knn(X2, X1, 9); w12, _ = capi.query_knn(X1, X2, 6)
    a, b = capi.find_mutual_nns(w21, w12), capi.ref_find_mutual_nns(w21, w12)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    d1 = rng.normal(scale=0.1, size=(12, 120)); d2 = rng.normal(scale=0.1, size=(12, 200)); cv = rng.uniform(size=(200, 12))
    r1 = rng.permutation(120)[:70]; r2 = rng.permutation(200)[:150]
    assert np.array_equal(capi.adjust_shift_variance(d1, d2, cv, 0.3, r1, r2), capi.ref_adjust_shift_variance(d1, d2, cv, 0.3, r1, r2))
    avg = rng.normal(size=(20, 30)); idx = rng.permutation(200)[:30]
    o1 = capi.smooth_gaussian_kernel(avg, idx, d2[:, :], 0.2); o2 = capi.ref_smooth_gaussian_kernel(avg, idx, d2, 0.2)
    assert np.max(np.abs(o1 - o2)) <= 1e-12 * np.max(np.abs(o2))
    # error behaviour of the reference (src/smooth_gaussian_kernel.cpp:18-20, src/adjust_shift_variance.cpp:33-44)
    with pytest.raises(capi.OracleError, match="must have length equal"):
        capi.ref_smooth_gaussian_kernel(avg, idx[:5], d2, 0.2)
    with pytest.raises(capi.OracleError, match="must have length equal"):
        capi.smooth_gaussian_kernel(avg, idx[:5], d2, 0.2)
    with pytest.raises(capi.OracleError, match="subset indices out of range"):
        capi.ref_adjust_shift_variance(d1, d2, cv, 0.3, np.array([500]), r2)
    with pytest.raises(capi.OracleError, match="subset indices out of range"):
        capi.adjust_shift_variance(d1, d2, cv, 0.3, np.array([500]), r2)
    with pytest.raises(capi.OracleError, match="number of genes"):
        capi.adjust_shift_variance(d1[:5], d2, cv, 0.3, r1, r2)


# ---------------------------------------------------------------------------------------------------------------
# kNN restatement (parity UNPINNED by the reference: BiocNeighbors is external) -- cross-checks
# ---------------------------------------------------------------------------------------------------------------
def test_knn_brute_vs_kdtree_and_kmknn():
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(11)
    X = rng.normal(size=(3000, 12)); Q = rng.normal(size=(500, 12))
    idx, dist = capi.query_knn(X, Q, 10)
    dd, ii = cKDTree(X).query(Q, 10)
    assert np.array_equal(idx, ii + 1)
    assert np.allclose(dist, dd, rtol=1e-12)
    ki, kd = capi.Kmknn(X).query(Q, 10)
    assert np.array_equal(ki, idx) and np.array_equal(kd, dist)


def test_knn_ties_broken_by_index():
    # integer grid: massive exact ties; expectation computed with a stable (distance, index) lexsort in numpy
    g = np.stack(np.meshgrid(np.arange(8.0), np.arange(8.0), indexing="ij"), -1).reshape(-1, 2)
    X = np.vstack([g, g])  # every point duplicated -> ties between index i and i+64
    idx, dist = capi.query_knn(X, g, 9)
    d2 = ((g[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    order = np.lexsort((np.broadcast_to(np.arange(X.shape[0]), d2.shape), d2), axis=1)[:, :9]
    assert np.array_equal(idx, order + 1)
    ki, _ = capi.Kmknn(X).query(g, 9)
    assert np.array_equal(ki, idx)


# ---------------------------------------------------------------------------------------------------------------
# restated testthat expectations
# ---------------------------------------------------------------------------------------------------------------
def _ref_correction_vectors(data1, data2, mnn1, mnn2, s2):
    """REF of tests/testthat/test-mnn-correct.R:36-65 (dense kernel)."""
    d = ((data2[:, None, :] - data2[None, :, :]) ** 2).sum(-1)
    w = np.exp(-d / s2)
    uniq = np.unique(mnn2) - 1
    dens = w[:, uniq].sum(axis=1)
    N = np.bincount(mnn2 - 1, minlength=data2.shape[0]).astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        kern = (w / (N * dens)[:, None]).T[:, mnn2 - 1]
    kern = kern / kern.sum(axis=1)[:, None]
    return kern @ (data1[mnn1 - 1] - data2[mnn2 - 1])


def test_correction_vectors_match_testthat_ref():
    rng = np.random.default_rng(10003)
    data1 = rng.normal(scale=0.1, size=(400, 25)); data2 = rng.normal(scale=0.1, size=(1000, 25))
    for mnn1, mnn2, s2 in [(np.arange(1, 11), np.arange(30, 20, -1), 0.1),
                           (np.r_[11, 12, 13, np.arange(1, 11)], np.r_[30, 30, 30, np.arange(30, 20, -1)], 0.1),
                           (np.arange(1, 201), np.arange(500, 300, -1), 0.1),
                           (np.arange(1, 11), np.arange(30, 20, -1), 0.5)]:
        got = ho.compute_correction_vectors(data1, data2, mnn1, mnn2, np.asfortranarray(data2.T), s2)
        assert np.allclose(got, _ref_correction_vectors(data1, data2, mnn1, mnn2, s2), rtol=1e-8, atol=1e-12)


def _ref_shift_variance(data1, data2, cell_vect, sigma):
    """REF of tests/testthat/test-mnn-correct.R:101-138."""
    d1, d2 = data1.T, data2.T
    out = np.zeros(cell_vect.shape[0])
    for c in range(cell_vect.shape[0]):
        v = cell_vect[c]; l2 = math.sqrt((v ** 2).sum()); v = v / l2
        c2, c1 = d2 @ v, d1 @ v
        def wts(mat):
            diff = d2[c][:, None] - mat.T
            diff = diff - np.outer(v, v @ diff)
            return np.exp(-(diff ** 2).sum(0) / sigma)
        w2, w1 = wts(d2), wts(d1)
        rank2 = np.argsort(np.argsort(c2, kind="stable"), kind="stable")
        prob2 = w2[rank2 <= rank2[c]].sum() / w2.sum()
        o1 = np.argsort(c1, kind="stable")
        cs = np.cumsum(w1[o1]); ecdf = cs / cs[-1]  # R's sum() and cumsum() agree on the last element
        out[c] = (c1[o1[np.nonzero(ecdf >= prob2)[0].min()]] - c2[c]) / l2
    return out


def test_shift_variance_matches_testthat_ref():
    rng = np.random.default_rng(100032)
    d1 = rng.normal(scale=0.1, size=(25, 120)); d2 = rng.normal(scale=0.1, size=(25, 200)); cv = rng.uniform(size=(200, 25))
    for s in (1.0, 0.1):
        got = capi.adjust_shift_variance(d1, d2, cv, s, np.arange(120), np.arange(200))
        want = _ref_shift_variance(d1, d2, cv, s)
        assert np.mean(np.isclose(got, want, rtol=1e-8, atol=1e-10)) >= 0.99  # discrete quantile pick: allow rare ties
    # subsetting (:156-160) and restriction identities (:162-173)
    i = np.arange(9, 20)
    t1 = ho.adjust_shift_variance(d1, d2, cv, 1.0, subset_row=i + 1)
    t2 = ho.adjust_shift_variance(d1[i], d2[i], cv[:, i], 1.0)
    assert np.allclose(t1[:, i], t2)
    i1 = np.arange(10, 21); i2 = np.arange(20, 9, -1)
    A1 = np.hstack([d1, d1[:, i1 - 1]]); A2 = np.hstack([d2, d2[:, i2 - 1]])
    a = ho.adjust_shift_variance(d1, d2, cv, 1.0)
    b = ho.adjust_shift_variance(A1, A2, np.vstack([cv, cv[i2 - 1]]), 1.0, restrict1=np.arange(1, 121), restrict2=np.arange(1, 201))
    assert np.array_equal(a, b[:200]) and np.array_equal(a[i2 - 1], b[200:])


def test_average_correction_matches_testthat():  # test-fast-mnn.R:6-32
    rng = np.random.default_rng(1200001)
    t1 = rng.normal(size=(100, 10)); t2 = rng.normal(size=(200, 10))
    m1 = rng.integers(1, 101, size=250); m2 = rng.integers(1, 101, size=250)
    averaged, second = ho.average_correction(t1, m1, t2, m2)
    correct = t1[m1 - 1] - t2[m2 - 1]
    ref = np.stack([correct[m2 == u].mean(axis=0) for u in np.unique(m2)])
    assert np.allclose(averaged, ref) and np.array_equal(second, np.unique(m2))
    e, s = ho.average_correction(t1, np.zeros(0, int), t2, np.zeros(0, int))
    assert e.shape == (0, 10) and s.size == 0


def test_centering_matches_testthat():  # test-fast-mnn.R:34-51
    rng = np.random.default_rng(1200002)
    t = rng.normal(size=(100, 10)); b = rng.normal(size=10)
    assert np.std(ho.center_along_batch_vector(t, b) @ b) < 1e-8
    t2 = np.vstack([t, t[:10]])
    assert np.array_equal(ho.center_along_batch_vector(t, b), ho.center_along_batch_vector(t2, b, restrict=np.arange(1, 101))[:100])


def test_tricube_matches_testthat():  # test-fast-mnn.R:53-92 and test-utils.R:82-115
    rng = np.random.default_rng(1200003)
    t = rng.normal(size=(100, 10)); corr = rng.normal(size=(50, 10)); involved = rng.permutation(100)[:50] + 1
    for k, nd in [(20, 3), (11, 3), (11, 1)]:
        out = ho.tricube_weighted_correction(t, corr, involved, k=k, ndist=nd)
        sub = t[involved - 1]; sk = min(k, 50); mid = math.ceil(sk / 2)
        idx, dist = capi.query_knn(sub, t, sk)
        ref = t.copy()
        for x in range(100):
            md = np.sort(dist[x])[mid - 1]
            w = (1 - np.minimum(1, dist[x] / (md * nd)) ** 3) ** 3; w = w / w.sum()
            ref[x] = t[x] + (corr[idx[x] - 1] * w[:, None]).sum(0)
        assert np.allclose(out, ref)
    # degenerate cases of test-utils.R:97-115
    vals = rng.normal(size=(30, 4)); idx = rng.integers(1, 31, size=(20, 1)); dist = rng.uniform(size=(20, 1))
    assert np.allclose(ho.compute_tricube_average(vals, idx, dist), vals[idx[:, 0] - 1])
    idx = rng.integers(1, 31, size=(20, 5)); dist = np.full((20, 5), 0.7)
    assert np.allclose(ho.compute_tricube_average(vals, idx, dist), np.stack([vals[r - 1].mean(0) for r in idx]))
    assert np.array_equal(ho.compute_tricube_average(vals, np.zeros((20, 0), int), np.zeros((20, 0))), np.zeros((30, 4)))


def test_merge_tree_order_and_reorder_utils():  # test-tree.R, test-utils.R:117-152, test-fast-mnn.R:365-366
    assert ho.binarize_tree([1, 2, 3, 4]) == [[[1, 2], 3], 4]
    assert ho.binarize_tree([[1, 2], [3, 4]]) == [[1, 2], [3, 4]]
    assert ho.binarize_tree([[1], [[2, 3]]]) == [1, [2, 3]]
    batches = [np.zeros((3, 2)) for _ in range(4)]
    tree = ho.create_tree_predefined(batches, None, [[1, 2], [3, 4]])
    l, r, path = ho.get_next_merge(tree)
    assert (l.index, r.index, path) == ([3], [4], (1,))  # right subtree first
    with pytest.raises(ValueError):
        ho.create_tree_predefined(batches, None, [1, 2, 2, 4])
    assert np.array_equal(ho.restore_original_order([3, 1, 2], [2, 3, 1]), [2, 3, 4, 5, 6, 1])
    pairs = ho.reindex_pairings([(np.array([1, 2]), np.array([3, 4]))], np.array([2, 3, 4, 1]))
    assert np.array_equal(pairs[0][0], [4, 1]) and np.array_equal(pairs[0][1], [2, 3])
    assert ho.choose_k(20, None, 1000) == 20 and ho.choose_k(10, 0.05, 1000) == 50 and ho.choose_k(10, 0.5, 15) == 10


def test_reduced_mnn_toy_known_answers():  # test-reduced-mnn.R:80-105
    core = np.stack([np.repeat(np.arange(1.0, 11), 10), np.tile(np.arange(1.0, 11), 10)], 1)
    b1 = core.copy(); b1[:, 0] += 20
    b2 = core.copy(); b2[:, 1] += 20
    o1 = ho.reduced_mnn([core, b1], k=1)["corrected"]
    assert np.allclose(o1[:, 0], 5.5) and np.allclose(o1[:, 1], np.r_[core[:, 1], b1[:, 1]])
    o2 = ho.reduced_mnn([core, b1, b2], k=1)["corrected"]
    assert np.allclose(o2, 5.5)
    oy = ho.reduced_mnn([core + 10, b2 + 10], k=1)["corrected"]
    assert np.allclose(oy[:, 0], np.r_[core[:, 0], b2[:, 0]] + 10) and np.allclose(oy[:, 1], 15.5)
    oz = ho.reduced_mnn([core, b1, core + 10, b2 + 10], merge_order=[[1, 2], [3, 4]], k=1)["corrected"]
    assert np.allclose(oz, 5.5)


def test_reduced_mnn_restrict_and_propk_identities():  # test-reduced-mnn.R:39-58, 107-145
    rng = np.random.default_rng(12000053)
    B1 = rng.normal(0, size=(300, 10)); B2 = rng.normal(1, size=(400, 10)); B3 = rng.normal(2, size=(200, 10))
    ref = ho.reduced_mnn([B1, B2, B3])
    i1 = np.arange(100, 49, -1); i2 = np.arange(1, 21); i3 = np.arange(50, 101)
    C1 = np.vstack([B1, B1[i1 - 1]]); C2 = np.vstack([B2, B2[i2 - 1]]); C3 = np.vstack([B3, B3[i3 - 1]])
    out = ho.reduced_mnn([C1, C2, C3], restrict=[np.arange(1, 301), np.arange(1, 401), np.arange(1, 201)])
    for b, keep, dup in [(1, 300, i1), (2, 400, i2), (3, 200, i3)]:
        r = ref["corrected"][ref["batch"] == b]; o = out["corrected"][out["batch"] == b]
        assert np.array_equal(r, o[:keep]) and np.array_equal(r[dup - 1], o[keep:])
    a = ho.reduced_mnn([B1, B2], k=10, prop_k=20 / 300)  # k1 = 20, k2 = round(26.7)=27 -> compare with explicit
    assert a["corrected"].shape == (700, 10)
    p1 = ho.reduced_mnn([B1, B1 + 1.0], k=10, prop_k=20 / 300)
    p2 = ho.reduced_mnn([B1, B1 + 1.0], k=20)
    assert np.array_equal(p1["corrected"], p2["corrected"])


def test_cosine_norm_matches_testthat():  # test-cos-norm.R:4-46
    rng = np.random.default_rng(3)
    X = rng.normal(size=(20, 30)); X[:, 4] = 0
    out, l2 = capi.cosine_norm(X)
    assert np.allclose(l2, np.sqrt((X ** 2).sum(0))) and np.allclose(out, X / np.maximum(1e-8, l2))
    assert np.all(out[:, 4] == 0)
    m, l = ho.cosine_norm(X, "all")
    assert np.allclose(m, out) and np.allclose(l, l2)


def test_mnn_correct_oracle_runs_and_restrict_identity():  # test-mnn-correct.R:379-441 (shape of the identity)
    rng = np.random.default_rng(10004)
    A = rng.normal(size=(15, 60)); B = rng.normal(size=(15, 80)) + 1
    ref = ho.mnn_correct([A, B], k=10)
    assert ref["corrected"].shape == (15, 140) and np.all(np.isfinite(ref["corrected"]))
    i1 = np.arange(5, 16); i2 = np.arange(20, 9, -1)
    A2 = np.hstack([A, A[:, i1 - 1]]); B2 = np.hstack([B, B[:, i2 - 1]])
    out = ho.mnn_correct([A2, B2], k=10, restrict=[np.arange(1, 61), np.arange(1, 81)])
    r = ref["corrected"]; o = out["corrected"]
    assert np.array_equal(r[:, :60], o[:, :60]) and np.array_equal(r[:, 60:], o[:, 71:151])
    assert np.array_equal(r[:, 60:][:, i2 - 1], o[:, 151:])


def test_shift_variance_cell_subset_equals_full_loop(golden):
    """The subset form used to check sampled cells of problems too large for a full CPU pass is the same per-cell loop."""
    g = golden["adjust_shift_variance"]
    cells = np.array([0, 7, 500, 999, 3], dtype=np.int64)
    sub = capi.adjust_shift_variance_cells(g["data1"], g["data2"], g["vect"][cells], cells, 0.1, np.arange(400), np.arange(1000))
    assert np.array_equal(sub, g["out_sigma_0.1"][cells])


def test_lost_var_is_recorded_before_the_tricube_step():
    """R/fastMNN.R:500-501 records the variance right after the centring along the batch vector.  Hand computation: the
    centring removes exactly the variance of the projections on the unit batch vector (no restriction, first merge)."""
    rng = np.random.default_rng(11)
    b1 = rng.normal(size=(300, 10)); b2 = rng.normal(size=(250, 10)) + 2.0
    res = ho.reduced_mnn([b1, b2], k=15)
    first, second = ho.restricted_mnn(b1, None, b2, None, 15)
    averaged, _ = ho.average_correction(b1, first, b2, second)
    v = averaged.mean(axis=0); v /= np.linalg.norm(v)
    want = [np.var(X @ v, ddof=1) / np.var(X, axis=0, ddof=1).sum() for X in (b1, b2)]
    assert np.allclose(res["merge_info"]["lost_var"][0], want, rtol=1e-9, atol=1e-12)


def test_auto_merge_picks_the_pair_with_most_mnn_pairs():
    """R/MNN_tree.R:154-226: with two overlapping batches and one distant batch the first merge must be the overlapping
    pair, and the result must equal the predefined order that auto.merge discovers."""
    rng = np.random.default_rng(5)
    A = rng.normal(size=(200, 6)); B = rng.normal(size=(180, 6)) + 0.3; Cc = rng.normal(size=(150, 6)); Cc[:, 0] += 6.0
    auto = ho.reduced_mnn([A, Cc, B], k=10, auto_merge=True)
    first = (auto["merge_info"]["left"][0], auto["merge_info"]["right"][0])
    assert sorted(first[0] + first[1]) == [1, 3]
    lhs, rhs = first[0][0], first[1][0]
    fixed = ho.reduced_mnn([A, Cc, B], k=10, merge_order=[lhs, rhs, 2])
    assert np.allclose(auto["corrected"], fixed["corrected"], rtol=0, atol=0)
