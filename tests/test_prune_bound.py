"""CPU check of the mathematics behind the cluster pruning of the kNN search (csrc/knn_cluster.cu): the per-cluster
lower bound must never exceed a true query-to-reference distance, whatever the clustering looks like.  numpy only --
this is the argument the CUDA path relies on for exactness, restated and tested without a GPU.

    LB(q, B) = -ext_B(A) - u.q,   u = (c_B - c_A) / |c_B - c_A|,   ext_B(A) = max_{x in B} (x.c_A - x.c_B) / |c_B - c_A|
"""
import numpy as np
import pytest


def _plan(X, Q, C, rng):
    cen = X[rng.choice(X.shape[0], C, replace=False)].copy()
    for _ in range(3):   # a few Lloyd steps; the bound must hold for ANY centroids and ANY assignment
        a = np.argmin(((X[:, None, :] - cen[None]) ** 2).sum(-1), axis=1)
        for c in range(C):
            if np.any(a == c):
                cen[c] = X[a == c].mean(0)
    ax = np.argmin(((X[:, None, :] - cen[None]) ** 2).sum(-1), axis=1)
    aq = np.argmin(((Q[:, None, :] - cen[None]) ** 2).sum(-1), axis=1)
    return cen, ax, aq


@pytest.mark.parametrize("d,C,seed", [(2, 8, 0), (10, 16, 1), (50, 16, 2), (50, 32, 3)])
def test_cluster_lower_bound_never_exceeds_a_true_distance(d, C, seed):
    rng = np.random.default_rng(seed)
    centres = rng.normal(scale=3.0, size=(6, d))
    X = centres[rng.integers(0, 6, 3000)] + rng.normal(size=(3000, d))
    Q = centres[rng.integers(0, 6, 500)] + rng.normal(size=(500, d)) + rng.normal(size=d) * 0.5
    cen, ax, aq = _plan(X, Q, C, rng)
    D = np.sqrt(((cen[:, None, :] - cen[None]) ** 2).sum(-1))
    gx, gq = X @ cen.T, Q @ cen.T
    # ext[B, A] = max over x in B of (x.c_A - x.c_B) / D[A, B]
    ext = np.full((C, C), -np.inf)
    for B in range(C):
        m = ax == B
        if m.any():
            with np.errstate(divide="ignore", invalid="ignore"):
                ext[B] = ((gx[m] - gx[m][:, [B]]) / D[B][None, :]).max(0)
    dist = np.sqrt(((Q[:, None, :] - X[None]) ** 2).sum(-1))           # [nq, n]
    checked = 0
    for B in range(C):
        m = ax == B
        if not m.any():
            continue
        true_min = dist[:, m].min(1)                                    # nearest member of B for every query
        A = aq
        with np.errstate(divide="ignore", invalid="ignore"):
            lb = -ext[B, A] - (gq[:, B] - gq[np.arange(len(Q)), A]) / D[A, B]
        ok = (A != B) & (D[A, B] > 1e-9)
        assert np.all(lb[ok] <= true_min[ok] + 1e-9), "lower bound exceeds a true distance"
        checked += int(ok.sum())
    assert checked > 0


def test_bound_is_useful_on_separated_clusters():
    """On well separated blobs the bound must actually exclude the far clusters (otherwise nothing would be pruned)."""
    rng = np.random.default_rng(5)
    d, C = 50, 8
    centres = rng.normal(scale=3.0, size=(C, d))
    lab = rng.integers(0, C, 4000)
    X = centres[lab] + rng.normal(size=(4000, d))
    Q = centres[lab[:400]] + rng.normal(size=(400, d))
    cen = np.stack([X[lab == c].mean(0) for c in range(C)])
    ax = np.argmin(((X[:, None, :] - cen[None]) ** 2).sum(-1), axis=1)
    aq = np.argmin(((Q[:, None, :] - cen[None]) ** 2).sum(-1), axis=1)
    D = np.sqrt(((cen[:, None, :] - cen[None]) ** 2).sum(-1)) + np.eye(C)
    gx, gq = X @ cen.T, Q @ cen.T
    ext = np.stack([((gx[ax == B] - gx[ax == B][:, [B]]) / D[B][None, :]).max(0) for B in range(C)])
    dist2 = ((Q[:, None, :] - X[None]) ** 2).sum(-1)
    kth = np.sort(dist2, axis=1)[:, 31]                                 # 32nd best squared distance of every query
    pruned = 0
    for B in range(C):
        lb = -ext[B, aq] - (gq[:, B] - gq[np.arange(len(Q)), aq]) / D[aq, B]
        pruned += int(((aq != B) & (lb > 0) & (lb ** 2 > kth)).sum())
    assert pruned > 0.8 * len(Q) * (C - 1), f"only {pruned} of {len(Q) * (C - 1)} (query, far cluster) pairs excluded"
