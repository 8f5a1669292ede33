"""TEST INFRASTRUCTURE ONLY -- numpy/fp64 restatement of the R host code around the MNN hot path.

Every function cites the reference lines (under /root/reference) it restates.  Indices are 1-BASED wherever
the R code's are, so the restated testthat cases in tests/ read like the originals.  The exact kNN and the three
native kernels come from ``oracle.capi`` (C restatement / the reference's own compiled kernels).

Parity status: the functions here are pinned by restating the reference's own tests
(tests/testthat/test-fast-mnn.R:6-92, test-utils.R:82-152, test-tree.R, test-reduced-mnn.R:80-145,
test-mnn-correct.R:28-174, test-cos-norm.R) in tests/test_oracle.py; the kNN they call is UNPINNED (see
mnn_oracle.c header): BiocNeighbors is not in /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

from . import capi


# ------------------------------------------------------------------------------------------------
# R/fastMNN.R:567-580  .average_correction
# ------------------------------------------------------------------------------------------------
def average_correction(refdata, mnn1, curdata, mnn2):
    refdata = np.asarray(refdata, dtype=np.float64)
    curdata = np.asarray(curdata, dtype=np.float64)
    mnn1 = np.asarray(mnn1, dtype=np.int64)
    mnn2 = np.asarray(mnn2, dtype=np.int64)
    corvec = refdata[mnn1 - 1, :] - curdata[mnn2 - 1, :]
    second, inv, counts = np.unique(mnn2, return_inverse=True, return_counts=True)
    summed = np.zeros((second.size, curdata.shape[1]))
    np.add.at(summed, inv, corvec)
    averaged = summed / counts[:, None] if second.size else summed
    return averaged, second.astype(np.int32)


# R/fastMNN.R:582-595  .get_batch_magnitude
def get_batch_magnitude(correction, ave=None):
    correction = np.asarray(correction, dtype=np.float64)
    if ave is None:
        ave = correction.mean(axis=0)
    ave_l2sq = float(np.sum(np.mean(correction ** 2, axis=0)))
    if ave_l2sq == 0:
        return 0.0
    return math.sqrt(float(np.sum(ave ** 2)) / ave_l2sq)


# R/fastMNN.R:626-640  .center_along_batch_vector   (restrict is 1-based or None)
def center_along_batch_vector(mat, batch_vec, restrict=None):
    mat = np.asarray(mat, dtype=np.float64)
    v = np.asarray(batch_vec, dtype=np.float64)
    v = v / math.sqrt(float(np.sum(v ** 2)))
    loc = (mat * v[None, :]).sum(axis=1)  # row-wise sums: identical rows give identical projections (BLAS gemv need not)
    central = loc.mean() if restrict is None else loc[np.asarray(restrict, dtype=np.int64) - 1].mean()
    return mat + np.outer(central - loc, v)


# R/fastMNN.R:642-647  .orthogonalize_other
def orthogonalize_other(data, restrict, vectors):
    for vec in vectors:
        data = center_along_batch_vector(data, vec, restrict)
    return data


# R/fastMNN.R:651-658  .compute_perbatch_var  (sum over dims of the sample variance of each original batch)
def compute_perbatch_var(data, index, origin):
    out = np.zeros(len(index))
    origin = np.asarray(origin)
    for i, b in enumerate(index):
        rows = data[origin == b]
        out[i] = float(np.sum(np.var(rows, axis=0, ddof=1))) if rows.shape[0] > 1 else float("nan")
    return out


# R/fastMNN.R:610-622  .combine_restrict
def combine_restrict(nleft, left_restrict, nright, right_restrict):
    if left_restrict is None and right_restrict is None:
        return None
    if left_restrict is None:
        left_restrict = np.arange(1, nleft + 1)
    if right_restrict is None:
        right_restrict = np.arange(1, nright + 1)
    return np.concatenate([np.asarray(left_restrict, dtype=np.int64), np.asarray(right_restrict, dtype=np.int64) + nleft])


# ------------------------------------------------------------------------------------------------
# R/utils_tricube.R:1-27  .compute_tricube_average
# ------------------------------------------------------------------------------------------------
def compute_tricube_average(vals, indices, distances, bandwidth=None, ndist=3):
    vals = np.asarray(vals, dtype=np.float64)
    indices = np.asarray(indices, dtype=np.int64)
    distances = np.asarray(distances, dtype=np.float64)
    if indices.ndim != 2 or indices.shape[1] == 0:
        return np.zeros((vals.shape[0], vals.shape[1]))  # utils_tricube.R:22-23 zero-column guard
    if bandwidth is None:
        middle = int(math.ceil(indices.shape[1] / 2))
        bandwidth = distances[:, middle - 1] * ndist
    bandwidth = np.maximum(1e-8, bandwidth)
    rel = distances / bandwidth[:, None]
    rel[rel > 1] = 1
    tricube = (1 - rel ** 3) ** 3
    weight = tricube / tricube.sum(axis=1)[:, None]
    out = np.zeros((indices.shape[0], vals.shape[1]))
    for kdx in range(indices.shape[1]):
        out = out + vals[indices[:, kdx] - 1, :] * weight[:, kdx][:, None]
    return out


# R/fastMNN.R:599-608  .tricube_weighted_correction
def tricube_weighted_correction(curdata, correction, in_mnn, k=20, ndist=3, knn=None):
    knn = knn or capi.query_knn
    curdata = np.asarray(curdata, dtype=np.float64)
    in_mnn = np.asarray(in_mnn, dtype=np.int64)
    cur_uniq = curdata[in_mnn - 1, :]
    safe_k = min(k, cur_uniq.shape[0])
    idx, dist = knn(cur_uniq, curdata, safe_k)
    return curdata + compute_tricube_average(correction, idx, dist, ndist=ndist)


# ------------------------------------------------------------------------------------------------
# R/MNN_tree.R:113-146  .restricted_mnn / .unrestrict_indices / .choose_k
# ------------------------------------------------------------------------------------------------
def r_round(x):
    """R's round(): IEC 60559 half-to-even, which is also Python's."""
    return int(round(x))


def choose_k(k, prop_k, N):
    if prop_k is None:
        return k
    return min(N, max(k, r_round(prop_k * N)))


def restricted_mnn(left, left_restrict, right, right_restrict, k, prop_k=None, mutual=None):
    mutual = mutual or capi.find_mutual_nn
    L = left if left_restrict is None else left[np.asarray(left_restrict, dtype=np.int64) - 1]
    R = right if right_restrict is None else right[np.asarray(right_restrict, dtype=np.int64) - 1]
    k1 = choose_k(k, prop_k, L.shape[0])
    k2 = choose_k(k, prop_k, R.shape[0])
    first, second = mutual(L, R, k1, k2)
    if left_restrict is not None:
        first = np.asarray(left_restrict, dtype=np.int64)[first - 1]
    if right_restrict is not None:
        second = np.asarray(right_restrict, dtype=np.int64)[second - 1]
    return np.asarray(first, dtype=np.int32), np.asarray(second, dtype=np.int32)


# ------------------------------------------------------------------------------------------------
# R/MNN_tree.R:2-109  MNN_treenode and the predefined merge tree
# ------------------------------------------------------------------------------------------------
class Node:
    def __init__(self, index, data, restrict, origin=None, extras=None):
        self.index = list(index) if isinstance(index, (list, tuple, np.ndarray)) else [index]
        self.data = data
        self.restrict = restrict
        self.origin = np.repeat(self.index[0], data.shape[0]) if origin is None else origin
        self.extras = [] if extras is None else extras


def binarize_tree(tree):  # R/MNN_tree.R:21-45
    if not isinstance(tree, (list, tuple)):
        return tree
    n = len(tree)
    if n == 0:
        raise ValueError("merge tree contains a node with no children")
    if n == 1:
        return binarize_tree(tree[0])
    cur = [binarize_tree(tree[0]), binarize_tree(tree[1])]
    for i in range(2, n):
        cur = [cur, binarize_tree(tree[i])]
    return cur


def _leaves(tree):
    if not isinstance(tree, list):
        return [tree]
    return _leaves(tree[0]) + _leaves(tree[1])


def create_tree_predefined(batches, restrict, merge_order):  # R/MNN_tree.R:80-109
    nb = len(batches)
    if merge_order is None:
        merge_order = list(range(1, nb + 1))
    if not any(isinstance(m, (list, tuple)) for m in merge_order) and len(merge_order) > 1:
        tree = [merge_order[0], merge_order[1]]
        for i in merge_order[2:]:
            tree = [tree, i]
    else:
        tree = merge_order
    tree = binarize_tree(list(tree) if isinstance(tree, tuple) else tree)
    leaves = _leaves(tree)
    if (any((not isinstance(l, (int, np.integer))) for l in leaves) or len(set(leaves)) != len(leaves)
            or any(l < 1 or l > nb for l in leaves)):
        raise ValueError("invalid leaf nodes specified in 'merge.order'")

    def fill(t):
        if not isinstance(t, list):
            r = None if restrict is None else restrict[t - 1]
            return Node(int(t), np.asarray(batches[t - 1], dtype=np.float64), r)
        return [fill(t[0]), fill(t[1])]

    return fill(tree)


def get_next_merge(tree, path=()):  # R/MNN_tree.R:61-69: right subtree first
    if not isinstance(tree[0], list) and not isinstance(tree[1], list):
        return tree[0], tree[1], path
    if isinstance(tree[1], list):
        return get_next_merge(tree[1], path + (1,))
    return get_next_merge(tree[0], path + (0,))


def update_tree(tree, path, node):  # R/MNN_tree.R:71-77
    if len(path) == 0:
        return node
    tree[path[0]] = update_tree(tree[path[0]], path[1:], node)
    return tree


# R/utils_reorder.R:1-36
def restore_original_order(batch_ordering, ncells_per_batch):
    if len(batch_ordering) != len(ncells_per_batch):
        raise ValueError("length of batch information vectors are not equal")
    reorder = [None] * len(batch_ordering)
    last = 0
    for idx in batch_ordering:
        n = ncells_per_batch[idx - 1]
        reorder[idx - 1] = last + np.arange(1, n + 1)
        last += n
    return np.concatenate(reorder) if reorder else np.zeros(0, dtype=np.int64)


def reindex_pairings(pairings, new_order):
    new_order = np.asarray(new_order, dtype=np.int64)
    rev = np.zeros(new_order.size, dtype=np.int64)
    rev[new_order - 1] = np.arange(1, new_order.size + 1)
    return [(rev[l - 1], rev[r - 1]) for (l, r) in pairings]


# ------------------------------------------------------------------------------------------------
# R/fastMNN.R:436-562  .fast_mnn_core  (predefined merge tree; reducedMNN = this on given PCs, R/reducedMNN.R:61-95)
# ------------------------------------------------------------------------------------------------
def _count_mnn_pairs(left, remainders, upto, k, prop_k, mutual):
    """R/MNN_tree.R:171-193 .count_mnn_pairs: note that left.data keeps the centrings of every earlier iteration."""
    ld = left.data
    n = np.zeros(upto, dtype=np.int64)
    for j in range(upto):
        right = remainders[j]
        rd = orthogonalize_other(right.data, right.restrict, left.extras)
        ld = orthogonalize_other(ld, left.restrict, right.extras)
        first, _ = restricted_mnn(ld, left.restrict, rd, right.restrict, k, prop_k, mutual=mutual)
        n[j] = len(first)
    return n


def reduced_mnn(batches, k=20, prop_k=None, restrict=None, ndist=3, merge_order=None, min_batch_skip=0.0,
                knn=None, mutual=None, auto_merge=False):
    batches = [np.asarray(b, dtype=np.float64) for b in batches]
    nb = len(batches)
    if auto_merge:   # R/MNN_tree.R:154-168 .initialize_auto_search
        remainders = [Node([i + 1], batches[i], None if restrict is None else restrict[i]) for i in range(nb)]
        pairwise = np.zeros((nb, nb), dtype=np.int64)
        for i in range(nb):
            pairwise[i, :i] = _count_mnn_pairs(remainders[i], remainders, i, k, prop_k, mutual)
        tree = None
    else:
        tree = create_tree_predefined(batches, restrict, merge_order)
    nmerges = nb - 1
    pairings, left_set, right_set = [], [], []
    batch_size = np.full(nmerges, np.nan)
    skipped = np.zeros(nmerges, dtype=bool)
    var_kept = np.ones((nmerges, nb))
    for mdx in range(nmerges):
        if auto_merge:   # .pick_best_merge (:196-202): which(stats == max(stats), arr.ind=TRUE)[1,] -> column-major first
            best = pairwise.max()
            cols, rows = np.nonzero(pairwise.T == best)
            chosen = (int(rows[0]), int(cols[0]))
            left, right = remainders[chosen[0]], remainders[chosen[1]]
        else:
            left, right, path = get_next_merge(tree)
        ld, rd = left.data, right.data
        left_old = compute_perbatch_var(ld, left.index, left.origin)
        right_old = compute_perbatch_var(rd, right.index, right.origin)
        left_set.append(list(left.index)); right_set.append(list(right.index))
        rd = orthogonalize_other(rd, right.restrict, left.extras)
        ld = orthogonalize_other(ld, left.restrict, right.extras)
        first, second = restricted_mnn(ld, left.restrict, rd, right.restrict, k, prop_k, mutual=mutual)
        averaged, _ = average_correction(ld, first, rd, second)
        overall = averaged.mean(axis=0) if averaged.shape[0] else np.full(ld.shape[1], np.nan)
        do_correct = True
        if min_batch_skip is not None and not (isinstance(min_batch_skip, float) and math.isnan(min_batch_skip)):
            mag = get_batch_magnitude(averaged, overall)
            batch_size[mdx] = mag
            if mag < min_batch_skip:
                do_correct = False
                skipped[mdx] = True
        if do_correct:
            ld = center_along_batch_vector(ld, overall, left.restrict)
            rd = center_along_batch_vector(rd, overall, right.restrict)
            # lost variance is recorded after the centring and BEFORE the tricube smoothing (R/fastMNN.R:500-501 vs :506)
            left_new = compute_perbatch_var(ld, left.index, left.origin)
            right_new = compute_perbatch_var(rd, right.index, right.origin)
            to_add = [overall]
            re_avg, re_second = average_correction(ld, first, rd, second)
            rd = tricube_weighted_correction(rd, re_avg, re_second, k=choose_k(k, prop_k, rd.shape[0]), ndist=ndist, knn=knn)
        else:
            to_add = []
            left_new = compute_perbatch_var(ld, left.index, left.origin)     # R/fastMNN.R:510-512
            right_new = compute_perbatch_var(rd, right.index, right.origin)
        var_kept[mdx, np.asarray(left.index) - 1] = left_new / left_old
        var_kept[mdx, np.asarray(right.index) - 1] = right_new / right_old
        pairings.append((first.astype(np.int64), second.astype(np.int64)))
        node = Node(left.index + right.index, np.vstack([ld, rd]),
                    combine_restrict(ld.shape[0], left.restrict, rd.shape[0], right.restrict),
                    origin=np.concatenate([left.origin, right.origin]), extras=left.extras + right.extras + to_add)
        if auto_merge:   # .update_remainders (:205-226)
            keep = [i for i in range(len(remainders)) if i not in chosen]
            remainders = [remainders[i] for i in keep]
            if remainders:
                old = pairwise[np.ix_(keep, keep)]
                new_stats = _count_mnn_pairs(node, remainders, len(remainders), k, prop_k, mutual)
                pairwise = np.hstack([np.vstack([old, new_stats[None, :]]), np.zeros((len(keep) + 1, 1), dtype=np.int64)])
                remainders.append(node)
            else:
                tree = node
        else:
            tree = update_tree(tree, path, node)
    return _finish(tree, tree.data, pairings, left_set, right_set,
                   dict(batch_size=batch_size, skipped=skipped, lost_var=1 - var_kept))


def _finish(tree, full_data, pairings, left_set, right_set, extra):
    full_order = tree.index
    full_origin = np.asarray(tree.origin)
    out_pairs = []
    for (l, r), ls, rs in zip(pairings, left_set, right_set):
        bonus1 = int(np.nonzero(full_origin == ls[0])[0][0])
        bonus2 = int(np.nonzero(full_origin == rs[0])[0][0])
        out_pairs.append((l + bonus1, r + bonus2))
    if any(full_order[i] > full_order[i + 1] for i in range(len(full_order) - 1)):
        ncells = np.bincount(full_origin, minlength=max(full_order) + 1)[1:]
        ordering = restore_original_order(full_order, ncells)
        full_data = full_data[ordering - 1]
        full_origin = full_origin[ordering - 1]
        out_pairs = reindex_pairings(out_pairs, ordering)
    info = dict(left=left_set, right=right_set, pairs=out_pairs)
    info.update(extra)
    return dict(corrected=full_data, batch=full_origin, merge_info=info)


# ------------------------------------------------------------------------------------------------
# R/cosineNorm.R:53-82
# ------------------------------------------------------------------------------------------------
def cosine_norm(x, mode="matrix"):
    """x [genes x cells].  mode in {'matrix','all','l2norm'} like the reference."""
    x = np.asarray(x, dtype=np.float64)
    l2 = np.sqrt(np.sum(x ** 2, axis=0))
    if mode == "l2norm":
        return l2
    mat = x / np.maximum(1e-8, l2)[None, :]
    return mat if mode == "matrix" else (mat, l2)


# ------------------------------------------------------------------------------------------------
# R/mnnCorrect.R:451-481  .compute_correction_vectors / .adjust_shift_variance
# ------------------------------------------------------------------------------------------------
def compute_correction_vectors(data1, data2, mnn1, mnn2, tdata2, sigma, smooth=None):
    """data1/data2 [cells x genes]; tdata2 [genes_for_dist x cells]; returns [cells2 x genes]."""
    smooth = smooth or capi.smooth_gaussian_kernel
    averaged, second = average_correction(data1, mnn1, data2, mnn2)  # == sumCountsAcrossCells(average=TRUE), :457
    cell_vect = smooth(np.asfortranarray(averaged.T), second - 1, tdata2, sigma)
    return np.ascontiguousarray(cell_vect.T)


def adjust_shift_variance(data1, data2, correction, sigma, subset_row=None, restrict1=None, restrict2=None, kernel=None):
    """data1/data2 [genes x cells]; correction [cells2 x genes]; restricts 1-based or None."""
    kernel = kernel or capi.adjust_shift_variance
    cell_vect = correction
    if subset_row is not None:
        sr = np.asarray(subset_row, dtype=np.int64) - 1
        cell_vect = cell_vect[:, sr]
        data1 = data1[sr, :]
        data2 = data2[sr, :]
    r1 = np.arange(data1.shape[1]) if restrict1 is None else np.asarray(restrict1, dtype=np.int64) - 1
    r2 = np.arange(data2.shape[1]) if restrict2 is None else np.asarray(restrict2, dtype=np.int64) - 1
    scaling = kernel(data1, data2, cell_vect, sigma, r1, r2)
    scaling = np.maximum(scaling, 1)  # pmax(scaling, 1): NaN propagates in R's pmax as in np.maximum
    return scaling[:, None] * correction


# R/mnnCorrect.R:179-393  .mnn_correct + .mnn_correct_core (svd.dim=0, same gene set in and out, predefined order)
def mnn_correct(batches, k=20, prop_k=None, sigma=0.1, cos_norm_in=True, cos_norm_out=True, var_adj=True,
                restrict=None, merge_order=None, knn=None, mutual=None, smooth=None, kernel=None, auto_merge=False):
    """batches: list of [genes x cells].  Returns dict(corrected [genes x cells], batch, merge_info)."""
    batches = [np.asarray(b, dtype=np.float64) for b in batches]
    in_b, out_b = list(batches), list(batches)
    same_set = True
    if cos_norm_in:
        norms = []
        for i, b in enumerate(in_b):
            m, l2 = cosine_norm(b, "all")
            in_b[i] = m
            norms.append(l2)
    if cos_norm_out:
        if not cos_norm_in:
            norms = [cosine_norm(b, "l2norm") for b in in_b]
        out_b = [b / np.maximum(1e-8, l2)[None, :] for b, l2 in zip(out_b, norms)]
    if cos_norm_out != cos_norm_in:
        same_set = False
    in_t = [b.T.copy() for b in in_b]
    out_t = [b.T.copy() for b in out_b]
    def add_out(t):
        if not isinstance(t, list):
            t.extras = [None if same_set else out_t[t.index[0] - 1]]
            return t
        return [add_out(t[0]), add_out(t[1])]

    def count_pairs(a, b):   # .count_mnn_pairs with orthogonalize=FALSE (R/mnnCorrect.R:212)
        return len(restricted_mnn(a.data, a.restrict, b.data, b.restrict, k, prop_k, mutual=mutual)[0])

    nb = len(batches)
    if auto_merge:   # R/mnnCorrect.R:211-223 + R/MNN_tree.R:154-168
        remainders = [add_out(Node([i + 1], in_t[i], None if restrict is None else restrict[i])) for i in range(nb)]
        pairwise = np.zeros((nb, nb), dtype=np.int64)
        for i in range(nb):
            for j in range(i):
                pairwise[i, j] = count_pairs(remainders[i], remainders[j])
        tree = None
    else:
        tree = add_out(create_tree_predefined(in_t, restrict, merge_order))
    nmerges = len(batches) - 1
    pairings, left_set, right_set = [], [], []
    for _ in range(nmerges):
        if auto_merge:
            cols, rows = np.nonzero(pairwise.T == pairwise.max())
            chosen = (int(rows[0]), int(cols[0]))
            left, right = remainders[chosen[0]], remainders[chosen[1]]
        else:
            left, right, path = get_next_merge(tree)
        ld, rd = left.data, right.data
        lx, rx = left.extras[0], right.extras[0]
        s1, s2 = restricted_mnn(ld, left.restrict, rd, right.restrict, k, prop_k, mutual=mutual)
        pairings.append((s1.astype(np.int64), s2.astype(np.int64)))
        left_set.append(list(left.index)); right_set.append(list(right.index))
        trans_right = np.asfortranarray(rd.T)
        cor_in = compute_correction_vectors(ld, rd, s1, s2, trans_right, sigma, smooth=smooth)
        if not same_set:
            cor_out = compute_correction_vectors(lx, rx, s1, s2, trans_right, sigma, smooth=smooth)
        if var_adj:
            cor_in = adjust_shift_variance(ld.T, rd.T, cor_in, sigma, restrict1=left.restrict, restrict2=right.restrict, kernel=kernel)
            if not same_set:
                cor_out = adjust_shift_variance(lx.T, rx.T, cor_out, sigma, restrict1=left.restrict, restrict2=right.restrict, kernel=kernel)
        rd = rd + cor_in
        if not same_set:
            rx = rx + cor_out
        node = Node(left.index + right.index, np.vstack([ld, rd]),
                    combine_restrict(ld.shape[0], left.restrict, rd.shape[0], right.restrict),
                    origin=np.concatenate([left.origin, right.origin]),
                    extras=[None if same_set else np.vstack([lx, rx])])
        if auto_merge:
            keep = [i for i in range(len(remainders)) if i not in chosen]
            remainders = [remainders[i] for i in keep]
            if remainders:
                old = pairwise[np.ix_(keep, keep)]
                new_stats = np.array([count_pairs(node, r) for r in remainders], dtype=np.int64)
                pairwise = np.hstack([np.vstack([old, new_stats[None, :]]), np.zeros((len(keep) + 1, 1), dtype=np.int64)])
                remainders.append(node)
            else:
                tree = node
        else:
            tree = update_tree(tree, path, node)
    full = tree.data if same_set else tree.extras[0]
    res = _finish(tree, full, pairings, left_set, right_set, {})
    res["corrected"] = res["corrected"].T
    return res


# ------------------------------------------------------------------------------------------------
# R/clusterMNN.R:267-312  sigma + .smooth_gaussian_from_centroids
# ------------------------------------------------------------------------------------------------
def smooth_gaussian_from_centroids(x, centers, sigma, delta):
    x = np.asarray(x, dtype=np.float64)
    weights = np.stack([-((x - centers[j]) ** 2).sum(axis=1) for j in range(centers.shape[0])], axis=1) / sigma ** 2
    top = weights.max(axis=1, keepdims=True)
    nw = np.exp(weights - top)
    nw /= nw.sum(axis=1, keepdims=True)
    out = x.copy()
    for j in range(centers.shape[0]):
        out = out + np.outer(nw[:, j], delta[j])
    return out


def propagate_to_cells(cells, centroids, corrected_centroids, restrict=None, knn=None):
    knn = knn or capi.query_knn
    q = cells if restrict is None else cells[np.asarray(restrict, dtype=np.int64) - 1]
    _, dist = knn(centroids, q, 1)
    sigma = float(np.median(dist[:, 0]))
    return smooth_gaussian_from_centroids(cells, centroids, sigma, corrected_centroids - centroids)


# ------------------------------------------------------------------------------------------------
# R/multiBatchPCA.R:211-322  .multi_pca_list with ExactParam (numpy SVD); rotation sign-normalised like the product
# ------------------------------------------------------------------------------------------------
def multi_batch_pca(batches, d=50, weights=None, get_variance=False):
    mats = [np.asarray(b, dtype=np.float64) for b in batches]
    ncells = np.array([m.shape[1] for m in mats], dtype=np.float64)
    w = np.ones(len(mats)) if weights is None else (ncells.copy() if weights is False else np.asarray(weights, dtype=np.float64))
    centers = sum(m.mean(axis=1) * wi for m, wi in zip(mats, w)) / w.sum()
    centred = [m - centers[:, None] for m in mats]
    scaled = np.hstack([c / math.sqrt(n / wi) for c, n, wi in zip(centred, ncells, w)])
    u, sv, _ = np.linalg.svd(scaled, full_matrices=False)
    u = u[:, :d]; sv = sv[:d]
    piv = np.argmax(np.abs(u), axis=0)
    u = u * np.sign(u[piv, np.arange(u.shape[1])])[None, :]
    out = {"pcs": [c.T @ u for c in centred], "rotation": u, "centers": centers}
    if get_variance:
        out["var_explained"] = sv ** 2 / len(mats)
        out["var_total"] = float((scaled ** 2).sum()) / len(mats)
    return out
