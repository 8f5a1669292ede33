# r/B200Param.R -- the GPU BNPARAM backend of batchelor's MNN path (goes into batchelor's R/ directory next to the shim).
#
# NOT executed in this repository (the image has no R).  Written against the S4 surface batchelor actually uses from
# BiocNeighbors (NAMESPACE:56-58): KmknnParam, queryKNN, findMutualNN.  The methods are registered for B200Param ONLY --
# no method is defined for signature ANY, so every other backend keeps BiocNeighbors' own dispatch.

#' @export
setClass("B200Param", contains = "BiocNeighborParam", slots = c(device = "integer"))

#' Exact Euclidean search on a B200 (same results as KmknnParam())
#' @export
B200Param <- function(device = 0L) new("B200Param", distance = "Euclidean", device = as.integer(device))

.b200_serial <- function(BPPARAM) {
    # a CUDA context does not survive MulticoreParam forks; the library shards over the visible GPUs itself
    if (!is(BPPARAM, "SerialParam")) stop("B200Param() needs BPPARAM=SerialParam(): the GPU backend shards across GPUs itself")
}

.b200_query_knn <- function(X, query, k, get.index = TRUE, get.distance = TRUE, BPPARAM = SerialParam(), ...) {
    .b200_serial(BPPARAM)
    X <- as.matrix(X); query <- as.matrix(query)
    if (k > nrow(X)) { warning("'k' capped at the number of observations"); k <- nrow(X) }
    out <- .Call("_batchelor_b200_query_knn", X, query, as.integer(k), isTRUE(get.distance), PACKAGE = "batchelor")
    names(out) <- c("index", "distance")
    if (!isTRUE(get.index)) out$index <- NULL
    if (!isTRUE(get.distance)) out$distance <- NULL
    out
}

# BiocNeighbors 1.x: the generics dispatch on (BNINDEX, BNPARAM); a method for (missing, B200Param) is all that is needed.
#' @export
setMethod("queryKNN", c("missing", "B200Param"), function(X, query, k, ..., BNINDEX, BNPARAM) .b200_query_knn(X, query, k, ...))

#' @export
setMethod("findKNN", c("missing", "B200Param"), function(X, k, ..., BNINDEX, BNPARAM) {
    # queryKNN of X against itself with k + 1 neighbours, the cell itself dropped (BiocNeighbors' findKNN contract)
    out <- .b200_query_knn(X, X, k + 1L, ...)
    self <- out$index == row(out$index)
    drop1 <- function(m) t(vapply(seq_len(nrow(m)), function(i) { j <- which(self[i, ])[1]; if (is.na(j)) j <- ncol(m); m[i, -j] }, m[1, -1]))
    if (!is.null(out$index)) idx <- drop1(out$index)
    if (!is.null(out$distance)) out$distance <- drop1(out$distance)
    if (!is.null(out$index)) out$index <- idx
    out
})
# BiocNeighbors >= 2.0 routes every search through defineBuilder(BNPARAM) -> a knncolle::Builder external pointer; the
# equivalent there is a ~40-line C++ Builder whose Prebuilt::search() batches its queries into b200mnn_query_knn.

# findMutualNN(): both searches and the pair extraction in ONE library call (R/MNN_tree.R:129 calls this through `...`)
findMutualNN.b200 <- function(data1, data2, k1, k2 = k1, BNPARAM, BPPARAM = SerialParam()) {
    .b200_serial(BPPARAM)
    out <- .Call("_batchelor_b200_find_mutual_nn", as.matrix(data1), as.matrix(data2), as.integer(k1), as.integer(k2), PACKAGE = "batchelor")
    list(first = out[[1]], second = out[[2]])
}
# ... and the one line of R/MNN_tree.R:129 that selects it:
#   FUN <- if (is(list(...)$BNPARAM, "B200Param")) findMutualNN.b200 else findMutualNN
#   pairs <- FUN(left.data, right.data, k1 = k1, k2 = k2, ...)

# reducedMNN()/fastMNN() post-PCA path as ONE call: replaces the body of .fast_mnn_core (R/fastMNN.R:436-562) when
# BNPARAM is a B200Param.  The merge ORDER stays R control flow: the tree is walked here exactly as .get_next_merge does
# (R/MNN_tree.R:61-69) and handed over as node ids.
.fast_mnn_core.b200 <- function(batches, k = 20, prop.k = NULL, restrict = NULL, ndist = 3, merge.order = NULL, auto.merge = FALSE,
                                min.batch.skip = 0, get.variance = TRUE) {
    nb <- length(batches)
    left <- right <- NULL
    sets <- list(left = list(), right = list())
    if (!auto.merge && nb > 1L) {
        tree <- .create_tree_predefined(lapply(seq_len(nb), function(i) i - 1L), NULL, merge.order)   # leaves carry node ids
        left <- right <- integer(nb - 1L)
        for (m in seq_len(nb - 1L)) {
            nxt <- .get_next_merge(tree)
            left[m] <- .get_node_data(nxt$left); right[m] <- .get_node_data(nxt$right)
            tree <- .update_tree(tree, nxt$chosen, data = nb + m - 1L, index = c(.get_node_index(nxt$left), .get_node_index(nxt$right)),
                                 restrict = NULL, origin = NULL, extras = list())
        }
    }
    out <- .Call("_batchelor_b200_reduced_mnn", lapply(batches, as.matrix), left, right, as.integer(k), prop.k, as.numeric(ndist),
                 as.numeric(min.batch.skip), restrict, isTRUE(get.variance), PACKAGE = "batchelor")
    names(out) <- c("corrected", "node.order", "node.ncells", "pairs", "batch.size", "skipped", "lost.var", "merge.left", "merge.right", "devices")
    out   # the caller reorders rows to input batch order and shifts the pair ids (R/fastMNN.R:533-547), as before
}
