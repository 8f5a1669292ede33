/* r/b200_shim.c -- the .Call shim that binds batchelor's R code to libb200mnn.so.
 *
 * Drop-in for batchelor's src/ directory (LTLA/batchelor v1.23.1): it replaces src/find_mutual_nns.cpp,
 * src/smooth_gaussian_kernel.cpp, src/adjust_shift_variance.cpp, src/utils.cpp and src/RcppExports.cpp, registers the SAME
 * three routines with the SAME arities (src/RcppExports.cpp:48-58) so that R/RcppExports.R:4-14 keeps working byte for
 * byte, and adds three routines for the GPU BNPARAM backend (r/B200Param.R).  Needs only Rinternals.h (no Rcpp).
 *
 * Build (where R exists; this repository's image has no R, so this file is NOT compiled here):
 *     cp r/b200_shim.c r/Makevars <batchelor>/src/ ; cp r/B200Param.R <batchelor>/R/ ; cp include/b200mnn.h <batchelor>/src/
 *     B200MNN_HOME=/path/to/this/repo R CMD INSTALL <batchelor>
 *
 * Conventions: inputs are R-owned and read-only; every output is allocated here with R's allocator and written by the
 * library (no second copy); a non-zero status becomes an R error carrying the library's message, which for argument errors
 * is the reference's own std::runtime_error text (what END_RCPP would have raised).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "b200mnn.h"

static void check(int rc) { if (rc != B200MNN_OK) Rf_error("%s", b200mnn_last_error()); }

/* Rcpp::traits::input_parameter coerces when the SEXP type differs; so does the shim. */
static SEXP as_real(SEXP x) { return Rf_isReal(x) ? x : Rf_coerceVector(x, REALSXP); }
static SEXP as_int(SEXP x) { return Rf_isInteger(x) ? x : Rf_coerceVector(x, INTSXP); }

/* Shrinks a freshly allocated vector in place (R >= 3.4 growable vectors): the pair lists are written straight into their
 * final INTSXPs, whose capacity is the a-priori bound, without a second buffer. */
static void shrink(SEXP x, R_xlen_t n) {
    const R_xlen_t cap = XLENGTH(x);
    if (n < cap) { SET_TRUELENGTH(x, cap); SET_GROWABLE_BIT(x); SETLENGTH(x, n); }
}

/* src/RcppExports.cpp:25-33 -> src/find_mutual_nns.cpp:8-41 */
SEXP _batchelor_find_mutual_nns(SEXP left, SEXP right) {
    left = PROTECT(as_int(left)); right = PROTECT(as_int(right));
    const R_xlen_t n1 = Rf_nrows(left), k2 = Rf_ncols(left), n2 = Rf_nrows(right), k1 = Rf_ncols(right);
    const R_xlen_t cap = n1 * k2 > 0 ? n1 * k2 : 1;
    SEXP f = PROTECT(Rf_allocVector(INTSXP, cap)), s = PROTECT(Rf_allocVector(INTSXP, cap));
    int64_t np = 0;
    check(b200mnn_find_mutual_nns(INTEGER(left), n1, (int) k2, INTEGER(right), n2, (int) k1, INTEGER(f), INTEGER(s), &np));
    shrink(f, (R_xlen_t) np); shrink(s, (R_xlen_t) np);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));       /* unnamed list of two integer vectors, as Rcpp::List::create */
    SET_VECTOR_ELT(out, 0, f); SET_VECTOR_ELT(out, 1, s);
    UNPROTECT(5);
    return out;
}

/* src/RcppExports.cpp:36-46 -> src/smooth_gaussian_kernel.cpp:11-117 */
SEXP _batchelor_smooth_gaussian_kernel(SEXP averaged, SEXP index, SEXP mat, SEXP sigma2) {
    averaged = PROTECT(as_real(averaged)); mat = PROTECT(as_real(mat)); index = PROTECT(as_int(index));
    const R_xlen_t G = Rf_nrows(averaged), nmnn = Rf_ncols(averaged), Gd = Rf_nrows(mat), nc = Rf_ncols(mat);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int) G, (int) nc));
    check(b200mnn_smooth_gaussian_kernel(REAL(averaged), G, nmnn, INTEGER(index), XLENGTH(index), REAL(mat), Gd, nc, Rf_asReal(sigma2), REAL(out)));
    UNPROTECT(4);
    return out;
}

/* src/RcppExports.cpp:10-22 -> src/adjust_shift_variance.cpp:30-164 */
SEXP _batchelor_adjust_shift_variance(SEXP data1, SEXP data2, SEXP vect, SEXP sigma2, SEXP restrict1, SEXP restrict2) {
    data1 = PROTECT(as_real(data1)); data2 = PROTECT(as_real(data2)); vect = PROTECT(as_real(vect));
    restrict1 = PROTECT(as_int(restrict1)); restrict2 = PROTECT(as_int(restrict2));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_ncols(data2)));
    check(b200mnn_adjust_shift_variance(REAL(data1), Rf_nrows(data1), Rf_ncols(data1), REAL(data2), Rf_nrows(data2), Rf_ncols(data2), REAL(vect),
                                        Rf_nrows(vect), Rf_ncols(vect), Rf_asReal(sigma2), INTEGER(restrict1), XLENGTH(restrict1),
                                        INTEGER(restrict2), XLENGTH(restrict2), REAL(out)));
    UNPROTECT(6);
    return out;
}

/* ---- GPU BNPARAM backend (r/B200Param.R) ---- */

/* queryKNN(X, query, k) for BNPARAM = B200Param(): list(index, distance), R layout, 1-based ids */
SEXP _batchelor_b200_query_knn(SEXP X, SEXP query, SEXP k, SEXP get_distance) {
    X = PROTECT(as_real(X)); query = PROTECT(as_real(query));
    const R_xlen_t n = Rf_nrows(X), d = Rf_ncols(X), nq = Rf_nrows(query);
    const int kk = Rf_asInteger(k), want = Rf_asLogical(get_distance);
    SEXP idx = PROTECT(Rf_allocMatrix(INTSXP, (int) nq, kk));
    SEXP dist = PROTECT(want ? Rf_allocMatrix(REALSXP, (int) nq, kk) : R_NilValue);
    check(b200mnn_query_knn(REAL(X), n, REAL(query), nq, (int) d, kk, /*col_major=*/1, INTEGER(idx), want ? REAL(dist) : NULL));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, idx); SET_VECTOR_ELT(out, 1, dist);
    UNPROTECT(5);
    return out;
}

/* findMutualNN(data1, data2, k1, k2): both searches and the pair extraction in one call */
SEXP _batchelor_b200_find_mutual_nn(SEXP data1, SEXP data2, SEXP k1, SEXP k2) {
    data1 = PROTECT(as_real(data1)); data2 = PROTECT(as_real(data2));
    const R_xlen_t n1 = Rf_nrows(data1), n2 = Rf_nrows(data2), d = Rf_ncols(data1);
    const int kk1 = Rf_asInteger(k1), kk2 = Rf_asInteger(k2);
    R_xlen_t cap = n1 * (kk2 < n2 ? kk2 : n2);
    if (cap < 1) cap = 1;
    SEXP f = PROTECT(Rf_allocVector(INTSXP, cap)), s = PROTECT(Rf_allocVector(INTSXP, cap));
    int64_t np = 0;
    check(b200mnn_find_mutual_nn(REAL(data1), n1, REAL(data2), n2, (int) d, kk1, kk2, 1, INTEGER(f), INTEGER(s), cap, &np));
    shrink(f, (R_xlen_t) np); shrink(s, (R_xlen_t) np);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, f); SET_VECTOR_ELT(out, 1, s);
    UNPROTECT(5);
    return out;
}

/* The whole merge loop of reducedMNN / fastMNN (.fast_mnn_core, R/fastMNN.R:436-562) as ONE call: every batch crosses
 * PCIe once, the tree nodes stay in HBM between merges, all visible GPUs are driven by the library.
 *   batches      list of numeric matrices [cells x d]
 *   merge_left / merge_right   integer vectors (0-based node ids: batches are 0..nb-1, merge m creates nb+m), or NULL for
 *                              auto.merge = TRUE
 *   restrict     list (NULL entries allowed) of 1-based integer row ids, or NULL
 * Returns list(corrected [ncells x d] in final node order, node_order, node_ncells, pairs = list of list(left, right)
 * (1-based rows within the merged nodes), batch_size, skipped, lost_var [(nb-1) x nb], merge_left, merge_right). */
SEXP _batchelor_b200_reduced_mnn(SEXP batches, SEXP merge_left, SEXP merge_right, SEXP k, SEXP prop_k, SEXP ndist, SEXP min_batch_skip,
                                 SEXP restrict_, SEXP get_variance) {
    const int nb = (int) XLENGTH(batches);
    if (nb < 1) Rf_error("at least one batch must be supplied");
    const double** ptr = (const double**) R_alloc(nb, sizeof(double*));
    int64_t* ncells = (int64_t*) R_alloc(nb, sizeof(int64_t));
    const int32_t** rptr = (const int32_t**) R_alloc(nb, sizeof(int32_t*));
    int64_t* rn = (int64_t*) R_alloc(nb, sizeof(int64_t));
    int nprot = 0, d = -1;
    for (int b = 0; b < nb; ++b) {
        SEXP m = PROTECT(as_real(VECTOR_ELT(batches, b))); ++nprot;
        if (d < 0) d = Rf_ncols(m);
        if (Rf_ncols(m) != d) Rf_error("number of columns is not the same across batches");
        ptr[b] = REAL(m); ncells[b] = Rf_nrows(m);
        rptr[b] = NULL; rn[b] = 0;
        if (!Rf_isNull(restrict_) && !Rf_isNull(VECTOR_ELT(restrict_, b))) {
            SEXP r = PROTECT(as_int(VECTOR_ELT(restrict_, b))); ++nprot;
            rptr[b] = (const int32_t*) INTEGER(r); rn[b] = XLENGTH(r);
        }
    }
    const int automerge = Rf_isNull(merge_left) || Rf_isNull(merge_right);
    b200mnn_merge_result* res = NULL;
    check(b200mnn_reduced_mnn(ptr, ncells, nb, d, /*col_major=*/1, automerge ? NULL : (const int32_t*) INTEGER(merge_left),
                              automerge ? NULL : (const int32_t*) INTEGER(merge_right), Rf_asInteger(k),
                              Rf_isNull(prop_k) ? -1.0 : Rf_asReal(prop_k), Rf_asReal(ndist), Rf_asReal(min_batch_skip) /* NA_real_ is a NaN */,
                              Rf_isNull(restrict_) ? NULL : rptr, rn, Rf_asLogical(get_variance), &res));
    const int nm = nb - 1;
    const int64_t ntotal = b200mnn_result_ncells(res);
    SEXP corrected = PROTECT(Rf_allocMatrix(REALSXP, (int) ntotal, d)); ++nprot;
    SEXP order = PROTECT(Rf_allocVector(INTSXP, nb)); ++nprot;
    SEXP bs = PROTECT(Rf_allocVector(REALSXP, nm)); ++nprot;
    SEXP sk = PROTECT(Rf_allocVector(LGLSXP, nm)); ++nprot;
    SEXP lv = PROTECT(Rf_allocMatrix(REALSXP, nm, nb)); ++nprot;
    SEXP ml = PROTECT(Rf_allocVector(INTSXP, nm)); ++nprot;
    SEXP mr = PROTECT(Rf_allocVector(INTSXP, nm)); ++nprot;
    SEXP cnt = PROTECT(Rf_allocVector(REALSXP, nb)); ++nprot;
    SEXP pairs = PROTECT(Rf_allocVector(VECSXP, nm)); ++nprot;
    int rc = b200mnn_result_corrected(res, REAL(corrected), 1);
    int64_t* counts = (int64_t*) R_alloc(nb, sizeof(int64_t));
    double* lv_rm = (double*) R_alloc((size_t) (nm > 0 ? nm : 1) * nb, sizeof(double));
    if (!rc) rc = b200mnn_result_info(res, INTEGER(order), counts, REAL(bs), LOGICAL(sk), lv_rm);
    if (!rc && nm > 0) rc = b200mnn_result_merges(res, INTEGER(ml), INTEGER(mr));
    for (int b = 0; b < nb; ++b) REAL(cnt)[b] = (double) counts[b];
    for (int m = 0; m < nm; ++m) for (int b = 0; b < nb; ++b) REAL(lv)[m + (size_t) nm * b] = lv_rm[(size_t) m * nb + b];   /* row-major -> R */
    for (int m = 0; m < nm && !rc; ++m) {
        const int64_t np = b200mnn_result_npairs(res, m);
        SEXP l = PROTECT(Rf_allocVector(INTSXP, np)), r = PROTECT(Rf_allocVector(INTSXP, np));
        rc = b200mnn_result_pairs(res, m, INTEGER(l), INTEGER(r));
        SEXP p = PROTECT(Rf_allocVector(VECSXP, 2));
        SET_VECTOR_ELT(p, 0, l); SET_VECTOR_ELT(p, 1, r);
        SET_VECTOR_ELT(pairs, m, p);
        UNPROTECT(3);
    }
    b200mnn_result_free(res);
    check(rc);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 10)); ++nprot;
    SET_VECTOR_ELT(out, 0, corrected); SET_VECTOR_ELT(out, 1, order); SET_VECTOR_ELT(out, 2, cnt); SET_VECTOR_ELT(out, 3, pairs);
    SET_VECTOR_ELT(out, 4, bs); SET_VECTOR_ELT(out, 5, sk); SET_VECTOR_ELT(out, 6, lv); SET_VECTOR_ELT(out, 7, ml); SET_VECTOR_ELT(out, 8, mr);
    SET_VECTOR_ELT(out, 9, Rf_ScalarInteger(b200mnn_device_count()));
    UNPROTECT(nprot);
    return out;
}

static const R_CallMethodDef CallEntries[] = {            /* src/RcppExports.cpp:48-53: same names, same arities */
    {"_batchelor_adjust_shift_variance", (DL_FUNC) &_batchelor_adjust_shift_variance, 6},
    {"_batchelor_find_mutual_nns", (DL_FUNC) &_batchelor_find_mutual_nns, 2},
    {"_batchelor_smooth_gaussian_kernel", (DL_FUNC) &_batchelor_smooth_gaussian_kernel, 4},
    {"_batchelor_b200_query_knn", (DL_FUNC) &_batchelor_b200_query_knn, 4},
    {"_batchelor_b200_find_mutual_nn", (DL_FUNC) &_batchelor_b200_find_mutual_nn, 4},
    {"_batchelor_b200_reduced_mnn", (DL_FUNC) &_batchelor_b200_reduced_mnn, 9},
    {NULL, NULL, 0}
};

void R_init_batchelor(DllInfo* dll) {                     /* src/RcppExports.cpp:55-58 */
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
