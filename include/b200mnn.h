/* b200mnn.h -- C ABI of the B200-native (sm_100a) MNN hot path of batchelor.
 *
 * This shared library (batchelor_b200/lib/libb200mnn.so) is the drop-in boundary for the reference's native
 * path.  Every entry point cites the reference interface it replaces (paths under LTLA/batchelor v1.23.1):
 *
 *   reference interface                                             replaced by
 *   --------------------------------------------------------------  ---------------------------------
 *   .Call _batchelor_find_mutual_nns   (src/RcppExports.cpp:25-33,   b200mnn_find_mutual_nns
 *          src/find_mutual_nns.cpp:8)
 *   .Call _batchelor_smooth_gaussian_kernel (src/RcppExports.cpp:36-46, b200mnn_smooth_gaussian_kernel
 *          src/smooth_gaussian_kernel.cpp:11)
 *   .Call _batchelor_adjust_shift_variance  (src/RcppExports.cpp:10-22, b200mnn_adjust_shift_variance
 *          src/adjust_shift_variance.cpp:30)
 *   BiocNeighbors::queryKNN(X, query, k, BNPARAM=KmknnParam())       b200mnn_query_knn
 *          (call sites R/fastMNN.R:605, R/clusterMNN.R:276)
 *   BiocNeighbors::findMutualNN(data1, data2, k1, k2, BNPARAM)       b200mnn_find_mutual_nn
 *          (call site R/MNN_tree.R:129; re-export R/findMutualNN.R:1-3)
 *   .average_correction            (R/fastMNN.R:567-580)             b200mnn_average_correction
 *   .center_along_batch_vector     (R/fastMNN.R:626-640)             b200mnn_center_along_batch_vector
 *   .tricube_weighted_correction   (R/fastMNN.R:599-608,             b200mnn_tricube_weighted_correction
 *          R/utils_tricube.R:1-27)
 *   cosineNorm / .apply_cosine_norm (R/cosineNorm.R:53-82)           b200mnn_cosine_norm
 *
 * Conventions
 *   - Plain pointers and sizes only.  Host-buffer entry points (no `_dev_` in the name) take R-layout host
 *     memory: column-major double / int32, neighbour and pair indices 1-BASED exactly where R's are, restrict /
 *     `index` arguments 0-BASED exactly where the reference's C++ takes them 0-based.  Inputs are read-only;
 *     outputs are caller-allocated (the R shim allocates them with Rf_allocMatrix / Rf_allocVector).
 *   - Every function returns 0 on success and a non-zero B200MNN_E* code on failure; b200mnn_last_error()
 *     returns the message (thread-local).  The messages for argument errors are the reference's own
 *     std::runtime_error strings so that an R shim can forward them verbatim with Rf_error().
 *   - There is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     B200MNN_ECUDA.
 *   - `_dev_` entry points take DEVICE pointers and a cudaStream_t (as void*); matrices are row-major
 *     [cells x dims] double (one cell contiguous), indices int32 0-based.  They never synchronise the host
 *     unless stated.  They are what the Python host mirror and bench.py drive (device memory owned by torch).
 */
#ifndef B200MNN_H
#define B200MNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MNN_OK 0
#define B200MNN_EINVAL 1   /* bad argument (message = the reference's error string where one exists) */
#define B200MNN_ECUDA 2    /* CUDA runtime/driver failure, or no device */
#define B200MNN_ENOMEM 3
#define B200MNN_ECAPACITY 4 /* caller-provided output capacity too small */

const char* b200mnn_last_error(void);
int b200mnn_version(void);
/* Kernels of this library launched so far by the calling process (every launch site counts itself). */
int64_t b200mnn_launch_count(void);
/* Number of visible CUDA devices (0 if none / no driver). */
int b200mnn_device_count(void);
/* Select the device used by this thread's subsequent calls (cudaSetDevice). */
int b200mnn_set_device(int device);

/* ------------------------------------------------------------------------------------------------------------
 * Host-buffer entry points (R layout)
 * ------------------------------------------------------------------------------------------------------------ */

/* queryKNN(X, query, k): exact Euclidean k nearest rows of X for every row of Q, ties broken by (distance, index).
 * X [n x d], Q [nq x d]; col_major != 0: R layout, else row-major.  k <= n required (R caller caps with a warning).
 * idx_out [nq x k] int32 1-based, dist_out [nq x k] double (may be NULL); both in the same major as the inputs. */
int b200mnn_query_knn(const double* X, int64_t n, const double* Q, int64_t nq, int d, int k, int col_major,
                      int32_t* idx_out, double* dist_out);

/* findMutualNN(data1, data2, k1, k2): W21 = kNN of data1 rows in data2 (k2 each), W12 = kNN of data2 rows in data1
 * (k1 each), then the mutual pairs in the order of src/find_mutual_nns.cpp:23-37.
 * first_out/second_out: capacity entries each (n1*min(k2,n2) always suffices); *np_out = number of pairs. 1-based.
 * From 262 144 rows per batch the uploads are pipelined: data1 crosses PCIe first, data2 follows in row chunks that are
 * searched against data1 as they land (same result; INTEGRATION.md section 6 lists the switches). */
int b200mnn_find_mutual_nn(const double* data1, int64_t n1, const double* data2, int64_t n2, int d, int k1, int k2,
                           int col_major, int32_t* first_out, int32_t* second_out, int64_t capacity, int64_t* np_out);

/* find_mutual_nns(left, right), src/find_mutual_nns.cpp:8-41.  left [n1 x k2], right [n2 x k1] int32 column-major,
 * 1-based.  first_out/second_out must hold n1*k2 entries; *np_out = number of pairs. */
int b200mnn_find_mutual_nns(const int32_t* left, int64_t n1, int k2, const int32_t* right, int64_t n2, int k1,
                            int32_t* first_out, int32_t* second_out, int64_t* np_out);

/* smooth_gaussian_kernel(averaged, index, mat, sigma2), src/smooth_gaussian_kernel.cpp:11-117.
 * averaged [G x nmnn], index0 int32[nindex] 0-based columns of mat, mat [Gdist x ncells], out [G x ncells];
 * column-major.  Fails with the reference's message if nindex != nmnn (:18-20). */
int b200mnn_smooth_gaussian_kernel(const double* averaged, int64_t G, int64_t nmnn, const int32_t* index0, int64_t nindex,
                                   const double* mat, int64_t Gdist, int64_t ncells, double sigma2, double* out);

/* adjust_shift_variance(data1, data2, vect, sigma2, restrict1, restrict2), src/adjust_shift_variance.cpp:30-164.
 * data1 [G1 x n1], data2 [G2 x n2], vect [vrows x vcols] column-major; restricts 0-based; out double[n2].
 * Fails with the reference's messages on G1!=G2||G1!=vcols (:34-36), n2!=vrows (:39-41), restrict out of range. */
int b200mnn_adjust_shift_variance(const double* data1, int64_t G1, int64_t n1, const double* data2, int64_t G2, int64_t n2,
                                  const double* vect, int64_t vrows, int64_t vcols, double sigma2,
                                  const int32_t* restrict1, int64_t nr1, const int32_t* restrict2, int64_t nr2, double* out);

/* cosineNorm, R/cosineNorm.R:53-82.  x [G x n] column-major (cells in columns); out [G x n] (may be NULL),
 * l2_out double[n] (may be NULL). */
int b200mnn_cosine_norm(const double* x, int64_t G, int64_t n, double* out, double* l2_out);

/* .average_correction(refdata, mnn1, curdata, mnn2), R/fastMNN.R:567-580.  refdata [n1 x d], curdata [n2 x d]
 * column-major, mnn1/mnn2 int32[np] 1-based.  averaged_out [n2 x d] column-major capacity (only the first
 * *nmnn_out rows... see below), second_out int32[min(np,n2)] ascending unique mnn2; *nmnn_out = their count.
 * averaged_out is written as a dense column-major [nmnn x d] matrix (leading dimension nmnn). */
int b200mnn_average_correction(const double* refdata, int64_t n1, const double* curdata, int64_t n2, int d,
                               const int32_t* mnn1, const int32_t* mnn2, int64_t np,
                               double* averaged_out, int32_t* second_out, int64_t* nmnn_out);

/* .center_along_batch_vector(mat, batch.vec, restrict), R/fastMNN.R:626-640.  mat [n x d] column-major,
 * restrict int32[nrestrict] 1-based or NULL; out [n x d]. */
int b200mnn_center_along_batch_vector(const double* mat, int64_t n, int d, const double* batch_vec,
                                      const int32_t* restrict1, int64_t nrestrict, double* out);

/* .tricube_weighted_correction(curdata, correction, in.mnn, k, ndist), R/fastMNN.R:599-608.  curdata [n x d],
 * correction [nmnn x d] column-major, in_mnn int32[nmnn] 1-based rows of curdata; out [n x d]. */
int b200mnn_tricube_weighted_correction(const double* curdata, int64_t n, int d, const double* correction,
                                        const int32_t* in_mnn, int64_t nmnn, int k, double ndist, double* out);

/* ------------------------------------------------------------------------------------------------------------
 * The whole merge loop as one call: reducedMNN / .fast_mnn_core, R/fastMNN.R:436-562
 * ------------------------------------------------------------------------------------------------------------ */

/* Opaque result of b200mnn_reduced_mnn: the final node stays in HBM until it is fetched. */
typedef struct b200mnn_merge_result b200mnn_merge_result;

/* batches[b]: host matrix of batch b, [ncells[b] x d] (col_major != 0: R layout).  The merge ORDER is host control flow
 * (the R side walks its MNN_treenode tree, R/MNN_tree.R:61-109) and arrives flattened: leaves are nodes 0..nb-1, merge m
 * (0-based) merges nodes merge_left[m] (reference side, "left") and merge_right[m] (corrected side) into node nb + m.
 * merge_left == NULL: auto.merge = TRUE -- the order is searched on the device as R/MNN_tree.R:154-226 does (MNN pair
 * counts of every pair of remaining nodes, largest first; b200mnn_result_merges reports what was chosen).
 * k, prop_k (< 0 or NaN: NULL), ndist, min_batch_skip (NaN: never test) as in R/fastMNN.R:283-287; restrict1[b]: 1-based
 * rows of batch b or NULL (restrict1 itself may be NULL); get_variance != 0 fills lost_var.
 * Every visible CUDA device that can peer with the current one is used (query rows sharded, reference rows replicated
 * over NVLink); B200MNN_DEVICES=n caps the number.  Error strings are the reference's where it has one. */
int b200mnn_reduced_mnn(const double* const* batches, const int64_t* ncells, int nb, int d, int col_major, const int32_t* merge_left,
                        const int32_t* merge_right, int k, double prop_k, double ndist, double min_batch_skip,
                        const int32_t* const* restrict1, const int64_t* nrestrict, int get_variance, b200mnn_merge_result** result_out);
/* Total number of cells / number of MNN pairs of merge m (-1: bad argument). */
int64_t b200mnn_result_ncells(const b200mnn_merge_result* res);
int64_t b200mnn_result_npairs(const b200mnn_merge_result* res, int merge);
/* Pairs of merge m: 1-based ROW NUMBERS WITHIN the left and the right node of that merge, in the order of
 * src/find_mutual_nns.cpp:23-37 (the R side shifts them to output positions, R/fastMNN.R:533-538). */
int b200mnn_result_pairs(const b200mnn_merge_result* res, int merge, int32_t* left_out, int32_t* right_out);
/* Corrected coordinates [ntotal x d] in the row order of the final node (see node_order below). */
int b200mnn_result_corrected(const b200mnn_merge_result* res, double* out, int col_major);
/* node_order int32[nb]: batches (1-based) in the row order of the final node, node_ncells int64[nb]: their row counts;
 * batch_size double[nb-1], skipped int32[nb-1], lost_var double[(nb-1) x nb] row-major (metadata(out)$merge.info,
 * R/fastMNN.R:549-560).  Any pointer may be NULL. */
int b200mnn_result_info(const b200mnn_merge_result* res, int32_t* node_order, int64_t* node_ncells, double* batch_size, int32_t* skipped,
                        double* lost_var);
/* Node ids (see above) merged at every step: the given order, or the one the auto-merge search chose.  int32[nb-1] each. */
int b200mnn_result_merges(const b200mnn_merge_result* res, int32_t* left_out, int32_t* right_out);
void b200mnn_result_free(b200mnn_merge_result* res);

/* ------------------------------------------------------------------------------------------------------------
 * Device-pointer entry points (row-major [cells x dims] double, int32 0-based ids, cudaStream_t as void*)
 * ------------------------------------------------------------------------------------------------------------ */

/* Exact kNN.  d_idx [nq x k] int32 0-based row-major; d_dist [nq x k] double or NULL.  Asynchronous on `stream`.
 * Path: (large searches) k-means grouping of both sides + rigorous per-cluster lower bounds, then fp16 tcgen05 candidate
 * scoring of the clusters that cannot be excluded -> exact fp64 re-rank + certificate -> three-term re-score and finally
 * an exact fp64 rescue of any uncertified query.  `d_stats` (may be NULL) receives int64[8]: {queries rescued,
 * candidate lists per query, path (2 = tensor + cluster pruning, 1 = tensor, 0 = generic), queries re-scored by the
 * second tier, 128x128 score tiles computed by tier 1, by tier 2, tiles of a dense scan, reserved}. */
int b200mnn_dev_query_knn(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k,
                          int32_t* d_idx, double* d_dist, int64_t* d_stats, void* stream);

/* Mutual pairs from two device index matrices (0-based, row-major): left [n1 x k2] ids into batch 2,
 * right [n2 x k1] ids into batch 1.  d_first/d_second: `capacity` entries (n1*k2 always suffices), 0-based;
 * d_np: device int64 receiving the pair count.  Pairs are emitted in the reference's order.  Asynchronous. */
int b200mnn_dev_find_mutual_nns(const int32_t* d_left, int64_t n1, int k2, const int32_t* d_right, int64_t n2, int k1,
                                int32_t* d_first, int32_t* d_second, int64_t capacity, int64_t* d_np, void* stream);

/* Group-average of pair differences (a4).  d_first/d_second int32[np] 0-based (second need not be sorted).
 * d_averaged [n2 x d] capacity, written densely as [nmnn x d]; d_second_unique int32[n2] capacity; d_nmnn device
 * int64.  Asynchronous. */
int b200mnn_dev_average_correction(const double* d_ref, int64_t n1, const double* d_cur, int64_t n2, int d,
                                   const int32_t* d_first, const int32_t* d_second, int64_t np,
                                   double* d_averaged, int32_t* d_second_unique, int64_t* d_nmnn, void* stream);

/* In-place centring along a unit-normalised batch vector (a8); d_restrict int32 0-based or NULL. */
int b200mnn_dev_center_along_batch_vector(double* d_mat, int64_t n, int d, const double* d_batch_vec,
                                          const int32_t* d_restrict, int64_t nrestrict, void* stream);

/* Tricube smoothing given the kNN result (a6, R/utils_tricube.R:1-27): d_out = d_cur + weighted mean of
 * d_correction rows d_idx[i, :] (0-based) with tricube weights from d_dist. */
int b200mnn_dev_tricube_apply(const double* d_cur, int64_t n, int d, const double* d_correction, int64_t nmnn,
                              const int32_t* d_idx, const double* d_dist, int k, double ndist, double* d_out, void* stream);

/* Gaussian smoothing (a5) on device: d_averaged [nmnn x G] row-major (one MNN cell contiguous), d_index0 int32[nmnn],
 * d_mat [ncells x Gdist] row-major, d_out [ncells x G].  Large problems run on the tensor cores (split-fp16 tcgen05 GEMMs,
 * <= 1e-5 relative, verified per call against an fp64 re-evaluation of sampled rows; synchronises the stream), small ones
 * in fp64 (<= 1e-10); B200MNN_SMOOTH=fp64|tensor forces one. */
int b200mnn_dev_smooth_gaussian_kernel(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0,
                                       const double* d_mat, int64_t Gdist, int64_t ncells, double sigma2, double* d_out,
                                       void* stream);

/* Which path the last smoothing call of this process took (1 = tensor cores, accepted by the fp64 sample check;
 * 2 = fp64; 3 = tensor result rejected by the check and redone in fp64) and the check's figures: largest relative
 * deviation of a sampled output row, largest absolute deviation of a sampled log-density. */
int b200mnn_smooth_last_check(int* path, double* row_err, double* dens_err);

/* Shift variance (a7) on device: d_data1 [n1 x G], d_data2 [n2 x G], d_vect [n2 x G] row-major; restricts 0-based
 * device int32; d_out double[n2].  Synchronises the stream (restrict validation, operand scale). */
int b200mnn_dev_adjust_shift_variance(const double* d_data1, int64_t n1, const double* d_data2, int64_t n2, int64_t G,
                                      const double* d_vect, double sigma2, const int32_t* d_r1, int64_t nr1,
                                      const int32_t* d_r2, int64_t nr2, double* d_out, void* stream);

/* clusterMNN's propagation of centroid corrections to cells, .smooth_gaussian_from_centroids (R/clusterMNN.R:289-312):
 * d_out[c,] = d_x[c,] + sum_j softmax_j(-||x_c - centre_j||^2 / sigma^2) delta[j,].  d_x, d_out [n x d], d_centers, d_delta
 * [nc x d], row-major device matrices (the small-K variant of the Gaussian smoothing). */
int b200mnn_dev_smooth_gaussian_from_centroids(const double* d_x, int64_t n, int d, const double* d_centers, const double* d_delta, int nc,
                                               double sigma, double* d_out, void* stream);

/* Cosine normalisation on device: d_x [n x G] row-major (one cell contiguous); d_out may alias d_x or be NULL. */
int b200mnn_dev_cosine_norm(const double* d_x, int64_t n, int64_t G, double* d_out, double* d_l2, void* stream);

/* Transposes between R's column-major [rows x cols] and row-major on device (used by the host-buffer layer). */
int b200mnn_dev_transpose_f64(const double* d_in, int64_t rows, int64_t cols, double* d_out, void* stream);

/* Debug/validation hook for tests: runs only the tensor-core scoring stage and returns, for every query, the
 * retained candidate ids [nq x ncand] (0-based, -1 = empty), their approximate squared distances
 * (||q||^2 + score, unscaled) and the per-query threshold; *ncand_out = candidates per query.  Synchronous. */
int b200mnn_dev_debug_candidates(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k,
                                 int32_t* d_cand_idx, double* d_cand_d2, double* d_thr, int64_t cand_capacity,
                                 int64_t* ncand_out, void* stream);

/* Debug/validation hook for tests: out[M x ldo] (fp32, ldo a multiple of 256 and >= N rounded up to 256) =
 * A[M x K] . B[N x K]^T (fp64 row-major device matrices) through the split-fp16 tcgen05 GEMM that the gene-space kernels
 * (Gaussian smoothing, wide-d kNN) are built on.  terms: 3 = Ah.Bh + Al.Bh + Ah.Bl, 1 = Ah.Bh; chunk_boxes: K boxes (64
 * columns) per TMEM accumulation chain (<= 0: one chain).  Synchronises once (operand scale). */
int b200mnn_dev_debug_gemm(const double* dA, int64_t M, const double* dB, int64_t N, int64_t K, int terms, int chunk_boxes,
                           float* d_out, int64_t ldo, void* stream);

/* Measurement hook (bench.py's roofline figure): while enabled, CUDA events are recorded on the launch stream around
 * every launch of the dominant kernel (the tcgen05 candidate-scoring kernel).  b200mnn_profile_collect synchronises
 * those events and returns their summed duration, the number of launches and the ALGORITHMIC flops (2*nq*n*d per
 * launch) they covered.  b200mnn_profile_enable(x) also clears earlier records. */
int b200mnn_profile_enable(int on);
int b200mnn_profile_collect(double* total_ms, int64_t* launches, double* algorithmic_flops);
/* Tensor-core flops actually executed by the profiled launches (cluster pruning skips most score tiles; the second
 * scoring tier adds some): issued tcgen05.mma instructions x 2*128*128*16. */
int b200mnn_profile_collect_executed(double* executed_flops);

/* Same for the split-fp16 tcgen05 GEMM launches of the gene-space kernels (smoothing, wide-d kNN): summed CUDA-event
 * time, launches and EXECUTED tensor flops (terms x padded tiles x K). */
int b200mnn_gemm_profile_enable(int on);
int b200mnn_gemm_profile_collect(double* total_ms, int64_t* launches, double* executed_flops);

#ifdef __cplusplus
}
#endif
#endif /* B200MNN_H */
