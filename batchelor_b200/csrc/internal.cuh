// Host-callable entry points of each translation unit (device pointers; `d_bad` is an optional device int that
// kernels raise when they meet an out-of-range index -- the host-buffer layer turns it into the reference's error).
#pragma once

#include "common.cuh"
#include "knn_cluster.cuh"

namespace b200 {

namespace knn {
struct DebugOut;
bool tensor_path_supported(int64_t n, int64_t nq, int d, int k);
// Reference side of a search -- row norms, the scale, the cluster plan, the grouped fp16 operand -- kept so that several
// calls can search the same reference set (b200mnn_find_mutual_nn searches batch 2 chunk by chunk against batch 1 while
// batch 2 is still crossing PCIe).  The first call with a cache fills it (nq == 0: that and nothing else); later calls must pass
// the same dX / n / d / k and
// run on the cache's stream.  With a cache the fp16 scale is derived from the reference rows alone (a query row more than
// 4x larger than every reference coordinate overflows fp16, fails its certificate and is answered by the exact rescue
// scan: slower, never wrong).
struct RefCache {
    explicit RefCache(cudaStream_t s) : ws(s) {}
    Scratch ws;              // owns the reference-side buffers
    bool ready = false;
    const double* dX = nullptr;
    int64_t n = 0;
    int d = 0, k = 0;
    bool use_prune = false;
    int64_t nq_hint = 0;     // query rows per search expected by a prepare-only call (nq == 0)
    void* opB = nullptr;
    double* xnorm = nullptr;
    unsigned char* scalars = nullptr;   // absmax_bits, scale_exp, maxnorm_bits, bmax_bits
    ClusterPlan plan;        // reference part
};
int query_knn_device(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                     int64_t* d_stats, cudaStream_t stream, const DebugOut* dbg, RefCache* cache = nullptr);
// knn_wide.cu: K-streamed tensor-core path for wide data / large k
bool wide_path_supported(int64_t n, int64_t nq, int d, int k);
int query_knn_wide(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                   int64_t* d_stats, cudaStream_t stream);
// knn_tc.cu: exact fp64 scan of the flagged queries (flag_dk2: exact squared distance of each one's k-th candidate), or of all
int launch_rescue(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, const int* flag_count, const int32_t* flag_list,
                  const double* flag_dk2, int32_t* d_idx, double* d_dist, cudaStream_t stream);
int launch_rescue_all(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                      int64_t* d_stats, cudaStream_t stream);
int write_stats(const int* flag_count, int64_t* d_stats, int64_t lists, int64_t path, cudaStream_t stream);
}  // namespace knn

namespace mutual {
int find_mutual_nns_device(const int32_t* d_left, int64_t n1, int k2, const int32_t* d_right, int64_t n2, int k1, int32_t* d_first,
                           int32_t* d_second, int64_t capacity, int64_t* d_np, int32_t out_base, int* d_bad, cudaStream_t stream);
}

namespace correct {
template <typename T>
int transpose_device(const T* d_in, int64_t rows, int64_t cols, T* d_out, cudaStream_t stream);
int average_correction_device(const double* d_ref, int64_t n1, const double* d_cur, int64_t n2, int d, const int32_t* d_first,
                              const int32_t* d_second, int64_t np, double* d_averaged, int32_t* d_second_unique, int64_t* d_nmnn,
                              int* d_bad, cudaStream_t stream);
int center_along_batch_vector_device(double* d_mat, int64_t n, int d, const double* d_batch_vec, const int32_t* d_restrict,
                                     int64_t nrestrict, int* d_bad, cudaStream_t stream);
int tricube_apply_device(const double* d_cur, int64_t n, int d, const double* d_correction, int64_t nmnn, const int32_t* d_idx,
                         const double* d_dist, int k, double ndist, double* d_out, int* d_bad, cudaStream_t stream);
int cosine_norm_device(const double* d_x, int64_t n, int64_t G, double* d_out, double* d_l2, cudaStream_t stream);
}  // namespace correct

namespace smooth {
int smooth_gaussian_kernel_device(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0, const double* d_mat,
                                  int64_t Gdist, int64_t ncells, double sigma2, double* d_out, int* d_bad, cudaStream_t stream);
}

namespace shiftvar {
int adjust_shift_variance_device(const double* d_data1, int64_t n1, const double* d_data2, int64_t n2, int64_t G, const double* d_vect,
                                 double sigma2, const int32_t* d_r1, int64_t nr1, const int32_t* d_r2, int64_t nr2, double* d_out,
                                 int* d_bad, cudaStream_t stream);
}

}  // namespace b200
