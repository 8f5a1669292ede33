// Cluster plan of the pruned exact kNN search (csrc/knn_cluster.cu): reference rows grouped by a coarse k-means,
// query rows grouped by their nearest reference centroid, and -- per tile of 128 grouped queries -- the reference
// clusters in ascending order of a rigorous lower bound on the query-to-cluster distance.  This is the B200 form of
// what KMKNN does on the CPU (k-means + triangle-inequality pruning inside BiocNeighbors::queryKNN, the call at
// R/MNN_tree.R:129): the search stays exact, whole 128x128 score tiles are skipped instead of single points.
#pragma once

#include "common.cuh"

namespace b200 {
namespace knn {

constexpr int CL_MAXC = 256;     // clusters (power of two)
constexpr int CL_TILE = 128;     // rows per tile: = TS_BN (reference tile) = BM (query tile)

struct ClusterPlan {
    // reference side (build_ref_plan); may be shared by several searches against the same reference set
    int C = 0;
    int64_t n_rows_max = 0;      // rows of the grouped + padded reference operand (multiple of CL_TILE, host-side upper bound)
    int32_t* refmap = nullptr;   // [n_rows_max]  grouped reference row -> original row (-1: padding)
    int* cl_tile0 = nullptr;     // [C + 1] first reference tile of every cluster
    int* cnt_ref = nullptr;      // [C] reference rows per cluster
    double* centroids = nullptr; // [C][d]
    double* centroids_t = nullptr;   // [d][C] the same, transposed (what the kernels stage in shared memory)
    double* cnorm = nullptr;     // [C] squared centroid norms
    double* cdist = nullptr;     // [C][C] centroid distances
    double* cinv = nullptr;      // [C][C] their reciprocals
    unsigned long long* vref = nullptr;   // [C][C] ordered-key maxima: extent of cluster B's rows towards centroid A
    // query side (build_query_plan)
    int64_t nslots_max = 0;      // query slots (multiple of CL_TILE, host-side upper bound)
    int32_t* qmap = nullptr;     // [nslots_max]  query slot -> original query (-1: padding)
    int* nslots = nullptr;       // device scalar: slots in use (multiple of CL_TILE)
    int2* cl_list = nullptr;     // [nslots_max / CL_TILE][C]  (cluster, float bits of S^2 * lower_bound^2), ascending
    float* qoff = nullptr;       // [nslots_max]  S^2 ||q||^2 rounded up (-inf for padding slots)
    int32_t* cid_q = nullptr;    // [nq] cluster of every query (original order)
    double* dots_q = nullptr;    // [nq][C] q . c for every centroid, kept from the assignment for the tile lists (may be null)
};

// Both parts are built on `stream` without host synchronisation.  Reference side: k-means on a sample, every reference row
// assigned and grouped, the C x C extent table.  Query side (needs the reference side): every query assigned to its nearest
// reference centroid and grouped, per query tile the sorted cluster list.  qnorm: fp64 squared norms of the queries
// (original order), xnorm: those of the references; scale_exp / maxnorm_bits: the device scalars of the scoring pipeline
// (knn_tc.cu).
int build_ref_plan(const double* dX, int64_t n, int d, int C, const double* xnorm, const unsigned long long* maxnorm_bits, Scratch& ws,
                   cudaStream_t stream, ClusterPlan* plan);
int build_query_plan(ClusterPlan* plan, const double* dQ, int64_t nq, int d, const double* qnorm, const int* scale_exp,
                     const unsigned long long* maxnorm_bits, Scratch& ws, cudaStream_t stream);

// Cluster lists (+ score offsets) for an arbitrary slot -> query map (used for the second scoring tier, whose slots are
// the uncertified queries in slot order): count is a device scalar, lists/qoff are sized for max_slots.
int build_tile_lists(const ClusterPlan& plan, const double* dQ, int d, const int32_t* qmap, const int* count, int64_t max_slots,
                     const double* qnorm, const int* scale_exp, const unsigned long long* maxnorm_bits, int2* lists, float* qoff,
                     cudaStream_t stream);

// Regroups an unordered device list of query ids (count on the device, at most max_items) into cluster-pure slots padded
// to multiples of CL_TILE: map[slot] = query or -1, *nslots = slots in use.  max_slots >= max_items + C * CL_TILE.
int regroup_query_list(const ClusterPlan& plan, const int32_t* list, const int* count, int64_t max_items, int32_t* map, int64_t max_slots,
                       int* nslots, int* work /* [3 * CL_MAXC] ints */, cudaStream_t stream);

}  // namespace knn
}  // namespace b200
