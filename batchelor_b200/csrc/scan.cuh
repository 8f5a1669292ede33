// Deterministic exclusive prefix sums over int32 counts (three small kernels, no atomics, no host sync).
// Used by the mutual-pair compaction (find_mutual_nns order contract) and the pair->cell grouping.
#pragma once

#include "common.cuh"

namespace b200 {
namespace scan {

constexpr int THREADS = 256;
constexpr int ITEMS = 16;               // consecutive items per thread
constexpr int TILE = THREADS * ITEMS;   // items per block

// Block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total receives the block sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[THREADS / 32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // protects warp_sums reuse across calls
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < THREADS / 32) ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < THREADS / 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == THREADS / 32 - 1) block_total = wi;
    }
    __syncthreads();
    if (total) *total = block_total;
    return incl - v + warp_sums[warp];
}

// Exclusive scan: d_out[i] = sum_{j<i} d_counts[j] (int64), *d_total = sum of all.  d_total may be null.
int exclusive_scan(const int32_t* d_counts, int64_t n, int64_t* d_out, int64_t* d_total, cudaStream_t stream);

}  // namespace scan
}  // namespace b200
