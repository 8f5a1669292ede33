// Deterministic exclusive scan (tile sums -> scan of tile sums -> per-tile apply).
#include "scan.cuh"

namespace b200 {
namespace scan {

__global__ void __launch_bounds__(THREADS) tile_sum_kernel(const int32_t* __restrict__ counts, int64_t n, int64_t* __restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
    int local = 0;
#pragma unroll
    for (int t = 0; t < ITEMS; ++t)
        if (base + t < n) local += counts[base + t];
    int total;
    block_exclusive_scan(local, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(THREADS) tile_scan_kernel(int64_t* __restrict__ tile_sums, int64_t ntiles, int64_t* __restrict__ total) {
    // single block; sequential over chunks of THREADS tiles, each chunk scanned cooperatively
    __shared__ long long carry;
    __shared__ long long vals[THREADS];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < ntiles; c0 += THREADS) {
        const int64_t i = c0 + threadIdx.x;
        vals[threadIdx.x] = (i < ntiles) ? tile_sums[i] : 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long run = carry;
            for (int t = 0; t < THREADS; ++t) { const long long v = vals[t]; vals[t] = run; run += v; }
            carry = run;
        }
        __syncthreads();
        if (i < ntiles) tile_sums[i] = vals[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(THREADS) tile_apply_kernel(const int32_t* __restrict__ counts, int64_t n, const int64_t* __restrict__ tile_sums,
                                                             int64_t* __restrict__ out) {
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
    int v[ITEMS];
    int local = 0;
#pragma unroll
    for (int t = 0; t < ITEMS; ++t) { v[t] = (base + t < n) ? counts[base + t] : 0; local += v[t]; }
    const int excl = block_exclusive_scan(local, nullptr);
    int64_t run = tile_sums[blockIdx.x] + excl;
#pragma unroll
    for (int t = 0; t < ITEMS; ++t) {
        if (base + t < n) out[base + t] = run;
        run += v[t];
    }
}

int exclusive_scan(const int32_t* d_counts, int64_t n, int64_t* d_out, int64_t* d_total, cudaStream_t stream) {
    if (n <= 0) {
        if (d_total) B200_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), stream));
        return 0;
    }
    const int64_t ntiles = ceil_div(n, TILE);
    Scratch ws(stream);
    int64_t* tile_sums = ws.get<int64_t>((size_t)ntiles);
    if (!ws.ok()) return B200MNN_ENOMEM;
    tile_sum_kernel<<<(unsigned)ntiles, THREADS, 0, stream>>>(d_counts, n, tile_sums);
    B200_LAUNCH_CHECK();
    tile_scan_kernel<<<1, THREADS, 0, stream>>>(tile_sums, ntiles, d_total);
    B200_LAUNCH_CHECK();
    tile_apply_kernel<<<(unsigned)ntiles, THREADS, 0, stream>>>(d_counts, n, tile_sums, d_out);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace scan
}  // namespace b200
