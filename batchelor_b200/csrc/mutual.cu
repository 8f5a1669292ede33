// Mutual-nearest-neighbour pair extraction for sm_100a -- replaces find_mutual_nns()
// (src/find_mutual_nns.cpp:8-41, .Call _batchelor_find_mutual_nns at src/RcppExports.cpp:25-33).
//
// left  [n1 x k2] : for every cell l of batch 1, the ids of its k2 nearest cells in batch 2
// right [n2 x k1] : for every cell v of batch 2, the ids of its k1 nearest cells in batch 1
// (device layout: row-major, 0-based).  A pair (l, v) is emitted iff v is in left[l,] and l is in right[v,].
// Order contract of the reference (its downstream rowsum depends on it): l ascending and, inside l, in the
// column order of left[l,].  The reference sorts each right row and binary-searches; a row is only k1 <= ~64
// ints (<= 3 sectors), so here every (l, v) probe is a straight scan of the gathered row, which lives in L2
// (n2*k1*4 B = 80 MB at 1M x 20).  Deterministic compaction: per-row counts -> exclusive scan -> ordered writes.
// HBM/L2-bound integer work: 4*(n1*k2 + n2*k1) B streamed + <= 4*n1*k2*k1 B of gathered rows + 8*np B written.
#include "common.cuh"
#include "scan.cuh"

namespace b200 {
namespace mutual {

__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ row, int k, int32_t want) {
    bool found = false;
    if ((k & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {   // 16-byte loads: a quarter of the load instructions
        const int4* r4 = reinterpret_cast<const int4*>(row);
        for (int c = 0; c < (k >> 2); ++c) {
            const int4 v = __ldg(r4 + c);
            found |= (v.x == want) | (v.y == want) | (v.z == want) | (v.w == want);
        }
        return found;
    }
    for (int c = 0; c < k; ++c) found |= (__ldg(row + c) == want);
    return found;
}

// counts[l] = number of mutual partners of l; masks[l] (optional, k2 <= 64): bit j set iff left[l, j] is a mutual partner, so
// that the write pass does not probe again.  bad[0] is raised if an id is out of range.
__global__ void count_kernel(const int32_t* __restrict__ left, int64_t n1, int k2, const int32_t* __restrict__ right, int64_t n2, int k1,
                             int32_t* __restrict__ counts, unsigned long long* __restrict__ masks, int* __restrict__ bad) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n1) return;
    int c = 0;
    unsigned long long m = 0ull;
    for (int j = 0; j < k2; ++j) {
        const int32_t v = left[l * k2 + j];
        if (v < 0 || v >= n2) { *bad = 1; continue; }
        const bool hit = row_contains(right + (int64_t)v * k1, k1, (int32_t)l);
        c += hit ? 1 : 0;
        if (hit && j < 64) m |= 1ull << j;
    }
    counts[l] = c;
    if (masks) masks[l] = m;
}

__global__ void write_kernel(const int32_t* __restrict__ left, int64_t n1, int k2, const int32_t* __restrict__ right, int64_t n2, int k1,
                             const int64_t* __restrict__ offsets, const unsigned long long* __restrict__ masks, int32_t* __restrict__ first,
                             int32_t* __restrict__ second, int64_t capacity, int32_t base) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n1) return;
    int64_t o = offsets[l];
    if (masks) {   // the count pass recorded which columns are mutual
        unsigned long long m = masks[l];
        while (m) {
            const int j = __ffsll((long long)m) - 1;
            m &= m - 1;
            if (o < capacity) { first[o] = (int32_t)l + base; second[o] = left[l * k2 + j] + base; }
            ++o;
        }
        return;
    }
    for (int j = 0; j < k2; ++j) {
        const int32_t v = left[l * k2 + j];
        if (v < 0 || v >= n2) continue;
        if (row_contains(right + (int64_t)v * k1, k1, (int32_t)l)) {
            if (o < capacity) { first[o] = (int32_t)l + base; second[o] = v + base; }
            ++o;
        }
    }
}

int find_mutual_nns_device(const int32_t* d_left, int64_t n1, int k2, const int32_t* d_right, int64_t n2, int k1, int32_t* d_first,
                           int32_t* d_second, int64_t capacity, int64_t* d_np, int32_t out_base, int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (n1 < 0 || n2 < 0 || k1 < 0 || k2 < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (n1 == 0 || k2 == 0 || n2 == 0 || k1 == 0) {
        B200_CUDA(cudaMemsetAsync(d_np, 0, sizeof(int64_t), stream));
        return 0;
    }
    Scratch ws(stream);
    int32_t* counts = ws.get<int32_t>((size_t)n1);
    int64_t* offsets = ws.get<int64_t>((size_t)n1);
    int* bad = d_bad ? d_bad : ws.get<int>(1);
    unsigned long long* masks = k2 <= 64 ? ws.get<unsigned long long>((size_t)n1) : nullptr;
    if (!ws.ok()) return B200MNN_ENOMEM;
    if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    const unsigned blocks = (unsigned)ceil_div(n1, 256);
    count_kernel<<<blocks, 256, 0, stream>>>(d_left, n1, k2, d_right, n2, k1, counts, masks, bad);
    B200_LAUNCH_CHECK();
    B200_TRY(scan::exclusive_scan(counts, n1, offsets, d_np, stream));
    write_kernel<<<blocks, 256, 0, stream>>>(d_left, n1, k2, d_right, n2, k1, offsets, masks, d_first, d_second, capacity, out_base);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace mutual
}  // namespace b200

extern "C" int b200mnn_dev_find_mutual_nns(const int32_t* d_left, int64_t n1, int k2, const int32_t* d_right, int64_t n2, int k1,
                                           int32_t* d_first, int32_t* d_second, int64_t capacity, int64_t* d_np, void* stream) {
    return b200::mutual::find_mutual_nns_device(d_left, n1, k2, d_right, n2, k1, d_first, d_second, capacity, d_np, 0, nullptr,
                                                static_cast<cudaStream_t>(stream));
}
