// HBM-bound correction kernels of the fastMNN merge step for sm_100a (all fp64, row-major [cells x dims]):
//   average_correction          R/fastMNN.R:567-580 (and R/mnnCorrect.R:456-457 via sumCountsAcrossCells)
//   center_along_batch_vector   R/fastMNN.R:626-640 (+ .orthogonalize_other :642-647 = repeated calls)
//   tricube_apply               R/utils_tricube.R:1-27 + R/fastMNN.R:607
//   cosine_norm                 R/cosineNorm.R:53-82
//   transpose                   column-major (R) <-> row-major staging for the host-buffer layer
// Everything is deterministic (no floating-point atomics): sums run in a fixed order so that duplicated cells
// get bit-identical results, as the reference's restrict tests demand (tests/testthat/test-reduced-mnn.R:107-145).
#include "common.cuh"
#include "scan.cuh"

namespace b200 {
namespace correct {

// ------------------------------------------------------------------------------------------------
// transpose: in [rows x cols] column-major  ->  out [rows x cols] row-major  (i.e. a plain 2-D transpose)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, int64_t rows, int64_t cols, int64_t tiles_r, T* __restrict__ out) {
    __shared__ T tile[32][33];
    const int64_t r0 = ((int64_t)blockIdx.x % tiles_r) * 32, c0 = ((int64_t)blockIdx.x / tiles_r) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + threadIdx.x, c = c0 + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[c * rows + r];  // coalesced along rows
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) out[r * cols + c] = tile[threadIdx.x][j];  // coalesced along cols
    }
}

template <typename T>
int transpose_device(const T* d_in, int64_t rows, int64_t cols, T* d_out, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t tiles_r = ceil_div(rows, 32), tiles_c = ceil_div(cols, 32);
    if (tiles_r * tiles_c > 2147483647LL) return fail(B200MNN_EINVAL, "matrix too large for the staging transpose");
    dim3 block(32, 8);
    transpose_kernel<T><<<(unsigned)(tiles_r * tiles_c), block, 0, stream>>>(d_in, rows, cols, tiles_r, d_out);
    B200_LAUNCH_CHECK();
    return 0;
}

template int transpose_device<double>(const double*, int64_t, int64_t, double*, cudaStream_t);
template int transpose_device<int32_t>(const int32_t*, int64_t, int64_t, int32_t*, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// average_correction: group pairs by their batch-2 cell, keep pair order inside a group (rowsum order)
// ------------------------------------------------------------------------------------------------
__global__ void pair_count_kernel(const int32_t* __restrict__ second, int64_t np, int64_t n2, int32_t* __restrict__ counts, int* __restrict__ bad) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int32_t s = second[p];
    if (s < 0 || s >= n2) { *bad = 1; return; }
    atomicAdd(&counts[s], 1);  // integer: order-independent
}

__global__ void present_kernel(const int32_t* __restrict__ counts, int64_t n2, int32_t* __restrict__ present) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) present[i] = counts[i] > 0 ? 1 : 0;
}

__global__ void pair_fill_kernel(const int32_t* __restrict__ second, int64_t np, int64_t n2, const int64_t* __restrict__ offsets,
                                 int32_t* __restrict__ cursor, int64_t* __restrict__ slots) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int32_t s = second[p];
    if (s < 0 || s >= n2) return;
    const int pos = atomicAdd(&cursor[s], 1);
    slots[offsets[s] + pos] = p;  // arbitrary order here; put back into pair order below
}

// One warp per batch-2 cell that has pairs: order its pair ids ascending (rank by counting), then accumulate
// ref[first] - cur[second] over them in that order, divide by the count.
__global__ void pair_average_kernel(const double* __restrict__ ref, const double* __restrict__ cur, int d, const int32_t* __restrict__ first,
                                    const int32_t* __restrict__ counts, const int64_t* __restrict__ offsets, const int64_t* __restrict__ rank,
                                    int64_t n2, int64_t n1, int64_t* __restrict__ slots, int64_t* __restrict__ sorted, double* __restrict__ averaged,
                                    int32_t* __restrict__ second_unique, int* __restrict__ bad) {
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n2) return;
    const int m = counts[s];
    if (m == 0) return;
    const int64_t base = offsets[s];
    for (int e = lane; e < m; e += 32) {
        const int64_t pe = slots[base + e];
        int r = 0;
        for (int o = 0; o < m; ++o) r += (slots[base + o] < pe) ? 1 : 0;
        sorted[base + r] = pe;
    }
    __syncwarp();
    const int64_t row = rank[s];
    if (lane == 0) second_unique[row] = (int32_t)s;
    const double* cs = cur + s * d;
    for (int t = lane; t < d; t += 32) {
        double acc = 0.0;
        const double cv = cs[t];
        for (int e = 0; e < m; ++e) {
            const int64_t p = sorted[base + e];
            const int32_t f = first[p];
            if (f < 0 || f >= n1) { *bad = 1; continue; }
            acc += ref[(int64_t)f * d + t] - cv;
        }
        averaged[row * d + t] = acc / (double)m;
    }
}

int average_correction_device(const double* d_ref, int64_t n1, const double* d_cur, int64_t n2, int d, const int32_t* d_first,
                              const int32_t* d_second, int64_t np, double* d_averaged, int32_t* d_second_unique, int64_t* d_nmnn,
                              int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (np < 0 || n1 < 0 || n2 < 0 || d < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (np == 0 || n2 == 0) {
        B200_CUDA(cudaMemsetAsync(d_nmnn, 0, sizeof(int64_t), stream));
        return 0;
    }
    Scratch ws(stream);
    int32_t* counts = ws.get<int32_t>((size_t)n2);
    int32_t* cursor = ws.get<int32_t>((size_t)n2);
    int32_t* present = ws.get<int32_t>((size_t)n2);
    int64_t* offsets = ws.get<int64_t>((size_t)n2);
    int64_t* rank = ws.get<int64_t>((size_t)n2);
    int64_t* slots = ws.get<int64_t>((size_t)np);
    int64_t* sorted = ws.get<int64_t>((size_t)np);
    int* bad = d_bad ? d_bad : ws.get<int>(1);
    if (!ws.ok()) return B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n2, stream));
    B200_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n2, stream));
    if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    pair_count_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, stream>>>(d_second, np, n2, counts, bad);
    B200_LAUNCH_CHECK();
    present_kernel<<<(unsigned)ceil_div(n2, 256), 256, 0, stream>>>(counts, n2, present);
    B200_LAUNCH_CHECK();
    B200_TRY(scan::exclusive_scan(counts, n2, offsets, nullptr, stream));
    B200_TRY(scan::exclusive_scan(present, n2, rank, d_nmnn, stream));
    pair_fill_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, stream>>>(d_second, np, n2, offsets, cursor, slots);
    B200_LAUNCH_CHECK();
    pair_average_kernel<<<(unsigned)ceil_div(n2 * 32, 256), 256, 0, stream>>>(d_ref, d_cur, d, d_first, counts, offsets, rank, n2, n1, slots, sorted,
                                                                             d_averaged, d_second_unique, bad);
    B200_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// centring along a batch vector
// ------------------------------------------------------------------------------------------------
// unit[t] = v[t] / sqrt(sum v^2)   (single warp; sequential-order sum like R's sum())
__global__ void unit_vector_kernel(const double* __restrict__ v, int d, double* __restrict__ unit) {
    __shared__ double nrm;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < d; ++t) s += v[t] * v[t];
        nrm = sqrt(s);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < d; t += blockDim.x) unit[t] = v[t] / nrm;
}

// loc[i] = mat[i,] . unit      (one warp per row; fixed shuffle tree -> deterministic, identical rows give identical loc)
__global__ void project_kernel(const double* __restrict__ mat, int64_t n, int d, const double* __restrict__ unit, double* __restrict__ loc) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    double acc = 0.0;
    for (int t = lane; t < d; t += 32) acc += mat[i * d + t] * unit[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) loc[i] = acc;
}

// partial[b] = sum over a fixed slice of loc[restrict] (or loc) -- fixed slicing and in-block tree: deterministic
constexpr int RED_BLOCKS = 256;
__global__ void mean_partial_kernel(const double* __restrict__ loc, const int32_t* __restrict__ restrict0, int64_t count, int64_t n,
                                    double* __restrict__ partial, int* __restrict__ bad) {
    __shared__ double sm[256];
    const int64_t per = (count + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * per, b1 = min(count, b0 + per);
    double acc = 0.0;
    for (int64_t j = b0 + threadIdx.x; j < b1; j += blockDim.x) {
        int64_t i = j;
        if (restrict0) {
            i = restrict0[j];
            if (i < 0 || i >= n) { *bad = 1; continue; }
        }
        acc += loc[i];
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void mean_final_kernel(const double* __restrict__ partial, int nb, int64_t count, double* __restrict__ central) {
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[b];
    *central = s / (double)count;
}

// mat[i,t] += (central - loc[i]) * unit[t]
__global__ void shift_kernel(double* __restrict__ mat, int64_t n, int d, const double* __restrict__ unit, const double* __restrict__ loc,
                             const double* __restrict__ central) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * d) return;
    const int64_t i = e / d;
    const int t = (int)(e - i * d);
    mat[e] = mat[e] + (*central - loc[i]) * unit[t];
}

int center_along_batch_vector_device(double* d_mat, int64_t n, int d, const double* d_batch_vec, const int32_t* d_restrict, int64_t nrestrict,
                                     int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (n <= 0 || d <= 0) return 0;
    Scratch ws(stream);
    double* unit = ws.get<double>((size_t)d);
    double* loc = ws.get<double>((size_t)n);
    double* partial = ws.get<double>(RED_BLOCKS);
    double* central = ws.get<double>(1);
    int* bad = d_bad ? d_bad : ws.get<int>(1);
    if (!ws.ok()) return B200MNN_ENOMEM;
    if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    unit_vector_kernel<<<1, 64, 0, stream>>>(d_batch_vec, d, unit);
    B200_LAUNCH_CHECK();
    project_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, stream>>>(d_mat, n, d, unit, loc);
    B200_LAUNCH_CHECK();
    const int64_t count = d_restrict ? nrestrict : n;
    mean_partial_kernel<<<RED_BLOCKS, 256, 0, stream>>>(loc, d_restrict, count, n, partial, bad);
    B200_LAUNCH_CHECK();
    mean_final_kernel<<<1, 1, 0, stream>>>(partial, RED_BLOCKS, count, central);
    B200_LAUNCH_CHECK();
    shift_kernel<<<(unsigned)ceil_div(n * d, 256), 256, 0, stream>>>(d_mat, n, d, unit, loc, central);
    B200_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// tricube smoothing given the neighbours among the MNN cells
// ------------------------------------------------------------------------------------------------
__global__ void tricube_kernel(const double* __restrict__ cur, int64_t n, int d, const double* __restrict__ correction, int64_t nmnn,
                               const int32_t* __restrict__ idx, const double* __restrict__ dist, int k, double ndist, double* __restrict__ out,
                               int* __restrict__ bad) {
    extern __shared__ double wbuf[];  // [warps][k]
    const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
    if (i >= n) return;
    double* w = wbuf + (size_t)warp_in_block * k;
    const int middle = (k + 1) / 2;  // ceiling(k/2), 1-based
    double bw = dist[i * k + (middle - 1)] * ndist;
    bw = fmax(1e-8, bw);
    for (int j = lane; j < k; j += 32) {
        double rel = dist[i * k + j] / bw;
        if (rel > 1.0) rel = 1.0;
        const double a = 1.0 - rel * rel * rel;
        w[j] = a * a * a;
    }
    __syncwarp();
    double total = 0.0;
    for (int j = 0; j < k; ++j) total += w[j];  // rowSums order
    for (int t = lane; t < d; t += 32) {
        double acc = 0.0;
        for (int j = 0; j < k; ++j) {
            const int32_t id = idx[i * k + j];
            if (id < 0 || id >= nmnn) { *bad = 1; continue; }
            acc += correction[(int64_t)id * d + t] * (w[j] / total);
        }
        out[i * d + t] = cur[i * d + t] + acc;
    }
}

int tricube_apply_device(const double* d_cur, int64_t n, int d, const double* d_correction, int64_t nmnn, const int32_t* d_idx,
                         const double* d_dist, int k, double ndist, double* d_out, int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (n <= 0 || d <= 0) return 0;
    if (k <= 0) {  // zero-column guard of R/utils_tricube.R:22-23: the weighted correction is all zeros
        if (d_out != d_cur) B200_CUDA(cudaMemcpyAsync(d_out, d_cur, sizeof(double) * n * d, cudaMemcpyDeviceToDevice, stream));
        return 0;
    }
    Scratch ws(stream);
    int* bad = d_bad ? d_bad : ws.get<int>(1);
    if (!ws.ok()) return B200MNN_ENOMEM;
    if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    const int warps = 8;
    const size_t smem = (size_t)warps * k * sizeof(double);
    if (smem > 48 * 1024) return fail(B200MNN_EINVAL, "tricube smoothing supports k up to 768");
    tricube_kernel<<<(unsigned)ceil_div(n, warps), warps * 32, smem, stream>>>(d_cur, n, d, d_correction, nmnn, d_idx, d_dist, k, ndist, d_out, bad);
    B200_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// cosine normalisation: x [n x G] row-major (one cell contiguous)
// ------------------------------------------------------------------------------------------------
__global__ void cosine_kernel(const double* __restrict__ x, int64_t n, int64_t G, double* __restrict__ out, double* __restrict__ l2) {
    __shared__ double sm[256];
    const int64_t i = blockIdx.x;
    const double* row = x + i * G;
    double acc = 0.0;
    for (int64_t g = threadIdx.x; g < G; g += blockDim.x) acc += row[g] * row[g];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    const double nrm = sqrt(sm[0]);
    if (threadIdx.x == 0 && l2) l2[i] = nrm;
    if (out) {
        const double dv = fmax(1e-8, nrm);
        for (int64_t g = threadIdx.x; g < G; g += blockDim.x) out[i * G + g] = row[g] / dv;
    }
}

int cosine_norm_device(const double* d_x, int64_t n, int64_t G, double* d_out, double* d_l2, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (n <= 0) return 0;
    if (n > 2147483647LL) return fail(B200MNN_EINVAL, "too many cells");
    cosine_kernel<<<(unsigned)n, 256, 0, stream>>>(d_x, n, G, d_out, d_l2);
    B200_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// clusterMNN's propagation of centroid corrections to cells: .smooth_gaussian_from_centroids (R/clusterMNN.R:289-312)
// ------------------------------------------------------------------------------------------------
// out[c,] = x[c,] + sum_j softmax_j(-||x_c - centre_j||^2 / sigma^2) * delta[j,]: a small-K variant of the Gaussian
// smoothing (tens of centroids, no density term, two-pass soft-max against the row maximum as the reference does).
// One warp per cell; centres and deltas are read through L1/L2 (a few KB); HBM-bound on x and out.
__global__ void __launch_bounds__(256)
centroid_smooth_kernel(const double* __restrict__ x, int64_t n, int d, const double* __restrict__ centers, const double* __restrict__ delta,
                       int nc, double inv_sigma2, double* __restrict__ out) {
    extern __shared__ double wsm[];   // [warps][nc] log-weights
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= n) return;
    double* w = wsm + (size_t)warp * nc;
    const double* xc = x + c * d;
    double top = -INFINITY;
    for (int j = lane; j < nc; j += 32) {
        double d2 = 0.0;
        for (int t = 0; t < d; ++t) { const double df = xc[t] - centers[(int64_t)j * d + t]; d2 += df * df; }   // colSums order
        const double l = -d2 * inv_sigma2;
        w[j] = l;
        top = fmax(top, l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) top = fmax(top, __shfl_xor_sync(0xffffffffu, top, o));
    __syncwarp();
    double total = 0.0;
    for (int j = 0; j < nc; ++j) total += exp(w[j] - top);   // rowSums order, identical on every lane
    for (int t = lane; t < d; t += 32) {
        double acc = xc[t];
        for (int j = 0; j < nc; ++j) acc += (exp(w[j] - top) / total) * delta[(int64_t)j * d + t];   // x + outer(norm.weights[,j], delta[j,]) in j order
        out[c * d + t] = acc;
    }
}

int centroid_smooth_device(const double* d_x, int64_t n, int d, const double* d_centers, const double* d_delta, int nc, double sigma,
                           double* d_out, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (n <= 0 || d <= 0) return 0;
    if (nc < 1) return fail(B200MNN_EINVAL, "at least one centroid is needed");
    if ((size_t)nc * 8 * sizeof(double) > (size_t)160 * 1024) return fail(B200MNN_EINVAL, "too many centroids for the propagation kernel");
    const size_t smem = (size_t)8 * nc * sizeof(double);
    if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(centroid_smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    centroid_smooth_kernel<<<(unsigned)ceil_div(n, 8), 256, smem, stream>>>(d_x, n, d, d_centers, d_delta, nc, 1.0 / (sigma * sigma), d_out);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace correct
}  // namespace b200

extern "C" {

int b200mnn_dev_smooth_gaussian_from_centroids(const double* d_x, int64_t n, int d, const double* d_centers, const double* d_delta, int nc,
                                               double sigma, double* d_out, void* stream) {
    return b200::correct::centroid_smooth_device(d_x, n, d, d_centers, d_delta, nc, sigma, d_out, static_cast<cudaStream_t>(stream));
}

int b200mnn_dev_average_correction(const double* d_ref, int64_t n1, const double* d_cur, int64_t n2, int d, const int32_t* d_first,
                                   const int32_t* d_second, int64_t np, double* d_averaged, int32_t* d_second_unique, int64_t* d_nmnn,
                                   void* stream) {
    return b200::correct::average_correction_device(d_ref, n1, d_cur, n2, d, d_first, d_second, np, d_averaged, d_second_unique, d_nmnn,
                                                    nullptr, static_cast<cudaStream_t>(stream));
}

int b200mnn_dev_center_along_batch_vector(double* d_mat, int64_t n, int d, const double* d_batch_vec, const int32_t* d_restrict,
                                          int64_t nrestrict, void* stream) {
    return b200::correct::center_along_batch_vector_device(d_mat, n, d, d_batch_vec, d_restrict, nrestrict, nullptr,
                                                           static_cast<cudaStream_t>(stream));
}

int b200mnn_dev_tricube_apply(const double* d_cur, int64_t n, int d, const double* d_correction, int64_t nmnn, const int32_t* d_idx,
                              const double* d_dist, int k, double ndist, double* d_out, void* stream) {
    return b200::correct::tricube_apply_device(d_cur, n, d, d_correction, nmnn, d_idx, d_dist, k, ndist, d_out, nullptr,
                                               static_cast<cudaStream_t>(stream));
}

int b200mnn_dev_cosine_norm(const double* d_x, int64_t n, int64_t G, double* d_out, double* d_l2, void* stream) {
    return b200::correct::cosine_norm_device(d_x, n, G, d_out, d_l2, static_cast<cudaStream_t>(stream));
}

int b200mnn_dev_transpose_f64(const double* d_in, int64_t rows, int64_t cols, double* d_out, void* stream) {
    B200_TRY(b200::ensure_device());
    return b200::correct::transpose_device<double>(d_in, rows, cols, d_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
