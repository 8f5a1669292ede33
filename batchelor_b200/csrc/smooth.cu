// Gaussian-kernel smoothing of per-MNN-cell correction vectors for sm_100a -- replaces smooth_gaussian_kernel()
// (src/smooth_gaussian_kernel.cpp:11-117, .Call _batchelor_smooth_gaussian_kernel at src/RcppExports.cpp:36-46).
//
//   logw(c,i) = -||m_c - m_{U_i}||^2 / sigma2        (reference :36-52; sigma2 is used un-squared)
//   dens_i    = logsumexp_j logw(U_j, i)              (:56-65)
//   out[c,:]  = sum_i softmax_i(logw(c,i) - dens_i) * averaged[i,:]      (:75-115)
//
// The reference walks MNN cells in the outer loop with one running exponent per (gene, cell); every gene of a cell
// sees the same log-multipliers, so this is blockwise attention with one running maximum per cell: a distance
// ("QK^T") tile, a row soft-max and a weighted accumulate ("PV").  Round-1 implementation: fp64 CUDA-core tiles
// (difference-form distances, so no cancellation at sigma = 0.1 on cosine-normalised data), rows processed in
// chunks so the [chunk x nmnn] weight block stays below ~1 GiB and is never materialised for all cells at once.
// Layout: mat [ncells x Gdist], averaged [nmnn x G], out [ncells x G], all row-major (one cell contiguous).
#include "common.cuh"

namespace b200 {
namespace smooth {

constexpr int T = 64;    // tile edge
constexpr int KC = 16;   // K chunk

// L[r, c] = -||mat[rowid(r)] - mat[colidx[c]]||^2 * inv_sigma - (dens ? dens[c] : 0)
__global__ void __launch_bounds__(256)
logit_tile_kernel(const double* __restrict__ mat, int64_t Gd, const int32_t* __restrict__ rowidx, int64_t row0, int64_t nrows,
                  const int32_t* __restrict__ colidx, int64_t ncols, double inv_sigma, const double* __restrict__ dens,
                  double* __restrict__ L) {
    __shared__ double As[KC][T + 1];
    __shared__ double Bs[KC][T + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t rb = (int64_t)blockIdx.y * T, cb = (int64_t)blockIdx.x * T;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;  // loader: row lr of the tile, 4 consecutive k
    int64_t arow = -1, brow = -1;
    if (rb + lr < nrows) arow = rowidx ? (int64_t)rowidx[rb + lr] : row0 + rb + lr;
    if (cb + lr < ncols) brow = colidx[cb + lr];
    for (int64_t k0 = 0; k0 < Gd; k0 += KC) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t k = k0 + lk + j;
            As[lk + j][lr] = (arow >= 0 && k < Gd) ? mat[arow * Gd + k] : 0.0;
            Bs[lk + j][lr] = (brow >= 0 && k < Gd) ? mat[brow * Gd + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[k][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[k][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double df = av[a] - bv[b];
                    acc[a][b] = fma(df, df, acc[a][b]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t r = rb + ty * 4 + a;
        if (r >= nrows) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t c = cb + tx * 4 + b;
            if (c < ncols) L[r * ncols + c] = -acc[a][b] * inv_sigma - (dens ? dens[c] : 0.0);
        }
    }
}

// Row soft-max in place.  lse_out (may be null) receives logsumexp of the row; when `normalise` the row becomes
// exp(l - max) / sum.
__global__ void __launch_bounds__(256)
row_softmax_kernel(double* __restrict__ L, int64_t nrows, int64_t ncols, bool normalise, double* __restrict__ lse_out) {
    __shared__ double sm[256];
    const int64_t r = blockIdx.x;
    double* row = L + r * ncols;
    double m = -INFINITY;
    for (int64_t c = threadIdx.x; c < ncols; c += blockDim.x) m = fmax(m, row[c]);
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    m = sm[0];
    __syncthreads();
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < ncols; c += blockDim.x) {
        const double e = exp(row[c] - m);
        s += e;
        if (normalise) row[c] = e;
    }
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    s = sm[0];
    if (lse_out && threadIdx.x == 0) lse_out[r] = m + log(s);
    if (normalise) {
        const double inv = 1.0 / s;
        for (int64_t c = threadIdx.x; c < ncols; c += blockDim.x) row[c] *= inv;
    }
}

// C[M x N] = A[M x K] * B[K x N], all row-major fp64.
__global__ void __launch_bounds__(256)
matmul_tile_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int64_t M, int64_t N, int64_t K) {
    __shared__ double As[KC][T + 1];
    __shared__ double Bs[KC][T + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t rb = (int64_t)blockIdx.y * T, cb = (int64_t)blockIdx.x * T;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;   // A loader
    const int bk = threadIdx.x >> 4, bc = (threadIdx.x & 15) * 4;  // B loader: row bk of the chunk, 4 consecutive cols
    for (int64_t k0 = 0; k0 < K; k0 += KC) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t k = k0 + lk + j;
            As[lk + j][lr] = (rb + lr < M && k < K) ? A[(rb + lr) * K + k] : 0.0;
            const int64_t c = cb + bc + j;
            Bs[bk][bc + j] = (k0 + bk < K && c < N) ? B[(k0 + bk) * N + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[k][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[k][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t r = rb + ty * 4 + a;
        if (r >= M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t c = cb + tx * 4 + b;
            if (c < N) C[r * N + c] = acc[a][b];
        }
    }
}

__global__ void check_index_kernel(const int32_t* __restrict__ idx, int64_t n, int64_t limit, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (idx[i] < 0 || idx[i] >= limit)) *bad = 1;
}

int smooth_gaussian_kernel_device(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0, const double* d_mat,
                                  int64_t Gdist, int64_t ncells, double sigma2, double* d_out, int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (G < 0 || nmnn < 0 || Gdist < 0 || ncells < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (ncells == 0 || G == 0) return 0;
    if (nmnn == 0) {  // reference: zero-initialised output times exp(-Inf - NA) (:105-115) = NaN everywhere
        B200_CUDA(cudaMemsetAsync(d_out, 0xFF, sizeof(double) * ncells * G, stream));  // all-ones bit pattern is a quiet NaN
        return 0;
    }
    Scratch ws(stream);
    const int64_t budget = (int64_t)1 << 27;  // doubles in one weight chunk (1 GiB)
    int64_t chunk = std::max<int64_t>(T, std::min<int64_t>(std::max<int64_t>(ncells, nmnn), budget / nmnn));
    chunk = std::min<int64_t>(round_up(chunk, T), (int64_t)1 << 21);  // grid.y limit
    double* W = ws.get<double>((size_t)chunk * nmnn);
    double* dens = ws.get<double>((size_t)nmnn);
    int* bad = d_bad ? d_bad : ws.get<int>(1);
    if (!ws.ok()) return B200MNN_ENOMEM;
    if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    check_index_kernel<<<(unsigned)ceil_div(nmnn, 256), 256, 0, stream>>>(d_index0, nmnn, ncells, bad);
    B200_LAUNCH_CHECK();
    const double inv_sigma = 1.0 / sigma2;
    // pass 1: density of every MNN cell among the MNN cells
    for (int64_t r0 = 0; r0 < nmnn; r0 += chunk) {
        const int64_t nr = std::min(chunk, nmnn - r0);
        dim3 grid((unsigned)ceil_div(nmnn, T), (unsigned)ceil_div(nr, T));
        logit_tile_kernel<<<grid, 256, 0, stream>>>(d_mat, Gdist, d_index0 + r0, 0, nr, d_index0, nmnn, inv_sigma, nullptr, W);
        B200_LAUNCH_CHECK();
        row_softmax_kernel<<<(unsigned)nr, 256, 0, stream>>>(W, nr, nmnn, false, dens + r0);
        B200_LAUNCH_CHECK();
    }
    // pass 2: weights of every cell, then the weighted average of the correction vectors
    for (int64_t r0 = 0; r0 < ncells; r0 += chunk) {
        const int64_t nr = std::min(chunk, ncells - r0);
        dim3 grid((unsigned)ceil_div(nmnn, T), (unsigned)ceil_div(nr, T));
        logit_tile_kernel<<<grid, 256, 0, stream>>>(d_mat, Gdist, nullptr, r0, nr, d_index0, nmnn, inv_sigma, dens, W);
        B200_LAUNCH_CHECK();
        row_softmax_kernel<<<(unsigned)nr, 256, 0, stream>>>(W, nr, nmnn, true, nullptr);
        B200_LAUNCH_CHECK();
        dim3 g2((unsigned)ceil_div(G, T), (unsigned)ceil_div(nr, T));
        matmul_tile_kernel<<<g2, 256, 0, stream>>>(W, d_averaged, d_out + r0 * G, nr, G, nmnn);
        B200_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace smooth
}  // namespace b200

extern "C" int b200mnn_dev_smooth_gaussian_kernel(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0,
                                                  const double* d_mat, int64_t Gdist, int64_t ncells, double sigma2, double* d_out,
                                                  void* stream) {
    return b200::smooth::smooth_gaussian_kernel_device(d_averaged, G, nmnn, d_index0, d_mat, Gdist, ncells, sigma2, d_out, nullptr,
                                                       static_cast<cudaStream_t>(stream));
}
