// Gaussian-kernel smoothing of per-MNN-cell correction vectors for sm_100a -- replaces smooth_gaussian_kernel()
// (src/smooth_gaussian_kernel.cpp:11-117, .Call _batchelor_smooth_gaussian_kernel at src/RcppExports.cpp:36-46).
//
//   logw(c,i) = -||m_c - m_{U_i}||^2 / sigma2        (reference :36-52; sigma2 is used un-squared)
//   dens_i    = logsumexp_j logw(U_j, i)              (:56-65)
//   out[c,:]  = sum_i softmax_i(logw(c,i) - dens_i) * averaged[i,:]      (:75-115)
//
// The reference walks MNN cells in the outer loop with one running exponent per (gene, cell); every gene of a cell
// sees the same log-multipliers, so this is blockwise attention with one running maximum per cell: a distance
// ("QK^T") tile, a row soft-max and a weighted accumulate ("PV").  Two paths, same contract:
//   tensor (large problems): both contractions run on the tcgen05 split-fp16 GEMM of gemm_tc.cu -- distance logits with
//          the norms, 1/sigma and the log-density folded into the GEMM epilogue, row soft-max written straight into the
//          fp16 hi/lo operand of the second GEMM, weighted accumulate against V^T.  With head dimension G ~ 2000 the
//          [chunk x nmnn] weight block costs 4 + 4 bytes per 24 000 tensor flops, so it is streamed through HBM in chunks
//          (never materialised for all cells) rather than kept on chip.  Every call re-evaluates
//          a sample of output rows and densities in fp64 difference form; if they do not agree with the tensor result
//          to 1e-5 (the contract) the whole call is redone by the fp64 path.
//   fp64   (small problems, fallback): CUDA-core tiles, difference-form distances (no cancellation at sigma = 0.1 on
//          cosine-normalised data), <= 1e-10 of the reference.
// Layout: mat [ncells x Gdist], averaged [nmnn x G], out [ncells x G], all row-major (one cell contiguous).
#include "common.cuh"
#include "gemm_tc.cuh"

#include <cstdlib>
#include <cstring>

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

static int smooth_fp64(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0, const double* d_mat, int64_t Gdist,
                       int64_t ncells, double sigma2, double* d_out, int* d_bad, cudaStream_t stream) {
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


// ------------------------------------------------------------------------------------------------
// tensor path
// ------------------------------------------------------------------------------------------------
constexpr int P_EXP = 14;   // soft-max weights are stored as p * 2^14 (fp16 hi + lo): weights ~1/nmnn stay normal numbers

__device__ __forceinline__ double block_max(double v, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, sm[w]);
    return r;
}
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sm[w];
    return r;
}

// logsumexp of every row of the fp32 logit block
__global__ void __launch_bounds__(256)
row_lse_f32_kernel(const float* __restrict__ L, int64_t ld, int64_t ncols, double* __restrict__ lse_out) {
    __shared__ double sm[8];
    const float* row = L + (int64_t)blockIdx.x * ld;
    float mf = -INFINITY;
    for (int64_t c = threadIdx.x; c < ncols; c += 256) mf = fmaxf(mf, row[c]);
    const double m = block_max((double)mf, sm);
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < ncols; c += 256) s += exp((double)row[c] - m);
    s = block_sum(s, sm);
    if (threadIdx.x == 0) lse_out[blockIdx.x] = m + log(s);
}

// row soft-max -> the fp16 hi/lo operand rows of the accumulate GEMM (p * 2^P_EXP), K padding zeroed
__global__ void __launch_bounds__(256)
row_softmax_split_kernel(const float* __restrict__ L, int64_t ld, int64_t ncols, __half* __restrict__ hi, __half* __restrict__ lo, int64_t Kp) {
    __shared__ double sm[8];
    const float* row = L + (int64_t)blockIdx.x * ld;
    float mf = -INFINITY;
    for (int64_t c = threadIdx.x; c < ncols; c += 256) mf = fmaxf(mf, row[c]);
    const double m = block_max((double)mf, sm);
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < ncols; c += 256) s += exp((double)row[c] - m);
    s = block_sum(s, sm);
    const double mul = scalbn(1.0, P_EXP) / s;
    __half* h = hi + (int64_t)blockIdx.x * Kp;
    __half* l = lo + (int64_t)blockIdx.x * Kp;
    for (int64_t c = threadIdx.x; c < Kp; c += 256) {
        double p = 0.0;
        if (c < ncols) p = exp((double)row[c] - m) * mul;
        const __half ph = __double2half(p);
        h[c] = ph;
        l[c] = __double2half(p - (double)__half2float(ph));
    }
}

__global__ void widen_rows_kernel(const float* __restrict__ in, int64_t ld, int64_t rows, int64_t cols, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols, c = i % cols;
    out[i] = (double)in[r * ld + c];
}

// fp64 re-evaluation of sample rows / sample densities (difference-form distances), compared with the tensor result.
// blocks [0, nrow_s): output rows of cells sample_cells[b]; blocks [nrow_s, nrow_s + ndens_s): densities of MNN cells.
// err[0] = max over sampled rows of max_g |ref - out| / max_g |ref| (double bits), err[1] = max |dens_ref - dens|.
__global__ void __launch_bounds__(1024)
smooth_check_kernel(const double* __restrict__ averaged, int64_t G, int64_t nmnn, const int32_t* __restrict__ index0,
                    const double* __restrict__ mat, int64_t Gd, int64_t ncells, double inv_sigma, const double* __restrict__ dens,
                    const double* __restrict__ out, int nrow_s, int ndens_s, double* __restrict__ wbuf /* [nrow_s + ndens_s][nmnn] */,
                    unsigned long long* __restrict__ err) {
    __shared__ double sm[32];
    const int b = blockIdx.x;
    const bool is_row = b < nrow_s;
    const int64_t cell = is_row ? (int64_t)((double)b * (double)(ncells - 1) / (double)max(nrow_s - 1, 1))
                                : (int64_t)index0[(int64_t)((double)(b - nrow_s) * (double)(nmnn - 1) / (double)max(ndens_s - 1, 1))];
    const int64_t dens_i = is_row ? -1 : (int64_t)((double)(b - nrow_s) * (double)(nmnn - 1) / (double)max(ndens_s - 1, 1));
    const double* x = mat + cell * Gd;
    double* w = wbuf + (int64_t)b * nmnn;
    double mloc = -INFINITY;
    for (int64_t i = threadIdx.x; i < nmnn; i += blockDim.x) {
        const double* y = mat + (int64_t)index0[i] * Gd;
        double d2 = 0.0;
        for (int64_t g = 0; g < Gd; ++g) { const double df = x[g] - y[g]; d2 = fma(df, df, d2); }
        const double l = -d2 * inv_sigma - (is_row ? dens[i] : 0.0);
        w[i] = l;
        mloc = fmax(mloc, l);
    }
    const double m = block_max(mloc, sm);
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nmnn; i += blockDim.x) s += exp(w[i] - m);
    s = block_sum(s, sm);
    if (!is_row) {
        if (threadIdx.x == 0) atomicMax(&err[1], (unsigned long long)__double_as_longlong(fabs((m + log(s)) - dens[dens_i])));
        return;
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < nmnn; i += blockDim.x) w[i] = exp(w[i] - m) / s;
    __syncthreads();
    double emax = 0.0, rmax = 0.0;
    for (int64_t g = threadIdx.x; g < G; g += blockDim.x) {
        double acc = 0.0;
        for (int64_t i = 0; i < nmnn; ++i) acc = fma(w[i], averaged[i * G + g], acc);
        emax = fmax(emax, fabs(acc - out[cell * G + g]));
        rmax = fmax(rmax, fabs(acc));
    }
    emax = block_max(emax, sm);
    rmax = block_max(rmax, sm);
    if (threadIdx.x == 0) {
        const double rel = (rmax > 0.0) ? emax / rmax : (emax > 0.0 ? INFINITY : 0.0);
        atomicMax(&err[0], (unsigned long long)__double_as_longlong(rel));
    }
}

static gemm::SplitMat row_view(const gemm::SplitMat& m, int64_t r0, int64_t rows) {
    gemm::SplitMat v = m;
    v.hi = m.hi + r0 * m.Kp;
    v.lo = m.lo + r0 * m.Kp;
    v.rows = rows;
    v.rows_pad = m.rows_pad - r0;
    return v;
}

// returns 0 and *accepted = 1 when the tensor result passed its fp64 sample check; *accepted = 0 asks for the fp64 path
static int smooth_tensor(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0, const double* d_mat, int64_t Gd,
                         int64_t ncells, double sigma2, double* d_out, cudaStream_t stream, int* accepted, double* check_err) {
    using namespace gemm;
    *accepted = 0;
    Scratch ws(stream);
    double* mean = ws.get<double>((size_t)Gd);
    double* norm_c = ws.get<double>((size_t)ncells);
    double* norm_m = ws.get<double>((size_t)nmnn);
    double* dens = ws.get<double>((size_t)nmnn);
    unsigned char* sc = ws.get<unsigned char>(64);
    if (!ws.ok()) return B200MNN_ENOMEM;
    unsigned int* amax = reinterpret_cast<unsigned int*>(sc);              // [0] cells, [1] averaged
    unsigned long long* nmax = reinterpret_cast<unsigned long long*>(sc + 8);   // [0] cells, [1] averaged
    unsigned long long* err = reinterpret_cast<unsigned long long*>(sc + 32);   // [2]
    B200_CUDA(cudaMemsetAsync(sc, 0, 64, stream));
    // distances are translation invariant: centring shrinks the norms the fp16 split has to carry
    B200_TRY(col_mean(d_mat, Gd, ncells, Gd, mean, stream));
    B200_TRY(row_stats(d_mat, Gd, ncells, Gd, mean, amax, nmax, stream));
    B200_TRY(row_stats(d_averaged, G, nmnn, G, nullptr, amax + 1, nmax + 1, stream));
    unsigned char h[64];
    B200_CUDA(cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    float am[2];
    double nm[2];
    memcpy(am, h, 8);
    memcpy(nm, h + 8, 16);
    if (!std::isfinite(am[0]) || !std::isfinite(am[1]) || !std::isfinite(nm[0])) return 0;   // non-finite input: fp64 path reproduces it
    const int e_m = pick_scale_exp(am[0], nm[0]);
    int e_v = 0;
    if (am[1] > 0.f) { int ex; frexpf(am[1], &ex); e_v = 12 - ex; }

    SplitMat Cop, Mop, Vop, Pop;
    B200_TRY(alloc_split(ws, ncells, Gd, &Cop));
    B200_TRY(alloc_split(ws, nmnn, Gd, &Mop));
    B200_TRY(alloc_split(ws, G, nmnn, &Vop));
    const int64_t ldL = round_up(nmnn, BN), ldG = round_up(G, BN);
    const int64_t budget = (int64_t)1 << 30;   // floats in the logit block (4 GiB)
    int64_t chunk = std::max<int64_t>(BM, std::min<int64_t>(round_up(std::max(ncells, nmnn), BM), (budget / ldL) / BM * BM));
    B200_TRY(alloc_split(ws, chunk, nmnn, &Pop));
    float* Lbuf = ws.get<float>((size_t)chunk * ldL);
    float* Obuf = ws.get<float>((size_t)chunk * ldG);
    if (!ws.ok()) return B200MNN_ENOMEM;
    B200_TRY(split_rows(d_mat, Gd, nullptr, ncells, Gd, mean, e_m, Cop, norm_c, stream));
    B200_TRY(split_rows(d_mat, Gd, d_index0, nmnn, Gd, mean, e_m, Mop, norm_m, stream));
    B200_TRY(split_transposed(d_averaged, G, G, nmnn, e_v, Vop, stream));

    // K boxes per tensor-core accumulation chain.  The accumulator truncates: a chain of n MMAs loses up to n ulp of its
    // running sum (measured: 6e-7 |a||b| for two boxes, half of that for one).  The logits are divided by sigma, so they
    // get the shortest chain; the weighted accumulate is an average of errors and runs two boxes per chain at full speed.
    const int chain_logit = 1, chain = 2;
    EpiArgs lg;
    lg.out = Lbuf; lg.ldo = ldL; lg.N = nmnn;
    lg.cold = norm_m;
    lg.beta = 2.0 * scalbn(1.0, -2 * e_m);
    lg.inv_sigma = 1.0 / sigma2;
    // pass 1: log-density of every MNN cell among the MNN cells (:56-65)
    for (int64_t r0 = 0; r0 < nmnn; r0 += chunk) {
        const int64_t nr = std::min(chunk, nmnn - r0);
        EpiArgs ep = lg;
        ep.M = nr; ep.rowd = norm_m + r0; ep.dens = nullptr;
        ep.diag = r0;   // an MNN cell is at distance exactly 0 from itself (the largest term of its density)
        B200_TRY(gemm_split(row_view(Mop, r0, nr), Mop, 3, EPI_LOGIT, ep, chain_logit, stream));
        row_lse_f32_kernel<<<(unsigned)nr, 256, 0, stream>>>(Lbuf, ldL, nmnn, dens + r0);
        B200_LAUNCH_CHECK();
    }
    // pass 2: weights of every cell and the weighted average of the correction vectors (:75-115)
    for (int64_t r0 = 0; r0 < ncells; r0 += chunk) {
        const int64_t nr = std::min(chunk, ncells - r0);
        EpiArgs ep = lg;
        ep.M = nr; ep.rowd = norm_c + r0; ep.dens = dens;
        B200_TRY(gemm_split(row_view(Cop, r0, nr), Mop, 3, EPI_LOGIT, ep, chain_logit, stream));
        row_softmax_split_kernel<<<(unsigned)nr, 256, 0, stream>>>(Lbuf, ldL, nmnn, Pop.hi, Pop.lo, Pop.Kp);
        B200_LAUNCH_CHECK();
        EpiArgs pv;
        pv.out = Obuf; pv.ldo = ldG; pv.M = nr; pv.N = G;
        pv.alpha = scalbn(1.0, -(P_EXP + e_v));
        B200_TRY(gemm_split(row_view(Pop, 0, nr), Vop, 3, EPI_PLAIN, pv, chain, stream));
        widen_rows_kernel<<<(unsigned)ceil_div(nr * G, 256), 256, 0, stream>>>(Obuf, ldG, nr, G, d_out + r0 * G);
        B200_LAUNCH_CHECK();
    }
    // fp64 sample check
    const int nrow_s = (int)std::min<int64_t>(24, ncells), ndens_s = (int)std::min<int64_t>(8, nmnn);
    double* wbuf = ws.get<double>((size_t)(nrow_s + ndens_s) * nmnn);
    if (!ws.ok()) return B200MNN_ENOMEM;
    smooth_check_kernel<<<nrow_s + ndens_s, 1024, 0, stream>>>(d_averaged, G, nmnn, d_index0, d_mat, Gd, ncells, 1.0 / sigma2, dens, d_out, nrow_s,
                                                             ndens_s, wbuf, err);
    B200_LAUNCH_CHECK();
    unsigned long long herr[2];
    B200_CUDA(cudaMemcpyAsync(herr, err, 16, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    double e0, e1;
    memcpy(&e0, &herr[0], 8);
    memcpy(&e1, &herr[1], 8);
    if (check_err) { check_err[0] = e0; check_err[1] = e1; }
    *accepted = (e0 <= 1e-5 && e1 <= 1e-5) ? 1 : 0;   // the contract's tolerance, measured on the sample
    return 0;
}

static double g_last_check[2] = {0.0, 0.0};
static int g_last_path = 0;   // 1 = tensor accepted, 2 = fp64, 3 = tensor rejected -> fp64

int smooth_gaussian_kernel_device(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0, const double* d_mat,
                                  int64_t Gdist, int64_t ncells, double sigma2, double* d_out, int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (G < 0 || nmnn < 0 || Gdist < 0 || ncells < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (ncells == 0 || G == 0) return 0;
    if (nmnn == 0) {  // reference: zero-initialised output times exp(-Inf - NA) (:105-115) = NaN everywhere
        B200_CUDA(cudaMemsetAsync(d_out, 0xFF, sizeof(double) * ncells * G, stream));  // all-ones bit pattern is a quiet NaN
        return 0;
    }
    // B200MNN_SMOOTH=fp64|tensor overrides; default: tensor cores once the contractions are worth a GEMM launch
    const char* env = getenv("B200MNN_SMOOTH");
    bool tensor = (double)ncells * (double)nmnn * (double)(Gdist + G) >= 2.0e10 && Gdist >= 1;
    if (env && strcmp(env, "fp64") == 0) tensor = false;
    if (env && strcmp(env, "tensor") == 0) tensor = Gdist >= 1;
    g_last_path = 2;
    if (tensor) {
        // the index vector is dereferenced by the operand kernels: validate it first (device-pointer callers included)
        Scratch ws(stream);
        int* bad = d_bad ? d_bad : ws.get<int>(1);
        if (!ws.ok()) return B200MNN_ENOMEM;
        if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
        check_index_kernel<<<(unsigned)ceil_div(nmnn, 256), 256, 0, stream>>>(d_index0, nmnn, ncells, bad);
        B200_LAUNCH_CHECK();
        int hb = 0;
        B200_CUDA(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
        B200_CUDA(cudaStreamSynchronize(stream));
        if (hb) return d_bad ? 0 : fail(B200MNN_EINVAL, "'index' entries out of range");
        int accepted = 0;
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        const bool dbg = getenv("B200MNN_SMOOTH_DEBUG") != nullptr;
        if (dbg) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, stream); }
        B200_TRY(smooth_tensor(d_averaged, G, nmnn, d_index0, d_mat, Gdist, ncells, sigma2, d_out, stream, &accepted, g_last_check));
        if (dbg) {
            float ms = 0.f;
            cudaEventRecord(t1, stream); cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1);
            cudaEventDestroy(t0); cudaEventDestroy(t1);
            fprintf(stderr, "b200mnn smoothing: tensor path %s in %.2f ms (sampled fp64 check: rows %.3g, log-density %.3g)\n",
                    accepted ? "accepted" : "REJECTED", ms, g_last_check[0], g_last_check[1]);
        }
        if (accepted) { g_last_path = 1; return 0; }
        g_last_path = 3;
    }
    return smooth_fp64(d_averaged, G, nmnn, d_index0, d_mat, Gdist, ncells, sigma2, d_out, d_bad, stream);
}

}  // namespace smooth
}  // namespace b200

extern "C" int b200mnn_smooth_last_check(int* path, double* row_err, double* dens_err) {
    if (path) *path = b200::smooth::g_last_path;
    if (row_err) *row_err = b200::smooth::g_last_check[0];
    if (dens_err) *dens_err = b200::smooth::g_last_check[1];
    return 0;
}

extern "C" int b200mnn_dev_smooth_gaussian_kernel(const double* d_averaged, int64_t G, int64_t nmnn, const int32_t* d_index0,
                                                  const double* d_mat, int64_t Gdist, int64_t ncells, double sigma2, double* d_out,
                                                  void* stream) {
    return b200::smooth::smooth_gaussian_kernel_device(d_averaged, G, nmnn, d_index0, d_mat, Gdist, ncells, sigma2, d_out, nullptr,
                                                       static_cast<cudaStream_t>(stream));
}
