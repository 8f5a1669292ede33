// K-streamed split-fp16 GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA) with fused epilogues -- the shared
// engine of the gene-space kernels: Gaussian smoothing (distance logits and the weighted accumulate,
// src/smooth_gaussian_kernel.cpp:32-115) and the wide-d exact kNN of mnnCorrect (R/mnnCorrect.R:288-289).
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace b200 {
namespace gemm {

constexpr int BM = 128;    // rows of A per tile (TMEM lanes)
constexpr int BN = 256;    // rows of B per tile (accumulator columns)
constexpr int KBOX = 64;   // fp16 columns per TMA box (128 bytes, one SWIZZLE_128B atom row)

// A matrix stored as two fp16 planes, x * scale = hi + lo + O(2^-22 |x * scale|), K-major rows of Kp = round_up(K, 64)
// columns, rows padded to `rows_pad` (a multiple of BN); padding columns are zero.
struct SplitMat {
    __half* hi = nullptr;
    __half* lo = nullptr;
    int64_t rows = 0, rows_pad = 0, K = 0, Kp = 0;
};

static inline int64_t pad_rows(int64_t rows) { return round_up(rows < 1 ? 1 : rows, BN); }
static inline int64_t pad_k(int64_t K) { return round_up(K < 1 ? 1 : K, KBOX); }

enum Epilogue : int {
    EPI_PLAIN = 0,   // out = alpha * acc
    EPI_SCORE = 1,   // out = colf[c] + alpha * acc                           (kNN score S^2 (||x||^2 - 2 q.x), alpha = -2)
    EPI_LOGIT = 2,   // out = (beta * acc - rowd[r] - cold[c]) * inv_sigma - dens[c]   (Gaussian log-weight minus log-density)
};

struct EpiArgs {
    float* out = nullptr;       // [M x ldo] row-major
    int64_t ldo = 0, M = 0, N = 0;
    double alpha = 1.0;
    const float* colf = nullptr;
    const double* rowd = nullptr;
    const double* cold = nullptr;
    const double* dens = nullptr;   // may be null (treated as 0)
    double beta = 1.0, inv_sigma = 1.0;
    int64_t diag = -1;              // EPI_LOGIT: column `diag + r` of row r is the row itself (distance exactly 0); -1: none
};

// D[M x N] = sum over the schedule's terms of A_t . B_t^T, fp32 accumulation in TMEM, re-based into registers every
// `chunk_boxes` K boxes (bounds the length of any single fp32 accumulation chain), then the epilogue.
//   terms == 3: Ah.Bh + Al.Bh + Ah.Bl (error ~2^-22 |a||b|);  terms == 1: Ah.Bh (error ~2^-10 |a||b|).
// Asynchronous on `stream`.
int gemm_split(const SplitMat& A, const SplitMat& B, int terms, int epilogue, const EpiArgs& ep, int chunk_boxes, cudaStream_t stream);

// fp64 rows [rows x K] (row i of the output = row (gather ? gather[i] : i) of X, minus `centre` when given) -> split
// planes scaled by 2^scale_exp; norm2 (optional) receives the fp64 squared norm of the centred row.  Rows >= rows and
// columns >= K are zero-filled.
int split_rows(const double* X, int64_t ldx, const int32_t* gather, int64_t rows, int64_t K, const double* centre, int scale_exp,
               const SplitMat& out, double* norm2, cudaStream_t stream);
// Transposing variant: X [K x rows_out ...] i.e. out row r, column k = X[k * ldx + r] (used for V^T in the accumulate).
int split_transposed(const double* X, int64_t ldx, int64_t rows, int64_t K, int scale_exp, const SplitMat& out, cudaStream_t stream);

int alloc_split(Scratch& ws, int64_t rows, int64_t K, SplitMat* out);

// Column means of X [rows x K] (fp64) into mean[K] (zeroed here).
int col_mean(const double* X, int64_t ldx, int64_t rows, int64_t K, double* mean, cudaStream_t stream);
// max |x - centre| (float bits, rounded up) and max squared norm of the centred rows (double bits) via atomicMax; the two
// words must be zeroed by the caller (several matrices can share them).
int row_stats(const double* X, int64_t ldx, int64_t rows, int64_t K, const double* centre, unsigned int* absmax_bits,
              unsigned long long* maxnorm2_bits, cudaStream_t stream);
// power-of-two exponent e such that |x| 2^e < 2^13 for every element and ||x|| 2^e < 2^15 for every row
int pick_scale_exp(float absmax, double maxnorm2);

// Timing hook (bench.py): CUDA events around gemm_split launches while enabled.
int profile_enable(int on);
int profile_collect(double* total_ms, int64_t* launches, double* executed_flops);

}  // namespace gemm
}  // namespace b200
