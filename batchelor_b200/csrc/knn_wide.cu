// Exact kNN for the shapes the resident-operand kernels of knn_tc.cu do not cover: wide data (mnnCorrect searches in
// gene space, d ~ 2000: R/mnnCorrect.R:288-289 -> R/MNN_tree.R:129) and large k (prop.k, R/MNN_tree.R:140-146).
// Same contract as every search of this library: the k smallest (squared distance accumulated in double in dimension
// order, index) pairs, ascending -- what BiocNeighbors::queryKNN(..., KmknnParam()) returns.
//
//   1. operands    column mean removed (distances are translation invariant; smaller norms = smaller scoring error),
//                  one power-of-two scale, fp16 hi/lo planes (gemm_tc.cu)
//   2. scores      three-term split-fp16 tcgen05 GEMM, K streamed: score(q, j) = S^2 (||x_j||^2 - 2 q.x_j) in fp32 for a
//                  chunk of queries against all references, written to HBM.  With d ~ 2000 a score costs 12 000 tensor
//                  flops and 4 bytes, so the selection is a separate HBM-bound pass instead of a fused epilogue.
//   3. select      one CTA per query: three-pass radix select of the KEEP-th smallest score (11 + 11 + 10 bits, integer
//                  shared-memory histograms), then collection of the KEEP best; threshold = that score.
//   4. re-rank     exact fp64 distances of the KEEP candidates in the reference's summation order, bitonic sort on
//                  (distance, index), first k out, and the CERTIFICATE d2_k < threshold - eps (eps bounds the scoring
//                  error, derivation in DESIGN.md section 4.1).  A certified result is provably the exact answer.
//   5. rescue      uncertified queries (ties across the threshold) go through the exact scan of knn_tc.cu.
#include <cstring>

#include "gemm_tc.cuh"
#include "internal.cuh"

namespace b200 {
namespace knn {

constexpr int WIDE_MAX_KEEP = 2048;

__global__ void score_norm_kernel(const double* __restrict__ norm2, int64_t n, double s2, float* __restrict__ colf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) colf[i] = (float)(norm2[i] * s2);
}

__device__ __forceinline__ uint32_t ord_key(float s) {
    const uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// All threads: the first bin whose inclusive prefix count reaches `rank` (1-based) and the rank inside that bin.
// hist has NB entries (NB <= 2048, multiple of 256); scratch holds 8 ints.
template <int NB>
__device__ __forceinline__ void find_rank_bin(const int* __restrict__ hist, int rank, int* __restrict__ scratch, int& bin, int& rank_in_bin) {
    constexpr int PER = NB / 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int loc[PER];
    int sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { loc[j] = hist[threadIdx.x * PER + j]; sum += loc[j]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __syncthreads();
    if (lane == 31) scratch[warp] = inc;
    if (threadIdx.x == 0) { scratch[8] = -1; scratch[9] = 0; }
    __syncthreads();
    int before = inc - sum;
    for (int w = 0; w < warp; ++w) before += scratch[w];
    if (before < rank && before + sum >= rank) {   // exactly one thread
        int run = before;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (run < rank && run + loc[j] >= rank) { scratch[8] = threadIdx.x * PER + j; scratch[9] = rank - run; }
            run += loc[j];
        }
    }
    __syncthreads();
    bin = scratch[8];
    rank_in_bin = scratch[9];
}

// One CTA per query row: the KEEP smallest scores of the row (ids, unsorted) and the KEEP-th smallest score.
__global__ void __launch_bounds__(256)
wide_select_kernel(const float* __restrict__ S, int64_t ld, int64_t n, int keep, int32_t* __restrict__ cand, float* __restrict__ thr) {
    __shared__ int hist[2048];
    __shared__ int scratch[16];
    __shared__ int s_cnt, s_ties;
    const float* row = S + (int64_t)blockIdx.x * ld;
    int32_t* out = cand + (int64_t)blockIdx.x * keep;
    if (n <= keep) {
        for (int64_t c = threadIdx.x; c < keep; c += 256) out[c] = (c < n) ? (int32_t)c : -1;
        if (threadIdx.x == 0) thr[blockIdx.x] = __int_as_float(0x7f800000);
        return;
    }
    int b1, r1, b2, r2, b3, r3;
    for (int i = threadIdx.x; i < 2048; i += 256) hist[i] = 0;
    __syncthreads();
    for (int64_t c = threadIdx.x; c < n; c += 256) atomicAdd(&hist[ord_key(row[c]) >> 21], 1);
    __syncthreads();
    find_rank_bin<2048>(hist, keep, scratch, b1, r1);
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += 256) hist[i] = 0;
    __syncthreads();
    for (int64_t c = threadIdx.x; c < n; c += 256) {
        const uint32_t k = ord_key(row[c]);
        if ((int)(k >> 21) == b1) atomicAdd(&hist[(k >> 10) & 0x7FFu], 1);
    }
    __syncthreads();
    find_rank_bin<2048>(hist, r1, scratch, b2, r2);
    __syncthreads();
    const uint32_t pre = ((uint32_t)b1 << 11) | (uint32_t)b2;
    for (int i = threadIdx.x; i < 1024; i += 256) hist[i] = 0;
    __syncthreads();
    for (int64_t c = threadIdx.x; c < n; c += 256) {
        const uint32_t k = ord_key(row[c]);
        if ((k >> 10) == pre) atomicAdd(&hist[k & 0x3FFu], 1);
    }
    __syncthreads();
    find_rank_bin<1024>(hist, r2, scratch, b3, r3);
    const uint32_t kth = (pre << 10) | (uint32_t)b3;
    if (threadIdx.x == 0) { s_cnt = 0; s_ties = 0; }
    __syncthreads();
    // everything strictly below the KEEP-th key (fewer than KEEP entries), then ties with it until the list is full
    for (int64_t c = threadIdx.x; c < n; c += 256) {
        const uint32_t k = ord_key(row[c]);
        if (k < kth) out[atomicAdd(&s_cnt, 1)] = (int32_t)c;
    }
    __syncthreads();
    const int less = s_cnt;
    for (int64_t c = threadIdx.x; c < n; c += 256) {
        const uint32_t k = ord_key(row[c]);
        if (k == kth) {
            const int t = atomicAdd(&s_ties, 1);
            if (less + t < keep) out[less + t] = (int32_t)c;
        }
    }
    if (threadIdx.x == 0) thr[blockIdx.x] = key_float(kth);
}

__device__ __forceinline__ bool pair_less_w(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

// One CTA (4 warps) per query: exact distances of the candidates, sort, first k, certificate.
__global__ void __launch_bounds__(128)
wide_rerank_kernel(const double* __restrict__ X, const double* __restrict__ Q, int64_t q0, int64_t nq, int d, int k, int keep,
                   const int32_t* __restrict__ cand, const float* __restrict__ thr, const double* __restrict__ qnorm2 /* centred */,
                   double inv_s2, double eps_q /* per unit |q| */, double eps_0, int32_t* __restrict__ out_idx,
                   double* __restrict__ out_dist, int* __restrict__ flag_count, int32_t* __restrict__ flag_list, double* __restrict__ flag_dk2) {
    extern __shared__ __align__(16) unsigned char wsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t jq = blockIdx.x;          // row of the chunk
    const int64_t q = q0 + jq;
    if (jq >= nq) return;
    const double* qv = Q + q * d;
    int np2 = 32;
    while (np2 < keep) np2 <<= 1;
    const int nthreads = (int)blockDim.x;   // min(128, np2): no idle warps for the usual 64 candidates
    // dynamic shared memory sized to the launch: [warps][32][17] staging + np2 (distance, id) pairs -> ~9.5 KB for 64
    // candidates, so that two dozen CTAs share an SM (the gather is latency-bound: occupancy is what hides it)
    double (*stage)[32][17] = reinterpret_cast<double (*)[32][17]>(wsm);
    double* cd = reinterpret_cast<double*>(wsm) + (size_t)(nthreads >> 5) * 32 * 17;   // [np2]
    int* ci = reinterpret_cast<int*>(cd + np2);                                        // [np2]
    for (int base = warp * 32; base < np2; base += nthreads) {
        const int c = base + lane;
        const int id = (c < keep) ? cand[jq * keep + c] : -1;
        double acc = 0.0;
        for (int t0 = 0; t0 < d; t0 += 16) {
            const int len = min(16, d - t0);
            const int sub = lane & 15, hw = lane >> 4;
            double v[16];   // 16 independent loads in flight per lane (the gather is latency-bound)
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
                const int rid = __shfl_sync(0xffffffffu, id, 2 * rr + hw);
                v[rr] = (rid >= 0 && sub < len) ? X[(int64_t)rid * d + t0 + sub] : 0.0;
            }
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) stage[warp][2 * rr + hw][sub] = v[rr];
            __syncwarp();
            for (int t = 0; t < len; ++t) {
                const double df = __dsub_rn(qv[t0 + t], stage[warp][lane][t]);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
            __syncwarp();
        }
        cd[c] = (id >= 0) ? acc : INFINITY;
        ci[c] = (id >= 0) ? id : 0x7fffffff;
    }
    __syncthreads();
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < np2 / 2; i += nthreads) {
                const int a = (i / stride) * (stride * 2) + (i % stride);
                const int b = a + stride;
                const bool up = ((a & size) == 0);
                const double da = cd[a], db = cd[b];
                const int ia = ci[a], ib = ci[b];
                const bool swap = up ? pair_less_w(db, ib, da, ia) : pair_less_w(da, ia, db, ib);
                if (swap) { cd[a] = db; ci[a] = ib; cd[b] = da; ci[b] = ia; }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < k; j += nthreads) {
        out_idx[q * k + j] = (ci[j] == 0x7fffffff) ? -1 : ci[j];
        if (out_dist) out_dist[q * k + j] = sqrt(cd[j]);
    }
    if (threadIdx.x == 0) {
        const float t = thr[jq];
        bool ok = true;
        const double dk = cd[k - 1];
        if (t < __int_as_float(0x7f800000)) {
            const double qn = qnorm2[q];
            const double bound = (double)t * inv_s2 + qn - (eps_q * sqrt(qn) + eps_0);
            ok = dk < bound;
        }
        if (!ok) {
            const int slot = atomicAdd(flag_count, 1);
            flag_list[slot] = (int32_t)q;
            flag_dk2[slot] = dk;
        }
    }
}

bool wide_path_supported(int64_t n, int64_t nq, int d, int k) {
    if (d < 1 || k < 1 || n < 1 || nq < 1) return false;
    if (n > (int64_t)INT32_MAX - 512 || nq > (int64_t)INT32_MAX - 512) return false;
    const int keep = (k <= 24) ? 64 : (int)round_up(k + std::max(32, k / 2), 32);
    return keep <= WIDE_MAX_KEEP;
}

int query_knn_wide(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                   int64_t* d_stats, cudaStream_t stream) {
    using namespace gemm;
    const int keep = (k <= 24) ? 64 : (int)round_up(k + std::max(32, k / 2), 32);
    Scratch ws(stream);
    double* mean = ws.get<double>((size_t)d);
    double* xnorm = ws.get<double>((size_t)n);
    double* qnorm = ws.get<double>((size_t)nq);
    float* colf = ws.get<float>((size_t)n);
    unsigned char* sc = ws.get<unsigned char>(64);
    int32_t* flag_list = ws.get<int32_t>((size_t)nq);
    double* flag_dk2 = ws.get<double>((size_t)nq);
    if (!ws.ok()) return B200MNN_ENOMEM;
    unsigned int* amax = reinterpret_cast<unsigned int*>(sc);
    unsigned long long* nmax = reinterpret_cast<unsigned long long*>(sc + 8);   // [0]: X, [1]: Q
    int* flag_count = reinterpret_cast<int*>(sc + 32);
    B200_CUDA(cudaMemsetAsync(sc, 0, 64, stream));
    B200_TRY(col_mean(dX, d, n, d, mean, stream));
    B200_TRY(row_stats(dX, d, n, d, mean, amax, nmax, stream));
    B200_TRY(row_stats(dQ, d, nq, d, mean, amax, nmax + 1, stream));
    unsigned char h[64];
    B200_CUDA(cudaMemcpyAsync(h, sc, 64, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    float am;
    double nm[2];
    memcpy(&am, h, 4);
    memcpy(nm, h + 8, 16);
    if (!std::isfinite(am) || !std::isfinite(nm[0]) || !std::isfinite(nm[1])) return launch_rescue_all(dX, n, dQ, nq, d, k, d_idx, d_dist, d_stats, stream);
    const int e = pick_scale_exp(am, std::max(nm[0], nm[1]));
    const double s2 = scalbn(1.0, 2 * e);

    SplitMat Xop, Qop;
    B200_TRY(alloc_split(ws, n, d, &Xop));
    B200_TRY(alloc_split(ws, nq, d, &Qop));
    const int64_t ldS = round_up(n, BN);
    const int64_t budget = (int64_t)1 << 30;   // floats in the score block (4 GiB)
    const int64_t chunk = std::max<int64_t>(BM, std::min<int64_t>(round_up(nq, BM), (budget / ldS) / BM * BM));
    float* Sbuf = ws.get<float>((size_t)chunk * ldS);
    int32_t* cand = ws.get<int32_t>((size_t)chunk * keep);
    float* thr = ws.get<float>((size_t)chunk);
    if (!ws.ok()) return B200MNN_ENOMEM;
    B200_TRY(split_rows(dX, d, nullptr, n, d, mean, e, Xop, xnorm, stream));
    B200_TRY(split_rows(dQ, d, nullptr, nq, d, mean, e, Qop, qnorm, stream));
    score_norm_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(xnorm, n, s2, colf);
    B200_LAUNCH_CHECK();

    // scoring error bound (unscaled squared-distance units), DESIGN.md section 4.1: with a = S q, b = S x (centred),
    // |score/S^2 + ||q||^2 - d^2| <= 2 c1 |q||x| + 2^-23 (||x||^2 + 2 |q||x|) + fp64 slack, where
    //   c1 = 3 * 2^-22 (dropped lo.lo and split residuals) + 2^-20 (products of one MMA) + chain * 2^-23 (truncating
    //        accumulation, `chain` MMAs per TMEM chain) + nchunks * 2^-24 (register re-basing, round to nearest)
    const int chain_boxes = 2;
    const double nchunks = (double)ceil_div(Xop.Kp / KBOX, chain_boxes);
    const double c1 = 3.0 * 2.384185791015625e-07 + 9.5367431640625e-07 + 12.0 * chain_boxes * 1.1920928955078125e-07 + nchunks * 5.9604644775390625e-08;
    const double M = sqrt(nm[0]);
    const double eps_q = 2.0 * (2.0 * c1 * M) + 2.0 * 1.1920928955078125e-07 * 2.0 * M;   // safety factor 2 on the derived bound
    const double eps_0 = 2.0 * 1.1920928955078125e-07 * nm[0] + 1e-12 * (nm[0] + nm[1]);

    int rr_threads = 32, rr_np2 = 32;
    while (rr_threads < keep && rr_threads < 128) rr_threads <<= 1;
    while (rr_np2 < keep) rr_np2 <<= 1;
    const size_t rr_smem = (size_t)(rr_threads / 32) * 32 * 17 * sizeof(double) + (size_t)rr_np2 * (sizeof(double) + sizeof(int));
    B200_CUDA(cudaFuncSetAttribute(wide_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem));
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t nr = std::min(chunk, nq - q0);
        SplitMat Qv = Qop;
        Qv.hi = Qop.hi + q0 * Qop.Kp;
        Qv.lo = Qop.lo + q0 * Qop.Kp;
        Qv.rows = nr;
        Qv.rows_pad = Qop.rows_pad - q0;
        EpiArgs ep;
        ep.out = Sbuf; ep.ldo = ldS; ep.M = nr; ep.N = n;
        ep.colf = colf;
        ep.alpha = -2.0;
        B200_TRY(gemm_split(Qv, Xop, 3, EPI_SCORE, ep, chain_boxes, stream));
        wide_select_kernel<<<(unsigned)nr, 256, 0, stream>>>(Sbuf, ldS, n, keep, cand, thr);
        B200_LAUNCH_CHECK();
        wide_rerank_kernel<<<(unsigned)nr, rr_threads, rr_smem, stream>>>(dX, dQ, q0, nr, d, k, keep, cand, thr, qnorm, 1.0 / s2, eps_q, eps_0, d_idx, d_dist,
                                                                  flag_count, flag_list, flag_dk2);
        B200_LAUNCH_CHECK();
    }
    B200_TRY(launch_rescue(dX, n, dQ, nq, d, k, flag_count, flag_list, flag_dk2, d_idx, d_dist, stream));
    if (d_stats) B200_TRY(write_stats(flag_count, d_stats, 1, 3, stream));
    return 0;
}

}  // namespace knn
}  // namespace b200
