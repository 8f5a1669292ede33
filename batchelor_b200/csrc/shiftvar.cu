// mnnCorrect's shift-variance adjustment for sm_100a -- replaces adjust_shift_variance()
// (src/adjust_shift_variance.cpp:9-164, .Call _batchelor_adjust_shift_variance at src/RcppExports.cpp:10-22).
//
// Per cell c of batch 2 (one thread block per cell):
//   g        = vect[c,] / ||vect[c,]||   (left un-normalised when the norm is 0, :62-68)
//   curproj  = g . x_c                                                      (:70)
//   own batch (restrict2): lw_j = -dist_to_line(x_c, g, x_j)^2 / sigma2 (:9-27, :88-90), self term fixed at 0 (:84);
//            prob2 = logsumexp{lw_j : j == c or g.x_j <= curproj} - logsumexp{lw_j}          (:91-111)
//   reference batch (restrict1): pairs (g.y_o, lw_o), sorted lexicographically (:133); walk the running
//            logsumexp until it reaches prob2 + logsumexp{lw_o}; that projection is the quantile (:139-157)
//   out[c]   = (quantile - curproj) / ||vect[c,]||                          (:160)
// Round-1 implementation: fp64 CUDA cores, one warp per comparison cell (lanes over genes, fixed shuffle tree so
// that duplicated cells give bit-identical projections and the `sameproj > curproj` test of :91 behaves), block
// bitonic sort of the (projection, logweight) pairs, chunked logsumexp scan for the quantile walk.
// Layout: data1 [n1 x G], data2 [n2 x G], vect [n2 x G], row-major (one cell contiguous).
#include "common.cuh"

namespace b200 {
namespace shiftvar {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

struct LSE {  // running log-sum-exp as (max, sum of exp(x - max)); empty when m == -inf and s == 0
    double m, s;
};
__device__ __forceinline__ void lse_add(LSE& a, double x) {
    if (x == -INFINITY && a.m == -INFINITY) { a.s += 1.0; return; }  // exp(-inf - -inf) convention: count it, value stays -inf
    if (x > a.m) { a.s = a.s * exp(a.m - x) + 1.0; a.m = x; }
    else a.s += exp(x - a.m);
}
__device__ __forceinline__ void lse_merge(LSE& a, const LSE& b) {
    if (b.s == 0.0) return;
    if (a.s == 0.0) { a = b; return; }
    if (b.m > a.m) { a.s = a.s * exp(a.m - b.m) + b.s; a.m = b.m; }
    else if (a.m == b.m) a.s += b.s;
    else a.s += b.s * exp(b.m - a.m);
}
__device__ __forceinline__ double lse_value(const LSE& a) { return a.m + log(a.s); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// projection of x on g and squared distance of x from the line through cur with direction g (warp-cooperative)
__device__ __forceinline__ void proj_and_dist(const double* __restrict__ x, const double* __restrict__ cur, const double* __restrict__ g,
                                              int64_t G, int lane, double& proj, double& dist) {
    double p = 0.0, sc = 0.0;
    for (int64_t t = lane; t < G; t += 32) {
        const double xv = x[t], gv = g[t];
        p += gv * xv;
        sc += (cur[t] - xv) * gv;
    }
    p = warp_sum(p);
    sc = warp_sum(sc);
    double ds = 0.0;
    for (int64_t t = lane; t < G; t += 32) {
        const double w = (cur[t] - x[t]) - sc * g[t];
        ds += w * w;
    }
    proj = p;
    dist = warp_sum(ds);
}

__device__ __forceinline__ bool pl_less(double pa, double la, double pb, double lb) { return pa < pb || (pa == pb && la < lb); }

__global__ void __launch_bounds__(THREADS)
shift_variance_kernel(const double* __restrict__ data1, int64_t n1, const double* __restrict__ data2, int64_t n2, int64_t G,
                      const double* __restrict__ vect, double sigma2, const int32_t* __restrict__ r1, int64_t nr1,
                      const int32_t* __restrict__ r2, int64_t nr2, int64_t nr1_pow2, double* __restrict__ scratch_proj,
                      double* __restrict__ scratch_lw, double* __restrict__ out) {
    extern __shared__ double sh[];
    double* g = sh;          // [G]
    double* cur = sh + G;    // [G]
    __shared__ double red[THREADS];
    __shared__ LSE wl_prob[WARPS], wl_tot[WARPS], wl_tot1[WARPS];
    __shared__ double s_l2, s_curproj, s_prob2, s_tot1;
    __shared__ LSE chunk_lse[THREADS];
    __shared__ long long s_first;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* P = scratch_proj + (int64_t)blockIdx.x * nr1_pow2;
    double* W = scratch_lw + (int64_t)blockIdx.x * nr1_pow2;

    for (int64_t c = blockIdx.x; c < n2; c += gridDim.x) {
        __syncthreads();
        // ---- unit gradient and the cell itself ----
        double part = 0.0;
        for (int64_t t = threadIdx.x; t < G; t += THREADS) {
            const double v = vect[c * G + t];
            g[t] = v;
            cur[t] = data2[c * G + t];
            part += v * v;
        }
        red[threadIdx.x] = part;
        __syncthreads();
        for (int o = THREADS / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) s_l2 = sqrt(red[0]);
        __syncthreads();
        const double l2 = s_l2;
        if (l2 != 0.0)
            for (int64_t t = threadIdx.x; t < G; t += THREADS) g[t] /= l2;
        __syncthreads();
        if (warp == 0) {  // same routine as for every other cell, so duplicates of c project identically
            double p, dd;
            proj_and_dist(cur, cur, g, G, lane, p, dd);
            if (lane == 0) s_curproj = p;
        }
        __syncthreads();
        const double curproj = s_curproj;

        // ---- own batch ----
        LSE lp = {-INFINITY, 0.0}, lt = {-INFINITY, 0.0};
        for (int64_t s = warp; s < nr2; s += WARPS) {
            const int64_t same = r2[s];
            bool add = true;
            double logp = 0.0;
            if (same != c) {
                double p, dd;
                proj_and_dist(data2 + same * G, cur, g, G, lane, p, dd);
                logp = -dd / sigma2;
                if (p > curproj) add = false;
            }
            if (add) lse_add(lp, logp);
            lse_add(lt, logp);
        }
        if (lane == 0) { wl_prob[warp] = lp; wl_tot[warp] = lt; }

        // ---- reference batch: projections and log-weights ----
        LSE l1 = {-INFINITY, 0.0};
        for (int64_t o = warp; o < nr1; o += WARPS) {
            double p, dd;
            proj_and_dist(data1 + (int64_t)r1[o] * G, cur, g, G, lane, p, dd);
            const double lw = -dd / sigma2;
            if (lane == 0) { P[o] = p; W[o] = lw; }
            lse_add(l1, lw);
        }
        if (lane == 0) wl_tot1[warp] = l1;
        for (int64_t o = nr1 + threadIdx.x; o < nr1_pow2; o += THREADS) { P[o] = INFINITY; W[o] = INFINITY; }  // padding sorts last
        __syncthreads();
        if (threadIdx.x == 0) {
            LSE a = wl_prob[0], b = wl_tot[0], e = wl_tot1[0];
            for (int w = 1; w < WARPS; ++w) { lse_merge(a, wl_prob[w]); lse_merge(b, wl_tot[w]); lse_merge(e, wl_tot1[w]); }
            // prob2 starts at 0 in the reference when nothing was added (:75); totals likewise
            const double pa = (a.s == 0.0) ? 0.0 : lse_value(a);
            const double pb = (b.s == 0.0) ? 0.0 : lse_value(b);
            s_prob2 = pa - pb;
            s_tot1 = (e.s == 0.0) ? 0.0 : lse_value(e);
        }
        __threadfence_block();
        __syncthreads();

        // ---- bitonic sort of (P, W) ascending, lexicographic ----
        for (int64_t size = 2; size <= nr1_pow2; size <<= 1) {
            for (int64_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (int64_t i = threadIdx.x; i < nr1_pow2 / 2; i += THREADS) {
                    const int64_t lo = (i / stride) * (stride * 2) + (i % stride);
                    const int64_t hi = lo + stride;
                    const bool up = ((lo & size) == 0);
                    const double pa = P[lo], wa = W[lo], pb = P[hi], wb = W[hi];
                    const bool swap = up ? pl_less(pb, wb, pa, wa) : pl_less(pa, wa, pb, wb);
                    if (swap) { P[lo] = pb; W[lo] = wb; P[hi] = pa; W[hi] = wa; }
                }
                __syncthreads();
            }
        }

        // ---- quantile walk: first sorted position whose running logsumexp reaches target ----
        double refq = NAN;
        if (nr1 > 0) {
            const double target = s_prob2 + s_tot1;
            const int64_t per = (nr1 + THREADS - 1) / THREADS;
            const int64_t b0 = (int64_t)threadIdx.x * per, b1 = min(nr1, b0 + per);
            LSE mine = {-INFINITY, 0.0};
            for (int64_t o = b0; o < b1; ++o) lse_add(mine, W[o]);
            chunk_lse[threadIdx.x] = mine;
            if (threadIdx.x == 0) s_first = (long long)nr1;  // "not found" -> default: the largest projection (:143)
            __syncthreads();
            LSE prefix = {-INFINITY, 0.0};
            for (int t = 0; t < (int)threadIdx.x; ++t) lse_merge(prefix, chunk_lse[t]);
            long long found = -1;
            for (int64_t o = b0; o < b1; ++o) {
                lse_add(prefix, W[o]);
                if (lse_value(prefix) >= target) { found = o; break; }
            }
            if (found >= 0) atomicMin(&s_first, found);
            __syncthreads();
            const long long first = s_first;
            refq = (first < (long long)nr1) ? P[first] : P[nr1 - 1];
        }
        if (threadIdx.x == 0) out[c] = (refq - curproj) / l2;
    }
}

__global__ void check_restrict_kernel(const int32_t* __restrict__ r, int64_t n, int64_t limit, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (r[i] == INT32_MIN || r[i] < 0 || r[i] >= limit)) *bad = 1;
}

int adjust_shift_variance_device(const double* d_data1, int64_t n1, const double* d_data2, int64_t n2, int64_t G, const double* d_vect,
                                 double sigma2, const int32_t* d_r1, int64_t nr1, const int32_t* d_r2, int64_t nr2, double* d_out,
                                 int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (n1 < 0 || n2 < 0 || G < 0 || nr1 < 0 || nr2 < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (n2 == 0) return 0;
    Scratch ws(stream);
    int64_t p2 = 2;
    while (p2 < nr1) p2 <<= 1;
    const int grid = (int)std::min<int64_t>(n2, (int64_t)sm_count() * 4);
    double* sp = ws.get<double>((size_t)grid * p2);
    double* sw = ws.get<double>((size_t)grid * p2);
    int* bad = d_bad ? d_bad : ws.get<int>(1);
    if (!ws.ok()) return B200MNN_ENOMEM;
    if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    if (nr1 > 0) { check_restrict_kernel<<<(unsigned)ceil_div(nr1, 256), 256, 0, stream>>>(d_r1, nr1, n1, bad); B200_LAUNCH_CHECK(); }
    if (nr2 > 0) { check_restrict_kernel<<<(unsigned)ceil_div(nr2, 256), 256, 0, stream>>>(d_r2, nr2, n2, bad); B200_LAUNCH_CHECK(); }
    const size_t smem = (size_t)2 * std::max<int64_t>(G, 1) * sizeof(double);
    int dev = 0, max_smem = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem + 16 * 1024 > (size_t)max_smem) return fail(B200MNN_EINVAL, "too many genes for the shift-variance kernel's shared-memory staging");
    B200_CUDA(cudaFuncSetAttribute(shift_variance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    shift_variance_kernel<<<grid, THREADS, smem, stream>>>(d_data1, n1, d_data2, n2, G, d_vect, sigma2, d_r1, nr1, d_r2, nr2, p2, sp, sw, d_out);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace shiftvar
}  // namespace b200

extern "C" int b200mnn_dev_adjust_shift_variance(const double* d_data1, int64_t n1, const double* d_data2, int64_t n2, int64_t G,
                                                 const double* d_vect, double sigma2, const int32_t* d_r1, int64_t nr1,
                                                 const int32_t* d_r2, int64_t nr2, double* d_out, void* stream) {
    return b200::shiftvar::adjust_shift_variance_device(d_data1, n1, d_data2, n2, G, d_vect, sigma2, d_r1, nr1, d_r2, nr2, d_out, nullptr,
                                                        static_cast<cudaStream_t>(stream));
}
