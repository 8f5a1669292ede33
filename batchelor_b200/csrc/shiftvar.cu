// mnnCorrect's shift-variance adjustment for sm_100a -- replaces adjust_shift_variance()
// (src/adjust_shift_variance.cpp:9-164, .Call _batchelor_adjust_shift_variance at src/RcppExports.cpp:10-22).
//
// Per cell c of batch 2:
//   g        = vect[c,] / ||vect[c,]||   (left un-normalised when the norm is 0, :62-68)
//   curproj  = g . x_c                                                      (:70)
//   own batch (restrict2): lw_j = -dist_to_line(x_c, g, x_j)^2 / sigma2 (:9-27, :88-90), self term fixed at 0 (:84);
//            prob2 = logsumexp{lw_j : j == c or g.x_j <= curproj} - logsumexp{lw_j}          (:91-111)
//   reference batch (restrict1): pairs (g.y_o, lw_o), sorted lexicographically (:133); walk the running
//            logsumexp until it reaches prob2 + logsumexp{lw_o}; that projection is the quantile (:139-157)
//   out[c]   = (quantile - curproj) / ||vect[c,]||                          (:160)
//
// The result is a DISCRETE pick (which reference cell is the quantile) followed by exact arithmetic on that cell, so it
// is reproduced bit for bit when (i) every projection / distance is accumulated in the reference's own order --
// sequentially over the genes, separate multiply and add, as std::inner_product and sq_distance_to_line do -- and
// (ii) the pick is the same.  The work is O(n2 (n1 + n2) G) fp64 operations; fp32-class tensor-core scoring cannot
// decide the pick (neighbouring cumulative weights differ by ~1/n1 relative), so this is an
// fp64 CUDA-core path:
//   sv_prep_kernel    one warp per cell: ||vect[c,]||, g, curproj in the reference's sequential order.
//   sv_pairs_kernel   GEMM-shaped tile kernel (64 cells x 64 comparison cells per CTA, 4 x 4 pairs per thread, genes
//                     staged through shared memory in chunks of 16).  Every pair keeps its own sequential-in-gene
//                     accumulators, so tiling does not change any rounding:
//                       EXACT: two sweeps over the genes (projection + scale, then the distance), 10 fp64 ops per gene
//                              and pair, bit-identical to the reference;
//                       FAST : one sweep of two FMAs per gene and pair (projection and Gram entry; Appendix A7 of
//                              SURVEY.md), error ~1e-13, for the bulk; every decision it feeds is certified below.  The
//                              default FAST kernel (sv_pairs_dmma_kernel) runs the two contractions on the fp64 tensor
//                              cores (mma.sync.m8n8k4.f64); B200MNN_SHIFTVAR=fast_simt keeps them on the CUDA cores.
//                     Own batch: running (masked, total) log-sum-exp per cell.  Reference batch: (projection,
//                     log-weight) rows written to a chunk buffer + running total, min and max projection.
//   sv_select_kernel  one CTA per cell: weighted-quantile SELECTION instead of a sort -- histogram of the weights over
//                     1 024 projection bins (refined until the crossing bin holds <= 2 048 cells), bitonic sort of that
//                     bin only, cumulative walk.  The pick is CERTIFIED when the target clears both neighbouring
//                     cumulative weights by a relative margin that covers the order-of-summation and libm differences
//                     (and, in FAST mode, the scoring error); uncertified cells are flagged.
//   sv_cell_kernel    flagged cells (a handful per 10^5): the reference's loop as it stands -- exact pairs, sequential
//                     logspace_add folds in restrict order, full sort, sequential cumulative walk.
// Layout: data1 [n1 x G], data2 [n2 x G], vect [n2 x G], row-major (one cell contiguous).
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace b200 {
namespace shiftvar {

constexpr int TS = 64;         // tile edge
constexpr int KC = 16;         // genes per staged chunk
constexpr int THREADS = 256;   // 16 x 16 threads, 4 x 4 pairs each
constexpr int SEL_BINS = 1024;
constexpr int SEL_CAP = 1024;

struct LSE {  // running log-sum-exp as (max, sum of exp(x - max)); empty when s == 0
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

// R::logspace_add (Rmath): max(x, y) + log1p(exp(-|x - y|))
__device__ __forceinline__ double logspace_add(double x, double y) { return fmax(x, y) + log1p(exp(-fabs(x - y))); }

__device__ __forceinline__ bool pl_less(double pa, double la, double pb, double lb) { return pa < pb || (pa == pb && la < lb); }

// ------------------------------------------------------------------------------------------------
// prep: l2 norm, unit gradient and the cell's own projection, all in the reference's sequential order
// ------------------------------------------------------------------------------------------------
constexpr int PREP_WARPS = 4;

__global__ void __launch_bounds__(PREP_WARPS * 32)
sv_prep_kernel(const double* __restrict__ vect, const double* __restrict__ data2, int64_t n2, int64_t G, double* __restrict__ grad,
               double* __restrict__ l2out, double* __restrict__ curproj) {
    __shared__ double buf[PREP_WARPS][2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * PREP_WARPS + warp;
    if (c >= n2) return;
    const double* v = vect + c * G;
    const double* x = data2 + c * G;
    double acc = 0.0;
    for (int64_t t0 = 0; t0 < G; t0 += 32) {
        buf[warp][0][lane] = (t0 + lane < G) ? v[t0 + lane] : 0.0;
        __syncwarp();
        if (lane == 0) {
            const int len = (int)min((int64_t)32, G - t0);
            for (int j = 0; j < len; ++j) acc = __dadd_rn(acc, __dmul_rn(buf[warp][0][j], buf[warp][0][j]));   // l2norm += g * g (:60)
        }
        __syncwarp();
    }
    acc = __shfl_sync(0xffffffffu, acc, 0);
    const double l2 = sqrt(acc);
    double p = 0.0;
    for (int64_t t0 = 0; t0 < G; t0 += 32) {
        double gv = 0.0, xv = 0.0;
        if (t0 + lane < G) {
            gv = v[t0 + lane];
            if (l2 != 0.0) gv = __ddiv_rn(gv, l2);   // g /= l2norm (:64-67)
            grad[c * G + t0 + lane] = gv;
            xv = x[t0 + lane];
        }
        buf[warp][0][lane] = gv;
        buf[warp][1][lane] = xv;
        __syncwarp();
        if (lane == 0) {
            const int len = (int)min((int64_t)32, G - t0);
            for (int j = 0; j < len; ++j) p = __dadd_rn(p, __dmul_rn(buf[warp][0][j], buf[warp][1][j]));       // inner_product (:70)
        }
        __syncwarp();
    }
    if (lane == 0) { l2out[c] = l2; curproj[c] = p; }
}

// exact (reference-order) projection of x on g and squared distance of x from the line through cur along g, one thread
__device__ __forceinline__ void exact_pair(const double* __restrict__ g, const double* __restrict__ cur, const double* __restrict__ x, int64_t G,
                                           double& proj, double& dist) {
    double p = 0.0, sc = 0.0;
    for (int64_t t = 0; t < G; ++t) {
        const double gv = g[t], xv = x[t];
        p = __dadd_rn(p, __dmul_rn(gv, xv));
        sc = __dadd_rn(sc, __dmul_rn(__dsub_rn(cur[t], xv), gv));
    }
    double ds = 0.0;
    for (int64_t t = 0; t < G; ++t) {
        const double gv = g[t];
        const double w = __dsub_rn(__dsub_rn(cur[t], x[t]), __dmul_rn(sc, gv));
        ds = __dadd_rn(ds, __dmul_rn(w, w));
    }
    proj = p;
    dist = ds;
}

// ------------------------------------------------------------------------------------------------
// pairs: tile kernel
// ------------------------------------------------------------------------------------------------
// Partial results per (cell of the chunk, column split): own batch {lp.m, lp.s, lt.m, lt.s}; reference batch
// {l1.m, l1.s, pmin, pmax}.
template <bool EXACT, bool OWN>
__global__ void __launch_bounds__(THREADS)
sv_pairs_kernel(const double* __restrict__ grad, const double* __restrict__ data2, const double* __restrict__ curproj,
                const double* __restrict__ norm2 /* FAST: ||x_c||^2 of batch-2 cells */, int64_t c0, int64_t nrows,
                const double* __restrict__ other, const double* __restrict__ onorm /* FAST: ||x||^2 of `other` rows */,
                const int32_t* __restrict__ ridx, int64_t ncols, int64_t G, double sigma2, double amb /* FAST: projection ambiguity */,
                double* __restrict__ Pout, double* __restrict__ Wout, int64_t ld, double* __restrict__ part, int nsplit,
                int tiles_per_split) {
    __shared__ double Gs[KC][TS + 1];
    __shared__ double Cs[KC][TS + 1];
    __shared__ double Xs[KC][TS + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;   // loader: row lr of the tile, 4 consecutive genes
    const int64_t rb = (int64_t)blockIdx.y * TS;                   // first row (cell of the chunk) of this CTA
    const int split = blockIdx.x;
    const int64_t ncoltiles = (ncols + TS - 1) / TS;
    const int64_t t_begin = (int64_t)split * tiles_per_split, t_end = min(ncoltiles, t_begin + tiles_per_split);

    const int64_t arow = (rb + lr < nrows) ? c0 + rb + lr : -1;    // cell this loader thread fetches
    double cp[4], cn[4];
    int64_t cell[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t r = rb + ty * 4 + a;
        cell[a] = (r < nrows) ? c0 + r : -1;
        cp[a] = (r < nrows) ? curproj[c0 + r] : 0.0;
        cn[a] = (!EXACT && r < nrows) ? norm2[c0 + r] : 0.0;
    }
    LSE l_a[4], l_b[4];   // OWN: masked / total; reference: total / unused
    double pmin[4], pmax[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { l_a[a] = {-INFINITY, 0.0}; l_b[a] = {-INFINITY, 0.0}; pmin[a] = INFINITY; pmax[a] = -INFINITY; }

    for (int64_t ct = t_begin; ct < t_end; ++ct) {
        const int64_t cb = ct * TS;
        const int64_t brow = (cb + lr < ncols) ? (int64_t)ridx[cb + lr] : -1;
        double acc0[4][4], acc1[4][4];   // EXACT: projection, scale (then distance); FAST: projection, Gram entry
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) { acc0[a][b] = 0.0; acc1[a][b] = 0.0; }
        // ---- sweep 1 (the next chunk's global loads are in flight while the current one is consumed) ----
        double pg[4], pc[4], px[4];
        auto fetch = [&](int64_t k0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + lk + j;
                const bool kin = k < G;
                pg[j] = (arow >= 0 && kin) ? grad[arow * G + k] : 0.0;
                pc[j] = (arow >= 0 && kin) ? data2[arow * G + k] : 0.0;
                px[j] = (brow >= 0 && kin) ? other[brow * G + k] : 0.0;
            }
        };
        auto stage = [&]() {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) { Gs[lk + j][lr] = pg[j]; Cs[lk + j][lr] = pc[j]; Xs[lk + j][lr] = px[j]; }
            __syncthreads();
        };
        fetch(0);
        for (int64_t k0 = 0; k0 < G; k0 += KC) {
            stage();
            if (k0 + KC < G) fetch(k0 + KC);
            const int kmax = (int)min((int64_t)KC, G - k0);
            for (int k = 0; k < kmax; ++k) {
                double gv[4], cv[4], xv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) { gv[a] = Gs[k][ty * 4 + a]; cv[a] = Cs[k][ty * 4 + a]; }
#pragma unroll
                for (int b = 0; b < 4; ++b) xv[b] = Xs[k][tx + 16 * b];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (EXACT) {
                            acc0[a][b] = __dadd_rn(acc0[a][b], __dmul_rn(gv[a], xv[b]));                      // inner_product(grad, other)
                            acc1[a][b] = __dadd_rn(acc1[a][b], __dmul_rn(__dsub_rn(cv[a], xv[b]), gv[a]));    // inner_product(working, grad)
                        } else {
                            acc0[a][b] = fma(gv[a], xv[b], acc0[a][b]);
                            acc1[a][b] = fma(cv[a], xv[b], acc1[a][b]);
                        }
                    }
            }
        }
        double proj[4][4], lw[4][4];
        if (EXACT) {
            // ---- sweep 2: distance to the line ----
            double dist[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dist[a][b] = 0.0;
            fetch(0);
            for (int64_t k0 = 0; k0 < G; k0 += KC) {
                stage();
                if (k0 + KC < G) fetch(k0 + KC);
                const int kmax = (int)min((int64_t)KC, G - k0);
                for (int k = 0; k < kmax; ++k) {
                    double gv[4], cv[4], xv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { gv[a] = Gs[k][ty * 4 + a]; cv[a] = Cs[k][ty * 4 + a]; }
#pragma unroll
                    for (int b = 0; b < 4; ++b) xv[b] = Xs[k][tx + 16 * b];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const double w = __dsub_rn(__dsub_rn(cv[a], xv[b]), __dmul_rn(acc1[a][b], gv[a]));   // w -= scale * grad
                            dist[a][b] = __dadd_rn(dist[a][b], __dmul_rn(w, w));                                // dist += w * w
                        }
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) { proj[a][b] = acc0[a][b]; lw[a][b] = __ddiv_rn(-dist[a][b], sigma2); }
        } else {
            // dl^2 = ||x_c||^2 + ||x||^2 - 2 x_c.x - (g.x_c - g.x)^2   
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int64_t j = cb + tx + 16 * b;
                const double xn = (j < ncols) ? onorm[ridx[j]] : 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const double dp = cp[a] - acc0[a][b];
                    double d2 = (cn[a] + xn) - 2.0 * acc1[a][b] - dp * dp;
                    d2 = fmax(d2, 0.0);
                    proj[a][b] = acc0[a][b];
                    lw[a][b] = -d2 / sigma2;
                }
            }
        }
        // ---- tile epilogue ----
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (cell[a] < 0) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int64_t j = cb + tx + 16 * b;
                if (j >= ncols) continue;
                if (OWN) {
                    const int64_t same = ridx[j];
                    bool add = true;
                    double logp = 0.0;
                    if (same != cell[a]) {
                        logp = lw[a][b];
                        double sp = proj[a][b];
                        if (!EXACT && fabs(sp - cp[a]) <= amb) {
                            // the FMA-order projection cannot decide `sameproj > curproj` (:91): redo this pair in the reference's order
                            double ep, ed;
                            exact_pair(grad + cell[a] * G, data2 + cell[a] * G, other + same * G, G, ep, ed);
                            sp = ep;
                            logp = __ddiv_rn(-ed, sigma2);
                        }
                        if (sp > cp[a]) add = false;
                    }
                    if (add) lse_add(l_a[a], logp);
                    lse_add(l_b[a], logp);
                } else {
                    const int64_t o = (rb + ty * 4 + a) * ld + j;
                    Pout[o] = proj[a][b];
                    Wout[o] = lw[a][b];
                    lse_add(l_a[a], lw[a][b]);
                    pmin[a] = fmin(pmin[a], proj[a][b]);
                    pmax[a] = fmax(pmax[a], proj[a][b]);
                }
            }
        }
    }
    // ---- per-row partials: the 16 threads (tx) that share rows ty*4 .. ty*4+3 are one half-warp ----
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            LSE oa, ob;
            oa.m = __shfl_xor_sync(0xffffffffu, l_a[a].m, o);
            oa.s = __shfl_xor_sync(0xffffffffu, l_a[a].s, o);
            lse_merge(l_a[a], oa);
            if (OWN) {
                ob.m = __shfl_xor_sync(0xffffffffu, l_b[a].m, o);
                ob.s = __shfl_xor_sync(0xffffffffu, l_b[a].s, o);
                lse_merge(l_b[a], ob);
            } else {
                pmin[a] = fmin(pmin[a], __shfl_xor_sync(0xffffffffu, pmin[a], o));
                pmax[a] = fmax(pmax[a], __shfl_xor_sync(0xffffffffu, pmax[a], o));
            }
        }
        const int64_t r = rb + ty * 4 + a;
        if (tx == 0 && r < nrows) {
            double* o = part + (r * nsplit + split) * 4;
            o[0] = l_a[a].m; o[1] = l_a[a].s;
            o[2] = OWN ? l_b[a].m : pmin[a];
            o[3] = OWN ? l_b[a].s : pmax[a];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pairs, FAST mode on the fp64 tensor cores (DMMA m8n8k4)
// ------------------------------------------------------------------------------------------------
// Same tile (64 cells x 64 comparison cells, genes staged in chunks of 16) and same epilogue as sv_pairs_kernel<false>;
// the two contractions (projection g_c . x and Gram entry x_c . x) run as mma.sync.m8n8k4.f64: 8 warps = 4 row groups
// (16 cells: two m8 blocks of the g rows and two of the x_c rows) x 2 column groups (32 cells: four n8 blocks), 16 DMMAs
// per 4 genes fed by 8 conflict-free LDS.64 -- the SIMT form needs 12 LDS.64 per 32 FMAs and is shared-memory bound.
// Partial results are written per (row, split, column group): the partial arrays hold 2 * nsplit entries per row.
constexpr int DM_KP = 20;   // shared-memory row pitch in doubles (rows of 16 staged genes): half-warps hit 16 distinct banks

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <bool OWN>
__global__ void __launch_bounds__(THREADS)
sv_pairs_dmma_kernel(const double* __restrict__ grad, const double* __restrict__ data2, const double* __restrict__ curproj,
                     const double* __restrict__ norm2, int64_t c0, int64_t nrows, const double* __restrict__ other,
                     const double* __restrict__ onorm, const int32_t* __restrict__ ridx, int64_t ncols, int64_t G, double sigma2, double amb,
                     double* __restrict__ Pout, double* __restrict__ Wout, int64_t ld, double* __restrict__ part, int nsplit, int tiles_per_split) {
    __shared__ __align__(16) double Gs[TS][DM_KP];
    __shared__ __align__(16) double Cs[TS][DM_KP];
    __shared__ __align__(16) double Xs[TS][DM_KP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rg = warp & 3, cg = warp >> 2;
    const int gid = lane >> 2, tig = lane & 3;
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;   // loader: row lr of the tile, 4 consecutive genes
    const int64_t rb = (int64_t)blockIdx.y * TS;
    const int split = blockIdx.x;
    const int64_t ncoltiles = (ncols + TS - 1) / TS;
    const int64_t t_begin = (int64_t)split * tiles_per_split, t_end = min(ncoltiles, t_begin + tiles_per_split);
    const int64_t arow = (rb + lr < nrows) ? c0 + rb + lr : -1;

    // this lane's two rows (m block 0 / 1)
    double cp[2], cn[2];
    int64_t cell[2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const int64_t r = rb + rg * 16 + mb * 8 + gid;
        cell[mb] = (r < nrows) ? c0 + r : -1;
        cp[mb] = (r < nrows) ? curproj[c0 + r] : 0.0;
        cn[mb] = (r < nrows) ? norm2[c0 + r] : 0.0;
    }
    LSE l_a[2], l_b[2];
    double pmin[2], pmax[2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) { l_a[mb] = {-INFINITY, 0.0}; l_b[mb] = {-INFINITY, 0.0}; pmin[mb] = INFINITY; pmax[mb] = -INFINITY; }

    for (int64_t ct = t_begin; ct < t_end; ++ct) {
        const int64_t cb = ct * TS;
        const int64_t brow = (cb + lr < ncols) ? (int64_t)ridx[cb + lr] : -1;
        double accP[2][4][2], accD[2][4][2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) { accP[mb][nb][0] = accP[mb][nb][1] = 0.0; accD[mb][nb][0] = accD[mb][nb][1] = 0.0; }
        double pg[4], pc[4], px[4];
        auto fetch = [&](int64_t k0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + lk + j;
                const bool kin = k < G;
                pg[j] = (arow >= 0 && kin) ? grad[arow * G + k] : 0.0;
                pc[j] = (arow >= 0 && kin) ? data2[arow * G + k] : 0.0;
                px[j] = (brow >= 0 && kin) ? other[brow * G + k] : 0.0;
            }
        };
        fetch(0);
        for (int64_t k0 = 0; k0 < G; k0 += KC) {
            __syncthreads();
            *reinterpret_cast<double2*>(&Gs[lr][lk]) = make_double2(pg[0], pg[1]);
            *reinterpret_cast<double2*>(&Gs[lr][lk + 2]) = make_double2(pg[2], pg[3]);
            *reinterpret_cast<double2*>(&Cs[lr][lk]) = make_double2(pc[0], pc[1]);
            *reinterpret_cast<double2*>(&Cs[lr][lk + 2]) = make_double2(pc[2], pc[3]);
            *reinterpret_cast<double2*>(&Xs[lr][lk]) = make_double2(px[0], px[1]);
            *reinterpret_cast<double2*>(&Xs[lr][lk + 2]) = make_double2(px[2], px[3]);
            __syncthreads();
            if (k0 + KC < G) fetch(k0 + KC);
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                double ag[2], ac[2], bx[4];
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    ag[mb] = Gs[rg * 16 + mb * 8 + gid][k4 * 4 + tig];
                    ac[mb] = Cs[rg * 16 + mb * 8 + gid][k4 * 4 + tig];
                }
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) bx[nb] = Xs[cg * 32 + nb * 8 + gid][k4 * 4 + tig];
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        dmma(accP[mb][nb][0], accP[mb][nb][1], ag[mb], bx[nb]);
                        dmma(accD[mb][nb][0], accD[mb][nb][1], ac[mb], bx[nb]);
                    }
            }
        }
        // ---- tile epilogue (same arithmetic as the SIMT form) ----
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int64_t j = cb + cg * 32 + nb * 8 + tig * 2 + i;
                if (j >= ncols) continue;
                const int64_t oidx = ridx[j];
                const double xn = onorm[oidx];
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    if (cell[mb] < 0) continue;
                    const double pr = accP[mb][nb][i];
                    const double dp = cp[mb] - pr;
                    double d2 = (cn[mb] + xn) - 2.0 * accD[mb][nb][i] - dp * dp;
                    d2 = fmax(d2, 0.0);
                    double lwv = -d2 / sigma2;
                    if (OWN) {
                        bool add = true;
                        double logp = 0.0;
                        if (oidx != cell[mb]) {
                            logp = lwv;
                            double sp = pr;
                            if (fabs(sp - cp[mb]) <= amb) {   // undecidable in FMA order: redo this pair in the reference's order
                                double ep, ed;
                                exact_pair(grad + cell[mb] * G, data2 + cell[mb] * G, other + oidx * G, G, ep, ed);
                                sp = ep;
                                logp = __ddiv_rn(-ed, sigma2);
                            }
                            if (sp > cp[mb]) add = false;
                        }
                        if (add) lse_add(l_a[mb], logp);
                        lse_add(l_b[mb], logp);
                    } else {
                        const int64_t o = (rb + rg * 16 + mb * 8 + gid) * ld + j;
                        Pout[o] = pr;
                        Wout[o] = lwv;
                        lse_add(l_a[mb], lwv);
                        pmin[mb] = fmin(pmin[mb], pr);
                        pmax[mb] = fmax(pmax[mb], pr);
                    }
                }
            }
    }
    // the four lanes of a group (tig) share the lane's two rows
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            LSE oa, ob;
            oa.m = __shfl_xor_sync(0xffffffffu, l_a[mb].m, o);
            oa.s = __shfl_xor_sync(0xffffffffu, l_a[mb].s, o);
            lse_merge(l_a[mb], oa);
            if (OWN) {
                ob.m = __shfl_xor_sync(0xffffffffu, l_b[mb].m, o);
                ob.s = __shfl_xor_sync(0xffffffffu, l_b[mb].s, o);
                lse_merge(l_b[mb], ob);
            } else {
                pmin[mb] = fmin(pmin[mb], __shfl_xor_sync(0xffffffffu, pmin[mb], o));
                pmax[mb] = fmax(pmax[mb], __shfl_xor_sync(0xffffffffu, pmax[mb], o));
            }
        }
        const int64_t r = rb + rg * 16 + mb * 8 + gid;
        if (tig == 0 && r < nrows) {
            double* o = part + (r * (2 * nsplit) + 2 * split + cg) * 4;
            o[0] = l_a[mb].m; o[1] = l_a[mb].s;
            o[2] = OWN ? l_b[mb].m : pmin[mb];
            o[3] = OWN ? l_b[mb].s : pmax[mb];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// select: certified weighted-quantile selection, one CTA per cell
// ------------------------------------------------------------------------------------------------
// The projection range is cut into 1 024 bins, level by level, until the bin in which the cumulative weight crosses the
// target holds <= SEL_CAP cells.  Bins are LOCATED with a 2^-40 fixed-point histogram (native 64-bit integer atomics in
// shared memory; fp64 atomics there are CAS loops); the cumulative weight in front of the chosen bin is then summed
// again in fp64 by a deterministic block reduction, the bin is sorted and walked, and the pick is certified with margins
// on those fp64 sums -- a bin mislocated by the fixed-point rounding simply fails the certificate.
constexpr int SEL_LEVELS = 6;

__device__ __forceinline__ int sel_bin(double p, double lo, double scale) {
    const double t = (p - lo) * scale;
    int b = (t >= (double)SEL_BINS) ? SEL_BINS - 1 : (int)t;   // (int)NaN == 0
    return b < 0 ? 0 : b;
}
// -1: in front of the chosen bin chain, 0: inside, +1: behind
__device__ __forceinline__ int sel_classify(double p, const double* lv_lo, const double* lv_scale, const int* lv_bin, int nlev) {
    for (int l = 0; l < nlev; ++l) {
        const int b = sel_bin(p, lv_lo[l], lv_scale[l]);
        if (b != lv_bin[l]) return b < lv_bin[l] ? -1 : 1;
    }
    return 0;
}

__device__ __forceinline__ double exact_proj(const double* __restrict__ g, const double* __restrict__ x, int64_t G) {
    double p = 0.0;
    for (int64_t t = 0; t < G; ++t) p = __dadd_rn(p, __dmul_rn(g[t], x[t]));
    return p;
}

__global__ void __launch_bounds__(THREADS)
sv_select_kernel(const double* __restrict__ P, const double* __restrict__ W, int64_t ld, int64_t nr1, const double* __restrict__ part2,
                 int nsplit2, int64_t nr2, const double* __restrict__ part1, int nsplit1, int64_t c0, int64_t nrows,
                 const double* __restrict__ grad, const double* __restrict__ data1, const int32_t* __restrict__ r1, int64_t G,
                 const double* __restrict__ curproj, const double* __restrict__ l2, double tol, double amb /* 0 for exact rows */,
                 double* __restrict__ out, int* __restrict__ flag_count, int32_t* __restrict__ flag_list) {
    __shared__ __align__(16) unsigned char raw[SEL_CAP * 28];
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(raw);     // histogram phase: [SEL_BINS] u64 + [SEL_BINS] int
    int* hcnt = reinterpret_cast<int*>(raw + SEL_BINS * 8);
    double* lP = reinterpret_cast<double*>(raw);                                // list phase: P, W, prefix sums, column ids
    double* lW = lP + SEL_CAP;
    double* lC = lW + SEL_CAP;
    int* lO = reinterpret_cast<int*>(lC + SEL_CAP);
    static_assert(SEL_BINS * 12 <= SEL_CAP * 28, "shared buffer too small for the histogram");
    __shared__ double lv_lo[SEL_LEVELS], lv_scale[SEL_LEVELS];
    __shared__ int lv_bin[SEL_LEVELS];
    __shared__ int s_nlev, s_state, s_n, s_found, s_pick, s_lastcount;   // state: 0 = refine further, 1 = collect, 2 = flag
    __shared__ unsigned long long s_basefx;
    __shared__ double s_tlin, s_M, s_base;
    __shared__ long long s_basecnt;
    __shared__ double wsum[THREADS / 32], wlo[THREADS / 32], whi[THREADS / 32];
    __shared__ long long wcnt[THREADS / 32];
    __shared__ double s_pbelow, s_pabove;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double FX = 1099511627776.0;   // 2^40

    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int64_t c = c0 + row;
        const double* Pr = P + row * ld;
        const double* Wr = W + row * ld;
        __syncthreads();
        if (threadIdx.x == 0) {
            LSE a = {-INFINITY, 0.0}, b = {-INFINITY, 0.0}, e = {-INFINITY, 0.0};
            double mn = INFINITY, mx = -INFINITY;
            if (nr2 > 0)
                for (int s = 0; s < nsplit2; ++s) {
                    const double* p = part2 + (row * nsplit2 + s) * 4;
                    lse_merge(a, LSE{p[0], p[1]});
                    lse_merge(b, LSE{p[2], p[3]});
                }
            for (int s = 0; s < nsplit1; ++s) {
                const double* p = part1 + (row * nsplit1 + s) * 4;
                lse_merge(e, LSE{p[0], p[1]});
                mn = fmin(mn, p[2]);
                mx = fmax(mx, p[3]);
            }
            // prob2 / totalprob2 start at 0 in the reference when nothing was added (:75-77)
            const double pa = (a.s == 0.0) ? 0.0 : lse_value(a);
            const double pb = (b.s == 0.0) ? 0.0 : lse_value(b);
            const double target = (pa - pb) + lse_value(e);
            s_M = e.m;
            s_tlin = exp(target - e.m);
            lv_lo[0] = mn;
            lv_scale[0] = (mx > mn) ? (double)SEL_BINS / (mx - mn) : 0.0;
            s_nlev = 0;
            s_basefx = 0ull;
            s_lastcount = 0x7fffffff;
            // linear-domain selection needs the target inside the double range relative to the largest weight
            const bool usable = isfinite(target) && isfinite(e.m) && target - e.m > -600.0 && isfinite(mn) && isfinite(mx) &&
                                isfinite(lv_scale[0]);
            s_state = usable ? 0 : 2;
        }
        __syncthreads();
        const double M = s_M;
        // ---- locate: refine the projection range until the crossing bin is small ----
        for (int level = 0; level < SEL_LEVELS; ++level) {
            if (s_state != 0) break;   // uniform: s_state only changes between the barriers below
            for (int i = threadIdx.x; i < SEL_BINS; i += THREADS) { hist[i] = 0ull; hcnt[i] = 0; }
            __syncthreads();
            const double lo = lv_lo[level], scale = lv_scale[level];
            for (int64_t o = threadIdx.x; o < nr1; o += THREADS) {
                const double p = Pr[o];
                if (sel_classify(p, lv_lo, lv_scale, lv_bin, level) == 0) {
                    const int b = sel_bin(p, lo, scale);
                    atomicAdd(&hist[b], (unsigned long long)(exp(Wr[o] - M) * FX));
                    atomicAdd(&hcnt[b], 1);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const double tfx_d = s_tlin * FX;
                const unsigned long long tfx = (tfx_d >= 9.0e18) ? 9000000000000000000ull : (unsigned long long)tfx_d;
                unsigned long long cum = s_basefx;
                int pick = -1, last = -1;
                unsigned long long cum_last = cum;
                for (int b = 0; b < SEL_BINS; ++b) {
                    if (hcnt[b] == 0) continue;
                    last = b;
                    cum_last = cum;
                    if (cum + hist[b] >= tfx) { pick = b; break; }
                    cum += hist[b];
                }
                if (pick < 0) { pick = last; cum = cum_last; }   // the total falls short of the target by rounding: the last bin
                if (pick < 0) s_state = 2;
                else {
                    const int count = hcnt[pick];
                    lv_bin[level] = pick;
                    s_basefx = cum;
                    s_nlev = level + 1;
                    if (count <= SEL_CAP) s_state = 1;
                    else if (level + 1 >= SEL_LEVELS || count >= s_lastcount) s_state = 2;   // thousands of tied projections: exact path
                    else {
                        lv_lo[level + 1] = lo + (double)pick / scale;
                        lv_scale[level + 1] = scale * (double)SEL_BINS;
                        if (!(scale > 0.0) || !isfinite(lv_scale[level + 1]) || !isfinite(lv_lo[level + 1])) s_state = 2;
                    }
                    s_lastcount = count;
                }
            }
            __syncthreads();
        }
        __syncthreads();
        const int state_after_locate = s_state;
        __syncthreads();
        if (state_after_locate == 1) {
            // ---- collect the crossing bin; sum what lies in front of it in fp64 ----
            if (threadIdx.x == 0) { s_n = 0; s_found = 0x7fffffff; }
            __syncthreads();
            const int nlev = s_nlev;
            double below = 0.0, pbelow = -INFINITY, pabove = INFINITY;   // weight in front; nearest projections outside the bin
            long long nbelow = 0;
            for (int64_t o = threadIdx.x; o < nr1; o += THREADS) {
                const double p = Pr[o];
                const int cls = sel_classify(p, lv_lo, lv_scale, lv_bin, nlev);
                if (cls < 0) { below += exp(Wr[o] - M); ++nbelow; pbelow = fmax(pbelow, p); }
                else if (cls > 0) pabove = fmin(pabove, p);
                else {
                    const int pos = atomicAdd(&s_n, 1);
                    if (pos < SEL_CAP) { lP[pos] = p; lW[pos] = Wr[o]; lO[pos] = (int)o; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                below += __shfl_xor_sync(0xffffffffu, below, o);
                nbelow += __shfl_xor_sync(0xffffffffu, nbelow, o);
                pbelow = fmax(pbelow, __shfl_xor_sync(0xffffffffu, pbelow, o));
                pabove = fmin(pabove, __shfl_xor_sync(0xffffffffu, pabove, o));
            }
            if (lane == 0) { wsum[warp] = below; wcnt[warp] = nbelow; wlo[warp] = pbelow; whi[warp] = pabove; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double b = 0.0, pl = -INFINITY, ph = INFINITY;
                long long nb = 0;
                for (int w = 0; w < THREADS / 32; ++w) { b += wsum[w]; nb += wcnt[w]; pl = fmax(pl, wlo[w]); ph = fmin(ph, whi[w]); }
                s_base = b;
                s_basecnt = nb;
                s_pbelow = pl;
                s_pabove = ph;
            }
            const int n = min(s_n, SEL_CAP);
            int np2 = 1;
            while (np2 < n) np2 <<= 1;
            __syncthreads();
            for (int i = n + threadIdx.x; i < np2; i += THREADS) { lP[i] = INFINITY; lW[i] = INFINITY; lO[i] = -1; }
            __syncthreads();
            for (int size = 2; size <= np2; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int i = threadIdx.x; i < np2 / 2; i += THREADS) {
                        const int a = (i / stride) * (stride * 2) + (i % stride);
                        const int b = a + stride;
                        const bool up = ((a & size) == 0);
                        const double pa = lP[a], wa = lW[a], pb = lP[b], wb = lW[b];
                        const bool swap = up ? pl_less(pb, wb, pa, wa) : pl_less(pa, wa, pb, wb);
                        if (swap) {
                            const int oa = lO[a], ob = lO[b];
                            lP[a] = pb; lW[a] = wb; lO[a] = ob;
                            lP[b] = pa; lW[b] = wa; lO[b] = oa;
                        }
                    }
                    __syncthreads();
                }
            }
            // inclusive prefix sums of the weights (each thread owns SEL_CAP / THREADS consecutive entries)
            constexpr int PER = SEL_CAP / THREADS;
            double loc[PER];
            double run = 0.0;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = threadIdx.x * PER + j;
                run += (i < n) ? exp(lW[i] - M) : 0.0;
                loc[j] = run;
            }
            double inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            double off = inc - run;
            for (int w = 0; w < warp; ++w) off += wsum[w];
            const double base = s_base, tlin = s_tlin;
            int first = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = threadIdx.x * PER + j;
                if (i < n) {
                    const double cv = off + loc[j];
                    lC[i] = cv;
                    if (first == 0x7fffffff && base + cv >= tlin) first = i;
                }
            }
            if (first != 0x7fffffff) atomicMin(&s_found, first);
            __syncthreads();
            if (threadIdx.x == 0) {
                bool ok = false;
                int i = s_found;
                const long long basecnt = s_basecnt;
                if (s_n <= SEL_CAP && n > 0) {
                    if (i == 0x7fffffff && basecnt + n == (long long)nr1) i = n - 1;   // short of the target by rounding only
                    if (i != 0x7fffffff) {
                        // cells whose projections cannot be told apart from the pick's form one group: whichever of them
                        // crosses first, the quantile is the same only if they are true duplicates (amb == 0: exact ties)
                        int gs = i, ge = i;
                        while (gs > 0 && lP[gs - 1] >= lP[i] - amb) --gs;
                        while (ge + 1 < n && lP[ge + 1] <= lP[i] + amb) ++ge;
                        bool same = true;
                        for (int j = gs; j <= ge; ++j) same = same && (lP[j] == lP[i]) && (amb == 0.0 || lW[j] == lW[i]);
                        const bool first_overall = (basecnt + gs == 0);
                        const bool last_overall = (basecnt + ge + 1 == (long long)nr1);
                        // (FMA-order rows) a cell of a neighbouring bin may be indistinguishable from the pick as well
                        const bool edge_unknown = amb > 0.0 && ((gs == 0 && !first_overall && s_pbelow >= lP[i] - amb) ||
                                                                (ge == n - 1 && !last_overall && s_pabove <= lP[i] + amb));
                        const double cum_before = base + (gs > 0 ? lC[gs - 1] : 0.0);
                        const double cum_after = base + lC[ge];
                        const bool lower_ok = first_overall || (tlin - cum_before) > tol * cum_after;
                        const bool upper_ok = last_overall || (cum_after - tlin) > tol * cum_after;
                        ok = same && !edge_unknown && lower_ok && upper_ok;
                    }
                }
                s_pick = ok ? lO[i] : -1;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            if (state_after_locate != 1 || s_pick < 0) {
                const int slot = atomicAdd(flag_count, 1);
                flag_list[slot] = (int32_t)c;
            } else {
                // FMA-order rows (amb > 0): the reported quantile is recomputed in the reference's order
                const double refq = (amb > 0.0) ? exact_proj(grad + c * G, data1 + (int64_t)r1[s_pick] * G, G) : Pr[s_pick];
                out[c] = __ddiv_rn(__dsub_rn(refq, curproj[c]), l2[c]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cell: the reference's loop for one cell per CTA (flagged cells, and the whole job when B200MNN_SHIFTVAR=cell)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
sv_cell_kernel(const double* __restrict__ data1, const double* __restrict__ data2, int64_t G, const double* __restrict__ grad,
               const double* __restrict__ curproj_all, const double* __restrict__ l2_all, double sigma2, const int32_t* __restrict__ r1,
               int64_t nr1, const int32_t* __restrict__ r2, int64_t nr2, int64_t nr1_pow2, const int* __restrict__ flag_count,
               const int32_t* __restrict__ flag_list, int64_t n2_all, double* __restrict__ sP, double* __restrict__ sW,
               double* __restrict__ sL2, unsigned char* __restrict__ sAdd, double* __restrict__ out) {
    __shared__ double s_prob2, s_tot1;
    const int64_t count = flag_list ? (int64_t)*flag_count : n2_all;
    double* P = sP + (int64_t)blockIdx.x * nr1_pow2;
    double* W = sW + (int64_t)blockIdx.x * nr1_pow2;
    double* L2 = sL2 + (int64_t)blockIdx.x * max(nr2, (int64_t)1);
    unsigned char* AD = sAdd + (int64_t)blockIdx.x * max(nr2, (int64_t)1);
    for (int64_t f = blockIdx.x; f < count; f += gridDim.x) {
        const int64_t c = flag_list ? (int64_t)flag_list[f] : f;
        const double* g = grad + c * G;
        const double* cur = data2 + c * G;
        const double curproj = curproj_all[c];
        __syncthreads();
        for (int64_t s = threadIdx.x; s < nr2; s += THREADS) {
            const int64_t same = r2[s];
            double logp = 0.0;
            unsigned char add = 1;
            if (same != c) {
                double p, dd;
                exact_pair(g, cur, data2 + same * G, G, p, dd);
                logp = __ddiv_rn(-dd, sigma2);
                if (p > curproj) add = 0;
            }
            L2[s] = logp;
            AD[s] = add;
        }
        for (int64_t o = threadIdx.x; o < nr1; o += THREADS) {
            double p, dd;
            exact_pair(g, cur, data1 + (int64_t)r1[o] * G, G, p, dd);
            P[o] = p;
            W[o] = __ddiv_rn(-dd, sigma2);
        }
        for (int64_t o = nr1 + threadIdx.x; o < nr1_pow2; o += THREADS) { P[o] = INFINITY; W[o] = INFINITY; }   // padding sorts last
        __threadfence_block();
        __syncthreads();
        if (threadIdx.x == 0) {   // sequential folds in restrict order (:96-131)
            double prob2 = 0.0, tot2 = 0.0;
            bool sp = true, st = true;
            for (int64_t s = 0; s < nr2; ++s) {
                const double lp = L2[s];
                if (AD[s]) { if (sp) { prob2 = lp; sp = false; } else prob2 = logspace_add(prob2, lp); }
                if (st) { tot2 = lp; st = false; } else tot2 = logspace_add(tot2, lp);
            }
            s_prob2 = prob2 - tot2;
        } else if (threadIdx.x == 32) {
            double tot1 = 0.0;
            bool st = true;
            for (int64_t o = 0; o < nr1; ++o) {
                if (st) { tot1 = W[o]; st = false; } else tot1 = logspace_add(tot1, W[o]);
            }
            s_tot1 = tot1;
        }
        __syncthreads();
        // bitonic sort of (P, W) ascending, lexicographic
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
        if (threadIdx.x == 0) {
            double refq = NAN;
            if (nr1 > 0) {
                const double target = s_prob2 + s_tot1;
                double cum = 0.0;
                bool st = true;
                refq = P[nr1 - 1];
                for (int64_t o = 0; o < nr1; ++o) {
                    if (st) { cum = W[o]; st = false; } else cum = logspace_add(cum, W[o]);
                    if (cum >= target) { refq = P[o]; break; }
                }
            }
            out[c] = __ddiv_rn(__dsub_rn(refq, curproj), l2_all[c]);
        }
    }
}

__global__ void check_restrict_kernel(const int32_t* __restrict__ r, int64_t n, int64_t limit, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (r[i] == INT32_MIN || r[i] < 0 || r[i] >= limit)) *bad = 1;
}

__global__ void fill_nan_kernel(double* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = NAN;
}

// ||x||^2 per row and the largest one (FAST mode's Gram form and its error scale)
__global__ void __launch_bounds__(128)
row_norm2_kernel(const double* __restrict__ X, int64_t n, int64_t G, double* __restrict__ norm2, unsigned long long* __restrict__ maxbits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 4 + warp;
    if (r >= n) return;
    double s = 0.0;
    for (int64_t t = lane; t < G; t += 32) { const double v = X[r * G + t]; s = fma(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        norm2[r] = s;
        atomicMax(maxbits, (unsigned long long)__double_as_longlong(s));   // non-negative doubles order like their bit patterns
    }
}

static int mode_from_env() {   // 0 = exact tiles, 1 = fast tiles on the fp64 tensor cores (default), 2 = per-cell loop, 3 = fast tiles, SIMT
    const char* e = getenv("B200MNN_SHIFTVAR");
    if (!e) return 1;
    if (strcmp(e, "exact") == 0) return 0;
    if (strcmp(e, "cell") == 0) return 2;
    if (strcmp(e, "fast_simt") == 0) return 3;
    return 1;
}

int adjust_shift_variance_device(const double* d_data1, int64_t n1, const double* d_data2, int64_t n2, int64_t G, const double* d_vect,
                                 double sigma2, const int32_t* d_r1, int64_t nr1, const int32_t* d_r2, int64_t nr2, double* d_out,
                                 int* d_bad, cudaStream_t stream) {
    B200_TRY(ensure_device());
    if (n1 < 0 || n2 < 0 || G < 0 || nr1 < 0 || nr2 < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (n2 == 0) return 0;
    Scratch ws(stream);
    int* bad = d_bad ? d_bad : ws.get<int>(1);
    double* grad = ws.get<double>((size_t)n2 * std::max<int64_t>(G, 1));
    double* l2 = ws.get<double>((size_t)n2);
    double* curproj = ws.get<double>((size_t)n2);
    int* flag_count = ws.get<int>(1);
    int32_t* flag_list = ws.get<int32_t>((size_t)n2);
    if (!ws.ok()) return B200MNN_ENOMEM;
    if (!d_bad) B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    B200_CUDA(cudaMemsetAsync(flag_count, 0, sizeof(int), stream));
    if (nr1 > 0) { check_restrict_kernel<<<(unsigned)ceil_div(nr1, 256), 256, 0, stream>>>(d_r1, nr1, n1, bad); B200_LAUNCH_CHECK(); }
    if (nr2 > 0) { check_restrict_kernel<<<(unsigned)ceil_div(nr2, 256), 256, 0, stream>>>(d_r2, nr2, n2, bad); B200_LAUNCH_CHECK(); }
    // a bad restrict index must not be dereferenced: the host-buffer layer reads `bad` and raises the reference's error;
    // device-pointer callers get the same protection by a synchronous check here
    {
        int h_bad = 0;
        B200_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
        B200_CUDA(cudaStreamSynchronize(stream));
        if (h_bad) return d_bad ? 0 : fail(B200MNN_EINVAL, "subset indices out of range");
    }
    sv_prep_kernel<<<(unsigned)ceil_div(n2, PREP_WARPS), PREP_WARPS * 32, 0, stream>>>(d_vect, d_data2, n2, G, grad, l2, curproj);
    B200_LAUNCH_CHECK();
    if (nr1 == 0) {   // ref_quan stays NA (:140): NA arithmetic gives NaN for every cell
        fill_nan_kernel<<<(unsigned)ceil_div(n2, 256), 256, 0, stream>>>(d_out, n2);
        B200_LAUNCH_CHECK();
        return 0;
    }
    const int mode = mode_from_env();
    int64_t p2 = 2;
    while (p2 < nr1) p2 <<= 1;
    const int cell_grid_all = (int)std::min<int64_t>(n2, (int64_t)sm_count() * 2);

    auto launch_cell = [&](const int* fc, const int32_t* fl, int grid) -> int {
        double* sp = ws.get<double>((size_t)grid * p2);
        double* sw = ws.get<double>((size_t)grid * p2);
        double* sl = ws.get<double>((size_t)grid * std::max<int64_t>(nr2, 1));
        unsigned char* sa = ws.get<unsigned char>((size_t)grid * std::max<int64_t>(nr2, 1));
        if (!ws.ok()) return B200MNN_ENOMEM;
        sv_cell_kernel<<<grid, THREADS, 0, stream>>>(d_data1, d_data2, G, grad, curproj, l2, sigma2, d_r1, nr1, d_r2, nr2, p2, fc, fl, n2, sp, sw,
                                                     sl, sa, d_out);
        B200_LAUNCH_CHECK();
        return 0;
    };
    if (mode == 2) return launch_cell(nullptr, nullptr, cell_grid_all);

    const bool exact = (mode == 0);
    double* norm_1 = nullptr;
    double* norm_2 = nullptr;
    double amb = 0.0, tol = 1e-10;
    if (!exact) {
        norm_1 = ws.get<double>((size_t)std::max<int64_t>(n1, 1));
        norm_2 = ws.get<double>((size_t)n2);
        unsigned long long* maxbits = ws.get<unsigned long long>(1);
        if (!ws.ok()) return B200MNN_ENOMEM;
        B200_CUDA(cudaMemsetAsync(maxbits, 0, sizeof(unsigned long long), stream));
        if (n1 > 0) { row_norm2_kernel<<<(unsigned)ceil_div(n1, 4), 128, 0, stream>>>(d_data1, n1, G, norm_1, maxbits); B200_LAUNCH_CHECK(); }
        row_norm2_kernel<<<(unsigned)ceil_div(n2, 4), 128, 0, stream>>>(d_data2, n2, G, norm_2, maxbits);
        B200_LAUNCH_CHECK();
        unsigned long long hb = 0;
        B200_CUDA(cudaMemcpyAsync(&hb, maxbits, sizeof(hb), cudaMemcpyDeviceToHost, stream));
        B200_CUDA(cudaStreamSynchronize(stream));
        double M2;
        memcpy(&M2, &hb, sizeof(double));
        // |fl(g.x) - g.x| <= G u |g||x| (sequential or FMA order alike, u = 2^-53, |g| <= 1): two such projections are compared
        amb = 4.0 * (double)std::max<int64_t>(G, 1) * 1.1102230246251565e-16 * sqrt(M2) + 1e-300;
        // relative error of a weight exp(-dl^2/sigma): the Gram form carries ~ (G + 8) u (||x_c||^2 + ||x||^2) absolute error in dl^2
        const double werr = 4.0 * (double)(G + 8) * 1.1102230246251565e-16 * 2.0 * M2 / fabs(sigma2);
        tol = std::max(tol, 8.0 * werr);
    }

    // chunk of cells whose (projection, log-weight) rows fit the buffer
    const int64_t budget = (int64_t)1 << 30;   // doubles per array (8 GiB each)
    int64_t chunk = std::max<int64_t>(TS, std::min<int64_t>(round_up(n2, TS), (budget / std::max<int64_t>(nr1, 1)) / TS * TS));
    chunk = std::min<int64_t>(chunk, (int64_t)TS * 32768);
    const int64_t ld = nr1;
    const int64_t ct1 = ceil_div(nr1, TS), ct2 = ceil_div(std::max<int64_t>(nr2, 1), TS);
    auto splits = [&](int64_t coltiles, int64_t rowtiles) {
        int64_t want = ceil_div((int64_t)sm_count() * 3, std::max<int64_t>(rowtiles, 1));
        return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, coltiles), 64));
    };
    const int64_t rowtiles_max = ceil_div(std::min(chunk, n2), TS);
    const int ns1 = splits(ct1, rowtiles_max), ns2 = splits(ct2, rowtiles_max);
    const int tps1 = (int)ceil_div(ct1, ns1), tps2 = (int)ceil_div(ct2, ns2);
    double* Pbuf = ws.get<double>((size_t)std::min(chunk, round_up(n2, TS)) * ld);
    double* Wbuf = ws.get<double>((size_t)std::min(chunk, round_up(n2, TS)) * ld);
    const bool use_dmma = (mode == 1);
    const int pm = use_dmma ? 2 : 1;   // the DMMA kernel writes one partial per (split, column group)
    double* part1 = ws.get<double>((size_t)chunk * ns1 * pm * 4);
    double* part2 = ws.get<double>((size_t)chunk * ns2 * pm * 4);
    if (!ws.ok()) return B200MNN_ENOMEM;
    for (int64_t c0 = 0; c0 < n2; c0 += chunk) {
        const int64_t nrows = std::min(chunk, n2 - c0);
        const unsigned rowtiles = (unsigned)ceil_div(nrows, TS);
        if (nr2 > 0) {
            dim3 grid((unsigned)ns2, rowtiles);
            if (exact)
                sv_pairs_kernel<true, true><<<grid, THREADS, 0, stream>>>(grad, d_data2, curproj, norm_2, c0, nrows, d_data2, norm_2, d_r2, nr2, G, sigma2,
                                                                          amb, nullptr, nullptr, 0, part2, ns2, tps2);
            else if (use_dmma)
                sv_pairs_dmma_kernel<true><<<grid, THREADS, 0, stream>>>(grad, d_data2, curproj, norm_2, c0, nrows, d_data2, norm_2, d_r2, nr2, G, sigma2,
                                                                         amb, nullptr, nullptr, 0, part2, ns2, tps2);
            else
                sv_pairs_kernel<false, true><<<grid, THREADS, 0, stream>>>(grad, d_data2, curproj, norm_2, c0, nrows, d_data2, norm_2, d_r2, nr2, G, sigma2,
                                                                           amb, nullptr, nullptr, 0, part2, ns2, tps2);
            B200_LAUNCH_CHECK();
        }
        {
            dim3 grid((unsigned)ns1, rowtiles);
            if (exact)
                sv_pairs_kernel<true, false><<<grid, THREADS, 0, stream>>>(grad, d_data2, curproj, norm_2, c0, nrows, d_data1, norm_1, d_r1, nr1, G, sigma2,
                                                                           amb, Pbuf, Wbuf, ld, part1, ns1, tps1);
            else if (use_dmma)
                sv_pairs_dmma_kernel<false><<<grid, THREADS, 0, stream>>>(grad, d_data2, curproj, norm_2, c0, nrows, d_data1, norm_1, d_r1, nr1, G, sigma2,
                                                                          amb, Pbuf, Wbuf, ld, part1, ns1, tps1);
            else
                sv_pairs_kernel<false, false><<<grid, THREADS, 0, stream>>>(grad, d_data2, curproj, norm_2, c0, nrows, d_data1, norm_1, d_r1, nr1, G, sigma2,
                                                                            amb, Pbuf, Wbuf, ld, part1, ns1, tps1);
            B200_LAUNCH_CHECK();
        }
        const int sgrid = (int)std::min<int64_t>(nrows, (int64_t)sm_count() * 8);
        sv_select_kernel<<<sgrid, THREADS, 0, stream>>>(Pbuf, Wbuf, ld, nr1, part2, ns2 * pm, nr2, part1, ns1 * pm, c0, nrows, grad, d_data1, d_r1, G, curproj, l2,
                                                        tol, exact ? 0.0 : amb, d_out, flag_count, flag_list);
        B200_LAUNCH_CHECK();
    }
    // flagged cells: the reference's loop as it stands
    B200_TRY(launch_cell(flag_count, flag_list, (int)std::min<int64_t>(n2, 64)));
    if (getenv("B200MNN_SV_DEBUG")) {   // measurement aid
        int h = 0;
        B200_CUDA(cudaMemcpyAsync(&h, flag_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
        B200_CUDA(cudaStreamSynchronize(stream));
        fprintf(stderr, "b200mnn shift variance: mode %s, %d of %lld cells went through the per-cell loop (tol %.3g, amb %.3g)\n",
                exact ? "exact" : "fast", h, (long long)n2, tol, amb);
    }
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
