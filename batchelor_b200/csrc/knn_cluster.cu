// Cluster plan for the pruned exact kNN search (see knn_cluster.cuh).
//
// Exactness never depends on the quality of the clustering.  For a query q assigned to centroid A and a reference
// cluster B (centroid c_B), let u = (c_B - c_A) / |c_B - c_A|.  Every reference x of B satisfies
//     |q - x| >= u.(x - q) = u.x - u.q >= min_{x in B} u.x - u.q                       (Cauchy-Schwarz),
// so with  ext_B(A) = max_{x in B} (x.c_A - x.c_B) / |c_B - c_A|   (how far B reaches towards A, = -min u.x) and
// v_q(B) = (q.c_B - q.c_A) / |c_B - c_A| (= u.q) the bound is  LB(q, B) = -ext_B(A) - v_q(B).  A tile of 128 queries
// takes the minimum over its rows.  All of it is computed in double with an explicit rounding margin; the scoring
// kernel skips a cluster only when S^2 LB^2 exceeds the current threshold score of EVERY row of the tile, which is
// exactly the condition under which the re-rank certificate (knn_tc.cu) treats an unseen reference like a rejected one.
//
// The k-means itself (farthest-point seeds + a few Lloyd steps on a strided sample) uses floating-point atomics, so
// centroids may differ in the last bits from run to run; only the amount of skipped work can depend on that.
#include "knn_cluster.cuh"

#include <algorithm>
#include <cmath>

namespace b200 {
namespace knn {

namespace {

constexpr int SAMPLE_MAX = 16384;   // rows of the k-means sample
constexpr int FPS_MAX = 2048;       // rows used by the farthest-point seeding (one block, serial in the seeds: 0.33 ms)
#ifndef B200_LLOYD_ITERS
#define B200_LLOYD_ITERS 4
#endif
constexpr int LLOYD_ITERS = B200_LLOYD_ITERS;

__device__ __forceinline__ unsigned long long dkey(double v) {
    const long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
    const long long b = (k & 0x8000000000000000ull) ? (long long)(k & 0x7fffffffffffffffull) : (long long)~k;
    return __longlong_as_double(b);
}

// ---------------------------------------------------------------------------------------------------------------
// k-means on a strided sample
// ---------------------------------------------------------------------------------------------------------------
__global__ void gather_sample_kernel(const double* __restrict__ X, int64_t n, int d, int m, double* __restrict__ S) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)m * d) return;
    const int64_t i = e / d;
    const int t = (int)(e - i * d);
    S[e] = X[((i * n) / m) * d + t];
}

// Transposed copy of every (m / mf)-th sample row: T[t][i], so that the seeding below reads coalesced.
__global__ void transpose_subset_kernel(const double* __restrict__ S, int m, int d, int mf, double* __restrict__ T) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= mf * d) return;
    const int t = e / mf, i = e - t * mf;
    T[e] = S[(size_t)i * (m / mf) * d + t];
}

// Farthest-point seeding over the subset: one block, FPS_ROWS rows per thread, running minimum distances in registers.
constexpr int FPS_ROWS = FPS_MAX / 1024;
__global__ void __launch_bounds__(1024) fps_seed_kernel(const double* __restrict__ T, int d, int mf, int C, double* cen) {
    __shared__ double wv[32];
    __shared__ int wi[32];
    __shared__ int pick;
    extern __shared__ double cj[];   // [d] the newest centroid
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double mind[FPS_ROWS];
#pragma unroll
    for (int r = 0; r < FPS_ROWS; ++r) mind[r] = INFINITY;
    if (tid == 0) pick = 0;
    __syncthreads();
    for (int j = 0; j < C; ++j) {
        const int pk = pick;
        for (int t = tid; t < d; t += 1024) {
            const double v = T[(size_t)t * mf + pk];
            cj[t] = v;
            cen[(size_t)j * d + t] = v;
        }
        __syncthreads();
        if (j == C - 1) break;
        double acc[FPS_ROWS];
#pragma unroll
        for (int r = 0; r < FPS_ROWS; ++r) acc[r] = 0.0;
        for (int t = 0; t < d; ++t) {
            const double c = cj[t];
#pragma unroll
            for (int r = 0; r < FPS_ROWS; ++r) {
                const int i = tid + r * 1024;
                const double df = (i < mf ? T[(size_t)t * mf + i] : c) - c;
                acc[r] = fma(df, df, acc[r]);
            }
        }
        double bv = -1.0;
        int bi = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < FPS_ROWS; ++r) {
            const int i = tid + r * 1024;
            if (i < mf) {
                mind[r] = fmin(mind[r], acc[r]);
                if (mind[r] > bv) { bv = mind[r]; bi = i; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { wv[warp] = bv; wi[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double fv = wv[0];
            int fi = wi[0];
            for (int w = 1; w < 32; ++w)
                if (wv[w] > fv || (wv[w] == fv && wi[w] < fi)) { fv = wv[w]; fi = wi[w]; }
            pick = (fi == 0x7fffffff) ? 0 : fi;
        }
        __syncthreads();
    }
}

// Centroids transposed into shared memory ([t][c], c fastest) and the dot products of one row with 16 of them: per
// dimension one row element and eight broadcast 16-byte loads feed sixteen FMAs.
__global__ void transpose_centroids_kernel(const double* __restrict__ cen, int C, int d, double* __restrict__ cenT_g) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= C * d) return;
    const int t = e / C, c = e - t * C;
    cenT_g[e] = cen[(size_t)c * d + t];
}
// cenT_g: the centroids already transposed ([t][c]) in global memory -> linear, conflict-free copy
__device__ __forceinline__ void load_centroids_t(const double* __restrict__ cenT_g, int C, int d, double* cenT) {
    for (int e = threadIdx.x; e < C * d; e += blockDim.x) cenT[e] = cenT_g[e];
}
template <int PASS>
__device__ __forceinline__ void dots(const double* __restrict__ x, int d, int C, const double* __restrict__ cenT, int c0, double (&a)[PASS]) {
#pragma unroll
    for (int u = 0; u < PASS; ++u) a[u] = 0.0;
    for (int t = 0; t < d; ++t) {
        const double xv = x[t];
        const double2* cp = reinterpret_cast<const double2*>(cenT + t * C + c0);
#pragma unroll
        for (int u = 0; u < PASS / 2; ++u) {
            const double2 cv = cp[u];
            a[2 * u] = fma(xv, cv.x, a[2 * u]);
            a[2 * u + 1] = fma(xv, cv.y, a[2 * u + 1]);
        }
    }
}

// Nearest centroid of one row (C is a multiple of PASS).  With PASS = 64 the row is read once per 64 centroids and
// the loop is FMA-bound (one row element + 32 broadcast 16-byte loads feed 64 FMAs).
template <int PASS>
__device__ __forceinline__ int nearest_centroid(const double* __restrict__ x, int d, int C, const double* __restrict__ cenT,
                                                const double* __restrict__ cnorm, double* __restrict__ dots_row = nullptr) {
    double best = INFINITY;
    int bi = 0;
    for (int c0 = 0; c0 < C; c0 += PASS) {
        double a[PASS];
        dots<PASS>(x, d, C, cenT, c0, a);
        if (dots_row) {   // kept for the projection kernels: the same C dot products, not computed twice
            double2* o = reinterpret_cast<double2*>(dots_row + c0);
#pragma unroll
            for (int u = 0; u < PASS / 2; ++u) o[u] = make_double2(a[2 * u], a[2 * u + 1]);
        }
#pragma unroll
        for (int u = 0; u < PASS; ++u) {
            const double sc = cnorm[c0 + u] - 2.0 * a[u];
            if (sc < best) { best = sc; bi = c0 + u; }
        }
    }
    return bi;
}

__global__ void centroid_norm_kernel(const double* __restrict__ cen, int C, int d, double* __restrict__ cnorm) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int t = 0; t < d; ++t) s = fma(cen[(size_t)c * d + t], cen[(size_t)c * d + t], s);
    cnorm[c] = s;
}

// Stages 128 consecutive rows (row-major, d doubles each) into shared memory with an odd pitch.
__device__ __forceinline__ void stage_rows(const double* __restrict__ X, int64_t row0, int64_t nrows, int d, int dp, double* rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = warp; r < CL_TILE; r += nw) {
        const int64_t i = row0 + r;
        for (int t = lane; t < d; t += 32) rows[r * dp + t] = (i < nrows) ? X[i * d + t] : 0.0;
    }
}

// One Lloyd step on the sample: nearest centroid, then sums / counts by atomics.
__global__ void __launch_bounds__(CL_TILE) lloyd_accum_kernel(const double* __restrict__ S, int m, int d, int dp, int C,
                                                            const double* __restrict__ cen, const double* __restrict__ cnorm,
                                                            double* __restrict__ sums, int* __restrict__ counts) {
    extern __shared__ double rows[];
    double* cenT = rows + CL_TILE * dp;
    const int64_t row0 = (int64_t)blockIdx.x * CL_TILE;
    stage_rows(S, row0, m, d, dp, rows);
    load_centroids_t(cen, C, d, cenT);
    __syncthreads();
    const int64_t i = row0 + threadIdx.x;
    if (i >= m) return;
    const double* x = rows + threadIdx.x * dp;
    const int c = nearest_centroid<16>(x, d, C, cenT, cnorm);
    for (int t = 0; t < d; ++t) atomicAdd(sums + (size_t)c * d + t, x[t]);
    atomicAdd(counts + c, 1);
}

__global__ void lloyd_update_kernel(double* __restrict__ cen, int C, int d, double* __restrict__ sums, int* __restrict__ counts) {
    const int c = blockIdx.x;
    const int cnt = counts[c];
    for (int t = threadIdx.x; t < d; t += blockDim.x) {
        if (cnt > 0) cen[(size_t)c * d + t] = sums[(size_t)c * d + t] / (double)cnt;
        sums[(size_t)c * d + t] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) counts[c] = 0;
}

// cdist[a][b] = |c_a - c_b| and its reciprocal (the projection kernels multiply; the rounding is inside their margin)
__global__ void centroid_dist_kernel(const double* __restrict__ cen, int C, int d, double* __restrict__ cdist, double* __restrict__ cinv) {
    const int a = blockIdx.x;
    for (int b = threadIdx.x; b < C; b += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < d; ++t) { const double df = cen[(size_t)a * d + t] - cen[(size_t)b * d + t]; s = fma(df, df, s); }
        const double D = sqrt(s);
        cdist[(size_t)a * C + b] = D;
        cinv[(size_t)a * C + b] = D > 0.0 ? 1.0 / D : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Assignment of every row, grouping
// ---------------------------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(CL_TILE) assign_kernel(const double* __restrict__ X, int64_t n, int d, int C,
                                                       const double* __restrict__ cen, const double* __restrict__ cnorm,
                                                       int32_t* __restrict__ cid, int* __restrict__ counts,
                                                       double* __restrict__ dots_out /* optional [n][C]: x . c for every centroid */) {
    extern __shared__ double cenT[];   // [d][C]
    load_centroids_t(cen, C, d, cenT);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * CL_TILE + threadIdx.x;
    const bool valid = i < n;
    int c = -1;
    if (valid) {
        // a warp's 32 rows are contiguous: L1 serves the strided reads
        c = nearest_centroid<PASS>(X + i * d, d, C, cenT, cnorm, dots_out ? dots_out + (size_t)i * C : nullptr);
        cid[i] = c;
    }
    // warp-aggregated histogram
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (valid && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + c, __popc(peers));
}

// Reference side: tile0[c] = first reference tile of cluster c, row_base[c] = its first grouped row.
__global__ void ref_offsets_kernel(const int* __restrict__ cnt_ref, int C, int* __restrict__ tile0, int* __restrict__ row_base) {
    if (threadIdx.x == 0) {
        int t = 0;
        for (int c = 0; c < C; ++c) {
            tile0[c] = t;
            row_base[c] = t * CL_TILE;
            t += (cnt_ref[c] + CL_TILE - 1) / CL_TILE;
        }
        tile0[C] = t;
    }
}
// Query side: slot_base[c] = first query slot of cluster c.  Query slots are laid out by DESCENDING reference count of
// their cluster: a query tile's work is roughly the size of its own cluster, and starting the heavy CTAs first shortens
// the tail of the scoring kernel.
__global__ void query_offsets_kernel(const int* __restrict__ cnt_ref, const int* __restrict__ cnt_q, int C, int* __restrict__ slot_base,
                                     int* __restrict__ nslots) {
    __shared__ int order[CL_MAXC];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {   // rank of cluster c by (reference count descending, id ascending)
        int rank = 0;
        for (int o = 0; o < C; ++o) rank += (cnt_ref[o] > cnt_ref[c] || (cnt_ref[o] == cnt_ref[c] && o < c)) ? 1 : 0;
        order[rank] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < C; ++i) {
            const int c = order[i];
            slot_base[c] = s;
            s += ((cnt_q[c] + CL_TILE - 1) / CL_TILE) * CL_TILE;
        }
        *nslots = s;
    }
}

__global__ void scatter_kernel(const int32_t* __restrict__ cid, int64_t n, const int* __restrict__ base, int* __restrict__ cursor,
                               int32_t* __restrict__ map) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const int c = valid ? cid[i] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int start = 0;
    if (valid && lane == leader) start = atomicAdd(cursor + c, __popc(peers));
    start = __shfl_sync(0xffffffffu, start, leader);
    if (valid) map[base[c] + start + __popc(peers & ((1u << lane) - 1u))] = (int32_t)i;
}

// ---------------------------------------------------------------------------------------------------------------
// Projections of a tile's rows on the centroid axes
// ---------------------------------------------------------------------------------------------------------------
// Block = 128 threads, one row each, every cluster in passes of PASS (dots).
// MODE 0 (reference tile): vref[P][c] = max over rows of (x.c_c - x.c_P) / D(P,c) + margin   (P = the tile's cluster)
// MODE 1 (query tile)    : LB[c] = min over rows of -ext_c(P_r) - (q.c_c - q.c_P_r) / D(P_r,c) - margin, then the sorted
//                          cluster list and the per-slot score offsets.
template <int MODE, int PASS>
__global__ void __launch_bounds__(CL_TILE) tile_project_kernel(const double* __restrict__ X, int d, int dp, int C, const int32_t* __restrict__ map,
                                                             const int* __restrict__ count, const int32_t* __restrict__ cid,
                                                             const double* __restrict__ cen, const double* __restrict__ cdist,
                                                             const double* __restrict__ cinv, unsigned long long* __restrict__ vref, const unsigned long long* __restrict__ maxnorm_bits,
                                                             const double* __restrict__ qnorm, const int* __restrict__ scale_exp,
                                                             int2* __restrict__ lists, float* __restrict__ qoff,
                                                             const double* __restrict__ dots_in /* optional [rows][C] from assign_kernel */,
                                                             const double* __restrict__ norms /* with dots_in: squared row norms */) {
    // With dots_in the C dot products of every row come from the assignment pass (bit-identical: same FMA order) and this
    // kernel only reduces them.  Otherwise:
    // Shared memory holds the transposed centroids only: every thread reads its (gathered) row straight from global
    // memory -- 400 contiguous bytes that stay in L1 across the passes -- so eight blocks fit an SM instead of two
    // (staging the 128 rows took 52 KB per block and left the kernel latency-bound at 8 warps per SM).
    extern __shared__ double cenT[];                 // [d][C] transposed centroids
    __shared__ unsigned long long red[CL_MAXC];
    __shared__ int src_s[CL_TILE];
    (void)dp;
    const int64_t row0 = (int64_t)blockIdx.x * CL_TILE;
    if (count && row0 >= (int64_t)*count) return;
    const int tid = threadIdx.x, lane = tid & 31;
    src_s[tid] = (!count || row0 + tid < (int64_t)*count) ? map[row0 + tid] : -1;
    for (int c = tid; c < CL_MAXC; c += CL_TILE) red[c] = (MODE == 0) ? dkey(-INFINITY) : dkey(INFINITY);
    __syncthreads();
    if (MODE == 0 && src_s[0] < 0) return;           // clusters are padded at their end: an empty first row = an unused tile
    if (!dots_in) {
        load_centroids_t(cen, C, d, cenT);
        __syncthreads();
    }
    const int src = src_s[tid];
    const bool valid = src >= 0;
    const int P = valid ? cid[src] : 0;
    const int64_t rsrc = valid ? src : (src_s[0] < 0 ? 0 : src_s[0]);   // padding rows read a valid row; their values are discarded
    const double* x = X + rsrc * d;
    const double* dr = dots_in ? dots_in + (size_t)rsrc * C : nullptr;
    const double M = sqrt(__longlong_as_double((long long)*maxnorm_bits));
    double gP = 0.0, xn = 0.0;
    if (dr) {
        gP = dr[P];
        xn = norms[rsrc];
    } else {
        for (int t = 0; t < d; ++t) { const double xv = x[t]; gP = fma(xv, cenT[t * C + P], gP); xn = fma(xv, xv, xn); }
    }
    const double mg_num = 1e-11 * (sqrt(xn) + M) * M;   // >= 1000x the rounding error of the two dot products and the reciprocal
    const double tiny = 1e-5 * M;
    for (int c0 = 0; c0 < C; c0 += PASS) {
        double a[PASS];
        if (dr) {
            const double2* ip = reinterpret_cast<const double2*>(dr + c0);
#pragma unroll
            for (int u = 0; u < PASS / 2; ++u) { const double2 v2 = ip[u]; a[2 * u] = v2.x; a[2 * u + 1] = v2.y; }
        } else {
            dots<PASS>(x, d, C, cenT, c0, a);
        }
#pragma unroll
        for (int u = 0; u < PASS; ++u) {
            const int c = c0 + u;
            const double D = cdist[(size_t)P * C + c];
            const double iD = cinv[(size_t)P * C + c];
            double val;
            if (MODE == 0) {
                // extent of this row towards centroid c, rounded up; coincident centroids give no usable axis
                val = (valid && c != P) ? ((D > tiny) ? (a[u] - gP + mg_num) * iD : INFINITY) : -INFINITY;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val = fmax(val, __shfl_xor_sync(0xffffffffu, val, o));
                if (lane == 0) atomicMax(&red[c], dkey(val));
            } else {
                if (!valid) val = INFINITY;
                else if (c == P || !(D > tiny)) val = -INFINITY;
                else {
                    const double ext = dkey_inv(vref[(size_t)c * C + P]);   // -inf for an empty cluster -> LB = +inf
                    val = -ext - (a[u] - gP + mg_num) * iD;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, o));
                if (lane == 0) atomicMin(&red[c], dkey(val));
            }
        }
    }
    __syncthreads();
    if (MODE == 0) {
        const int Pt = cid[src_s[0]];
        for (int c = tid; c < C; c += CL_TILE)
            if (c != Pt) atomicMax(&vref[(size_t)Pt * C + c], red[c]);
        return;
    }
    // MODE 1: keys (float bits of S^2 LB^2 rounded down, cluster), ascending
    const double S2 = scalbn(1.0, 2 * (*scale_exp));
    int CP = 1;
    while (CP < C) CP <<= 1;
    unsigned long long keys[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = tid + h * CL_TILE;
        keys[h] = ~0ull;
        if (c < C) {
            const double LB = dkey_inv(red[c]);
            float lb;
            if (!(LB > 0.0)) lb = 0.f;
            else if (isinf(LB)) lb = __int_as_float(0x7f800000);
            else lb = __double2float_rd(S2 * LB * LB * (1.0 - 1e-6));
            keys[h] = ((unsigned long long)__float_as_uint(lb) << 32) | (unsigned)c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (tid + h * CL_TILE < CP) red[tid + h * CL_TILE] = keys[h];
    __syncthreads();
    for (int k = 2; k <= CP; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = tid + h * CL_TILE;
                const int partner = i ^ j;
                if (i < CP && partner > i) {
                    const unsigned long long a0 = red[i], b0 = red[partner];
                    const bool up = (i & k) == 0;
                    if ((a0 > b0) == up) { red[i] = b0; red[partner] = a0; }
                }
            }
            __syncthreads();
        }
    }
    for (int c = tid; c < C; c += CL_TILE) {
        const unsigned long long kk = red[c];
        lists[(size_t)blockIdx.x * C + c] = make_int2((int)(unsigned)kk, (int)(unsigned)(kk >> 32));
    }
    {
        const int s2 = src_s[tid];
        qoff[row0 + tid] = (s2 >= 0) ? __double2float_ru(S2 * qnorm[s2] * (1.0 + 1e-6)) : __int_as_float(0xff800000);
    }
}

// Second tier: the uncertified queries (an unordered device list) regrouped into cluster-pure, padded slots.
__global__ void list_hist_kernel(const int32_t* __restrict__ list, const int* __restrict__ count, const int32_t* __restrict__ cid,
                                 int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *count) atomicAdd(cnt + cid[list[i]], 1);
}
__global__ void list_offsets_kernel(const int* __restrict__ cnt, int C, int* __restrict__ base, int* __restrict__ nslots, int* __restrict__ cursors) {
    if (threadIdx.x == 0) {
        int s = 0;
        for (int c = 0; c < C; ++c) {
            base[c] = s;
            s += ((cnt[c] + CL_TILE - 1) / CL_TILE) * CL_TILE;
        }
        *nslots = s;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) cursors[c] = 0;
}
__global__ void list_scatter_kernel(const int32_t* __restrict__ list, const int* __restrict__ count, const int32_t* __restrict__ cid,
                                    const int* __restrict__ base, int* __restrict__ cursors, int32_t* __restrict__ map) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    const int q = list[i];
    const int c = cid[q];
    map[base[c] + atomicAdd(cursors + c, 1)] = q;
}

__global__ void fill_vref_kernel(unsigned long long* __restrict__ vref, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vref[i] = dkey(-INFINITY);
}

}  // namespace

static int set_smem(const void* fn, size_t bytes) {
    B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// the dot products of every row with every centroid are kept between the assignment and the projection kernels unless
// that would take more than this many bytes (then the projection kernels compute them again)
static const size_t kMaxDotsBytes = (size_t)6 << 30;

int build_ref_plan(const double* dX, int64_t n, int d, int C, const double* xnorm, const unsigned long long* maxnorm_bits, Scratch& ws,
                   cudaStream_t stream, ClusterPlan* plan) {
    if (C < 16 || C > CL_MAXC || (C & (C - 1)) != 0) return fail(B200MNN_EINVAL, "internal: cluster count must be a power of two in [16, 256]");
    int m = (int)std::min<int64_t>(n, SAMPLE_MAX);
    const int mf = std::min(m, FPS_MAX);
    m = (m / mf) * mf;          // the seeding walks the sample with an integer stride
    if (mf < C) return fail(B200MNN_EINVAL, "internal: too few rows for the requested number of clusters");
    const int dp = d | 1;
    const size_t row_smem = ((size_t)CL_TILE * dp + (size_t)C * d) * sizeof(double);   // staged rows + transposed centroids (Lloyd steps)
    if (row_smem > (size_t)200 * 1024) return fail(B200MNN_EINVAL, "internal: too many dimensions for the cluster plan");

    ClusterPlan& p = *plan;
    p.C = C;
    p.n_rows_max = round_up(n, CL_TILE) + (int64_t)C * CL_TILE;
    p.refmap = ws.get<int32_t>((size_t)p.n_rows_max);
    p.centroids = ws.get<double>((size_t)C * d);
    p.cdist = ws.get<double>((size_t)C * C);
    p.cinv = ws.get<double>((size_t)C * C);
    p.centroids_t = ws.get<double>((size_t)C * d);
    p.vref = ws.get<unsigned long long>((size_t)C * C);
    p.cnorm = ws.get<double>((size_t)C);
    int32_t* cid_ref = ws.get<int32_t>((size_t)n);
    double* sample = ws.get<double>((size_t)m * d);
    double* subset_t = ws.get<double>((size_t)mf * d);
    double* sums = ws.get<double>((size_t)C * d);
    int* ints = ws.get<int>((size_t)5 * CL_MAXC + 16);
    if (!ws.ok()) return B200MNN_ENOMEM;
    Scratch tmp(stream);                   // released (stream-ordered) when this function returns
    double* dots_ref = ((size_t)n * C * sizeof(double) <= kMaxDotsBytes) ? tmp.get<double>((size_t)n * C) : nullptr;
    int* counts = ints;                    // [C]   Lloyd counts
    p.cnt_ref = ints + CL_MAXC;            // [C]
    int* row_base = ints + 2 * CL_MAXC;    // [C]
    int* cursors = ints + 3 * CL_MAXC;     // [C]
    p.cl_tile0 = ints + 4 * CL_MAXC;       // [C + 1]
    B200_CUDA(cudaMemsetAsync(ints, 0, sizeof(int) * (5 * CL_MAXC + 16), stream));
    B200_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)C * d, stream));
    B200_CUDA(cudaMemsetAsync(p.refmap, 0xFF, sizeof(int32_t) * (size_t)p.n_rows_max, stream));

    const size_t cen_smem = (size_t)C * d * sizeof(double);
    B200_TRY(set_smem((const void*)lloyd_accum_kernel, row_smem));
    B200_TRY(set_smem((const void*)assign_kernel<16>, cen_smem));
    B200_TRY(set_smem((const void*)assign_kernel<64>, cen_smem));
    B200_TRY(set_smem((const void*)tile_project_kernel<0, 16>, cen_smem));
    B200_TRY(set_smem((const void*)tile_project_kernel<1, 16>, cen_smem));
    B200_TRY(set_smem((const void*)tile_project_kernel<0, 64>, cen_smem));
    B200_TRY(set_smem((const void*)tile_project_kernel<1, 64>, cen_smem));

    gather_sample_kernel<<<(unsigned)ceil_div((int64_t)m * d, 256), 256, 0, stream>>>(dX, n, d, m, sample);
    B200_LAUNCH_CHECK();
    transpose_subset_kernel<<<(unsigned)ceil_div((int64_t)mf * d, 256), 256, 0, stream>>>(sample, m, d, mf, subset_t);
    B200_LAUNCH_CHECK();
    fps_seed_kernel<<<1, 1024, (size_t)d * sizeof(double), stream>>>(subset_t, d, mf, C, p.centroids);
    B200_LAUNCH_CHECK();
    for (int it = 0; it < LLOYD_ITERS; ++it) {
        centroid_norm_kernel<<<1, CL_MAXC, 0, stream>>>(p.centroids, C, d, p.cnorm);
        B200_LAUNCH_CHECK();
        transpose_centroids_kernel<<<(unsigned)ceil_div((int64_t)C * d, 256), 256, 0, stream>>>(p.centroids, C, d, p.centroids_t);
        B200_LAUNCH_CHECK();
        lloyd_accum_kernel<<<(unsigned)ceil_div(m, CL_TILE), CL_TILE, row_smem, stream>>>(sample, m, d, dp, C, p.centroids_t, p.cnorm, sums, counts);
        B200_LAUNCH_CHECK();
        lloyd_update_kernel<<<C, 64, 0, stream>>>(p.centroids, C, d, sums, counts);
        B200_LAUNCH_CHECK();
    }
    centroid_norm_kernel<<<1, CL_MAXC, 0, stream>>>(p.centroids, C, d, p.cnorm);
    B200_LAUNCH_CHECK();
    centroid_dist_kernel<<<C, 64, 0, stream>>>(p.centroids, C, d, p.cdist, p.cinv);
    B200_LAUNCH_CHECK();
    transpose_centroids_kernel<<<(unsigned)ceil_div((int64_t)C * d, 256), 256, 0, stream>>>(p.centroids, C, d, p.centroids_t);
    B200_LAUNCH_CHECK();

    if (C % 64 == 0) assign_kernel<64><<<(unsigned)ceil_div(n, CL_TILE), CL_TILE, cen_smem, stream>>>(dX, n, d, C, p.centroids_t, p.cnorm, cid_ref, p.cnt_ref, dots_ref);
    else assign_kernel<16><<<(unsigned)ceil_div(n, CL_TILE), CL_TILE, cen_smem, stream>>>(dX, n, d, C, p.centroids_t, p.cnorm, cid_ref, p.cnt_ref, dots_ref);
    B200_LAUNCH_CHECK();
    ref_offsets_kernel<<<1, 32, 0, stream>>>(p.cnt_ref, C, p.cl_tile0, row_base);
    B200_LAUNCH_CHECK();
    scatter_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(cid_ref, n, row_base, cursors, p.refmap);
    B200_LAUNCH_CHECK();

    fill_vref_kernel<<<(unsigned)ceil_div((int64_t)C * C, 256), 256, 0, stream>>>(p.vref, C * C);
    B200_LAUNCH_CHECK();
    if (C % 64 == 0)
        tile_project_kernel<0, 64><<<(unsigned)(p.n_rows_max / CL_TILE), CL_TILE, cen_smem, stream>>>(dX, d, dp, C, p.refmap, nullptr, cid_ref, p.centroids_t,
                                                                                                p.cdist, p.cinv, p.vref, maxnorm_bits, nullptr, nullptr, nullptr, nullptr,
                                                                                                dots_ref, xnorm);
    else
        tile_project_kernel<0, 16><<<(unsigned)(p.n_rows_max / CL_TILE), CL_TILE, cen_smem, stream>>>(dX, d, dp, C, p.refmap, nullptr, cid_ref, p.centroids_t,
                                                                                                p.cdist, p.cinv, p.vref, maxnorm_bits, nullptr, nullptr, nullptr, nullptr,
                                                                                                dots_ref, xnorm);
    B200_LAUNCH_CHECK();
    return 0;
}

int build_query_plan(ClusterPlan* plan, const double* dQ, int64_t nq, int d, const double* qnorm, const int* scale_exp,
                     const unsigned long long* maxnorm_bits, Scratch& ws, cudaStream_t stream) {
    ClusterPlan& p = *plan;
    const int C = p.C;
    p.nslots_max = round_up(nq, CL_TILE) + (int64_t)C * CL_TILE;
    p.qmap = ws.get<int32_t>((size_t)p.nslots_max);
    p.cl_list = ws.get<int2>((size_t)(p.nslots_max / CL_TILE) * C);
    p.qoff = ws.get<float>((size_t)p.nslots_max);
    p.cid_q = ws.get<int32_t>((size_t)nq);
    int* ints = ws.get<int>((size_t)3 * CL_MAXC + 16);
    p.dots_q = ((size_t)nq * C * sizeof(double) <= kMaxDotsBytes) ? ws.get<double>((size_t)nq * C) : nullptr;
    if (!ws.ok()) return B200MNN_ENOMEM;
    int* cnt_q = ints;                     // [C]
    int* slot_base = ints + CL_MAXC;       // [C]
    int* cursors = ints + 2 * CL_MAXC;     // [C]
    p.nslots = ints + 3 * CL_MAXC + 8;
    B200_CUDA(cudaMemsetAsync(ints, 0, sizeof(int) * (3 * CL_MAXC + 16), stream));
    B200_CUDA(cudaMemsetAsync(p.qmap, 0xFF, sizeof(int32_t) * (size_t)p.nslots_max, stream));
    const size_t cen_smem = (size_t)C * d * sizeof(double);
    if (C % 64 == 0) assign_kernel<64><<<(unsigned)ceil_div(nq, CL_TILE), CL_TILE, cen_smem, stream>>>(dQ, nq, d, C, p.centroids_t, p.cnorm, p.cid_q, cnt_q, p.dots_q);
    else assign_kernel<16><<<(unsigned)ceil_div(nq, CL_TILE), CL_TILE, cen_smem, stream>>>(dQ, nq, d, C, p.centroids_t, p.cnorm, p.cid_q, cnt_q, p.dots_q);
    B200_LAUNCH_CHECK();
    query_offsets_kernel<<<1, 256, 0, stream>>>(p.cnt_ref, cnt_q, C, slot_base, p.nslots);
    B200_LAUNCH_CHECK();
    scatter_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, stream>>>(p.cid_q, nq, slot_base, cursors, p.qmap);
    B200_LAUNCH_CHECK();
    B200_TRY(build_tile_lists(p, dQ, d, p.qmap, p.nslots, p.nslots_max, qnorm, scale_exp, maxnorm_bits, p.cl_list, p.qoff, stream));
    return 0;
}

int regroup_query_list(const ClusterPlan& p, const int32_t* list, const int* count, int64_t max_items, int32_t* map, int64_t max_slots,
                       int* nslots, int* work /* [3 * CL_MAXC] */, cudaStream_t stream) {
    B200_CUDA(cudaMemsetAsync(work, 0, sizeof(int) * 3 * CL_MAXC, stream));
    B200_CUDA(cudaMemsetAsync(map, 0xFF, sizeof(int32_t) * (size_t)max_slots, stream));
    const unsigned grid = (unsigned)ceil_div(max_items, 256);
    list_hist_kernel<<<grid, 256, 0, stream>>>(list, count, p.cid_q, work);
    B200_LAUNCH_CHECK();
    list_offsets_kernel<<<1, 256, 0, stream>>>(work, p.C, work + CL_MAXC, nslots, work + 2 * CL_MAXC);
    B200_LAUNCH_CHECK();
    list_scatter_kernel<<<grid, 256, 0, stream>>>(list, count, p.cid_q, work + CL_MAXC, work + 2 * CL_MAXC, map);
    B200_LAUNCH_CHECK();
    return 0;
}

int build_tile_lists(const ClusterPlan& p, const double* dQ, int d, const int32_t* qmap, const int* count, int64_t max_slots,
                     const double* qnorm, const int* scale_exp, const unsigned long long* maxnorm_bits, int2* lists, float* qoff,
                     cudaStream_t stream) {
    const int dp = d | 1;
    const size_t row_smem = (size_t)p.C * d * sizeof(double);   // transposed centroids
    if (p.C % 64 == 0)
        tile_project_kernel<1, 64><<<(unsigned)(max_slots / CL_TILE), CL_TILE, row_smem, stream>>>(dQ, d, dp, p.C, qmap, count, p.cid_q, p.centroids_t, p.cdist,
                                                                                             p.cinv, p.vref, maxnorm_bits, qnorm, scale_exp, lists, qoff, p.dots_q, qnorm);
    else
        tile_project_kernel<1, 16><<<(unsigned)(max_slots / CL_TILE), CL_TILE, row_smem, stream>>>(dQ, d, dp, p.C, qmap, count, p.cid_q, p.centroids_t, p.cdist,
                                                                                             p.cinv, p.vref, maxnorm_bits, qnorm, scale_exp, lists, qoff, p.dots_q, qnorm);
    B200_LAUNCH_CHECK();
    return 0;
}

}  // namespace knn
}  // namespace b200
