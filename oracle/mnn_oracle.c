/* TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
 *
 * Plain-C, fp64 restatement of the MNN hot path of LTLA/batchelor (reference v1.23.1 under
 * /root/reference).  It is the parity checker for the CUDA path and (bench.py only) the "port"
 * CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it; the product library (batchelor_b200/csrc) never does.
 *
 * Parity status of each function:
 *   oracle_find_mutual_nns        follows src/find_mutual_nns.cpp:8-41; PINNED against the reference's own
 *                                 object code (oracle/_ref) in tests/test_oracle.py and tests/golden/.
 *   oracle_smooth_gaussian_kernel follows src/smooth_gaussian_kernel.cpp:11-117; PINNED the same way and
 *                                 against the REF formula of tests/testthat/test-mnn-correct.R:36-65.
 *   oracle_adjust_shift_variance  follows src/adjust_shift_variance.cpp:9-164; PINNED the same way and against
 *                                 the REF formula of tests/testthat/test-mnn-correct.R:101-138.
 *   oracle_query_knn              PARITY UNPINNED: the arithmetic lives in BiocNeighbors (KMKNN on knncolle;
 *                                 Imports, no version pin, DESCRIPTION:17), which is absent from /root/reference
 *                                 and from this image.  Restated from its published contract as used at
 *                                 R/MNN_tree.R:129 and R/fastMNN.R:605: exact Euclidean search, squared
 *                                 distance accumulated in double in dimension order, the k smallest under the
 *                                 total order (distance, index), reported in ascending order, distance = sqrt.
 *
 * All matrices are column-major like R's unless a function says otherwise.  Compile with
 * -ffp-contract=off so that x86-64 and this restatement agree bit for bit (no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Rmath's logspace_add (used at src/smooth_gaussian_kernel.cpp:63,79 and
 * src/adjust_shift_variance.cpp:101,108,129,150). */
static double lse2(double a, double b) {
    return (a > b ? a : b) + log1p(exp(-fabs(a - b)));
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * a1. exact kNN (what BiocNeighbors::queryKNN(X, query, k, BNPARAM=KmknnParam()) returns).
 * X [n x d], Q [nq x d]; col_major!=0: R layout (element (i,t) at [i + t*n]); else row-major.
 * idx_out [nq x k] int32 1-BASED, dist_out [nq x k] (may be NULL), both column-major like R.
 * k is capped to n by the caller (BiocNeighbors warns and caps).
 * ------------------------------------------------------------------------------------------ */
int oracle_query_knn(const double* X, int64_t n, const double* Q, int64_t nq, int d, int k, int col_major,
                     int32_t* idx_out, double* dist_out, int nthreads) {
    if (k > n || k < 0 || d < 0) return 1;
    if (k == 0 || nq == 0) return 0;
    /* Work on row-major copies so the inner loop streams. */
    double* Xr = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (size_t)(d > 0 ? d : 1));
    double* Qr = (double*)malloc(sizeof(double) * (size_t)nq * (size_t)(d > 0 ? d : 1));
    if (!Xr || !Qr) { free(Xr); free(Qr); return 2; }
    for (int64_t i = 0; i < n; ++i)
        for (int t = 0; t < d; ++t) Xr[i * d + t] = col_major ? X[i + (int64_t)t * n] : X[i * d + t];
    for (int64_t i = 0; i < nq; ++i)
        for (int t = 0; t < d; ++t) Qr[i * d + t] = col_major ? Q[i + (int64_t)t * nq] : Q[i * d + t];
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        double* bd = (double*)malloc(sizeof(double) * (size_t)k);
        int64_t* bi = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
#pragma omp for schedule(dynamic, 16)
        for (int64_t q = 0; q < nq; ++q) {
            const double* qv = Qr + q * d;
            int have = 0;
            for (int64_t j = 0; j < n; ++j) {
                const double* xv = Xr + j * d;
                double s = 0.0;
                for (int t = 0; t < d; ++t) {
                    const double df = qv[t] - xv[t];
                    s += df * df;
                }
                /* sorted insertion under (distance, index); j ascends so ties keep the earlier index */
                if (have < k) {
                    int p = have++;
                    while (p > 0 && bd[p - 1] > s) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                    bd[p] = s; bi[p] = j;
                } else if (s < bd[k - 1]) {
                    int p = k - 1;
                    while (p > 0 && bd[p - 1] > s) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                    bd[p] = s; bi[p] = j;
                }
            }
            for (int r = 0; r < k; ++r) {
                idx_out[q + (int64_t)r * nq] = (int32_t)(bi[r] + 1);
                if (dist_out) dist_out[q + (int64_t)r * nq] = sqrt(bd[r]);
            }
        }
        free(bd); free(bi);
    }
    free(Xr); free(Qr);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a3. mutual pairs, src/find_mutual_nns.cpp:8-41.
 * left [n1 x k2] (ids of batch-2 cells, 1-based), right [n2 x k1] (ids of batch-1 cells, 1-based),
 * both column-major.  Emits (l+1, v) for l ascending and, within l, in left's column order, iff
 * l+1 occurs in row v of right.  Outputs must hold n1*k2 entries.
 * ------------------------------------------------------------------------------------------ */
static int cmp_i32(const void* a, const void* b) {
    const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

int oracle_find_mutual_nns(const int32_t* left, int64_t n1, int k2, const int32_t* right, int64_t n2, int k1,
                           int32_t* first_out, int32_t* second_out, int64_t* np_out) {
    int32_t* sorted = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n2 * k1 + 1));
    if (!sorted) return 2;
    for (int64_t r = 0; r < n2; ++r) {
        for (int c = 0; c < k1; ++c) sorted[r * k1 + c] = right[r + (int64_t)c * n2];
        qsort(sorted + r * k1, (size_t)k1, sizeof(int32_t), cmp_i32);
    }
    int64_t np = 0;
    for (int64_t l = 0; l < n1; ++l) {
        const int32_t want = (int32_t)(l + 1);
        for (int c = 0; c < k2; ++c) {
            const int32_t v = left[l + (int64_t)c * n1];
            if (v < 1 || v > n2) { free(sorted); return 1; }
            const int32_t* row = sorted + (int64_t)(v - 1) * k1;
            int lo = 0, hi = k1; /* lower bound */
            while (lo < hi) { int mid = (lo + hi) / 2; if (row[mid] < want) lo = mid + 1; else hi = mid; }
            if (lo < k1 && row[lo] == want) { first_out[np] = want; second_out[np] = v; ++np; }
        }
    }
    *np_out = np;
    free(sorted);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a5. Gaussian smoothing, src/smooth_gaussian_kernel.cpp:11-117.
 * averaged [G x nmnn], index0 int32[nmnn] 0-based columns of mat, mat [Gdist x ncells] -> out [G x ncells].
 * Restated per output cell instead of per MNN cell:
 *   logw(c,i) = -||m_c - m_{U_i}||^2 / sigma2            (:36-52; sigma2 is NOT squared again)
 *   dens_i    = logsumexp_j logw(U_j, i)                  (:56-65)
 *   out[:,c]  = sum_i exp(logw(c,i) - dens_i) avg[:,i] / sum_i exp(logw(c,i) - dens_i)   (:75-115)
 * The reference keeps one running exponent per (gene, cell); all genes of a cell share the same
 * sequence of log-multipliers, so a per-cell running maximum is the same computation.
 * Returns 1 on the reference's size check failure (:18-20).
 * ------------------------------------------------------------------------------------------ */
int oracle_smooth_gaussian_kernel(const double* averaged, int64_t G, int64_t nmnn, const int32_t* index0, int64_t nidx,
                                  const double* mat, int64_t Gdist, int64_t ncells, double sigma2, double* out,
                                  int nthreads) {
    if (nmnn != nidx) return 1;
    for (int64_t i = 0; i < nmnn; ++i)
        if (index0[i] < 0 || index0[i] >= ncells) return 3;
    double* dens = (double*)malloc(sizeof(double) * (size_t)(nmnn > 0 ? nmnn : 1));
    if (!dens) return 2;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    /* density of every MNN cell among the MNN cells */
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t i = 0; i < nmnn; ++i) {
        const double* mi = mat + (int64_t)index0[i] * Gdist;
        double acc = 0.0;
        for (int64_t j = 0; j < nmnn; ++j) {
            const double* mj = mat + (int64_t)index0[j] * Gdist;
            double s = 0.0;
            for (int64_t g = 0; g < Gdist; ++g) { const double t = mi[g] - mj[g]; s += t * t; }
            s /= -sigma2;
            acc = (j == 0) ? s : lse2(acc, s);
        }
        dens[i] = acc;
    }
#pragma omp parallel num_threads(nthreads)
    {
        double* lm = (double*)malloc(sizeof(double) * (size_t)(nmnn > 0 ? nmnn : 1));
#pragma omp for schedule(static)
        for (int64_t c = 0; c < ncells; ++c) {
            const double* mc = mat + c * Gdist;
            double* oc = out + c * G;
            for (int64_t g = 0; g < G; ++g) oc[g] = (nmnn == 0) ? NAN : 0.0; /* no MNN cell: 0 * exp(-Inf - NA) = NaN (:105-115) */
            if (nmnn == 0) continue;
            double total = 0.0, top = -INFINITY;
            for (int64_t i = 0; i < nmnn; ++i) {
                const double* mi = mat + (int64_t)index0[i] * Gdist;
                double s = 0.0;
                for (int64_t g = 0; g < Gdist; ++g) { const double t = mi[g] - mc[g]; s += t * t; }
                s /= -sigma2;
                lm[i] = s - dens[i];
                total = (i == 0) ? lm[i] : lse2(total, lm[i]);
                if (lm[i] > top) top = lm[i];
            }
            for (int64_t i = 0; i < nmnn; ++i) {
                const double w = exp(lm[i] - top);
                const double* av = averaged + i * G;
                for (int64_t g = 0; g < G; ++g) oc[g] += w * av[g];
            }
            const double sc = exp(top - total);
            for (int64_t g = 0; g < G; ++g) oc[g] *= sc;
        }
        free(lm);
    }
    free(dens);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a7. shift-variance adjustment, src/adjust_shift_variance.cpp:9-164.
 * data1 [G x n1], data2 [G x n2] (cell-contiguous), vect [n2 x G] column-major (row c strided, :57),
 * restricts 0-based.  out[n2].  Returns 1 / 4 on the dimension checks (:33-41), 3 on the restrict
 * check (src/utils.cpp:6-13).  Follows the reference's operation order so the discrete quantile
 * pick (:145-156) agrees: same projections (inner products in gene order), same distance-to-line
 * recipe (:9-27), same (projection, logweight) lexicographic sort (:133).
 * ------------------------------------------------------------------------------------------ */
typedef struct { double proj, lw; } pl_t;
static int cmp_pl(const void* a, const void* b) {
    const pl_t* x = (const pl_t*)a; const pl_t* y = (const pl_t*)b;
    if (x->proj < y->proj) return -1;
    if (x->proj > y->proj) return 1;
    if (x->lw < y->lw) return -1;
    if (x->lw > y->lw) return 1;
    return 0;
}

static double sqdist_to_line(const double* ref, const double* grad, const double* pt, double* work, int64_t G) {
    for (int64_t g = 0; g < G; ++g) work[g] = ref[g] - pt[g];
    double scale = 0.0;
    for (int64_t g = 0; g < G; ++g) scale += work[g] * grad[g];
    double dist = 0.0;
    for (int64_t g = 0; g < G; ++g) { work[g] -= scale * grad[g]; dist += work[g] * work[g]; }
    return dist;
}

/* `cells` (may be NULL = every cell of batch 2): the cells whose scaling is wanted; out[i] belongs to cells[i].  The
 * subset form exists so that tests can check sampled cells of a problem whose full O(n2 (n1 + n2) G) loop would take
 * days on a CPU; the arithmetic per cell is the same loop. */
int oracle_adjust_shift_variance_cells(const double* data1, int64_t G1, int64_t n1, const double* data2, int64_t G2, int64_t n2,
                                       const double* vect, int64_t vr, int64_t vc, double sigma2,
                                       const int32_t* r1, int64_t nr1, const int32_t* r2, int64_t nr2,
                                       const int64_t* cells, int64_t ncells_sub, int vect_rows_are_cells, double* out, int nthreads) {
    const int64_t G = G1;
    if (G != G2 || G != vc) return 1;
    if (n2 != vr) return 4;
    for (int64_t i = 0; i < nr1; ++i) if (r1[i] == INT32_MIN || r1[i] < 0 || r1[i] >= n1) return 3;
    for (int64_t i = 0; i < nr2; ++i) if (r2[i] == INT32_MIN || r2[i] < 0 || r2[i] >= n2) return 3;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        double* work = (double*)malloc(sizeof(double) * (size_t)(G > 0 ? G : 1));
        double* grad = (double*)malloc(sizeof(double) * (size_t)(G > 0 ? G : 1));
        pl_t* d1 = (pl_t*)malloc(sizeof(pl_t) * (size_t)(nr1 > 0 ? nr1 : 1));
        const int64_t ncount = cells ? ncells_sub : n2;
#pragma omp for schedule(dynamic, 1)
        for (int64_t ci = 0; ci < ncount; ++ci) {
            const int64_t c = cells ? cells[ci] : ci;
            const double* cur = data2 + c * G;
            double l2 = 0.0;
            /* vect is [n2 x G] column-major as in R, or (subset form) one contiguous G-vector per requested cell */
            for (int64_t g = 0; g < G; ++g) { grad[g] = vect_rows_are_cells ? vect[ci * G + g] : vect[c + g * n2]; l2 += grad[g] * grad[g]; }
            l2 = sqrt(l2);
            if (l2 != 0.0) for (int64_t g = 0; g < G; ++g) grad[g] /= l2;
            double curproj = 0.0;
            for (int64_t g = 0; g < G; ++g) curproj += grad[g] * cur[g];

            /* own batch: cumulative probability of this cell along the line */
            double prob2 = 0.0, tot2 = 0.0;
            int have_p = 0, have_t = 0;
            for (int64_t s = 0; s < nr2; ++s) {
                const int64_t same = r2[s];
                int add = 1;
                double lp = 0.0;
                if (same != c) {
                    const double* sc = data2 + same * G;
                    double sp = 0.0;
                    for (int64_t g = 0; g < G; ++g) sp += grad[g] * sc[g];
                    const double sd = sqdist_to_line(cur, grad, sc, work, G);
                    lp = -sd / sigma2;
                    if (sp > curproj) add = 0;
                }
                if (add) { prob2 = have_p ? lse2(prob2, lp) : lp; have_p = 1; }
                tot2 = have_t ? lse2(tot2, lp) : lp; have_t = 1;
            }
            prob2 -= tot2;

            /* reference batch: weighted quantile along the same line */
            double tot1 = 0.0;
            for (int64_t o = 0; o < nr1; ++o) {
                const double* oc = data1 + (int64_t)r1[o] * G;
                double p = 0.0;
                for (int64_t g = 0; g < G; ++g) p += grad[g] * oc[g];
                const double od = sqdist_to_line(cur, grad, oc, work, G);
                d1[o].proj = p; d1[o].lw = -od / sigma2;
                tot1 = (o == 0) ? d1[o].lw : lse2(tot1, d1[o].lw);
            }
            qsort(d1, (size_t)nr1, sizeof(pl_t), cmp_pl);
            double refq = NAN;
            if (nr1 > 0) {
                const double target = prob2 + tot1;
                double cum = 0.0;
                refq = d1[nr1 - 1].proj;
                for (int64_t o = 0; o < nr1; ++o) {
                    cum = (o == 0) ? d1[o].lw : lse2(cum, d1[o].lw);
                    if (cum >= target) { refq = d1[o].proj; break; }
                }
            }
            out[cells ? ci : c] = (refq - curproj) / l2;
        }
        free(work); free(grad); free(d1);
    }
    return 0;
}

int oracle_adjust_shift_variance(const double* data1, int64_t G1, int64_t n1, const double* data2, int64_t G2, int64_t n2,
                                 const double* vect, int64_t vr, int64_t vc, double sigma2,
                                 const int32_t* r1, int64_t nr1, const int32_t* r2, int64_t nr2, double* out,
                                 int nthreads) {
    return oracle_adjust_shift_variance_cells(data1, G1, n1, data2, G2, n2, vect, vr, vc, sigma2, r1, nr1, r2, nr2, NULL, 0, 0, out, nthreads);
}

/* a9. cosine normalisation, R/cosineNorm.R:53-82: l2 = sqrt(colSums(x^2)); x / max(1e-8, l2).
 * x [G x n] column-major (cells are columns); l2_out may be NULL; out may be NULL. */
int oracle_cosine_norm(const double* x, int64_t G, int64_t n, double* out, double* l2_out) {
    for (int64_t c = 0; c < n; ++c) {
        double s = 0.0;
        for (int64_t g = 0; g < G; ++g) s += x[c * G + g] * x[c * G + g];
        const double l2 = sqrt(s);
        if (l2_out) l2_out[c] = l2;
        const double dv = l2 > 1e-8 ? l2 : 1e-8;
        if (out) for (int64_t g = 0; g < G; ++g) out[c * G + g] = x[c * G + g] / dv;
    }
    return 0;
}
