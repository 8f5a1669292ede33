// TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
//
// CPU port of the exact KMKNN search (Wang, 2012) that the reference reaches through
// BiocNeighbors::queryKNN(..., BNPARAM=KmknnParam(), BPPARAM=MulticoreParam(...)) at
// R/MNN_tree.R:129 and R/fastMNN.R:605.  BiocNeighbors itself is NOT in /root/reference and
// not in this image (no R), so this is a restatement of the published algorithm:
//   build : k-means with ceil(sqrt(n)) centres; per cluster, members sorted by distance to the centre;
//   query : visit clusters by increasing centre distance; inside a cluster only members whose centre
//           distance lies in [dc - thr, dc + thr] (triangle inequality) get an exact distance;
//           keep the k best under the total order (squared distance, index).
// Exact squared distances are accumulated in double in dimension order, exactly as
// oracle_query_knn does, so both give identical (index, distance) results; the pruning window is
// widened by a relative 1e-9 so that rounding in sqrt never drops a boundary candidate.
// Two uses: (1) a faster exact checker for parity tests at sizes where brute force is too slow,
// (2) bench.py's cpu_baseline / --impl reference leg ("kind": "port"), query-parallel over all host
// threads the way MulticoreParam splits queries across workers.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <thread>
#include <utility>
#include <vector>

namespace {

struct Index {
    int64_t n = 0;
    int d = 0;
    int nc = 0;
    std::vector<double> pts;      // [n x d] row-major, reordered cluster by cluster
    std::vector<int64_t> orig;    // original row of each reordered point
    std::vector<double> centres;  // [nc x d]
    std::vector<int64_t> start;   // [nc + 1] offsets into pts
    std::vector<double> cdist;    // distance of each reordered point to its centre (ascending inside a cluster)
};

inline double sqdist(const double* a, const double* b, int d) {
    double s = 0.0;
    for (int t = 0; t < d; ++t) { const double df = a[t] - b[t]; s += df * df; }
    return s;
}

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
    if (nthreads <= 1 || n < 2) { f(0, n, 0); return; }
    std::atomic<int64_t> next(0);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(256, n / (nthreads * 8) + 1));
    std::vector<std::thread> th;
    for (int w = 0; w < nthreads; ++w)
        th.emplace_back([&, w]() {
            for (;;) {
                const int64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                f(b, std::min(n, b + chunk), w);
            }
        });
    for (auto& t : th) t.join();
}

void assign(const double* X, int64_t n, int d, const std::vector<double>& C, int nc, std::vector<int32_t>& lab, int nthreads) {
    parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
        for (int64_t i = b; i < e; ++i) {
            double best = std::numeric_limits<double>::infinity();
            int32_t arg = 0;
            for (int c = 0; c < nc; ++c) {
                const double s = sqdist(X + i * d, C.data() + (int64_t)c * d, d);
                if (s < best) { best = s; arg = c; }
            }
            lab[i] = arg;
        }
    });
}

}  // namespace

extern "C" {

int kmknn_hw_threads() {
    const unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

// X row-major [n x d].  Lloyd iterations run on a subsample (the clustering only steers pruning; results are exact
// for any clustering), then every point is assigned once.
void* kmknn_build(const double* X, int64_t n, int d, int nthreads, uint64_t seed) {
    Index* ix = new Index();
    ix->n = n; ix->d = d;
    if (nthreads <= 0) nthreads = kmknn_hw_threads();
    const int nc = (int)std::max<int64_t>(1, (int64_t)std::ceil(std::sqrt((double)std::max<int64_t>(n, 1))));
    ix->nc = nc;
    std::mt19937_64 rng(seed);
    // subsample for Lloyd
    const int64_t ns = std::min<int64_t>(n, std::max<int64_t>(20 * (int64_t)nc, 20000));
    std::vector<int64_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    for (int64_t i = 0; i < ns && n > 1; ++i) { std::uniform_int_distribution<int64_t> u(i, n - 1); std::swap(perm[i], perm[u(rng)]); }
    std::vector<double> S((size_t)ns * d);
    for (int64_t i = 0; i < ns; ++i) std::memcpy(&S[i * d], X + perm[i] * d, sizeof(double) * d);
    ix->centres.assign((size_t)nc * d, 0.0);
    for (int c = 0; c < nc; ++c) std::memcpy(&ix->centres[(size_t)c * d], &S[(size_t)(c % ns) * d], sizeof(double) * d);
    std::vector<int32_t> lab(ns);
    for (int it = 0; it < 6 && n > 0; ++it) {
        assign(S.data(), ns, d, ix->centres, nc, lab, nthreads);
        std::vector<double> sum((size_t)nc * d, 0.0);
        std::vector<int64_t> cnt(nc, 0);
        for (int64_t i = 0; i < ns; ++i) { cnt[lab[i]]++; for (int t = 0; t < d; ++t) sum[(size_t)lab[i] * d + t] += S[i * d + t]; }
        for (int c = 0; c < nc; ++c)
            if (cnt[c]) for (int t = 0; t < d; ++t) ix->centres[(size_t)c * d + t] = sum[(size_t)c * d + t] / (double)cnt[c];
    }
    std::vector<int32_t> full(n);
    assign(X, n, d, ix->centres, nc, full, nthreads);
    std::vector<int64_t> cnt(nc, 0);
    for (int64_t i = 0; i < n; ++i) cnt[full[i]]++;
    ix->start.assign(nc + 1, 0);
    for (int c = 0; c < nc; ++c) ix->start[c + 1] = ix->start[c] + cnt[c];
    std::vector<std::pair<double, int64_t>> keyed(n);
    std::vector<int64_t> fill(ix->start.begin(), ix->start.end() - 1);
    for (int64_t i = 0; i < n; ++i) {
        const int c = full[i];
        keyed[fill[c]++] = std::make_pair(std::sqrt(sqdist(X + i * d, &ix->centres[(size_t)c * d], d)), i);
    }
    for (int c = 0; c < nc; ++c) std::sort(keyed.begin() + ix->start[c], keyed.begin() + ix->start[c + 1]);
    ix->pts.resize((size_t)n * d);
    ix->orig.resize(n);
    ix->cdist.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        ix->cdist[i] = keyed[i].first;
        ix->orig[i] = keyed[i].second;
        std::memcpy(&ix->pts[(size_t)i * d], X + keyed[i].second * d, sizeof(double) * d);
    }
    return ix;
}

void kmknn_free(void* h) { delete static_cast<Index*>(h); }

// Q row-major [nq x d]; idx_out row-major [nq x k] 0-BASED; dist_out row-major [nq x k] (may be NULL).
int kmknn_query(void* h, const double* Q, int64_t nq, int k, int32_t* idx_out, double* dist_out, int nthreads) {
    const Index& ix = *static_cast<Index*>(h);
    if (k > ix.n || k <= 0) return 1;
    if (nthreads <= 0) nthreads = kmknn_hw_threads();
    const int d = ix.d, nc = ix.nc;
    parallel_for(nq, nthreads, [&](int64_t b, int64_t e, int) {
        typedef std::pair<double, int64_t> cand;  // (squared distance, original index); max-heap on the pair
        std::vector<cand> heap;
        heap.reserve(k + 1);
        std::vector<std::pair<double, int>> order(nc);
        for (int64_t q = b; q < e; ++q) {
            const double* qv = Q + q * d;
            for (int c = 0; c < nc; ++c) order[c] = std::make_pair(std::sqrt(sqdist(qv, &ix.centres[(size_t)c * d], d)), c);
            std::sort(order.begin(), order.end());
            heap.clear();
            double thr = std::numeric_limits<double>::infinity();  // un-squared distance of the current k-th best
            for (int oc = 0; oc < nc; ++oc) {
                const int c = order[oc].second;
                const double dc = order[oc].first;
                const int64_t s = ix.start[c], t = ix.start[c + 1];
                if (s == t) continue;
                int64_t lo = s, hi = t;
                if (heap.size() == (size_t)k) {
                    const double slack = 1e-9 * (dc + thr) + 1e-300;
                    const double lower = dc - thr - slack, upper = dc + thr + slack;
                    if (ix.cdist[t - 1] < lower) continue;
                    lo = std::lower_bound(ix.cdist.begin() + s, ix.cdist.begin() + t, lower) - ix.cdist.begin();
                    hi = std::upper_bound(ix.cdist.begin() + s, ix.cdist.begin() + t, upper) - ix.cdist.begin();
                }
                for (int64_t p = lo; p < hi; ++p) {
                    const cand cur(sqdist(qv, &ix.pts[(size_t)p * d], d), ix.orig[p]);
                    if (heap.size() < (size_t)k) {
                        heap.push_back(cur);
                        std::push_heap(heap.begin(), heap.end());
                        if (heap.size() == (size_t)k) thr = std::sqrt(heap.front().first);
                    } else if (cur < heap.front()) {
                        std::pop_heap(heap.begin(), heap.end());
                        heap.back() = cur;
                        std::push_heap(heap.begin(), heap.end());
                        thr = std::sqrt(heap.front().first);
                    }
                }
            }
            std::sort_heap(heap.begin(), heap.end());
            for (int r = 0; r < k; ++r) {
                idx_out[q * k + r] = (int32_t)heap[r].second;
                if (dist_out) dist_out[q * k + r] = std::sqrt(heap[r].first);
            }
        }
    });
    return 0;
}

}  // extern "C"
