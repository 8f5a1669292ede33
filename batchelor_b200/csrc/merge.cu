// Device-resident merge loop of reducedMNN / fastMNN -- .fast_mnn_core (R/fastMNN.R:436-562) as ONE C-ABI call
// (SURVEY.md section 8f, N1): every batch is uploaded once, the nodes of the merge tree live in HBM across merges, and the
// host receives the corrected matrix, the MNN pairs of every merge and the diagnostics (batch.size, skipped, lost.var).
// Per merge (all steps are kernels of this library):
//   .compute_perbatch_var (:467-468, :651-658)  -> segment_var
//   .orthogonalize_other  (:473-474, :642-647)  -> correct::center_along_batch_vector_device per earlier batch vector
//   .restricted_mnn       (:476, R/MNN_tree.R:113-133) -> gather restricted rows, two exact searches, mutual pairs
//   .average_correction   (:480, :505)          -> correct::average_correction_device
//   .get_batch_magnitude  (:484, :582-595)      -> col_mean / col_mean_sq
//   .center_along_batch_vector (:496-497)       -> both sides
//   .tricube_weighted_correction (:506-507, :599-608) -> exact search of all right cells among the MNN cells + tricube
// The merge ORDER stays host control flow (a10: the R side walks its MNN_treenode tree, R/MNN_tree.R:61-109) and arrives
// as a list of (left node, right node) ids: leaves are 0..nb-1, the node created by merge m is nb + m.
//
// Multi-GPU behind the boundary: an R process cannot be a torch.distributed rank, so this entry drives every visible
// device itself (B200MNN_DEVICES caps it).  The two searches of a merge and the tricube search shard their QUERY rows
// over the devices; the reference rows are replicated with peer copies over NVLink and the per-shard index blocks come
// back the same way (north_star's split: queries sharded, reference replicated, top-k gathered).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "gemm_tc.cuh"
#include "internal.cuh"

namespace b200 {
namespace merge {

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const double* __restrict__ src, int d, const int32_t* __restrict__ rows, int64_t nrows, double* __restrict__ dst) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nrows * d) return;
    const int64_t i = e / d;
    dst[e] = src[(int64_t)rows[i] * d + (e - i * d)];
}
__global__ void map_ids_kernel(int32_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ map) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = map[ids[i]];
}
__global__ void add_ids_kernel(const int32_t* __restrict__ in, int64_t n, int32_t off, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + off;
}
__global__ void iota_kernel(int32_t* __restrict__ out, int64_t n, int32_t off) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)i + off;
}
// Column moments in a FIXED summation order (no floating-point atomics: results are bit-identical run to run and for any
// number of devices): partial[slab][t] = sum over a slab of 1 024 rows of f(X[r, t]), f = identity (SQ = 0), square
// (SQ = 1) or squared deviation from mean[t] (SQ = 2); then out[t] = inv * sum over the slabs in slab order.
template <int SQ>
__global__ void __launch_bounds__(256)
col_moment_kernel(const double* __restrict__ X, int64_t rows, int d, const double* __restrict__ mean, double* __restrict__ partial) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= d) return;
    const int64_t r0 = (int64_t)blockIdx.y * 1024, r1 = min(rows, r0 + 1024);
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const double v = X[r * d + t];
        if (SQ == 0) s += v;
        else if (SQ == 1) s += v * v;
        else { const double dv = v - mean[t]; s += dv * dv; }
    }
    partial[(int64_t)blockIdx.y * d + t] = s;
}
__global__ void col_moment_final_kernel(const double* __restrict__ partial, int64_t nslabs, int d, double inv, double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d) return;
    double s = 0.0;
    for (int64_t b = 0; b < nslabs; ++b) s += partial[b * d + t];
    out[t] = s * inv;
}
__global__ void sum_vec_kernel(const double* __restrict__ v, int d, double* __restrict__ out) {
    double s = 0.0;
    for (int t = 0; t < d; ++t) s += v[t];
    *out = s;
}
__global__ void sumsq_vec_kernel(const double* __restrict__ v, int d, double* __restrict__ out) {
    double s = 0.0;
    for (int t = 0; t < d; ++t) s += v[t] * v[t];
    *out = s;
}

// ------------------------------------------------------------------------------------------------
// host-side state
// ------------------------------------------------------------------------------------------------
struct DevBuf {   // stream-ordered device allocation, freed explicitly or at destruction
    void* p = nullptr;
    cudaStream_t s = nullptr;
    int dev = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    int alloc(size_t bytes, cudaStream_t stream) {
        release();
        s = stream;
        cudaGetDevice(&dev);
        B200_CUDA(cudaMallocAsync(&p, bytes ? bytes : 16, stream));
        return 0;
    }
    void release() {
        if (p) {
            int cur = 0;
            cudaGetDevice(&cur);
            if (cur != dev) cudaSetDevice(dev);
            cudaFreeAsync(p, s);
            if (cur != dev) cudaSetDevice(cur);
            p = nullptr;
        }
    }
    ~DevBuf() { release(); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct Node {
    DevBuf data;                   // [n x d] row-major, when the node owns its rows ...
    double* ptr = nullptr;         // ... or a window of the arena (rows()) -- children of a merge are adjacent there
    double* rows() const { return ptr ? ptr : data.as<double>(); }
    int64_t n = 0;
    DevBuf restrict_rows;          // int32 0-based rows, or empty
    int64_t nres = -1;             // -1: no restriction
    std::vector<int> index;        // batch ids (0-based) in row order
    std::vector<int64_t> seg;      // rows per batch of `index`
    std::vector<DevBuf*> extras;   // batch vectors applied inside this node (owned by the result's pool)
    bool alive = false;
};

struct MergeResult {
    int nb = 0, d = 0;
    int64_t ntotal = 0;
    int device0 = 0;
    cudaStream_t stream = nullptr;
    std::deque<Node> nodes;                        // 2 nb - 1 (deque: nodes are neither copied nor moved)
    DevBuf arena;                                  // [ntotal x d]: leaves in the left-to-right leaf order of the merge tree
    std::vector<DevBuf*> vec_pool;
    std::vector<std::vector<int32_t>> pl, pr;      // pairs per merge, 1-based rows within the left / right node
    std::vector<double> batch_size, lost_var;
    std::vector<int32_t> skipped;
    std::vector<int32_t> merged_left, merged_right;   // node ids of every merge (given, or chosen by the auto-merge search)
    int final_node = -1;
    ~MergeResult() {
        if (stream) { cudaSetDevice(device0); cudaStreamSynchronize(stream); }
        nodes.clear();
        arena.release();
        for (auto* v : vec_pool) delete v;
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    }
};

static int choose_k(int k, double prop_k, int64_t N) {   // R/MNN_tree.R:140-146
    if (!(prop_k >= 0.0)) return k;
    const int64_t byprop = (int64_t)std::nearbyint(prop_k * (double)N);   // R's round(): half to even
    return (int)std::min<int64_t>(N, std::max<int64_t>(k, byprop));
}

// --- multi-device exact search: queries sharded, references replicated over NVLink peer copies ---
struct DeviceSet {
    std::vector<int> devs;
    std::vector<cudaStream_t> streams;
    int init(int primary) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) n = 1;
        int want = n;
        if (const char* e = getenv("B200MNN_DEVICES")) want = std::max(1, std::min(n, atoi(e)));
        devs.push_back(primary);
        for (int g = 0; g < n && (int)devs.size() < want; ++g) {
            if (g == primary) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, primary, g) != cudaSuccess || !can) continue;
            devs.push_back(g);
        }
        for (size_t i = 0; i < devs.size(); ++i) {
            B200_CUDA(cudaSetDevice(devs[i]));
            for (size_t j = 0; j < devs.size(); ++j)
                if (i != j) { cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0); if (e != cudaSuccess) cudaGetLastError(); }
            cudaStream_t s;
            B200_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            streams.push_back(s);
            cudaMemPool_t pool;   // keep freed scratch cached on the helper devices too
            if (cudaDeviceGetDefaultMemPool(&pool, devs[i]) == cudaSuccess) {
                uint64_t thr = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
        B200_CUDA(cudaSetDevice(primary));
        return 0;
    }
    void destroy() {
        for (size_t i = 0; i < devs.size(); ++i) { cudaSetDevice(devs[i]); cudaStreamSynchronize(streams[i]); cudaStreamDestroy(streams[i]); }
        if (!devs.empty()) cudaSetDevice(devs[0]);
        devs.clear(); streams.clear();
    }
};

// X [n x d], Q [nq x d] on the primary device (ready on `stream`); results on the primary device, ready on `stream`.
// A helper device is worth its fixed cost (replica of X over NVLink, its own cluster plan, ~150 launches) only for a large
// block of queries: the search uses min(G, nq / 131072) devices (B200MNN_SHARD_MIN overrides the block size), each
// helper driven by its own host thread so that the launch streams of the devices fill concurrently.
static int sharded_query_knn(DeviceSet& ds, const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                             cudaStream_t stream) {
    int64_t block = 131072;
    if (const char* e = getenv("B200MNN_SHARD_MIN")) block = std::max<int64_t>(1024, atoll(e));
    const int G = (int)std::min<int64_t>((int64_t)ds.devs.size(), std::max<int64_t>(1, nq / block));
    if (G <= 1) return knn::query_knn_device(dX, n, dQ, nq, d, k, d_idx, d_dist, nullptr, stream, nullptr);
    const int primary = ds.devs[0];
    cudaEvent_t ready;
    B200_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    B200_CUDA(cudaEventRecord(ready, stream));
    const int64_t per = ceil_div(nq, G);
    std::vector<cudaEvent_t> done(G, nullptr);
    std::vector<int> rcs(G, 0);
    std::vector<std::string> errs(G);
    std::vector<std::thread> workers;
    for (int g = 1; g < G; ++g) {
        const int64_t lo = std::min(nq, g * per), hi = std::min(nq, (g + 1) * per);
        if (hi <= lo) continue;
        workers.emplace_back([&, g, lo, hi]() {
            const int dev = ds.devs[g];
            cudaStream_t s = ds.streams[g];
            auto body = [&]() -> int {
                B200_CUDA(cudaSetDevice(dev));
                B200_CUDA(cudaStreamWaitEvent(s, ready, 0));
                double *x = nullptr, *q = nullptr, *dd = nullptr;
                int32_t* ii = nullptr;
                B200_CUDA(cudaMallocAsync(&x, sizeof(double) * n * d, s));
                B200_CUDA(cudaMallocAsync(&q, sizeof(double) * (hi - lo) * d, s));
                B200_CUDA(cudaMallocAsync(&ii, sizeof(int32_t) * (hi - lo) * k, s));
                if (d_dist) B200_CUDA(cudaMallocAsync(&dd, sizeof(double) * (hi - lo) * k, s));
                B200_CUDA(cudaMemcpyPeerAsync(x, dev, dX, primary, sizeof(double) * n * d, s));
                B200_CUDA(cudaMemcpyPeerAsync(q, dev, dQ + lo * d, primary, sizeof(double) * (hi - lo) * d, s));
                B200_TRY(knn::query_knn_device(x, n, q, hi - lo, d, k, ii, dd, nullptr, s, nullptr));
                B200_CUDA(cudaMemcpyPeerAsync(d_idx + lo * k, primary, ii, dev, sizeof(int32_t) * (hi - lo) * k, s));
                if (d_dist) B200_CUDA(cudaMemcpyPeerAsync(d_dist + lo * k, primary, dd, dev, sizeof(double) * (hi - lo) * k, s));
                cudaFreeAsync(x, s); cudaFreeAsync(q, s); cudaFreeAsync(ii, s);
                if (dd) cudaFreeAsync(dd, s);
                B200_CUDA(cudaEventCreateWithFlags(&done[g], cudaEventDisableTiming));
                B200_CUDA(cudaEventRecord(done[g], s));
                return 0;
            };
            rcs[g] = body();
            if (rcs[g]) errs[g] = b200mnn_last_error();   // the message is thread-local: carry it to the caller's thread
        });
    }
    int rc = knn::query_knn_device(dX, n, dQ, std::min(nq, per), d, k, d_idx, d_dist, nullptr, stream, nullptr);
    for (auto& w : workers) w.join();
    cudaSetDevice(primary);
    for (int g = 1; g < G; ++g) {
        if (done[g]) { cudaStreamWaitEvent(stream, done[g], 0); cudaEventDestroy(done[g]); }
        if (!rc && rcs[g]) rc = fail(rcs[g], errs[g]);
    }
    cudaEventDestroy(ready);
    return rc;
}

template <int SQ>
static int col_moment(const double* X, int64_t rows, int d, const double* mean, double inv, double* out, cudaStream_t s) {
    const int64_t nslabs = ceil_div(rows, 1024);
    DevBuf partial;
    B200_TRY(partial.alloc(sizeof(double) * nslabs * d, s));
    dim3 grid((unsigned)ceil_div(d, 256), (unsigned)nslabs);
    col_moment_kernel<SQ><<<grid, 256, 0, s>>>(X, rows, d, mean, partial.as<double>());
    B200_LAUNCH_CHECK();
    col_moment_final_kernel<<<(unsigned)ceil_div(d, 128), 128, 0, s>>>(partial.as<double>(), nslabs, d, inv, out);
    B200_LAUNCH_CHECK();
    return 0;
}

static int total_var(const double* X, int64_t rows, int d, double* scratch /* [2 d + 1] device */, double* host_out, cudaStream_t s) {
    // .compute_perbatch_var: sum over dimensions of the unbiased column variance (NaN for a single row, like R's var)
    if (rows < 2) { *host_out = NAN; return 0; }
    double* mean = scratch;
    double* var = scratch + d;
    double* tot = scratch + 2 * d;
    B200_TRY(col_moment<0>(X, rows, d, nullptr, 1.0 / (double)rows, mean, s));
    B200_TRY(col_moment<2>(X, rows, d, mean, 1.0 / (double)(rows - 1), var, s));
    sum_vec_kernel<<<1, 1, 0, s>>>(var, d, tot);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(host_out, tot, sizeof(double), cudaMemcpyDeviceToHost, s));
    return 0;
}

static int perbatch_var(const Node& nd, int d, double* scratch, std::vector<double>& out, cudaStream_t s) {
    out.assign(nd.index.size(), 0.0);
    int64_t r0 = 0;
    for (size_t i = 0; i < nd.index.size(); ++i) {
        B200_TRY(total_var(nd.rows() + r0 * d, nd.seg[i], d, scratch, &out[i], s));
        B200_CUDA(cudaStreamSynchronize(s));   // `out[i]` is pageable host memory
        r0 += nd.seg[i];
    }
    return 0;
}

}  // namespace merge
}  // namespace b200

using b200::merge::MergeResult;

extern "C" {

struct b200mnn_merge_result { MergeResult r; };

int b200mnn_reduced_mnn(const double* const* batches, const int64_t* ncells, int nb, int d, int col_major, const int32_t* merge_left,
                        const int32_t* merge_right, int k, double prop_k, double ndist, double min_batch_skip,
                        const int32_t* const* restrict1, const int64_t* nrestrict, int get_variance, b200mnn_merge_result** result_out) {
    using namespace b200;
    using namespace b200::merge;
    B200_TRY(ensure_device());
    if (!result_out) return fail(B200MNN_EINVAL, "result_out is NULL");
    *result_out = nullptr;
    if (nb < 1 || d < 1) return fail(B200MNN_EINVAL, "at least one batch and one dimension are needed");
    for (int b = 0; b < nb; ++b)
        if (ncells[b] < 1) return fail(B200MNN_EINVAL, "every batch needs at least one cell");
    b200mnn_merge_result* holder = new b200mnn_merge_result();
    MergeResult& R = holder->r;
    struct Guard { b200mnn_merge_result* h; ~Guard() { delete h; } } guard{holder};
    R.nb = nb; R.d = d;
    B200_CUDA(cudaGetDevice(&R.device0));
    B200_CUDA(cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking));
    cudaStream_t s = R.stream;
    DeviceSet ds;
    B200_TRY(ds.init(R.device0));
    struct DsGuard { DeviceSet* d; ~DsGuard() { d->destroy(); } } dsg{&ds};
    R.nodes.resize((size_t)2 * nb - 1);
    const int nm = nb - 1;
    R.pl.resize(nm); R.pr.resize(nm);
    R.batch_size.assign(nm, NAN);
    R.skipped.assign(nm, 0);
    R.lost_var.assign((size_t)nm * nb, 0.0);
    int* bad = nullptr;
    DevBuf badbuf, scratch;
    B200_TRY(badbuf.alloc(sizeof(int), s));
    bad = badbuf.as<int>();
    B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    B200_TRY(scratch.alloc(sizeof(double) * (4 * (size_t)d + 8), s));

    // With a given merge order every merge joins two nodes that are neighbours in the left-to-right leaf order of the
    // tree, so the leaves are uploaded into ONE arena in that order and rbind(left, right) (:520-525) costs nothing.
    const bool given_order = (merge_left != nullptr && merge_right != nullptr);
    std::vector<int64_t> leaf_offset(nb, -1);
    for (int b = 0; b < nb; ++b) R.ntotal += ncells[b];
    if (given_order && nm > 0) {
        std::vector<std::vector<int>> leaves((size_t)2 * nb - 1);
        for (int b = 0; b < nb; ++b) leaves[b] = {b};
        bool ok = true;
        for (int m = 0; m < nm && ok; ++m) {
            const int li = merge_left[m], ri = merge_right[m];
            ok = li >= 0 && ri >= 0 && li < nb + m && ri < nb + m && li != ri && !leaves[li].empty() && !leaves[ri].empty();
            if (!ok) break;
            leaves[nb + m] = leaves[li];
            leaves[nb + m].insert(leaves[nb + m].end(), leaves[ri].begin(), leaves[ri].end());
            leaves[li].clear(); leaves[ri].clear();
        }
        if (ok && (int)leaves[(size_t)2 * nb - 2].size() == nb) {
            int64_t off = 0;
            for (int b : leaves[(size_t)2 * nb - 2]) { leaf_offset[b] = off; off += ncells[b]; }
            B200_TRY(R.arena.alloc(sizeof(double) * R.ntotal * d, s));
        }
    }
    // leaves: upload (transposing R's column-major [cells x d] on device)
    for (int b = 0; b < nb; ++b) {
        Node& nd = R.nodes[b];
        nd.n = ncells[b];
        if (R.arena.p && leaf_offset[b] >= 0) nd.ptr = R.arena.as<double>() + leaf_offset[b] * d;
        else B200_TRY(nd.data.alloc(sizeof(double) * nd.n * d, s));
        if (col_major) {
            DevBuf raw;
            B200_TRY(raw.alloc(sizeof(double) * nd.n * d, s));
            B200_CUDA(cudaMemcpyAsync(raw.p, batches[b], sizeof(double) * nd.n * d, cudaMemcpyHostToDevice, s));
            B200_TRY(correct::transpose_device<double>(raw.as<double>(), nd.n, d, nd.rows(), s));   // column-major [n x d] -> row-major
        } else {
            B200_CUDA(cudaMemcpyAsync(nd.rows(), batches[b], sizeof(double) * nd.n * d, cudaMemcpyHostToDevice, s));
        }
        if (restrict1 && restrict1[b]) {
            const int64_t nr = nrestrict[b];
            if (nr < 1) return fail(B200MNN_EINVAL, "no cells remaining in a batch after restriction");
            std::vector<int32_t> rows((size_t)nr);
            for (int64_t i = 0; i < nr; ++i) {
                const int32_t v = restrict1[b][i];
                if (v < 1 || v > nd.n) return fail(B200MNN_EINVAL, "subset indices out of range");
                rows[i] = v - 1;
            }
            B200_TRY(nd.restrict_rows.alloc(sizeof(int32_t) * nr, s));
            B200_CUDA(cudaMemcpyAsync(nd.restrict_rows.p, rows.data(), sizeof(int32_t) * nr, cudaMemcpyHostToDevice, s));
            B200_CUDA(cudaStreamSynchronize(s));
            nd.nres = nr;
        }
        nd.index = {b};
        nd.seg = {nd.n};
        nd.alive = true;
    }

    // .restricted_mnn (R/MNN_tree.R:113-133): mutual pairs between the (restricted) rows of two nodes; ids are mapped back
    // to node rows.  first / second may be NULL when only the count is wanted (auto-merge search).
    auto find_pairs = [&](const double* ld, const Node& L, const double* rd, const Node& Rt, DevBuf* first_out, DevBuf* second_out,
                          int64_t* np_out) -> int {
        const int64_t nL = L.nres >= 0 ? L.nres : L.n, nR = Rt.nres >= 0 ? Rt.nres : Rt.n;
        DevBuf lsub, rsub, w21, w12, dnp, f_local, s_local;
        DevBuf& first = first_out ? *first_out : f_local;
        DevBuf& second = second_out ? *second_out : s_local;
        const double* lq = ld;
        const double* rq = rd;
        if (L.nres >= 0) {
            B200_TRY(lsub.alloc(sizeof(double) * nL * d, s));
            gather_rows_kernel<<<(unsigned)ceil_div(nL * d, 256), 256, 0, s>>>(ld, d, L.restrict_rows.as<int32_t>(), nL, lsub.as<double>());
            B200_LAUNCH_CHECK();
            lq = lsub.as<double>();
        }
        if (Rt.nres >= 0) {
            B200_TRY(rsub.alloc(sizeof(double) * nR * d, s));
            gather_rows_kernel<<<(unsigned)ceil_div(nR * d, 256), 256, 0, s>>>(rd, d, Rt.restrict_rows.as<int32_t>(), nR, rsub.as<double>());
            B200_LAUNCH_CHECK();
            rq = rsub.as<double>();
        }
        const int k1 = (int)std::min<int64_t>(choose_k(k, prop_k, nL), nL), k2 = (int)std::min<int64_t>(choose_k(k, prop_k, nR), nR);
        B200_TRY(w21.alloc(sizeof(int32_t) * nL * k2, s));
        B200_TRY(w12.alloc(sizeof(int32_t) * nR * k1, s));
        B200_TRY(sharded_query_knn(ds, rq, nR, lq, nL, d, k2, w21.as<int32_t>(), nullptr, s));   // neighbours of left cells in the right node
        B200_TRY(sharded_query_knn(ds, lq, nL, rq, nR, d, k1, w12.as<int32_t>(), nullptr, s));
        const int64_t cap = nL * (int64_t)k2;
        B200_TRY(first.alloc(sizeof(int32_t) * cap, s));
        B200_TRY(second.alloc(sizeof(int32_t) * cap, s));
        B200_TRY(dnp.alloc(sizeof(int64_t), s));
        B200_TRY(mutual::find_mutual_nns_device(w21.as<int32_t>(), nL, k2, w12.as<int32_t>(), nR, k1, first.as<int32_t>(), second.as<int32_t>(), cap,
                                                dnp.as<int64_t>(), 0, bad, s));
        int64_t np = 0;
        B200_CUDA(cudaMemcpyAsync(&np, dnp.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        *np_out = np;
        if (np > 0 && first_out) {
            if (L.nres >= 0) { map_ids_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, s>>>(first.as<int32_t>(), np, L.restrict_rows.as<int32_t>()); B200_LAUNCH_CHECK(); }
            if (Rt.nres >= 0) { map_ids_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, s>>>(second.as<int32_t>(), np, Rt.restrict_rows.as<int32_t>()); B200_LAUNCH_CHECK(); }
        }
        return 0;
    };

    // auto.merge (R/MNN_tree.R:154-226): .count_mnn_pairs of `left` against the listed nodes, with the reference's
    // orthogonalisation -- the left copy accumulates the centrings of every right node it has met so far (:186-188).
    auto count_pairs = [&](const Node& left, const std::vector<int>& rights, std::vector<int64_t>& counts) -> int {
        counts.assign(rights.size(), 0);
        DevBuf lcopy;
        bool lcopied = false;
        const double* ldata = left.rows();
        for (size_t j = 0; j < rights.size(); ++j) {
            const Node& right = R.nodes[rights[j]];
            const double* rdata = right.rows();
            DevBuf rcopy;
            if (!left.extras.empty()) {
                B200_TRY(rcopy.alloc(sizeof(double) * right.n * d, s));
                B200_CUDA(cudaMemcpyAsync(rcopy.p, right.rows(), sizeof(double) * right.n * d, cudaMemcpyDeviceToDevice, s));
                for (DevBuf* v : left.extras)
                    B200_TRY(correct::center_along_batch_vector_device(rcopy.as<double>(), right.n, d, v->as<double>(), right.restrict_rows.as<int32_t>(),
                                                                       std::max<int64_t>(right.nres, 0), bad, s));
                rdata = rcopy.as<double>();
            }
            if (!right.extras.empty()) {
                if (!lcopied) {
                    B200_TRY(lcopy.alloc(sizeof(double) * left.n * d, s));
                    B200_CUDA(cudaMemcpyAsync(lcopy.p, left.rows(), sizeof(double) * left.n * d, cudaMemcpyDeviceToDevice, s));
                    lcopied = true;
                    ldata = lcopy.as<double>();
                }
                for (DevBuf* v : right.extras)
                    B200_TRY(correct::center_along_batch_vector_device(lcopy.as<double>(), left.n, d, v->as<double>(), left.restrict_rows.as<int32_t>(),
                                                                       std::max<int64_t>(left.nres, 0), bad, s));
            }
            B200_TRY(find_pairs(ldata, left, rdata, right, nullptr, nullptr, &counts[j]));
        }
        return 0;
    };
    // B200MNN_MERGE_DEBUG=1: wall-clock per stage (each bracketed by a stream synchronisation), printed at the end
    const bool dbg = getenv("B200MNN_MERGE_DEBUG") != nullptr;
    std::vector<std::pair<std::string, double>> stage_s;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_last = now();
    auto tick = [&](const char* name) {
        if (!dbg) return;
        cudaStreamSynchronize(s);
        const double t = now();
        for (auto& e : stage_s) if (e.first == name) { e.second += t - t_last; t_last = t; return; }
        stage_s.emplace_back(name, t - t_last);
        t_last = t;
    };
    tick("upload");
    const bool auto_merge = (merge_left == nullptr || merge_right == nullptr);
    std::vector<int> remainders;                      // node ids still to be merged (auto mode)
    std::vector<std::vector<int64_t>> pairwise;       // lower-triangular counts, [i][j < i]
    if (auto_merge && nm > 0) {   // .initialize_auto_search (:154-168)
        for (int b = 0; b < nb; ++b) remainders.push_back(b);
        pairwise.assign(nb, std::vector<int64_t>(nb, 0));
        for (int i = 1; i < nb; ++i) {
            std::vector<int> below(remainders.begin(), remainders.begin() + i);
            std::vector<int64_t> c;
            B200_TRY(count_pairs(R.nodes[remainders[i]], below, c));
            for (int j = 0; j < i; ++j) pairwise[i][j] = c[j];
        }
    }
    R.merged_left.assign(nm, -1);
    R.merged_right.assign(nm, -1);

    for (int m = 0; m < nm; ++m) {
        int li, ri, pos_l = -1, pos_r = -1;
        if (auto_merge) {   // .pick_best_merge (:196-202): first maximum in column-major order; row = left, column = right
            int64_t best = -1;
            const int cnt = (int)remainders.size();
            for (int c = 0; c < cnt; ++c)
                for (int r = 0; r < cnt; ++r)
                    if (pairwise[r][c] > best) { best = pairwise[r][c]; pos_l = r; pos_r = c; }
            li = remainders[pos_l];
            ri = remainders[pos_r];
        } else {
            li = merge_left[m];
            ri = merge_right[m];
        }
        R.merged_left[m] = li;
        R.merged_right[m] = ri;
        if (li < 0 || ri < 0 || li >= nb + m || ri >= nb + m || li == ri || !R.nodes[li].alive || !R.nodes[ri].alive)
            return fail(B200MNN_EINVAL, "invalid merge order: a node is merged twice or before it exists");
        Node& L = R.nodes[li];
        Node& Rt = R.nodes[ri];
        double* ld = L.rows();
        double* rd = Rt.rows();
        std::vector<double> lold, rold, lnew, rnew;
        if (get_variance) {
            B200_TRY(perbatch_var(L, d, scratch.as<double>(), lold, s));
            B200_TRY(perbatch_var(Rt, d, scratch.as<double>(), rold, s));
        }
        tick("perbatch_var");
        // orthogonalise each side along the other side's earlier batch vectors (:473-474)
        for (DevBuf* v : L.extras) B200_TRY(correct::center_along_batch_vector_device(rd, Rt.n, d, v->as<double>(), Rt.restrict_rows.as<int32_t>(), std::max<int64_t>(Rt.nres, 0), bad, s));
        for (DevBuf* v : Rt.extras) B200_TRY(correct::center_along_batch_vector_device(ld, L.n, d, v->as<double>(), L.restrict_rows.as<int32_t>(), std::max<int64_t>(L.nres, 0), bad, s));

        tick("orthogonalize");
        // .restricted_mnn
        const int64_t nL = L.nres >= 0 ? L.nres : L.n, nR = Rt.nres >= 0 ? Rt.nres : Rt.n;
        DevBuf first, second;
        int64_t np = 0;
        B200_TRY(find_pairs(ld, L, rd, Rt, &first, &second, &np));
        if (np == 0) return fail(B200MNN_EINVAL, "no MNN pairs found between the batches being merged");

        tick("mnn_search");
        // correction vectors, overall batch vector, magnitude
        const int64_t acap = std::min<int64_t>(np, Rt.n);
        DevBuf averaged, uniq, dnm;
        B200_TRY(averaged.alloc(sizeof(double) * acap * d, s));
        B200_TRY(uniq.alloc(sizeof(int32_t) * acap, s));
        B200_TRY(dnm.alloc(sizeof(int64_t), s));
        B200_TRY(correct::average_correction_device(ld, L.n, rd, Rt.n, d, first.as<int32_t>(), second.as<int32_t>(), np, averaged.as<double>(),
                                                    uniq.as<int32_t>(), dnm.as<int64_t>(), bad, s));
        int64_t nmnn = 0;
        B200_CUDA(cudaMemcpyAsync(&nmnn, dnm.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        DevBuf* overall = new DevBuf();
        R.vec_pool.push_back(overall);
        B200_TRY(overall->alloc(sizeof(double) * d, s));
        B200_TRY(col_moment<0>(averaged.as<double>(), nmnn, d, nullptr, 1.0 / (double)nmnn, overall->as<double>(), s));   // colMeans(averaged) (:481)
        bool do_correct = true;
        if (!std::isnan(min_batch_skip)) {
            double* sq = scratch.as<double>();   // [d] mean of squares, then two scalars
            B200_TRY(col_moment<1>(averaged.as<double>(), nmnn, d, nullptr, 1.0 / (double)nmnn, sq, s));
            sum_vec_kernel<<<1, 1, 0, s>>>(sq, d, sq + d);
            B200_LAUNCH_CHECK();
            sumsq_vec_kernel<<<1, 1, 0, s>>>(overall->as<double>(), d, sq + d + 1);
            B200_LAUNCH_CHECK();
            double h[2];
            B200_CUDA(cudaMemcpyAsync(h, sq + d, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
            B200_CUDA(cudaStreamSynchronize(s));
            const double mag = (h[0] == 0.0) ? 0.0 : std::sqrt(h[1] / h[0]);   // .get_batch_magnitude (:582-595)
            R.batch_size[m] = mag;
            if (mag < min_batch_skip) { do_correct = false; R.skipped[m] = 1; }
        }
        tick("average_magnitude");
        DevBuf newright;
        if (do_correct) {
            B200_TRY(correct::center_along_batch_vector_device(ld, L.n, d, overall->as<double>(), L.restrict_rows.as<int32_t>(), std::max<int64_t>(L.nres, 0), bad, s));
            B200_TRY(correct::center_along_batch_vector_device(rd, Rt.n, d, overall->as<double>(), Rt.restrict_rows.as<int32_t>(), std::max<int64_t>(Rt.nres, 0), bad, s));
            if (get_variance) {   // recorded after the centring, before the tricube step (:500-501)
                B200_TRY(perbatch_var(L, d, scratch.as<double>(), lnew, s));
                B200_TRY(perbatch_var(Rt, d, scratch.as<double>(), rnew, s));
            }
            B200_TRY(correct::average_correction_device(ld, L.n, rd, Rt.n, d, first.as<int32_t>(), second.as<int32_t>(), np, averaged.as<double>(),
                                                        uniq.as<int32_t>(), dnm.as<int64_t>(), bad, s));
            tick("center_reaverage");
            // .tricube_weighted_correction: all right cells among the MNN cells of the right node
            const int kk = (int)std::min<int64_t>(std::min<int64_t>(choose_k(k, prop_k, Rt.n), Rt.n), nmnn);
            DevBuf sub, idx, dist;
            B200_TRY(sub.alloc(sizeof(double) * nmnn * d, s));
            gather_rows_kernel<<<(unsigned)ceil_div(nmnn * d, 256), 256, 0, s>>>(rd, d, uniq.as<int32_t>(), nmnn, sub.as<double>());
            B200_LAUNCH_CHECK();
            B200_TRY(idx.alloc(sizeof(int32_t) * Rt.n * kk, s));
            B200_TRY(dist.alloc(sizeof(double) * Rt.n * kk, s));
            B200_TRY(sharded_query_knn(ds, sub.as<double>(), nmnn, rd, Rt.n, d, kk, idx.as<int32_t>(), dist.as<double>(), s));
            B200_TRY(newright.alloc(sizeof(double) * Rt.n * d, s));
            B200_TRY(correct::tricube_apply_device(rd, Rt.n, d, averaged.as<double>(), nmnn, idx.as<int32_t>(), dist.as<double>(), kk, ndist,
                                                   newright.as<double>(), bad, s));
            rd = newright.as<double>();
        } else if (get_variance) {
            B200_TRY(perbatch_var(L, d, scratch.as<double>(), lnew, s));
            B200_TRY(perbatch_var(Rt, d, scratch.as<double>(), rnew, s));
        }
        tick("tricube");
        if (get_variance) {
            for (int b = 0; b < nb; ++b) R.lost_var[(size_t)m * nb + b] = 0.0;
            for (size_t i = 0; i < L.index.size(); ++i) R.lost_var[(size_t)m * nb + L.index[i]] = 1.0 - lnew[i] / lold[i];
            for (size_t i = 0; i < Rt.index.size(); ++i) R.lost_var[(size_t)m * nb + Rt.index[i]] = 1.0 - rnew[i] / rold[i];
        }
        // pairs back to the host (1-based rows within the left / right node, the reference's order)
        R.pl[m].resize((size_t)np); R.pr[m].resize((size_t)np);
        {
            DevBuf tmp;
            B200_TRY(tmp.alloc(sizeof(int32_t) * np, s));
            add_ids_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, s>>>(first.as<int32_t>(), np, 1, tmp.as<int32_t>());
            B200_LAUNCH_CHECK();
            B200_CUDA(cudaMemcpyAsync(R.pl[m].data(), tmp.p, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
            B200_CUDA(cudaStreamSynchronize(s));
            add_ids_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, s>>>(second.as<int32_t>(), np, 1, tmp.as<int32_t>());
            B200_LAUNCH_CHECK();
            B200_CUDA(cudaMemcpyAsync(R.pr[m].data(), tmp.p, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
            B200_CUDA(cudaStreamSynchronize(s));
        }
        tick("pairs_to_host");
        // the new node: rbind(left, right) (:520-525)
        Node& N = R.nodes[nb + m];
        N.n = L.n + Rt.n;
        if (L.ptr && Rt.ptr && L.ptr + L.n * d == Rt.ptr) {   // neighbours in the arena: the new node is the joint window
            N.ptr = L.ptr;
            if (rd != Rt.ptr) B200_CUDA(cudaMemcpyAsync(Rt.ptr, rd, sizeof(double) * Rt.n * d, cudaMemcpyDeviceToDevice, s));   // tricube output back in place
        } else {
            B200_TRY(N.data.alloc(sizeof(double) * N.n * d, s));
            B200_CUDA(cudaMemcpyAsync(N.data.p, ld, sizeof(double) * L.n * d, cudaMemcpyDeviceToDevice, s));
            B200_CUDA(cudaMemcpyAsync(N.data.as<double>() + L.n * d, rd, sizeof(double) * Rt.n * d, cudaMemcpyDeviceToDevice, s));
        }
        if (L.nres >= 0 || Rt.nres >= 0) {   // .combine_restrict (:610-622)
            N.nres = nL + nR;
            B200_TRY(N.restrict_rows.alloc(sizeof(int32_t) * N.nres, s));
            int32_t* o = N.restrict_rows.as<int32_t>();
            if (L.nres >= 0) B200_CUDA(cudaMemcpyAsync(o, L.restrict_rows.p, sizeof(int32_t) * nL, cudaMemcpyDeviceToDevice, s));
            else { iota_kernel<<<(unsigned)ceil_div(nL, 256), 256, 0, s>>>(o, nL, 0); B200_LAUNCH_CHECK(); }
            if (Rt.nres >= 0) { add_ids_kernel<<<(unsigned)ceil_div(nR, 256), 256, 0, s>>>(Rt.restrict_rows.as<int32_t>(), nR, (int32_t)L.n, o + nL); B200_LAUNCH_CHECK(); }
            else { iota_kernel<<<(unsigned)ceil_div(nR, 256), 256, 0, s>>>(o + nL, nR, (int32_t)L.n); B200_LAUNCH_CHECK(); }
        }
        N.index = L.index; N.index.insert(N.index.end(), Rt.index.begin(), Rt.index.end());
        N.seg = L.seg; N.seg.insert(N.seg.end(), Rt.seg.begin(), Rt.seg.end());
        N.extras = L.extras; N.extras.insert(N.extras.end(), Rt.extras.begin(), Rt.extras.end());
        if (do_correct) N.extras.push_back(overall);
        N.alive = true;
        B200_CUDA(cudaStreamSynchronize(s));   // buffers of this merge are released below
        L.data.release(); L.restrict_rows.release(); L.alive = false;
        Rt.data.release(); Rt.restrict_rows.release(); Rt.alive = false;
        R.final_node = nb + m;
        tick("rbind");
        if (auto_merge && m + 1 < nm) {   // .update_remainders (:205-226)
            std::vector<int> rest;
            std::vector<std::vector<int64_t>> meta;
            const int cnt = (int)remainders.size();
            for (int r = 0; r < cnt; ++r) {
                if (r == pos_l || r == pos_r) continue;
                rest.push_back(remainders[r]);
                std::vector<int64_t> row;
                for (int c = 0; c < cnt; ++c)
                    if (c != pos_l && c != pos_r) row.push_back(pairwise[r][c]);
                row.push_back(0);
                meta.push_back(row);
            }
            std::vector<int64_t> c;
            B200_TRY(count_pairs(R.nodes[nb + m], rest, c));
            c.push_back(0);
            meta.push_back(c);
            rest.push_back(nb + m);
            remainders.swap(rest);
            pairwise.swap(meta);
        }
    }
    if (nm == 0) R.final_node = 0;
    if (dbg) {
        fprintf(stderr, "b200mnn_reduced_mnn: %d devices;", (int)ds.devs.size());
        for (auto& e : stage_s) fprintf(stderr, " %s %.3f s;", e.first.c_str(), e.second);
        fprintf(stderr, "\n");
    }
    int hb = 0;
    B200_CUDA(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    if (hb) return fail(B200MNN_EINVAL, "subset indices out of range");
    guard.h = nullptr;
    *result_out = holder;
    return 0;
}

int64_t b200mnn_result_ncells(const b200mnn_merge_result* res) { return res ? res->r.ntotal : -1; }

int64_t b200mnn_result_npairs(const b200mnn_merge_result* res, int merge) {
    if (!res || merge < 0 || merge >= (int)res->r.pl.size()) return -1;
    return (int64_t)res->r.pl[merge].size();
}

int b200mnn_result_pairs(const b200mnn_merge_result* res, int merge, int32_t* left_out, int32_t* right_out) {
    if (!res || merge < 0 || merge >= (int)res->r.pl.size()) return b200::fail(B200MNN_EINVAL, "merge index out of range");
    const auto& l = res->r.pl[merge];
    const auto& r = res->r.pr[merge];
    if (!l.empty()) { memcpy(left_out, l.data(), sizeof(int32_t) * l.size()); memcpy(right_out, r.data(), sizeof(int32_t) * r.size()); }
    return 0;
}

int b200mnn_result_corrected(const b200mnn_merge_result* res, double* out, int col_major) {
    using namespace b200;
    if (!res || res->r.final_node < 0) return fail(B200MNN_EINVAL, "no result");
    const MergeResult& R = res->r;
    B200_CUDA(cudaSetDevice(R.device0));
    const b200::merge::Node& nd = R.nodes[R.final_node];
    cudaStream_t s = R.stream;
    if (col_major) {
        b200::merge::DevBuf t;
        B200_TRY(t.alloc(sizeof(double) * nd.n * R.d, s));
        B200_TRY(correct::transpose_device<double>(nd.rows(), R.d, nd.n, t.as<double>(), s));   // row-major [n x d] == column-major [d x n]
        B200_CUDA(cudaMemcpyAsync(out, t.p, sizeof(double) * nd.n * R.d, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
    } else {
        B200_CUDA(cudaMemcpyAsync(out, nd.rows(), sizeof(double) * nd.n * R.d, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int b200mnn_result_info(const b200mnn_merge_result* res, int32_t* node_order, int64_t* node_ncells, double* batch_size, int32_t* skipped,
                        double* lost_var) {
    if (!res || res->r.final_node < 0) return b200::fail(B200MNN_EINVAL, "no result");
    const MergeResult& R = res->r;
    const b200::merge::Node& nd = R.nodes[R.final_node];
    for (size_t i = 0; i < nd.index.size(); ++i) {
        if (node_order) node_order[i] = nd.index[i] + 1;
        if (node_ncells) node_ncells[i] = nd.seg[i];
    }
    for (size_t m = 0; m < R.batch_size.size(); ++m) {
        if (batch_size) batch_size[m] = R.batch_size[m];
        if (skipped) skipped[m] = R.skipped[m];
    }
    if (lost_var && !R.lost_var.empty()) memcpy(lost_var, R.lost_var.data(), sizeof(double) * R.lost_var.size());
    return 0;
}

int b200mnn_result_merges(const b200mnn_merge_result* res, int32_t* left_out, int32_t* right_out) {
    if (!res) return b200::fail(B200MNN_EINVAL, "no result");
    for (size_t m = 0; m < res->r.merged_left.size(); ++m) { left_out[m] = res->r.merged_left[m]; right_out[m] = res->r.merged_right[m]; }
    return 0;
}

void b200mnn_result_free(b200mnn_merge_result* res) { delete res; }

}  // extern "C"
