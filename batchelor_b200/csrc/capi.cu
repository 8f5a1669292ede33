// Host-buffer C ABI of libb200mnn (the entry points an R .Call shim binds; see include/b200mnn.h and INTEGRATION.md).
// Each call: stage the R-layout host buffers on the device, run the CUDA path, copy the result back into the
// caller-provided output.  There is no CPU computation here beyond argument checks: without a device every call fails.
#include "internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

namespace b200 {
namespace host {

static cudaStream_t lib_stream() {
    // one non-blocking stream per device, created lazily
    static std::mutex mu;
    static std::vector<cudaStream_t> streams;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if ((int)streams.size() <= dev) streams.resize(dev + 1, nullptr);
    if (!streams[dev]) {
        if (cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    return streams[dev];
}

// Two side streams and three events per device (created lazily): the two searches of findMutualNN are independent, so
// their kernels are enqueued side by side -- the serial stretches of one direction's cluster plan overlap with the other.
struct SideStreams {
    cudaStream_t a = nullptr, b = nullptr;
    cudaEvent_t start = nullptr, done_a = nullptr, done_b = nullptr;
};
static SideStreams* side_streams() {
    static std::mutex mu;
    static std::vector<SideStreams> all;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if ((int)all.size() <= dev) all.resize(dev + 1);
    SideStreams& ss = all[dev];
    if (!ss.a) {
        bool ok = cudaStreamCreateWithFlags(&ss.a, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&ss.b, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&ss.start, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&ss.done_a, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&ss.done_b, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); ss = SideStreams(); return nullptr; }
    }
    return &all[dev];
}

template <typename T>
static T* upload(Scratch& ws, const T* host, size_t count, cudaStream_t s) {
    T* d = ws.get<T>(count);
    if (!d) return nullptr;
    if (count && cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, s) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(H2D)", __FILE__, __LINE__);
        return nullptr;
    }
    return d;
}

__global__ void add_i32_kernel(int32_t* __restrict__ v, int64_t n, int32_t delta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += delta;
}

__global__ void gather_rows_kernel(const double* __restrict__ src, int64_t nsrc, int d, const int32_t* __restrict__ rows0, int64_t nrows,
                                   double* __restrict__ dst, int* __restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nrows * d) return;
    const int64_t r = e / d;
    const int t = (int)(e - r * d);
    const int64_t s = rows0[r];
    if (s < 0 || s >= nsrc) { *bad = 1; dst[e] = 0.0; return; }
    dst[e] = src[s * d + t];
}

static int add_i32(int32_t* d, int64_t n, int32_t delta, cudaStream_t s) {
    if (n <= 0) return 0;
    add_i32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(d, n, delta);
    B200_LAUNCH_CHECK();
    return 0;
}

// Stage a [rows x cols] R matrix (column-major) as row-major on the device; if !col_major it is copied as is.
static double* stage_matrix(Scratch& ws, const double* host, int64_t rows, int64_t cols, bool col_major, cudaStream_t s, int* rc) {
    *rc = 0;
    double* raw = upload(ws, host, (size_t)rows * cols, s);
    if (!raw) { *rc = ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM; return nullptr; }
    if (!col_major || rows <= 1 || cols <= 1) return raw;
    double* rm = ws.get<double>((size_t)rows * cols);
    if (!rm) { *rc = B200MNN_ENOMEM; return nullptr; }
    *rc = correct::transpose_device<double>(raw, rows, cols, rm, s);
    return *rc ? nullptr : rm;
}

// Copy a device row-major [rows x cols] matrix back into an R column-major (or row-major) host buffer.
template <typename T>
static int unstage_matrix(Scratch& ws, const T* d_rm, int64_t rows, int64_t cols, bool col_major, T* host, cudaStream_t s) {
    if (rows * cols == 0) return 0;
    const T* src = d_rm;
    if (col_major && rows > 1 && cols > 1) {
        T* cm = ws.get<T>((size_t)rows * cols);
        if (!cm) return B200MNN_ENOMEM;
        // row-major [rows x cols] == column-major [cols x rows]; its "row-major form" is column-major [rows x cols]
        B200_TRY(correct::transpose_device<T>(d_rm, cols, rows, cm, s));
        src = cm;
    }
    B200_CUDA(cudaMemcpyAsync(host, src, sizeof(T) * rows * cols, cudaMemcpyDeviceToHost, s));
    return 0;
}

static int finish(cudaStream_t s, const int* d_bad, const char* bad_msg) {
    int bad = 0;
    if (d_bad) B200_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    if (bad) return fail(B200MNN_EINVAL, bad_msg);
    return 0;
}

}  // namespace host
}  // namespace b200

using namespace b200;
using namespace b200::host;

extern "C" {

int b200mnn_query_knn(const double* X, int64_t n, const double* Q, int64_t nq, int d, int k, int col_major, int32_t* idx_out,
                      double* dist_out) {
    B200_TRY(ensure_device());
    if (n < 0 || nq < 0 || d < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (k < 0 || k > n) return fail(B200MNN_EINVAL, "'k' must be non-negative and no larger than the number of points in 'X'");
    if (nq == 0 || k == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    int rc;
    double* dX = stage_matrix(ws, X, n, d, col_major != 0, s, &rc);
    if (rc) return rc;
    double* dQ = stage_matrix(ws, Q, nq, d, col_major != 0, s, &rc);
    if (rc) return rc;
    int32_t* d_idx = ws.get<int32_t>((size_t)nq * k);
    double* d_dist = dist_out ? ws.get<double>((size_t)nq * k) : nullptr;
    if (!ws.ok()) return B200MNN_ENOMEM;
    B200_TRY(knn::query_knn_device(dX, n, dQ, nq, d, k, d_idx, d_dist, nullptr, s, nullptr));
    B200_TRY(add_i32(d_idx, nq * k, 1, s));
    B200_TRY(unstage_matrix<int32_t>(ws, d_idx, nq, k, col_major != 0, idx_out, s));
    if (dist_out) B200_TRY(unstage_matrix<double>(ws, d_dist, nq, k, col_major != 0, dist_out, s));
    return finish(s, nullptr, "");
}

// Host buffers -> device, pipelined (large row counts): batch 1 crosses PCIe first; while batch 2 follows in row chunks,
// the GPU already prepares batch 1 as the reference set of the second search (norms, cluster plan, grouped operand) and
// then answers every chunk of batch-2 rows as soon as it has landed (knn::RefCache).  The first search (batch-1 rows in
// batch 2) needs all of batch 2 and starts when the last chunk is in.  Results are identical to the unpipelined order of
// operations: every search is exact.
static const int64_t kPipelineMinRows = 262144;
static const int kPipelineChunks = 3;   // measured at 1M x 1M x 50: 2-3 chunks 55.9 ms, 4: 56.4, 6: 57.5, unpipelined 61.5

struct EventPool {   // events of one call, destroyed at scope exit
    std::vector<cudaEvent_t> ev;
    ~EventPool() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
    cudaEvent_t make() {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ev.push_back(e);
        return e;
    }
};

// Uploads rows [r0, r1) of a host [rows x cols] matrix (column-major: R layout; else row-major) into the row-major device
// matrix `dst` ([rows x cols]); `raw` is a device staging buffer of (r1 - r0) * cols doubles for the column-major case.
static int upload_rows(const double* host, int64_t rows, int64_t cols, bool col_major, int64_t r0, int64_t r1, double* dst, double* raw,
                       cudaStream_t s) {
    const int64_t nr = r1 - r0;
    if (nr <= 0) return 0;
    if (!col_major || cols <= 1) {
        B200_CUDA(cudaMemcpyAsync(dst + r0 * cols, host + r0 * cols, sizeof(double) * (size_t)nr * cols, cudaMemcpyHostToDevice, s));
        return 0;
    }
    // column t of the chunk: host[t * rows + r0 ...] -> raw[t * nr ...] (a column-major [nr x cols] block), then transpose
    B200_CUDA(cudaMemcpy2DAsync(raw, sizeof(double) * (size_t)nr, host + r0, sizeof(double) * (size_t)rows, sizeof(double) * (size_t)nr, (size_t)cols,
                                cudaMemcpyHostToDevice, s));
    return correct::transpose_device<double>(raw, nr, cols, dst + r0 * cols, s);
}

int b200mnn_find_mutual_nn(const double* data1, int64_t n1, const double* data2, int64_t n2, int d, int k1, int k2, int col_major,
                           int32_t* first_out, int32_t* second_out, int64_t capacity, int64_t* np_out) {
    B200_TRY(ensure_device());
    if (n1 < 0 || n2 < 0 || d < 0 || k1 < 0 || k2 < 0) return fail(B200MNN_EINVAL, "negative dimension");
    // BiocNeighbors caps k at the number of points (with a warning on the R side)
    k1 = (int)std::min<int64_t>(k1, n1);
    k2 = (int)std::min<int64_t>(k2, n2);
    *np_out = 0;
    if (n1 == 0 || n2 == 0 || k1 == 0 || k2 == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    SideStreams* ss = getenv("B200MNN_SERIAL") ? nullptr : side_streams();   // B200MNN_SERIAL: both searches on the one stream
    int64_t min_rows = kPipelineMinRows;
    if (const char* e = getenv("B200MNN_PIPELINE_MIN_ROWS")) min_rows = std::max<int64_t>(256, atoll(e));   // tests exercise the path on small inputs
    const bool pipelined = ss && !getenv("B200MNN_NO_PIPELINE") && n1 >= min_rows && n2 >= min_rows && d > 0 &&
                           knn::tensor_path_supported(n1, n2 / kPipelineChunks, d, k1);
    int rc;
    double* d1 = stage_matrix(ws, data1, n1, d, col_major != 0, s, &rc);
    if (rc) return rc;
    double* d2 = nullptr;
    if (!pipelined) {
        d2 = stage_matrix(ws, data2, n2, d, col_major != 0, s, &rc);
        if (rc) return rc;
    }
    int32_t* w21 = ws.get<int32_t>((size_t)n1 * k2);  // neighbours of batch-1 cells in batch 2
    int32_t* w12 = ws.get<int32_t>((size_t)n2 * k1);  // neighbours of batch-2 cells in batch 1
    const int64_t cap = std::min<int64_t>(capacity, n1 * (int64_t)k2);
    int32_t* dfirst = ws.get<int32_t>((size_t)std::max<int64_t>(cap, 1));
    int32_t* dsecond = ws.get<int32_t>((size_t)std::max<int64_t>(cap, 1));
    int64_t* dnp = ws.get<int64_t>(1);
    if (!ws.ok()) return B200MNN_ENOMEM;
    EventPool events;
    if (pipelined) {
        knn::RefCache cache(ss->a);   // batch 1 as the reference set of the second search; lives on stream a
        int nchunks = kPipelineChunks;
        if (const char* e = getenv("B200MNN_PIPELINE_CHUNKS")) nchunks = std::max(1, std::min(64, atoi(e)));
        const int64_t chunk = round_up(ceil_div(n2, nchunks), 128);
        const bool cm = col_major != 0 && d > 1;
        d2 = ws.get<double>((size_t)n2 * d);
        double* raw = cm ? ws.get<double>((size_t)chunk * d) : nullptr;
        if (!ws.ok()) return B200MNN_ENOMEM;
        auto bail = [&](int code) {   // the scratch buffers must outlive whatever was enqueued on the side streams
            const std::string msg = b200mnn_last_error();
            cudaDeviceSynchronize();
            cudaGetLastError();
            return fail(code, msg);
        };
        cudaEvent_t e1 = events.make();
        if (!e1) return fail(B200MNN_ECUDA, "cudaEventCreate failed");
        B200_CUDA(cudaEventRecord(e1, s));                 // batch 1 (and every buffer allocated above) is ready
        B200_CUDA(cudaStreamWaitEvent(ss->a, e1, 0));
        cache.nq_hint = std::min(chunk, n2);
        rc = knn::query_knn_device(d1, n1, nullptr, 0, d, k1, nullptr, nullptr, nullptr, ss->a, nullptr, &cache);   // reference side only
        if (rc) return bail(rc);
        for (int64_t r0 = 0; r0 < n2; r0 += chunk) {
            const int64_t r1 = std::min(n2, r0 + chunk);
            rc = upload_rows(data2, n2, d, cm, r0, r1, d2, raw, s);
            if (rc) return bail(rc);
            cudaEvent_t ec = events.make();
            if (!ec) return bail(B200MNN_ECUDA);
            if (cudaEventRecord(ec, s) != cudaSuccess || cudaStreamWaitEvent(ss->a, ec, 0) != cudaSuccess) return bail(B200MNN_ECUDA);
            // (copies and transposes share stream s: the next chunk's copy into `raw` is ordered after this chunk's transpose)
            rc = knn::query_knn_device(d1, n1, d2 + r0 * d, r1 - r0, d, k1, w12 + r0 * k1, nullptr, nullptr, ss->a, nullptr, &cache);
            if (rc) return bail(rc);
        }
        cudaEvent_t e2 = events.make();
        if (!e2) return bail(B200MNN_ECUDA);
        if (cudaEventRecord(e2, s) != cudaSuccess || cudaStreamWaitEvent(ss->b, e2, 0) != cudaSuccess) return bail(B200MNN_ECUDA);
        rc = knn::query_knn_device(d2, n2, d1, n1, d, k2, w21, nullptr, nullptr, ss->b, nullptr);
        if (rc) return bail(rc);
        if (cudaEventRecord(ss->done_a, ss->a) != cudaSuccess || cudaEventRecord(ss->done_b, ss->b) != cudaSuccess ||
            cudaStreamWaitEvent(s, ss->done_a, 0) != cudaSuccess || cudaStreamWaitEvent(s, ss->done_b, 0) != cudaSuccess)
            return bail(B200MNN_ECUDA);
        rc = mutual::find_mutual_nns_device(w21, n1, k2, w12, n2, k1, dfirst, dsecond, cap, dnp, 1, nullptr, s);
        if (rc) return bail(rc);
        int64_t np = 0;
        if (cudaMemcpyAsync(&np, dnp, sizeof(int64_t), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
            return bail(B200MNN_ECUDA);
        // everything on the side streams is complete here: the cache may release its buffers at scope exit
        *np_out = np;
        if (np > capacity) return fail(B200MNN_ECAPACITY, "pair output capacity too small");
        if (np > 0) {
            B200_CUDA(cudaMemcpyAsync(first_out, dfirst, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
            B200_CUDA(cudaMemcpyAsync(second_out, dsecond, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
        }
        return finish(s, nullptr, "");
    }
    if (ss) {
        B200_CUDA(cudaEventRecord(ss->start, s));
        B200_CUDA(cudaStreamWaitEvent(ss->a, ss->start, 0));
        B200_CUDA(cudaStreamWaitEvent(ss->b, ss->start, 0));
        const int rc1 = knn::query_knn_device(d2, n2, d1, n1, d, k2, w21, nullptr, nullptr, ss->a, nullptr);
        const int rc2 = rc1 ? 0 : knn::query_knn_device(d1, n1, d2, n2, d, k1, w12, nullptr, nullptr, ss->b, nullptr);
        if (rc1 || rc2) {   // the scratch buffers must outlive whatever was enqueued on the side streams
            const std::string msg = b200mnn_last_error();
            cudaDeviceSynchronize();
            cudaGetLastError();
            return fail(rc1 ? rc1 : rc2, msg);
        }
        B200_CUDA(cudaEventRecord(ss->done_a, ss->a));
        B200_CUDA(cudaEventRecord(ss->done_b, ss->b));
        B200_CUDA(cudaStreamWaitEvent(s, ss->done_a, 0));
        B200_CUDA(cudaStreamWaitEvent(s, ss->done_b, 0));
    } else {
        B200_TRY(knn::query_knn_device(d2, n2, d1, n1, d, k2, w21, nullptr, nullptr, s, nullptr));
        B200_TRY(knn::query_knn_device(d1, n1, d2, n2, d, k1, w12, nullptr, nullptr, s, nullptr));
    }
    B200_TRY(mutual::find_mutual_nns_device(w21, n1, k2, w12, n2, k1, dfirst, dsecond, cap, dnp, 1, nullptr, s));
    int64_t np = 0;
    B200_CUDA(cudaMemcpyAsync(&np, dnp, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    *np_out = np;
    if (np > capacity) return fail(B200MNN_ECAPACITY, "pair output capacity too small");
    if (np > 0) {
        B200_CUDA(cudaMemcpyAsync(first_out, dfirst, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaMemcpyAsync(second_out, dsecond, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
    }
    return finish(s, nullptr, "");
}

int b200mnn_find_mutual_nns(const int32_t* left, int64_t n1, int k2, const int32_t* right, int64_t n2, int k1, int32_t* first_out,
                            int32_t* second_out, int64_t* np_out) {
    B200_TRY(ensure_device());
    if (n1 < 0 || n2 < 0 || k1 < 0 || k2 < 0) return fail(B200MNN_EINVAL, "negative dimension");
    *np_out = 0;
    if (n1 == 0 || k2 == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    int32_t* lraw = upload(ws, left, (size_t)n1 * k2, s);
    int32_t* rraw = upload(ws, right, (size_t)n2 * k1, s);
    int32_t* l = ws.get<int32_t>((size_t)n1 * k2);
    int32_t* r = ws.get<int32_t>((size_t)std::max<int64_t>(n2 * k1, 1));
    int32_t* dfirst = ws.get<int32_t>((size_t)n1 * k2);
    int32_t* dsecond = ws.get<int32_t>((size_t)n1 * k2);
    int64_t* dnp = ws.get<int64_t>(1);
    int* bad = ws.get<int>(1);
    if (!ws.ok() || !lraw || !rraw) return ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    B200_TRY(correct::transpose_device<int32_t>(lraw, n1, k2, l, s));
    B200_TRY(correct::transpose_device<int32_t>(rraw, n2, k1, r, s));
    B200_TRY(add_i32(l, n1 * k2, -1, s));
    B200_TRY(add_i32(r, n2 * k1, -1, s));
    B200_TRY(mutual::find_mutual_nns_device(l, n1, k2, r, n2, k1, dfirst, dsecond, n1 * (int64_t)k2, dnp, 1, bad, s));
    int64_t np = 0;
    B200_CUDA(cudaMemcpyAsync(&np, dnp, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B200_TRY(finish(s, bad, "neighbour index out of range in 'left'"));
    *np_out = np;
    if (np > 0) {
        B200_CUDA(cudaMemcpyAsync(first_out, dfirst, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaMemcpyAsync(second_out, dsecond, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, s));
    }
    return finish(s, nullptr, "");
}

int b200mnn_smooth_gaussian_kernel(const double* averaged, int64_t G, int64_t nmnn, const int32_t* index0, int64_t nindex,
                                   const double* mat, int64_t Gdist, int64_t ncells, double sigma2, double* out) {
    if (nmnn != nindex) return fail(B200MNN_EINVAL, "'index' must have length equal to number of rows in 'averaged'");
    B200_TRY(ensure_device());
    if (G < 0 || nmnn < 0 || Gdist < 0 || ncells < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (G * ncells == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    // R's column-major [G x nmnn] is exactly the device layout "row-major [nmnn x G]": no transposes needed
    double* dA = upload(ws, averaged, (size_t)G * nmnn, s);
    int32_t* dI = upload(ws, index0, (size_t)nmnn, s);
    double* dM = upload(ws, mat, (size_t)Gdist * ncells, s);
    double* dO = ws.get<double>((size_t)G * ncells);
    int* bad = ws.get<int>(1);
    if (!ws.ok() || !dA || !dI || !dM) return ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    B200_TRY(smooth::smooth_gaussian_kernel_device(dA, G, nmnn, dI, dM, Gdist, ncells, sigma2, dO, bad, s));
    B200_CUDA(cudaMemcpyAsync(out, dO, sizeof(double) * G * ncells, cudaMemcpyDeviceToHost, s));
    return finish(s, bad, "'index' entries out of range");
}

int b200mnn_adjust_shift_variance(const double* data1, int64_t G1, int64_t n1, const double* data2, int64_t G2, int64_t n2,
                                  const double* vect, int64_t vrows, int64_t vcols, double sigma2, const int32_t* restrict1, int64_t nr1,
                                  const int32_t* restrict2, int64_t nr2, double* out) {
    if (G1 != G2 || G1 != vcols) return fail(B200MNN_EINVAL, "number of genes do not match up between matrices");
    if (n2 != vrows) return fail(B200MNN_EINVAL, "number of cells do not match up between matrices");
    for (int64_t i = 0; i < nr1; ++i)
        if (restrict1[i] == INT32_MIN || restrict1[i] < 0 || restrict1[i] >= n1) return fail(B200MNN_EINVAL, "subset indices out of range");
    for (int64_t i = 0; i < nr2; ++i)
        if (restrict2[i] == INT32_MIN || restrict2[i] < 0 || restrict2[i] >= n2) return fail(B200MNN_EINVAL, "subset indices out of range");
    B200_TRY(ensure_device());
    if (n2 == 0) return 0;
    const int64_t G = G1;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    int rc;
    double* d1 = upload(ws, data1, (size_t)G * n1, s);
    double* d2 = upload(ws, data2, (size_t)G * n2, s);
    double* dv = stage_matrix(ws, vect, n2, G, true, s, &rc);  // [n2 x G] column-major -> one cell contiguous
    if (rc) return rc;
    int32_t* r1 = upload(ws, restrict1, (size_t)nr1, s);
    int32_t* r2 = upload(ws, restrict2, (size_t)nr2, s);
    double* dO = ws.get<double>((size_t)n2);
    int* bad = ws.get<int>(1);
    if (!ws.ok() || !d1 || !d2 || !r1 || !r2) return ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    B200_TRY(shiftvar::adjust_shift_variance_device(d1, n1, d2, n2, G, dv, sigma2, r1, nr1, r2, nr2, dO, bad, s));
    B200_CUDA(cudaMemcpyAsync(out, dO, sizeof(double) * n2, cudaMemcpyDeviceToHost, s));
    return finish(s, bad, "subset indices out of range");
}

int b200mnn_cosine_norm(const double* x, int64_t G, int64_t n, double* out, double* l2_out) {
    B200_TRY(ensure_device());
    if (G < 0 || n < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (n == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    double* dx = upload(ws, x, (size_t)G * n, s);
    double* dout = out ? ws.get<double>((size_t)G * n) : nullptr;
    double* dl2 = l2_out ? ws.get<double>((size_t)n) : nullptr;
    if (!ws.ok() || !dx) return ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM;
    B200_TRY(correct::cosine_norm_device(dx, n, G, dout, dl2, s));
    if (out && G > 0) B200_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * G * n, cudaMemcpyDeviceToHost, s));
    if (l2_out) B200_CUDA(cudaMemcpyAsync(l2_out, dl2, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    return finish(s, nullptr, "");
}

int b200mnn_average_correction(const double* refdata, int64_t n1, const double* curdata, int64_t n2, int d, const int32_t* mnn1,
                               const int32_t* mnn2, int64_t np, double* averaged_out, int32_t* second_out, int64_t* nmnn_out) {
    B200_TRY(ensure_device());
    if (n1 < 0 || n2 < 0 || d < 0 || np < 0) return fail(B200MNN_EINVAL, "negative dimension");
    *nmnn_out = 0;
    if (np == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    int rc;
    double* dref = stage_matrix(ws, refdata, n1, d, true, s, &rc);
    if (rc) return rc;
    double* dcur = stage_matrix(ws, curdata, n2, d, true, s, &rc);
    if (rc) return rc;
    int32_t* f = upload(ws, mnn1, (size_t)np, s);
    int32_t* g = upload(ws, mnn2, (size_t)np, s);
    const int64_t cap = std::min<int64_t>(np, n2);
    double* davg = ws.get<double>((size_t)cap * std::max(d, 1));
    int32_t* dsec = ws.get<int32_t>((size_t)cap);
    int64_t* dn = ws.get<int64_t>(1);
    int* bad = ws.get<int>(1);
    if (!ws.ok() || !f || !g) return ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    B200_TRY(add_i32(f, np, -1, s));
    B200_TRY(add_i32(g, np, -1, s));
    B200_TRY(correct::average_correction_device(dref, n1, dcur, n2, d, f, g, np, davg, dsec, dn, bad, s));
    int64_t nmnn = 0;
    B200_CUDA(cudaMemcpyAsync(&nmnn, dn, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B200_TRY(finish(s, bad, "MNN pair index out of range"));
    *nmnn_out = nmnn;
    if (nmnn > 0) {
        B200_TRY(add_i32(dsec, nmnn, 1, s));
        B200_CUDA(cudaMemcpyAsync(second_out, dsec, sizeof(int32_t) * nmnn, cudaMemcpyDeviceToHost, s));
        B200_TRY(unstage_matrix<double>(ws, davg, nmnn, d, true, averaged_out, s));
    }
    return finish(s, nullptr, "");
}

int b200mnn_center_along_batch_vector(const double* mat, int64_t n, int d, const double* batch_vec, const int32_t* restrict1,
                                      int64_t nrestrict, double* out) {
    B200_TRY(ensure_device());
    if (n < 0 || d < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (n * d == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    int rc;
    double* dm = stage_matrix(ws, mat, n, d, true, s, &rc);
    if (rc) return rc;
    double* dv = upload(ws, batch_vec, (size_t)d, s);
    int32_t* dr = restrict1 ? upload(ws, restrict1, (size_t)nrestrict, s) : nullptr;
    int* bad = ws.get<int>(1);
    if (!ws.ok() || !dv || (restrict1 && !dr)) return ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    if (dr) B200_TRY(add_i32(dr, nrestrict, -1, s));
    // stage_matrix may hand back the raw upload (no transpose needed); centring is in place either way
    B200_TRY(correct::center_along_batch_vector_device(dm, n, d, dv, dr, nrestrict, bad, s));
    B200_TRY(unstage_matrix<double>(ws, dm, n, d, true, out, s));
    return finish(s, bad, "restrict indices out of range");
}

int b200mnn_tricube_weighted_correction(const double* curdata, int64_t n, int d, const double* correction, const int32_t* in_mnn,
                                        int64_t nmnn, int k, double ndist, double* out) {
    B200_TRY(ensure_device());
    if (n < 0 || d < 0 || nmnn < 0 || k < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (n * d == 0) return 0;
    cudaStream_t s = lib_stream();
    Scratch ws(s);
    int rc;
    double* dcur = stage_matrix(ws, curdata, n, d, true, s, &rc);
    if (rc) return rc;
    const int safe_k = (int)std::min<int64_t>(k, nmnn);  // R/fastMNN.R:604
    double* dout = ws.get<double>((size_t)n * d);
    int* bad = ws.get<int>(1);
    if (!ws.ok()) return B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    if (safe_k > 0) {
        double* dcorr = stage_matrix(ws, correction, nmnn, d, true, s, &rc);
        if (rc) return rc;
        int32_t* dmnn = upload(ws, in_mnn, (size_t)nmnn, s);
        double* dsub = ws.get<double>((size_t)nmnn * d);
        int32_t* didx = ws.get<int32_t>((size_t)n * safe_k);
        double* ddist = ws.get<double>((size_t)n * safe_k);
        if (!ws.ok() || !dmnn) return ws.ok() ? B200MNN_ECUDA : B200MNN_ENOMEM;
        B200_TRY(add_i32(dmnn, nmnn, -1, s));
        gather_rows_kernel<<<(unsigned)ceil_div(nmnn * d, 256), 256, 0, s>>>(dcur, n, d, dmnn, nmnn, dsub, bad);
        B200_LAUNCH_CHECK();
        B200_TRY(knn::query_knn_device(dsub, nmnn, dcur, n, d, safe_k, didx, ddist, nullptr, s, nullptr));
        B200_TRY(correct::tricube_apply_device(dcur, n, d, dcorr, nmnn, didx, ddist, safe_k, ndist, dout, bad, s));
    } else {
        B200_CUDA(cudaMemcpyAsync(dout, dcur, sizeof(double) * n * d, cudaMemcpyDeviceToDevice, s));
    }
    B200_TRY(unstage_matrix<double>(ws, dout, n, d, true, out, s));
    return finish(s, bad, "'in.mnn' indices out of range");
}

}  // extern "C"
