// Error plumbing, device selection and stream-ordered scratch memory for libb200mnn.
#include "common.cuh"

#include <atomic>
#include <mutex>

namespace b200 {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[1024];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    cudaGetLastError();  // clear the sticky-free error state
    return e == cudaErrorMemoryAllocation ? B200MNN_ENOMEM : B200MNN_ECUDA;
}

static std::once_flag g_pool_once;

int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(B200MNN_ECUDA, "no usable CUDA device: libb200mnn has no CPU fallback");
    }
    // Keep freed scratch cached in the default pool instead of returning it to the driver on every sync.
    std::call_once(g_pool_once, []() {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t thr = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
        cudaGetLastError();
    });
    return 0;
}

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaGetLastError();
    cached = n > 0 ? n : 148;
    return cached;
}

void* Scratch::alloc(size_t bytes) {
    if (!ok_) return nullptr;
    if (n_ >= kMax) { ok_ = false; set_error("internal: too many scratch allocations"); return nullptr; }
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(&p, bytes, stream_);
    if (e != cudaSuccess) {
        ok_ = false;
        cuda_fail(e, "cudaMallocAsync(scratch)", __FILE__, __LINE__);
        return nullptr;
    }
    ptrs_[n_++] = p;
    return p;
}

Scratch::~Scratch() {
    for (int i = n_ - 1; i >= 0; --i) cudaFreeAsync(ptrs_[i], stream_);
}

}  // namespace b200

extern "C" {

const char* b200mnn_last_error(void) { return b200::g_last_error.c_str(); }
int b200mnn_version(void) { return 100; }
int64_t b200mnn_launch_count(void) { return (int64_t)b200::g_launches.load(); }

int b200mnn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int b200mnn_set_device(int device) {
    B200_CUDA(cudaSetDevice(device));
    return 0;
}

}  // extern "C"
