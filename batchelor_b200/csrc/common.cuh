// Shared helpers for libb200mnn: error plumbing, stream-ordered scratch memory, small PTX wrappers.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/b200mnn.h"

namespace b200 {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define B200_CUDA(expr)                                                         \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) return ::b200::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define B200_TRY(expr)            \
    do {                          \
        int _rc = (expr);         \
        if (_rc != 0) return _rc; \
    } while (0)

// Checks the launch that just happened (configuration errors surface here, execution errors at the next sync) and
// counts it (b200mnn_launch_count: what bench.py reports as gpu_launches).
void count_launch();
#define B200_LAUNCH_CHECK()             \
    do {                                \
        ::b200::count_launch();         \
        B200_CUDA(cudaGetLastError());  \
    } while (0)

int ensure_device();  // B200MNN_ECUDA if no usable device

// Stream-ordered scratch allocations that free themselves (cudaFreeAsync) at scope exit.
class Scratch {
public:
    explicit Scratch(cudaStream_t s) : stream_(s) {}
    ~Scratch();
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    // Returns nullptr and records the error on failure.
    void* alloc(size_t bytes);
    template <typename T>
    T* get(size_t count) { return static_cast<T*>(alloc(count * sizeof(T))); }
    bool ok() const { return ok_; }

private:
    cudaStream_t stream_;
    static const int kMax = 64;
    void* ptrs_[kMax];
    int n_ = 0;
    bool ok_ = true;
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int sm_count();

}  // namespace b200
