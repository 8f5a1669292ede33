// Microbenchmark (B200): tcgen05.ld throughput by shape / warps / outstanding loads, and tcgen05.mma issue rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tmem tools/ubench_tmem.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, v)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18," \
                 "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                            \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),      \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),     \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                   \
                 : "r"(taddr)                                                                                                \
                 : "memory")

#define LD16(taddr, v)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"   \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                    \
                 : "r"(taddr)                                                                                                \
                 : "memory")

// 16 lanes x 256 bit: each repeat = 8 columns x 16 lanes; .x8 -> 64 columns x 16 lanes = 32 regs per thread
#define LD16x256(taddr, v)                                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18," \
                 "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                            \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),      \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),     \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                   \
                 : "r"(taddr)                                                                                                \
                 : "memory")

#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// mode 0: x32, wait after every load.  1: x32, two in flight.  2: x32, four in flight.  3: x16 wait each.  4: 16x256b.x8 wait each
template <int MODE>
__global__ void __launch_bounds__(256, 1) ld_kernel(int iters, unsigned long long* out_cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) { uint32_t v[32]; LD32(base + c * 32, v); WAIT_LD(); acc += v[lane & 31 ? 1 : 0] ^ v[31]; }
        } else if (MODE == 1) {
#pragma unroll 1
            for (int c = 0; c < 8; c += 2) { uint32_t v[32], w[32]; LD32(base + c * 32, v); LD32(base + c * 32 + 32, w); WAIT_LD(); acc += v[0] ^ w[31]; }
        } else if (MODE == 2) {
#pragma unroll 1
            for (int c = 0; c < 8; c += 4) {
                uint32_t v[32], w[32], x[32], y[32];
                LD32(base + c * 32, v); LD32(base + c * 32 + 32, w); LD32(base + c * 32 + 64, x); LD32(base + c * 32 + 96, y);
                WAIT_LD();
                acc += v[0] ^ w[31] ^ x[5] ^ y[7];
            }
        } else if (MODE == 3) {
#pragma unroll 1
            for (int c = 0; c < 16; ++c) { uint32_t v[16]; LD16(base + c * 16, v); WAIT_LD(); acc += v[0] ^ v[15]; }
        } else if (MODE == 4) {
            // each 16x256b.x8 covers 16 lanes x 64 columns; two per 32 lanes
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32], w[32];
                LD16x256(base + c * 64, v);
                LD16x256(base + (16u << 16) + c * 64, w);
                WAIT_LD();
                acc += v[0] ^ w[31];
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512));
}

// ---- MMA issue rate: M=128, N=256, K=16 kind::f16 from (garbage) smem, NMMA per commit, one thread issues ----
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, int nmma, int nval, unsigned long long* out_cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nval >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (warp == 1 && lane == 0) {
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem) + 16384);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int m = 0; m < nmma; ++m) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase + (it & 1) * 256),
                             "l"(da + (uint64_t)((m & 3) * 2)), "l"(db + (uint64_t)((m & 3) * 2)), "r"(idesc), "r"(m > 0 ? 1u : 0u)
                             : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[it & 1])) : "memory");
            if (it >= 1) {  // wait for the previous batch: at most two batches in flight, no barrier is ever over-run
                const int pit = it - 1;
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar[pit & 1])), "r"((uint32_t)((pit >> 1) & 1)) : "memory");
            }
        }
        {
            const int pit = iters - 1;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[pit & 1])), "r"((uint32_t)((pit >> 1) & 1)) : "memory");
        }
        t1 = clock64();
        out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512));
}

// ---- contention: warp 1 issues MMAs (N=256, accumulate into columns 0..255) while warps 2..5 read columns 256..511 ----
__global__ void __launch_bounds__(192, 1) both_kernel(int iters, int do_mma, int do_ld, int swap, unsigned long long* out_cycles, uint32_t* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1 && lane == 0 && do_mma) {
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem) + 16384);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int m = 0; m < 10; ++m)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase + (swap ? 256 : 0)),
                             "l"(da + (uint64_t)((m & 3) * 2)), "l"(db + (uint64_t)((m & 3) * 2)), "r"(idesc), "r"(m > 0 ? 1u : 0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[it & 1])) : "memory");
            if (it >= 1) {
                const int pit = it - 1; uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar[pit & 1])), "r"((uint32_t)((pit >> 1) & 1)) : "memory");
            }
        }
        { const int pit = iters - 1; uint32_t ok = 0;
          while (!ok)
              asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                           : "=r"(ok) : "r"(smem_u32(&bar[pit & 1])), "r"((uint32_t)((pit >> 1) & 1)) : "memory"); }
        out_cycles[blockIdx.x * 2] = (unsigned long long)(clock64() - t0);
    }
    if (warp >= 2 && do_ld) {
        const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (swap ? 0 : 256);
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) { uint32_t v[32]; LD32(base + c * 32, v); WAIT_LD(); acc += v[1] ^ v[31]; }
        }
        if (threadIdx.x == 64) out_cycles[blockIdx.x * 2 + 1] = (unsigned long long)(clock64() - t0);
        if (acc == 0x12345678u) sink[0] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512));
}

// ---- commit pattern: per "tile" 10 MMAs split as `split` + rest with `ncommit` commits, waiting on the tile two back ----
__global__ void __launch_bounds__(128, 1) pattern_kernel(int iters, int split, int ncommit, int nval, int nstage, unsigned long long* out_cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar[8];
    __shared__ uint64_t junk[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1));
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&junk[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nval >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1 && lane == 0) {
        const uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem) + 16384);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (it >= nstage) {  // accumulator stage free? (tile `nstage` back completed)
                const int pit = it - nstage; uint32_t ok = 0;
                while (!ok)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar[pit % nstage])), "r"((uint32_t)((pit / nstage) & 1)) : "memory");
            }
            for (int m = 0; m < 10; ++m) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase + (it % nstage) * nval),
                             "l"(da + (uint64_t)((m & 3) * 2)), "l"(db + (uint64_t)((m & 3) * 2)), "r"(idesc), "r"(m > 0 ? 1u : 0u) : "memory");
                if (m + 1 == split && ncommit >= 2)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&junk[0])) : "memory");
            }
            if (ncommit >= 3)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&junk[1])) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[it % nstage])) : "memory");
        }
        for (int pit = iters - nstage; pit < iters; ++pit) {
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[pit % nstage])), "r"((uint32_t)((pit / nstage) & 1)) : "memory");
        }
        out_cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512));
}

template <int MODE>
static void run_ld(const char* name, int threads, int grid) {
    unsigned long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, sizeof(unsigned long long) * grid); cudaMalloc(&sink, 64);
    const int iters = 2000;
    ld_kernel<MODE><<<grid, threads>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    unsigned long long h[1024]; cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
    const double warps = threads / 32.0;
    const double bytes = (double)iters * 8 * 32 * 32 * 4 * warps;  // each warp reads 32 lanes x 256 columns per iteration
    printf("%-34s threads=%3d grid=%3d : %8.1f cycles/iter, %6.1f B/cycle/SM\n", name, threads, grid, (double)h[0] / iters, bytes / (double)h[0]);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaSetDevice(0);
    for (int threads : {128, 256}) {
        run_ld<0>("32x32b.x32 wait-each", threads, 148);
        run_ld<1>("32x32b.x32 2 in flight", threads, 148);
        run_ld<2>("32x32b.x32 4 in flight", threads, 148);
        run_ld<3>("32x32b.x16 wait-each", threads, 148);
        run_ld<4>("16x256b.x8 pair wait", threads, 148);
    }
    unsigned long long* cyc; cudaMalloc(&cyc, sizeof(unsigned long long) * 148);
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int nval : {256, 128}) {
        for (int nmma : {1, 4, 10, 32}) {
            const int iters = 2000;
            mma_kernel<<<148, 128, 64 * 1024>>>(iters, nmma, nval, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mma: %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("mma M=128 N=%3d K=16: %2d per commit : %7.1f cycles per MMA\n", nval, nmma, (double)h / iters / nmma);
        }
    }
    {
        unsigned long long* c2; uint32_t* sink; cudaMalloc(&c2, sizeof(unsigned long long) * 2 * 148); cudaMalloc(&sink, 64);
        cudaFuncSetAttribute(both_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        const int iters = 2000;
        for (int cfg = 0; cfg < 4; ++cfg) {
            const int do_mma = cfg != 1, do_ld = cfg != 0, swap = cfg == 3;
            cudaMemset(c2, 0, sizeof(unsigned long long) * 2 * 148);
            both_kernel<<<148, 192, 64 * 1024>>>(iters, do_mma, do_ld, swap, c2, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("both: %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long h[2]; cudaMemcpy(h, c2, sizeof(h), cudaMemcpyDeviceToHost);
            printf("concurrent mma=%d ld=%d swap=%d : %7.1f cycles per 10-MMA tile, %7.1f cycles per 128KB TMEM readout\n", do_mma, do_ld, swap,
                   (double)h[0] / iters, (double)h[1] / iters);
        }
    }
    {
        unsigned long long* c3; cudaMalloc(&c3, sizeof(unsigned long long) * 148);
        cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        const int iters = 2000;
        const int cfgs[][2] = {{256, 2}, {128, 4}, {128, 3}, {128, 2}, {64, 8}, {64, 4}, {256, 1}, {128, 1}};
        for (auto& c : cfgs) {
            pattern_kernel<<<148, 128, 64 * 1024>>>(iters, 6, 1, c[0], c[1], c3);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("pattern: %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long h; cudaMemcpy(&h, c3, sizeof(h), cudaMemcpyDeviceToHost);
            printf("pattern N=%3d, %d accumulator stages, 10 MMAs/tile : %7.1f cycles per tile = %7.1f per 256 columns\n", c[0], c[1],
                   (double)h / iters, (double)h / iters * 256.0 / c[0]);
        }
    }
    return 0;
}
