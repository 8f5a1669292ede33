// PTX wrappers shared by the tcgen05 kernels of libb200mnn (mbarrier, TMA, TMEM, UMMA descriptors).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

namespace b200 {
namespace tc {

// ------------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier, TMA, tcgen05)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifndef B200_WAIT_TESTWAIT
#define B200_WAIT_TESTWAIT 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if B200_WAIT_TESTWAIT
    // non-blocking probe: the warp never parks inside the shared-memory pipe; callers back off with nanosleep
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok) __nanosleep(B200_WAIT_TESTWAIT);
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// Spin with a watchdog: a protocol bug must surface as a trapped kernel (-> CUDA error -> B200MNN_ECUDA),
// never as a hung GPU.
#ifndef B200_SPIN_BACKOFF
#define B200_SPIN_BACKOFF 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        ++spins;
        // long waits (the producer and the MMA issuers while the epilogue works through a hit-heavy tile) back off, so
        // that the polling loop does not take issue slots from the epilogue warps of the same scheduler
        if (B200_SPIN_BACKOFF && spins > 16u) __nanosleep(B200_SPIN_BACKOFF);
        if ((spins & 0x3FFu) == 0 && clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz
            printf("b200mnn: mbarrier watchdog fired (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// Warp-uniform wait for code executed by all 32 lanes.  Lanes polling on their own see the phase flip in different
// iterations and leave the loop at different times; on sm_100 the warp then stays split, and every later
// __shfl_sync/__ballot_sync/__syncwarp takes the compiler's divergent fallback (WARPSYNC.COLLECTIVE per shuffle --
// measured: a 45-shuffle row merge took 8-11k cycles instead of 1.1k, a quiet chunk scan 3x longer).  Here the vote
// makes every lane leave in the same iteration, so the warp never diverges.
__device__ __forceinline__ void mbar_wait_u(uint32_t bar, uint32_t parity) {
    if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        ++spins;
        if (B200_SPIN_BACKOFF && spins > 16u) __nanosleep(B200_SPIN_BACKOFF);
        if ((spins & 0x3FFu) == 0 && clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz
            if ((threadIdx.x & 31) == 0)
                printf("b200mnn: mbarrier watchdog fired (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// Waits of the helper warps (TMA producer, MMA issuers) of a kernel whose epilogue warps are issue-bound: try_wait with
// a suspend-time hint parks the warp in hardware until the phase completes (or the hint expires), so the polling loop
// does not take issue slots from the epilogue warps of the same scheduler (measured in the kNN kernel: 20 % of all
// issued instructions were these loops).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// (Measured at the end of round 2: with the final epilogue the plain polling waits are 1 % (pruned) to 3.5 % (dense scan) faster
// again, so they are the default; -DB200_NO_PARK=0 brings the hinted waits back.)
#ifndef B200_NO_PARK
#define B200_NO_PARK 1
#endif
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
    if (B200_NO_PARK) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 20000000000LL) {
            printf("b200mnn: mbarrier watchdog fired (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
#ifndef B200_PARK_NS
#define B200_PARK_NS 0
#endif
__device__ __forceinline__ void mbar_wait_parked_u(uint32_t bar, uint32_t parity) {
    if (B200_NO_PARK) { mbar_wait_u(bar, parity); return; }
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!__all_sync(0xffffffffu, mbar_try_wait_hint(bar, parity, 20000u))) {
        if (B200_PARK_NS) __nanosleep(B200_PARK_NS);
        if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 20000000000LL) {
            if ((threadIdx.x & 31) == 0)
                printf("b200mnn: mbarrier watchdog fired (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// Elects one lane of a fully converged warp (warp-uniform control flow keeps descriptors in uniform registers).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// Multicast variant: the box lands at the same CTA-relative shared-memory offset in every CTA of `cta_mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16, fp32 accumulate, issued by one thread.  ACC is a compile-time flag so
// the issue path carries no predicate arithmetic.
template <bool ACC>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Same, arriving on the barrier at this offset in every CTA of `cta_mask` (slot recycling under TMA multicast).
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives columns [col, col+32) of lane base+i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// 16 lanes x (2 x 32) consecutive fp32 columns: thread i < 16 receives columns [col, col+32) of lane base+i, thread
// i >= 16 columns [col+32, col+64) of lane base+(i-16) (cute: SM100_TMEM_LOAD_16dp32b32x).  Lets the two epilogue warps
// of a TMEM lane quarter split its ROWS (16 each) instead of the columns.
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x32bx2.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32], 32;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64).
// Rows are 128 B apart, 8-row groups 1024 B apart (dense [rows][128 B] tile as written by TMA).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0,
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

}  // namespace tc
}  // namespace b200
