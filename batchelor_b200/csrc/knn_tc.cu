// Exact k-nearest-neighbour search for sm_100a -- replaces BiocNeighbors::queryKNN(..., BNPARAM=KmknnParam())
// as batchelor calls it (R/MNN_tree.R:129 inside findMutualNN, R/fastMNN.R:605 for the tricube search).
//
// Contract (same as the exact KMKNN search): for every query the k reference rows with the smallest
// (squared Euclidean distance accumulated in double in dimension order, row index) pairs, ascending.
//
// Pipeline (all on one stream, no host synchronisation):
//   K0 absmax         : one power-of-two scale S so that every |x*S| < 2^15 (fp16 range, lo parts stay normal)
//   K1 prep_operand   : fp64 row -> fp16 "hi" and "lo" K-slices (x*S = hi + lo + O(2^-22)), fp32 scaled norm,
//                       fp64 norm.  Operand rows are K-major, 64 fp16 (128 B) per TMA box, SWIZZLE_128B.
//   K2 knn_candidates : tcgen05 kernel.  score(q,j) = ||x_j||^2 - 2 q.x_j with q.x_j ~ qh.xh + ql.xh + qh.xl
//                       (three fp16 MMAs per 16 dims, fp32 accumulation in TMEM; the remainder dims are packed
//                       as one virtual slice).  One CTA = 128 queries (TMEM lanes) x a stream of 256-reference
//                       tiles; TMA producer warp, single-thread MMA issuer, 4 epilogue warps that read the
//                       accumulators straight out of TMEM (tcgen05.ld), compare against a per-query running
//                       threshold held in a register and keep the 32*E best approximate scores per query in a
//                       warp-cooperative sorted list in shared memory.
//   K3 rerank         : exact fp64 distances of the <= nsplit*32*E candidates (same summation order as the
//                       reference arithmetic), (distance, index) selection of the k best, and a CERTIFICATE:
//                       the result is provably exact if d2_k < (smallest retained threshold) - eps, where eps
//                       bounds the fp16x3/fp32 scoring error.  Uncertified queries are flagged.
//   K4 rescue         : flagged queries are recomputed by an exact fp64 scan of all references.  This is also
//                       the generic path for shapes the tensor path does not cover (k > 56, d > 192).
#include "common.cuh"

#include <cuda.h>

#include <cmath>
#include <cstring>
#include <mutex>

namespace b200 {
namespace knn {

// ------------------------------------------------------------------------------------------------
// Tiling constants
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128;               // queries per CTA (TMEM lanes)
constexpr int BN = 256;               // references per tile (TMEM columns per accumulator stage)
constexpr int KBOX = 64;              // fp16 columns per TMA box (128 B, one swizzle-128B atom row)
constexpr int SLICE = 16;             // K of one tcgen05.mma kind::f16
constexpr int A_BOX_BYTES = BM * 128; // 16 KB
constexpr int B_BOX_BYTES = BN * 128; // 32 KB
constexpr int MAX_NBOX = 6;
constexpr int MAX_MMA_PER_BOX = 6;
constexpr int MAX_SLOTS = 8;
constexpr int NUM_THREADS = 192;      // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr int TMEM_COLS = 512;        // two accumulator stages of BN columns
constexpr int MAX_K_TENSOR = 56;      // k <= 32*E - 8
constexpr int MAX_SPLIT = 8;

struct MmaSched {
    int nbox;
    int nmma[MAX_NBOX];
    unsigned char a_slice[MAX_NBOX][MAX_MMA_PER_BOX];  // stored slice of A (absolute)
    unsigned char b_sub[MAX_NBOX][MAX_MMA_PER_BOX];    // stored slice of B inside the current box (0..3)
};

// Stored K layout of one operand row (both operands): for every full group g of 16 dims the "hi" slice at
// 2g and the "lo" slice at 2g+1; then, if r = d % 16 dims remain and 3r <= 16, one packed remainder slice
//   query row : [qh(r) | qh(r) | ql(r) | 0]      reference row : [xh(r) | xl(r) | xh(r) | 0]
// whose single MMA yields qh.xh + qh.xl + ql.xh for those dims.  If 3r > 16 the remainder is zero-padded into
// one more full group.
struct KLayout {
    int d, groups, rem, nslices, nbox;
};

static KLayout make_layout(int d) {
    KLayout L;
    L.d = d;
    L.groups = d / SLICE;
    L.rem = d % SLICE;
    if (L.rem * 3 > SLICE) { L.groups += 1; L.rem = 0; }
    L.nslices = 2 * L.groups + (L.rem ? 1 : 0);
    L.nbox = (L.nslices + 3) / 4;
    return L;
}

static MmaSched make_sched(const KLayout& L) {
    MmaSched s;
    memset(&s, 0, sizeof(s));
    s.nbox = L.nbox;
    for (int g = 0; g < L.groups; ++g) {
        const int sh = 2 * g, sl = 2 * g + 1, box = sh / 4;
        int& m = s.nmma[box];
        s.a_slice[box][m] = (unsigned char)sh; s.b_sub[box][m] = (unsigned char)(sh % 4); ++m;  // qh.xh
        s.a_slice[box][m] = (unsigned char)sl; s.b_sub[box][m] = (unsigned char)(sh % 4); ++m;  // ql.xh
        s.a_slice[box][m] = (unsigned char)sh; s.b_sub[box][m] = (unsigned char)(sl % 4); ++m;  // qh.xl
    }
    if (L.rem) {
        const int sr = 2 * L.groups, box = sr / 4;
        int& m = s.nmma[box];
        s.a_slice[box][m] = (unsigned char)sr; s.b_sub[box][m] = (unsigned char)(sr % 4); ++m;
    }
    return s;
}

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
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spin with a watchdog: a protocol bug must surface as a trapped kernel (-> CUDA error -> B200MNN_ECUDA),
// never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFu) == 0 && clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz
            printf("b200mnn: mbarrier watchdog fired (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
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

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16, fp32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__host__ __device__ constexpr uint32_t make_idesc() { return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

__device__ __forceinline__ float entry_score(unsigned long long e) { return __uint_as_float((uint32_t)(e >> 32)); }
__device__ __forceinline__ unsigned long long make_entry(float s, int idx) {
    return ((unsigned long long)__float_as_uint(s) << 32) | (uint32_t)idx;
}

// ------------------------------------------------------------------------------------------------
// K2: tensor-core candidate scoring + per-query top-(32*E) of the approximate scores
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(NUM_THREADS, 1)
knn_candidates_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const float* __restrict__ norms,  // [n_pad] scaled ||x||^2, +inf on padding
                      const MmaSched sched, const int nslot, const int64_t nq, const int ntiles, const int tiles_per_split,
                      int32_t* __restrict__ cand_idx,   // [nsplit][nq][32E]
                      float* __restrict__ cand_score,   // [nsplit][nq][32E] (may be null)
                      float* __restrict__ thr_out)      // [nsplit][nq]
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nbox = sched.nbox;

    uint8_t* smA = smem;
    uint8_t* smB = smA + (size_t)nbox * A_BOX_BYTES;
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(smB + (size_t)nslot * B_BOX_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(lists + BM * 32 * E);
    uint64_t* full = bars;                  // [MAX_SLOTS]
    uint64_t* empty = bars + MAX_SLOTS;     // [MAX_SLOTS]
    uint64_t* afull = bars + 2 * MAX_SLOTS; // [1]
    uint64_t* tfull = afull + 1;            // [2]
    uint64_t* tempty = tfull + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tile0 = blockIdx.y * tiles_per_split;
    const int tile1 = min(ntiles, tile0 + tiles_per_split);
    const int my_tiles = tile1 - tile0;
    const int m0 = blockIdx.x * BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < nslot; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        mbar_init(smem_u32(afull), 1);
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull[i]), 1); mbar_init(smem_u32(&tempty[i]), 4); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(afull), (uint32_t)(nbox * A_BOX_BYTES));
            for (int b = 0; b < nbox; ++b) tma_load_2d(smem_u32(smA + (size_t)b * A_BOX_BYTES), &tmA, smem_u32(afull), b * KBOX, m0);
            int slot = 0;
            uint32_t phase = 0;
            for (int t = tile0; t < tile1; ++t) {
                for (int b = 0; b < nbox; ++b) {
                    mbar_wait(smem_u32(&empty[slot]), phase ^ 1);
                    mbar_arrive_expect_tx(smem_u32(&full[slot]), (uint32_t)B_BOX_BYTES);
                    tma_load_2d(smem_u32(smB + (size_t)slot * B_BOX_BYTES), &tmB, smem_u32(&full[slot]), b * KBOX, t * BN);
                    if (++slot == nslot) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc();
            mbar_wait(smem_u32(afull), 0);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smA);
            const uint32_t b_base = smem_u32(smB);
            int slot = 0;
            uint32_t phase = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int stage = tl & 1;
                const uint32_t use = (uint32_t)(tl >> 1);
                mbar_wait(smem_u32(&tempty[stage]), (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(stage * BN);
                uint32_t acc = 0;
                for (int b = 0; b < nbox; ++b) {
                    mbar_wait(smem_u32(&full[slot]), phase);
                    tc_fence_after();
                    const uint32_t b_addr = b_base + (uint32_t)slot * B_BOX_BYTES;
                    const int nm = sched.nmma[b];
                    for (int m = 0; m < nm; ++m) {
                        const int as = sched.a_slice[b][m];
                        const uint64_t da = make_sw128_desc(a_base + (uint32_t)(as >> 2) * A_BOX_BYTES + (uint32_t)(as & 3) * 32u);
                        const uint64_t db = make_sw128_desc(b_addr + (uint32_t)sched.b_sub[b][m] * 32u);
                        umma_f16(tmem_d, da, db, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(smem_u32(&empty[slot]));  // frees the smem slot once these MMAs retire
                    if (++slot == nslot) { slot = 0; phase ^= 1; }
                }
                umma_commit(smem_u32(&tfull[stage]));      // accumulator stage complete
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> threshold filter -> sorted lists =====================
        const int q4 = warp & 3;  // TMEM lane quarter this warp may access
        unsigned long long* mylist = lists + (size_t)(q4 * 32) * (32 * E);
        const unsigned long long empty_entry = make_entry(__int_as_float(0x7f800000), -1);
#pragma unroll 1
        for (int r = 0; r < 32; ++r)
#pragma unroll
            for (int e = 0; e < E; ++e) mylist[(r * E + e) * 32 + lane] = empty_entry;
        float thr = __int_as_float(0x7f800000);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);

        for (int tl = 0; tl < my_tiles; ++tl) {
            const int stage = tl & 1;
            const uint32_t use = (uint32_t)(tl >> 1);
            mbar_wait(smem_u32(&tfull[stage]), use & 1);
            tc_fence_after();
            const int colbase = (tile0 + tl) * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(lane_base + (uint32_t)(stage * BN + c * 32), v);
                tmem_ld_wait();
                if (c == BN / 32 - 1) {  // every column of this stage is now in registers: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tempty[stage]));
                }
                const float4* np4 = reinterpret_cast<const float4*>(norms + colbase + c * 32);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 na = __ldg(np4 + 2 * g), nb = __ldg(np4 + 2 * g + 1);
                    float s[8];
                    s[0] = fmaf(__uint_as_float(v[8 * g + 0]), -2.f, na.x);
                    s[1] = fmaf(__uint_as_float(v[8 * g + 1]), -2.f, na.y);
                    s[2] = fmaf(__uint_as_float(v[8 * g + 2]), -2.f, na.z);
                    s[3] = fmaf(__uint_as_float(v[8 * g + 3]), -2.f, na.w);
                    s[4] = fmaf(__uint_as_float(v[8 * g + 4]), -2.f, nb.x);
                    s[5] = fmaf(__uint_as_float(v[8 * g + 5]), -2.f, nb.y);
                    s[6] = fmaf(__uint_as_float(v[8 * g + 6]), -2.f, nb.z);
                    s[7] = fmaf(__uint_as_float(v[8 * g + 7]), -2.f, nb.w);
                    const bool hit = (s[0] < thr) | (s[1] < thr) | (s[2] < thr) | (s[3] < thr) | (s[4] < thr) | (s[5] < thr) |
                                     (s[6] < thr) | (s[7] < thr);
                    if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            unsigned ballot = __ballot_sync(0xffffffffu, s[i] < thr);
                            const int col = colbase + c * 32 + g * 8 + i;
                            while (ballot) {  // warp-uniform: one cooperative insertion per hitting row
                                const int r = __ffs(ballot) - 1;
                                ballot &= ballot - 1;
                                const float sr = __shfl_sync(0xffffffffu, s[i], r);
                                const unsigned long long ne = make_entry(sr, col);
                                float newthr;
                                if (E == 1) {
                                    const unsigned long long cur = mylist[r * 32 + lane];
                                    const bool gt = entry_score(cur) > sr;
                                    const unsigned long long up = __shfl_up_sync(0xffffffffu, cur, 1);
                                    const int pos = __ffs(__ballot_sync(0xffffffffu, gt)) - 1;
                                    const unsigned long long nw = lane < pos ? cur : (lane == pos ? ne : up);
                                    mylist[r * 32 + lane] = nw;
                                    newthr = entry_score(__shfl_sync(0xffffffffu, nw, 31));
                                } else {
                                    const unsigned long long a = mylist[(r * 2 + 0) * 32 + lane];  // sorted position 2*lane
                                    const unsigned long long b = mylist[(r * 2 + 1) * 32 + lane];  // sorted position 2*lane+1
                                    const bool ga = entry_score(a) > sr, gb = entry_score(b) > sr;
                                    const unsigned long long pb = __shfl_up_sync(0xffffffffu, b, 1);
                                    const bool gp = lane > 0 && entry_score(pb) > sr;
                                    const unsigned long long na2 = !ga ? a : (gp ? pb : ne);
                                    const unsigned long long nb2 = !gb ? b : (ga ? a : ne);
                                    mylist[(r * 2 + 0) * 32 + lane] = na2;
                                    mylist[(r * 2 + 1) * 32 + lane] = nb2;
                                    newthr = entry_score(__shfl_sync(0xffffffffu, nb2, 31));
                                }
                                if (lane == r) thr = newthr;
                            }
                        }
                    }
                }
            }
        }
        // write the retained candidates: row r of this warp, 32*E entries, coalesced
        const int64_t rowbase = (int64_t)m0 + q4 * 32;
        const int64_t sbase = (int64_t)blockIdx.y * nq;
#pragma unroll 1
        for (int r = 0; r < 32; ++r) {
            const int64_t row = rowbase + r;
            if (row < nq) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long en = mylist[(r * E + e) * 32 + lane];
                    const int64_t o = (sbase + row) * (32 * E) + (E == 1 ? lane : lane * 2 + e);
                    cand_idx[o] = (int32_t)(uint32_t)en;
                    if (cand_score) cand_score[o] = entry_score(en);
                }
            }
        }
        if (rowbase + lane < nq) thr_out[sbase + rowbase + lane] = thr;
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// K0: absmax of an fp64 array (non-negative floats order like their bit patterns)
// ------------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const double* __restrict__ x, int64_t count, unsigned int* __restrict__ out_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = fabsf((float)x[i]);
        m = (a > m) ? a : m;  // NaN never wins
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) {
        // round up one ulp so the double value can never exceed it
        atomicMax(out_bits, __float_as_uint(m) + 1u);
    }
}

// scale exponent e such that |x| * 2^e < 2^15 for all x; stored as int
__global__ void scale_kernel(const unsigned int* __restrict__ absmax_bits, int* __restrict__ scale_exp) {
    const float m = __uint_as_float(*absmax_bits);
    int e = 0;
    if (m > 0.f && isfinite(m)) {
        int ex;
        frexpf(m, &ex);  // m = f * 2^ex, f in [0.5, 1)  =>  m < 2^ex
        e = 15 - ex;
    }
    *scale_exp = e;
}

// ------------------------------------------------------------------------------------------------
// K1: operand preparation.  One thread per (padded) row.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_half(double xs, __half& hi, __half& lo) {
    hi = __double2half(xs);
    lo = __double2half(xs - (double)__half2float(hi));
}

template <bool IS_QUERY>
__global__ void prep_operand_kernel(const double* __restrict__ X, int64_t n, int64_t n_pad, int d, KLayout L,
                                    const int* __restrict__ scale_exp, __half* __restrict__ op, float* __restrict__ norm_f32,
                                    double* __restrict__ norm_f64, unsigned long long* __restrict__ maxnorm_bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const int KS = L.nbox * KBOX;
    uint4* row = reinterpret_cast<uint4*>(op + i * KS);
    if (i >= n) {
        for (int c = 0; c < KS / 8; ++c) row[c] = make_uint4(0, 0, 0, 0);
        if (norm_f32) norm_f32[i] = __int_as_float(0x7f800000);
        return;
    }
    const double S = scalbn(1.0, *scale_exp);
    const double* x = X + i * d;
    double nrm = 0.0;
    __align__(16) __half hbuf[16];
    __align__(16) __half lbuf[16];
    for (int g = 0; g < L.groups; ++g) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int t = g * 16 + j;
            const double xv = (t < d) ? x[t] : 0.0;
            nrm = __dadd_rn(nrm, __dmul_rn(xv, xv));
            split_half(xv * S, hbuf[j], lbuf[j]);
        }
        const uint4* hp = reinterpret_cast<const uint4*>(hbuf);
        const uint4* lp = reinterpret_cast<const uint4*>(lbuf);
        row[(2 * g) * 2 + 0] = hp[0];
        row[(2 * g) * 2 + 1] = hp[1];
        row[(2 * g + 1) * 2 + 0] = lp[0];
        row[(2 * g + 1) * 2 + 1] = lp[1];
    }
    int written = 2 * L.groups;
    if (L.rem) {
#pragma unroll
        for (int j = 0; j < 16; ++j) hbuf[j] = __float2half(0.f);
        const int r = L.rem;
        for (int j = 0; j < r; ++j) {
            const double xv = x[L.groups * 16 + j];
            nrm = __dadd_rn(nrm, __dmul_rn(xv, xv));
            __half h, l;
            split_half(xv * S, h, l);
            hbuf[j] = h;
            if (IS_QUERY) { hbuf[r + j] = h; hbuf[2 * r + j] = l; }
            else          { hbuf[r + j] = l; hbuf[2 * r + j] = h; }
        }
        const uint4* hp = reinterpret_cast<const uint4*>(hbuf);
        row[written * 2 + 0] = hp[0];
        row[written * 2 + 1] = hp[1];
        ++written;
    }
    for (int s = written; s < L.nbox * 4; ++s) { row[s * 2 + 0] = make_uint4(0, 0, 0, 0); row[s * 2 + 1] = make_uint4(0, 0, 0, 0); }
    if (norm_f32) norm_f32[i] = (float)(nrm * S * S);
    if (norm_f64) norm_f64[i] = nrm;
    if (maxnorm_bits) {
        // warp-aggregate then one atomic (non-negative doubles order like their bit patterns)
        unsigned long long b = (unsigned long long)__double_as_longlong(nrm);
        if (!(nrm >= 0.0)) b = 0;  // NaN
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(__activemask(), b, o);
            b = ob > b ? ob : b;
        }
        if ((threadIdx.x & 31) == 0) atomicMax(maxnorm_bits, b);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: exact fp64 re-rank + certificate.  One warp per query, one candidate per lane per round.
// ------------------------------------------------------------------------------------------------
constexpr int RR_WARPS = 4;
constexpr int RR_DCH = 32;  // dims staged per chunk
constexpr int RR_MAXROUNDS = MAX_SPLIT * 2;

__device__ __forceinline__ bool pair_less(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

__global__ void __launch_bounds__(RR_WARPS * 32)
rerank_kernel(const double* __restrict__ X, const double* __restrict__ Q, int64_t nq, int d, int k,
              const int32_t* __restrict__ cand_idx, const float* __restrict__ thr, int nsplit, int ncand_per_split,
              const int* __restrict__ scale_exp, const double* __restrict__ qnorm, const unsigned long long* __restrict__ maxnorm_bits,
              int32_t* __restrict__ out_idx, double* __restrict__ out_dist, int* __restrict__ flag_count, int32_t* __restrict__ flag_list,
              double* __restrict__ dbg_d2 /* [nq][ncand] or null */) {
    __shared__ double stage[RR_WARPS][32][RR_DCH + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * RR_WARPS + warp;
    if (q >= nq) return;
    const int ncand = nsplit * ncand_per_split;
    const int rounds = ncand / 32;
    double cd[RR_MAXROUNDS];
    int ci[RR_MAXROUNDS];
#pragma unroll
    for (int r = 0; r < RR_MAXROUNDS; ++r) { cd[r] = INFINITY; ci[r] = -1; }
    const double* qv = Q + q * d;

#pragma unroll
    for (int r = 0; r < RR_MAXROUNDS; ++r) {
        if (r < rounds) {
            const int c = r * 32 + lane;
            const int sp = c / ncand_per_split, within = c % ncand_per_split;
            const int id = cand_idx[((int64_t)sp * nq + q) * ncand_per_split + within];
            ci[r] = id;
            double acc = 0.0;
            for (int t0 = 0; t0 < d; t0 += RR_DCH) {
                const int len = min(RR_DCH, d - t0);
                // cooperative, coalesced staging of 32 candidate rows (chunk of dims) into shared memory
                for (int rr = 0; rr < 32; ++rr) {
                    const int rid = __shfl_sync(0xffffffffu, id, rr);
                    if (lane < len) stage[warp][rr][lane] = (rid >= 0) ? X[(int64_t)rid * d + t0 + lane] : 0.0;
                }
                __syncwarp();
                for (int t = 0; t < len; ++t) {
                    const double df = __dsub_rn(qv[t0 + t], stage[warp][lane][t]);
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
                __syncwarp();
            }
            cd[r] = (id >= 0) ? acc : INFINITY;
            if (dbg_d2) dbg_d2[q * ncand + c] = cd[r];
        }
    }
    // k rounds of warp arg-min under (distance, index)
    double dk = INFINITY;
    for (int j = 0; j < k; ++j) {
        double bd = INFINITY;
        int bi = 0x7fffffff, br = -1;
#pragma unroll
        for (int r = 0; r < RR_MAXROUNDS; ++r)
            if (r < rounds && ci[r] >= 0 && pair_less(cd[r], ci[r], bd, bi)) { bd = cd[r]; bi = ci[r]; br = r; }
        double wd = bd;
        int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, wd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (pair_less(od, oi, wd, wi)) { wd = od; wi = oi; }
        }
        if (wi == bi && br >= 0) {  // this lane owned the winner: retire it
#pragma unroll
            for (int r = 0; r < RR_MAXROUNDS; ++r)
                if (r == br) ci[r] = -1;
        }
        if (lane == 0) {
            out_idx[q * k + j] = (wi == 0x7fffffff) ? -1 : wi;
            if (out_dist) out_dist[q * k + j] = sqrt(wd);
        }
        dk = wd;
    }
    // certificate: every non-candidate j has score_j >= min_split thr, and |score_j/S^2 + ||q||^2 - d2_j| <= eps
    if (lane == 0) {
        float tmin = __int_as_float(0x7f800000);
        for (int sp = 0; sp < nsplit; ++sp) tmin = fminf(tmin, thr[(int64_t)sp * nq + q]);
        bool ok = true;
        if (tmin < __int_as_float(0x7f800000)) {
            const double inv = scalbn(1.0, -2 * (*scale_exp));
            const double qn = qnorm[q];
            const double M2 = __longlong_as_double((long long)*maxnorm_bits);
            const double eps = 1.52587890625e-05 * (sqrt(qn * M2) + M2);  // 2^-16 (|q| M + M^2)
            const double bound = (double)tmin * inv + qn - eps;
            ok = dk < bound;
        }
        if (!(ok)) {
            const int slot = atomicAdd(flag_count, 1);
            flag_list[slot] = (int32_t)q;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4: exact rescue / generic path.  One block per flagged query, k rounds of block arg-min over the
// pairs strictly greater than the previous pick.  O(k n d) per query -- only for rare queries.
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;

__global__ void __launch_bounds__(RS_THREADS)
rescue_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ Q, int d, int k,
              const int* __restrict__ flag_count, const int32_t* __restrict__ flag_list, int64_t nq_all,
              int32_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    __shared__ double sd[RS_THREADS / 32];
    __shared__ int si[RS_THREADS / 32];
    __shared__ double pick_d;
    __shared__ int pick_i;
    extern __shared__ double qs[];  // [d]
    const int count = flag_list ? *flag_count : (int)nq_all;
    for (int f = blockIdx.x; f < count; f += gridDim.x) {
        const int64_t q = flag_list ? flag_list[f] : f;
        __syncthreads();
        for (int t = threadIdx.x; t < d; t += blockDim.x) qs[t] = Q[q * d + t];
        if (threadIdx.x == 0) { pick_d = -1.0; pick_i = -1; }
        __syncthreads();
        for (int j = 0; j < k; ++j) {
            const double pd = pick_d;
            const int pi = pick_i;
            double bd = INFINITY;
            int bi = 0x7fffffff;
            for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
                const double* xv = X + r * d;
                double acc = 0.0;
                for (int t = 0; t < d; ++t) {
                    const double df = __dsub_rn(qs[t], xv[t]);
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
                const int ri = (int)r;
                const bool after = acc > pd || (acc == pd && ri > pi);  // strictly after the previous pick
                if (after && pair_less(acc, ri, bd, bi)) { bd = acc; bi = ri; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (pair_less(od, oi, bd, bi)) { bd = od; bi = oi; }
            }
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { sd[threadIdx.x >> 5] = bd; si[threadIdx.x >> 5] = bi; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double fd = sd[0];
                int fi = si[0];
                for (int w = 1; w < RS_THREADS / 32; ++w)
                    if (pair_less(sd[w], si[w], fd, fi)) { fd = sd[w]; fi = si[w]; }
                pick_d = fd;
                pick_i = fi;
                out_idx[q * k + j] = (fi == 0x7fffffff) ? -1 : fi;
                if (out_dist) out_dist[q * k + j] = sqrt(fd);
            }
            __syncthreads();
        }
    }
}

__global__ void write_stats_kernel(const int* __restrict__ flag_count, int64_t* __restrict__ stats, int64_t lists, int64_t path) {
    stats[0] = flag_count ? *flag_count : 0;
    stats[1] = lists;
    stats[2] = path;
    stats[3] = 0;
}

// squared distance debug view: cand score -> unscaled approximate squared distance
__global__ void debug_convert_kernel(const float* __restrict__ cand_score, const float* __restrict__ thr, const double* __restrict__ qnorm,
                                     const int* __restrict__ scale_exp, int64_t nq, int nsplit, int per,
                                     const int32_t* __restrict__ cand_idx, int32_t* __restrict__ out_idx,
                                     double* __restrict__ out_d2, double* __restrict__ out_thr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ncand = nsplit * per;
    if (i >= nq * ncand) return;
    const int64_t q = i / ncand;
    const int c = (int)(i % ncand), sp = c / per, w = c % per;
    const double inv = scalbn(1.0, -2 * (*scale_exp));
    const int64_t src = ((int64_t)sp * nq + q) * per + w;
    out_idx[i] = cand_idx[src];
    out_d2[i] = (double)cand_score[src] * inv + qnorm[q];
    if (c == 0) {
        float tmin = __int_as_float(0x7f800000);
        for (int s2 = 0; s2 < nsplit; ++s2) tmin = fminf(tmin, thr[(int64_t)s2 * nq + q]);
        out_thr[q] = (double)tmin * inv + qnorm[q];
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
    });
    return fn;
}

// rows x KS fp16, row-major; box = 64 columns x box_rows rows; 128-byte swizzle.
static int make_operand_map(CUtensorMap* map, const __half* base, int64_t rows, int KS, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(B200MNN_ECUDA, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    cuuint64_t dims[2] = {(cuuint64_t)KS, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)KS * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)KBOX, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return fail(B200MNN_ECUDA, buf);
    }
    return 0;
}

static size_t candidates_smem_bytes(int nbox, int nslot, int E) {
    return 1024 /* alignment slack */ + (size_t)nbox * A_BOX_BYTES + (size_t)nslot * B_BOX_BYTES + (size_t)BM * 32 * E * 8 +
           (2 * MAX_SLOTS + 5) * 8 + 16;
}

bool tensor_path_supported(int64_t n, int64_t nq, int d, int k) {
    if (d < 1 || k < 1 || n < 1 || nq < 1) return false;
    if (k > MAX_K_TENSOR) return false;
    if (n > (int64_t)INT32_MAX - 512 || nq > (int64_t)INT32_MAX - 512) return false;
    return make_layout(d).nbox <= MAX_NBOX;
}

struct DebugOut {
    int32_t* cand_idx = nullptr;  // [nq x ncand]
    double* cand_d2 = nullptr;    // approximate squared distances
    double* thr = nullptr;        // [nq]
    int64_t capacity = 0;         // entries available per query in the arrays above
    int64_t* ncand_host = nullptr;
};

int query_knn_device(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                     int64_t* d_stats, cudaStream_t stream, const DebugOut* dbg) {
    B200_TRY(ensure_device());
    if (n < 0 || nq < 0 || d < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (k < 0 || k > n) return fail(B200MNN_EINVAL, "'k' must be positive and no larger than the number of points in 'X'");
    if (nq == 0 || k == 0) return 0;
    if (n > (int64_t)INT32_MAX - 512 || nq > (int64_t)INT32_MAX - 512) return fail(B200MNN_EINVAL, "more than 2^31 points are not supported");
    Scratch ws(stream);

    if (!tensor_path_supported(n, nq, d, k) || d == 0) {
        if (dbg) return fail(B200MNN_EINVAL, "debug candidates requested for a shape outside the tensor path");
        // generic exact path: every query through the rescue kernel
        const int grid = (int)std::min<int64_t>(nq, (int64_t)sm_count() * 8);
        rescue_kernel<<<grid, RS_THREADS, (size_t)std::max(d, 1) * sizeof(double), stream>>>(dX, n, dQ, d, k, nullptr, nullptr, nq, d_idx, d_dist);
        B200_LAUNCH_CHECK();
        if (d_stats) { write_stats_kernel<<<1, 1, 0, stream>>>(nullptr, d_stats, 0, 0); B200_LAUNCH_CHECK(); }
        return 0;
    }

    const KLayout L = make_layout(d);
    const MmaSched sched = make_sched(L);
    const int KS = L.nbox * KBOX;
    const int E = (k <= 24) ? 1 : 2;
    const int per = 32 * E;
    const int64_t n_pad = round_up(n, BN), nq_pad = round_up(nq, BM);
    const int ntiles = (int)(n_pad / BN);
    const int mtiles = (int)(nq_pad / BM);
    int nsplit = 1;
    if (mtiles < 2 * sm_count()) nsplit = (int)std::min<int64_t>(std::min<int64_t>(MAX_SPLIT, ntiles), ceil_div(2 * sm_count(), mtiles));
    const int tiles_per_split = (int)ceil_div(ntiles, nsplit);
    nsplit = (int)ceil_div(ntiles, tiles_per_split);

    __half* opB = ws.get<__half>((size_t)n_pad * KS);
    __half* opA = ws.get<__half>((size_t)nq_pad * KS);
    float* norms = ws.get<float>((size_t)n_pad);
    double* qnorm = ws.get<double>((size_t)nq_pad);
    int32_t* cand_idx = ws.get<int32_t>((size_t)nsplit * nq * per);
    float* cand_score = dbg ? ws.get<float>((size_t)nsplit * nq * per) : nullptr;
    float* thr = ws.get<float>((size_t)nsplit * nq);
    int32_t* flag_list = ws.get<int32_t>((size_t)nq);
    unsigned char* scalars = ws.get<unsigned char>(64);
    if (!ws.ok()) return B200MNN_ENOMEM;
    unsigned int* absmax_bits = reinterpret_cast<unsigned int*>(scalars);
    int* scale_exp = reinterpret_cast<int*>(scalars + 8);
    unsigned long long* maxnorm_bits = reinterpret_cast<unsigned long long*>(scalars + 16);
    int* flag_count = reinterpret_cast<int*>(scalars + 24);
    B200_CUDA(cudaMemsetAsync(scalars, 0, 64, stream));

    {
        const int blocks = sm_count() * 8;
        absmax_kernel<<<blocks, 256, 0, stream>>>(dX, n * d, absmax_bits);
        B200_LAUNCH_CHECK();
        absmax_kernel<<<blocks, 256, 0, stream>>>(dQ, nq * d, absmax_bits);
        B200_LAUNCH_CHECK();
        scale_kernel<<<1, 1, 0, stream>>>(absmax_bits, scale_exp);
        B200_LAUNCH_CHECK();
        prep_operand_kernel<false><<<(unsigned)ceil_div(n_pad, 128), 128, 0, stream>>>(dX, n, n_pad, d, L, scale_exp, opB, norms, nullptr, maxnorm_bits);
        B200_LAUNCH_CHECK();
        prep_operand_kernel<true><<<(unsigned)ceil_div(nq_pad, 128), 128, 0, stream>>>(dQ, nq, nq_pad, d, L, scale_exp, opA, nullptr, qnorm, nullptr);
        B200_LAUNCH_CHECK();
    }

    CUtensorMap tmA, tmB;
    B200_TRY(make_operand_map(&tmA, opA, nq_pad, KS, BM));
    B200_TRY(make_operand_map(&tmB, opB, n_pad, KS, BN));

    int dev = 0, max_smem = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int nslot = MAX_SLOTS;
    while (nslot > 2 && candidates_smem_bytes(L.nbox, nslot, E) > (size_t)max_smem) --nslot;
    const size_t smem = candidates_smem_bytes(L.nbox, nslot, E);
    if (smem > (size_t)max_smem) return fail(B200MNN_ECUDA, "device does not offer enough shared memory per block for the kNN kernel");
    dim3 grid((unsigned)mtiles, (unsigned)nsplit);
    if (E == 1) {
        B200_CUDA(cudaFuncSetAttribute(knn_candidates_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_candidates_kernel<1><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, norms, sched, nslot, nq, ntiles, tiles_per_split, cand_idx, cand_score, thr);
    } else {
        B200_CUDA(cudaFuncSetAttribute(knn_candidates_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_candidates_kernel<2><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, norms, sched, nslot, nq, ntiles, tiles_per_split, cand_idx, cand_score, thr);
    }
    B200_LAUNCH_CHECK();

    if (dbg) {
        const int64_t ncand = (int64_t)nsplit * per;
        if (dbg->capacity < ncand) return fail(B200MNN_ECAPACITY, "debug candidate capacity too small");
        debug_convert_kernel<<<(unsigned)ceil_div(nq * ncand, 256), 256, 0, stream>>>(cand_score, thr, qnorm, scale_exp, nq, nsplit, per, cand_idx,
                                                                                   dbg->cand_idx, dbg->cand_d2, dbg->thr);
        B200_LAUNCH_CHECK();
        if (dbg->ncand_host) *dbg->ncand_host = ncand;
        B200_CUDA(cudaStreamSynchronize(stream));
        return 0;
    }

    rerank_kernel<<<(unsigned)ceil_div(nq, RR_WARPS), RR_WARPS * 32, 0, stream>>>(dX, dQ, nq, d, k, cand_idx, thr, nsplit, per, scale_exp, qnorm,
                                                                                maxnorm_bits, d_idx, d_dist, flag_count, flag_list, nullptr);
    B200_LAUNCH_CHECK();
    rescue_kernel<<<sm_count() * 4, RS_THREADS, (size_t)d * sizeof(double), stream>>>(dX, n, dQ, d, k, flag_count, flag_list, nq, d_idx, d_dist);
    B200_LAUNCH_CHECK();
    if (d_stats) { write_stats_kernel<<<1, 1, 0, stream>>>(flag_count, d_stats, nsplit, 1); B200_LAUNCH_CHECK(); }
    return 0;
}

}  // namespace knn
}  // namespace b200

extern "C" {

int b200mnn_dev_query_knn(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                          int64_t* d_stats, void* stream) {
    return b200::knn::query_knn_device(dX, n, dQ, nq, d, k, d_idx, d_dist, d_stats, static_cast<cudaStream_t>(stream), nullptr);
}

int b200mnn_dev_debug_candidates(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_cand_idx,
                                 double* d_cand_d2, double* d_thr, int64_t cand_capacity, int64_t* ncand_out, void* stream) {
    b200::knn::DebugOut dbg;
    dbg.cand_idx = d_cand_idx;
    dbg.cand_d2 = d_cand_d2;
    dbg.thr = d_thr;
    dbg.capacity = cand_capacity;
    dbg.ncand_host = ncand_out;
    return b200::knn::query_knn_device(dX, n, dQ, nq, d, k, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream), &dbg);
}

}  // extern "C"
