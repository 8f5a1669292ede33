// Exact k-nearest-neighbour search for sm_100a -- replaces BiocNeighbors::queryKNN(..., BNPARAM=KmknnParam())
// as batchelor calls it (R/MNN_tree.R:129 inside findMutualNN, R/fastMNN.R:605 for the tricube search).
//
// Contract (same as the exact KMKNN search): for every query the k reference rows with the smallest
// (squared Euclidean distance accumulated in double in dimension order, row index) pairs, ascending.
//
// Pipeline (all on one stream, no host synchronisation):
//   K0 rowstat/scale  : fp64 norms; one power-of-two scale S so that every |x*S| < 2^15 (fp16 range, lo parts normal)
//   P  cluster plan   : (large searches, knn_cluster.cu) k-means grouping of references and queries, per query tile the
//                       reference clusters in ascending order of a rigorous lower bound on their distance
//   K1 prep_operand   : fp64 row -> fp16 "hi" and "lo" K-slices (x*S = hi + lo + O(2^-22)) + the folded norm columns;
//                       operand rows are K-major, 64 fp16 (128 B) per TMA box, SWIZZLE_128B, references in grouped order
//   K2 knn_candidates : tcgen05 kernels (knn_candidates_ts_kernel: query operand in TMEM, the default; knn_candidates_kernel:
//                       both operands in shared memory).  score(q,j) = S^2 (||x_j||^2 - 2 q.x_j) accumulated in fp32 in
//                       TMEM, first with one fp16 term per 16 dims, for the queries that tier cannot certify with three
//                       (qh.xh + ql.xh + qh.xl).  One CTA = 128 queries (TMEM lanes) x the stream of 128-reference
//                       tiles its producer warp selects; the epilogue warps read the accumulators straight out of TMEM
//                       (tcgen05.ld), compare against per-row running thresholds and keep the 32*E best approximate
//                       scores of every row in a replace-the-maximum list in shared memory.
//   K3 rerank         : exact fp64 distances of the <= nsplit*32*E candidates (same summation order as the
//                       reference arithmetic), (distance, index) selection of the k best, and a CERTIFICATE:
//                       the result is provably exact if d2_k < (smallest retained threshold) - eps, where eps
//                       bounds the scoring error of the schedule.  Uncertified queries are flagged.
//   K4 rescue         : queries no tier certifies are recomputed by an exact fp64 scan of the references.  This is also
//                       the generic path for shapes the tensor path does not cover (k > 56, d > ~190).
#include "common.cuh"
#include "knn_cluster.cuh"
#include "internal.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

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
constexpr int MAX_MMA_PER_BOX = 8;   // a box holds 4 slices; a hi slice of B feeds two MMAs (ah.bh, al.bh)
constexpr int MAX_SLOTS = 8;
constexpr int NUM_THREADS = 224;      // warp 0: TMA, warps 1 and 6: MMA issuers (even / odd tiles), warps 2..5: epilogue
constexpr int TMEM_COLS = 512;        // two accumulator stages of BN columns
constexpr int MAX_K_TENSOR = 56;      // k <= 32*E - 8
constexpr int MAX_SPLIT = 8;

struct MmaSched {
    int nbox;
    int nmma[MAX_NBOX];
    unsigned char a_slice[MAX_NBOX][MAX_MMA_PER_BOX];  // stored slice of A (absolute)
    unsigned char b_sub[MAX_NBOX][MAX_MMA_PER_BOX];    // stored slice of B inside the current box (0..3)
};

// Stored K layout of one operand row (both operands).  With S the common power-of-two scale:
//   query row     a = -2*S*q  split as a = ah + al (+O(2^-22));     reference row  b = S*x = bh + bl (+O(2^-22))
// PRIMARY slices come first: the "hi" slice of every full group g of 16 dims (slice g), then -- if r = d % 16 dims
// remain and 3r <= 16 -- one remainder slice
//   query row : [ah(r) | ah(r) | al(r) | ...]      reference row : [bh(r) | bl(r) | bh(r) | ...]
// whose single MMA yields ah.bh + ah.bl + al.bh for those dims (if 3r > 16 the remainder is zero-padded into one more
// full group).  Three more columns fold the reference norm into the same accumulation:
//   query row : [2^15, 2^4, 2^-7]                   reference row : N = S^2 ||x||^2 split as 2^15 n1 + 2^4 n2 + 2^-7 n3
// so the accumulator IS the score S^2 (||x||^2 - 2 q.x): the epilogue needs neither loads nor FMAs.  The norm
// columns share the remainder slice when 3r + 3 <= 16, else they get a slice (and one MMA) of their own.
// The "lo" slices of the full groups follow the primary slices (slice nprim + g).  Two schedules use this layout:
//   precise : per group ah.bh, al.bh, ah.bl (3 MMAs) + remainder/norm        -> error ~2^-22 relative
//   fast    : per group ah.bh only (1 MMA) + remainder/norm, reading only the boxes that hold primary slices
//             -> error <= ~2^-9 |q| |x|, 2.5x fewer MMAs and half the operand traffic for d = 50
struct KLayout {
    int d, groups, rem, nslices, nbox;
    int nprim, nbox_fast;       // primary slices (hi, remainder, norm) and the boxes that hold them
    int norm_slice, norm_col;   // stored slice and first column (0..13) of the three norm columns
    int rem_slice;              // stored slice of the packed remainder dims (-1 if none)
};

static KLayout make_layout(int d) {
    KLayout L;
    L.d = d;
    L.groups = d / SLICE;
    L.rem = d % SLICE;
    if (L.rem * 3 > SLICE) { L.groups += 1; L.rem = 0; }
    int ns = L.groups;
    L.rem_slice = -1;
    if (L.rem) L.rem_slice = ns++;
    if (L.rem && 3 * L.rem + 3 <= SLICE) { L.norm_slice = L.rem_slice; L.norm_col = 3 * L.rem; }
    else { L.norm_slice = ns++; L.norm_col = 0; }
    L.nprim = ns;
    L.nslices = L.nprim + L.groups;
    L.nbox = (L.nslices + 3) / 4;
    L.nbox_fast = (L.nprim + 3) / 4;
    return L;
}

static MmaSched make_sched(const KLayout& L, bool fast) {
    MmaSched s;
    memset(&s, 0, sizeof(s));
    s.nbox = fast ? L.nbox_fast : L.nbox;
    auto add = [&](int a_slice, int b_slice) {
        const int box = b_slice / 4;
        int& m = s.nmma[box];
        s.a_slice[box][m] = (unsigned char)a_slice;
        s.b_sub[box][m] = (unsigned char)(b_slice % 4);
        ++m;
    };
    for (int g = 0; g < L.groups; ++g) {
        const int sh = g, sl = L.nprim + g;
        add(sh, sh);                       // ah.bh
        if (!fast) {
            add(sl, sh);                   // al.bh
            add(sh, sl);                   // ah.bl
        }
    }
    if (L.rem_slice >= 0) add(L.rem_slice, L.rem_slice);
    if (L.norm_slice != L.rem_slice) add(L.norm_slice, L.norm_slice);
    return s;
}

using namespace tc;

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0,
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc() { return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

// ------------------------------------------------------------------------------------------------
// Per-query candidate buffers (epilogue)
// ------------------------------------------------------------------------------------------------
// Every query row owns CAP = KEEP + PEND 64-bit keys in shared memory: the KEEP = 32*E best approximate scores seen so
// far (sorted after the first compaction) followed by up to PEND pending hits.  A key is (order-preserving score bits
// << 32 | reference id), so unsigned comparison orders by (score, id).  A hit (score < the row's threshold) is appended
// by the row's own lane with one shared-memory store -- no warp cooperation on the hot path.  When a row is about to
// run out of room the warp compacts it cooperatively (bitonic sort of the pending keys across lanes, bitonic merge with
// the retained keys), which also tightens the threshold to the KEEP-th best score.  Thresholds only ever decrease, so
// every key that was ever rejected or dropped has score >= the final threshold: the certificate of the re-rank kernel.
constexpr int ROWPITCH = 33;   // buffer position stride (keys): compaction reads of one row are bank-conflict free
template <int E>
struct Cand {
    static constexpr int KEEP = 32 * E;
    static constexpr int PEND = 32;
    static constexpr int CAP = KEEP + PEND;
    static constexpr int WARP_KEYS = CAP * ROWPITCH;
};
constexpr unsigned long long EMPTY_KEY = ~0ull;

__device__ __forceinline__ unsigned long long make_key(float s, int col) {
    uint32_t u = __float_as_uint(s);
    u ^= (uint32_t)((int32_t)u >> 31) | 0x80000000u;
    return ((unsigned long long)u << 32) | (uint32_t)col;
}
__device__ __forceinline__ float key_score(unsigned long long k) {
    uint32_t u = (uint32_t)(k >> 32);
    u ^= ((u >> 31) - 1u) | 0x80000000u;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

// Bitonic networks over NR independent key sets at once (one key of every set per lane).  Measured on B200: one
// network stage costs ~61 cycles for one row (SHFL latency + 64-bit compare/select) but ~175 cycles for four
// interleaved rows -- the integer compare/select work runs on the half-rate ALU pipe and dominates, so batching rows
// buys little and hurts when a call has fewer rows than the batch; NR = 1 is the default.
#ifndef B200_MERGE_ROWS
#define B200_MERGE_ROWS 1
#endif
#ifndef B200_SORT_ROLLED
#define B200_SORT_ROLLED 0
#endif
#if B200_SORT_ROLLED
#define B200_SORT_UNROLL _Pragma("unroll 1")
#else
#define B200_SORT_UNROLL _Pragma("unroll")
#endif
constexpr int MERGE_ROWS = B200_MERGE_ROWS;
// bitonic sort of one key per lane, ascending by lane
template <int NR>
__device__ __forceinline__ void sort32n(unsigned long long (&key)[NR], const int lane) {
    B200_SORT_UNROLL
    for (int k = 2; k <= 32; k <<= 1) {
        B200_SORT_UNROLL
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool up = (lane & k) == 0;                     // ascending block? (k == 32: always)
            const bool low = (lane & j) == 0;                    // lower partner of the pair?
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[i], j);
                key[i] = ((key[i] < other) == (up == low)) ? key[i] : other;   // one 64-bit compare, one select
            }
        }
    }
}
// bitonic sequence (one key per lane) -> ascending
template <int NR>
__device__ __forceinline__ void clean32n(unsigned long long (&key)[NR], const int lane) {
    B200_SORT_UNROLL
    for (int j = 16; j > 0; j >>= 1) {
        const bool low = (lane & j) == 0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[i], j);
            key[i] = ((key[i] < other) == low) ? key[i] : other;
        }
    }
}
__device__ __forceinline__ unsigned long long reverse32(unsigned long long key) { return __shfl_xor_sync(0xffffffffu, key, 31); }

// Merges, for NR rows at once, the pending keys p[] (unsorted, one per lane, EMPTY_KEY where none) into the kept keys
// a0[] (E == 1) or a0[] <= a1[] (E == 2); sort_kept: the kept keys are not sorted yet (first merge of a row).
template <int E, int NR>
__device__ __forceinline__ void merge_keys(unsigned long long (&a0)[NR], unsigned long long (&a1)[NR], unsigned long long (&p)[NR],
                                           const bool sort_kept, const int lane) {
    if (sort_kept) {   // warp-uniform
        sort32n<NR>(a0, lane);
        if (E == 2) {
            sort32n<NR>(a1, lane);
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const unsigned long long ra = reverse32(a1[i]);
                const unsigned long long lo = umin64(a0[i], ra), hi = umax64(a0[i], ra);
                a0[i] = lo;
                a1[i] = hi;
            }
            clean32n<NR>(a0, lane);
            clean32n<NR>(a1, lane);                                 // now a0 <= a1 elementwise, both sorted
        }
    }
    sort32n<NR>(p, lane);
    if (E == 1) {
#pragma unroll
        for (int i = 0; i < NR; ++i) a0[i] = umin64(a0[i], reverse32(p[i]));
        clean32n<NR>(a0, lane);                                     // the 32 smallest of both, sorted
    } else {
#pragma unroll
        for (int i = 0; i < NR; ++i) p[i] = umin64(a1[i], reverse32(p[i]));
        clean32n<NR>(p, lane);                                      // 32 smallest of (a1 u p); the rest is dropped
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const unsigned long long rl = reverse32(p[i]);
            const unsigned long long lo = umin64(a0[i], rl), hi = umax64(a0[i], rl);
            a0[i] = lo;
            a1[i] = hi;
        }
        clean32n<NR>(a0, lane);
        clean32n<NR>(a1, lane);
    }
}

// Compacts every row whose lane has cnt > limit.  buf: this warp's CAP x ROWPITCH keys.
// (Inlined on purpose: as a call it forced thr/cnt and the live TMEM registers of the caller into local memory.)
template <int E>
__device__ __forceinline__ void compact_rows(unsigned long long* __restrict__ buf, const int lane, const int limit, float& thr, int& cnt,
                                             bool& sorted) {
    constexpr int NR = MERGE_ROWS;
    __syncwarp();
    unsigned todo = __ballot_sync(0xffffffffu, cnt > limit);
    const bool sort_kept = __any_sync(0xffffffffu, cnt > limit && !sorted);   // sorting sorted keys again is harmless
    while (todo) {
        int r[NR], c[NR];
        unsigned long long a0[NR], a1[NR], p[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            r[i] = todo ? __ffs(todo) - 1 : -1;                    // -1: batch slot unused (computed on row r[0], not stored)
            todo &= todo - 1;
            const int rr = r[i] < 0 ? r[0] : r[i];
            c[i] = __shfl_sync(0xffffffffu, cnt, rr);
            a0[i] = (lane < c[i]) ? buf[lane * ROWPITCH + rr] : EMPTY_KEY;
            if (E == 1) {
                a1[i] = EMPTY_KEY;
                p[i] = (lane + 32 < c[i]) ? buf[(lane + 32) * ROWPITCH + rr] : EMPTY_KEY;
            } else {
                a1[i] = (lane + 32 < c[i]) ? buf[(lane + 32) * ROWPITCH + rr] : EMPTY_KEY;
                p[i] = (lane + 64 < c[i]) ? buf[(lane + 64) * ROWPITCH + rr] : EMPTY_KEY;
            }
        }
        merge_keys<E, NR>(a0, a1, p, sort_kept, lane);
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (r[i] >= 0) {   // warp-uniform
                buf[lane * ROWPITCH + r[i]] = a0[i];
                if (E == 2) buf[(lane + 32) * ROWPITCH + r[i]] = a1[i];
                const int keep = 32 * E;
                const int newcnt = c[i] < keep ? c[i] : keep;
                const unsigned long long last = __shfl_sync(0xffffffffu, E == 1 ? a0[i] : a1[i], 31);
                if (lane == r[i]) {
                    cnt = newcnt;
                    sorted = true;
                    thr = (newcnt == keep) ? key_score(last) : __int_as_float(0x7f800000);
                }
            }
        }
    }
    __syncwarp();
}

// One 32-column chunk of accumulators (registers v[], lane = query row) against this lane's running threshold.
// Hot path: one min tree (3-input FMNMX3: 16 instructions for 32 scores) and ONE warp-uniform vote/branch per chunk.
// Rare path (some lane has a score below its threshold): the warp stages the chunk in shared memory (XOR-swizzled so
// that both the row-wise writes and the column-wise reads below are bank-conflict free), then for every hitting row L
// the 32 lanes each look at ONE score of that row, vote, and the hitting lanes store their keys side by side at the
// end of row L's buffer.  No divergent branches and no per-lane search: a chunk with one hit costs ~20 instructions.
// A row whose buffer is full drops the hit and is marked dirty; dirty rows are handed to the exact rescue kernel, so
// exactness never depends on buffer capacity.
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float chunk_min(const uint32_t (&v)[32]) {
#define VF(i) __uint_as_float(v[i])
    const float t0 = fmin3(VF(0), VF(1), VF(2)), t1 = fmin3(VF(3), VF(4), VF(5)), t2 = fmin3(VF(6), VF(7), VF(8));
    const float t3 = fmin3(VF(9), VF(10), VF(11)), t4 = fmin3(VF(12), VF(13), VF(14)), t5 = fmin3(VF(15), VF(16), VF(17));
    const float t6 = fmin3(VF(18), VF(19), VF(20)), t7 = fmin3(VF(21), VF(22), VF(23)), t8 = fmin3(VF(24), VF(25), VF(26));
    const float t9 = fmin3(VF(27), VF(28), VF(29)), t10 = fminf(VF(30), VF(31));
    const float u0 = fmin3(t0, t1, t2), u1 = fmin3(t3, t4, t5), u2 = fmin3(t6, t7, t8), u3 = fminf(t9, t10);
    return fminf(fminf(u0, u1), fminf(u2, u3));
#undef VF
}
// Rare path of a chunk: `hit` = lanes (rows) that have at least one score below their threshold.
template <int CAPV>
__device__ __forceinline__ void append_hits(const uint32_t (&v)[32], unsigned hit, const int col0, const float thr, int& cnt, bool& dirty,
                                            unsigned long long* __restrict__ buf, float4* __restrict__ stg, const int lane, int* stat) {
#define VF(i) __uint_as_float(v[i])
    if (stat) { stat[0] += 1; stat[1] += __popc(hit); }
#pragma unroll
    for (int j = 0; j < 8; ++j) stg[j * 32 + (lane ^ j)] = make_float4(VF(4 * j), VF(4 * j + 1), VF(4 * j + 2), VF(4 * j + 3));
    __syncwarp();
    const float* stgf = reinterpret_cast<const float*>(stg);
    const int g = lane >> 2;
    do {
        const int L = __ffs(hit) - 1;
        hit &= hit - 1;
        const float thrL = __shfl_sync(0xffffffffu, thr, L);
        const int cntL = __shfl_sync(0xffffffffu, cnt, L);
        const float sc = stgf[((g * 32 + (L ^ g)) << 2) + (lane & 3)];   // score of (row L, column col0 + lane)
        const bool h = sc < thrL;
        const unsigned m = __ballot_sync(0xffffffffu, h);
        const int pos = cntL + __popc(m & ((1u << lane) - 1u));
        if (h && pos < CAPV) buf[pos * ROWPITCH + L] = make_key(sc, col0 + lane);
        if (stat) stat[2] += __popc(m);
        if (lane == L) {
            const int n = cntL + __popc(m);
            if (n > CAPV) dirty = true;
            cnt = n > CAPV ? CAPV : n;
        }
    } while (hit);
    __syncwarp();   // staging area is reused by the next chunk; appended keys become visible to the whole warp
#undef VF
}
template <int E>
__device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], const int col0, const float thr, int& cnt, bool& dirty,
                                           unsigned long long* __restrict__ buf, float4* __restrict__ stg, const int lane, int* stat = nullptr) {
    const float mm = chunk_min(v);
    const unsigned hit = __ballot_sync(0xffffffffu, mm < thr);
    if (hit) append_hits<Cand<E>::CAP>(v, hit, col0, thr, cnt, dirty, buf, stg, lane, stat);
}
// Two chunks behind ONE vote/branch: both min trees interleave (a lone warp per scheduler has nothing else to hide
// the ALU latency behind) and the quiet path pays one vote instead of two.
template <int CAPV>
__device__ __forceinline__ bool scan_pair(const uint32_t (&va)[32], const uint32_t (&vb)[32], const int col0, const float thr, int& cnt,
                                          bool& dirty, unsigned long long* __restrict__ buf, float4* __restrict__ stg, const int lane,
                                          int* stat = nullptr) {
    const float ma = chunk_min(va), mb = chunk_min(vb);
    const bool any = __any_sync(0xffffffffu, fminf(ma, mb) < thr);
    if (any) {
        const unsigned ha = __ballot_sync(0xffffffffu, ma < thr);
        if (ha) append_hits<CAPV>(va, ha, col0, thr, cnt, dirty, buf, stg, lane, stat);
        const unsigned hb = __ballot_sync(0xffffffffu, mb < thr);
        if (hb) append_hits<CAPV>(vb, hb, col0 + 32, thr, cnt, dirty, buf, stg, lane, stat);
    }
    return any;   // warp-uniform
}

// ------------------------------------------------------------------------------------------------
// K2: tensor-core candidate scoring + per-query top-(32*E) of the approximate scores
// ------------------------------------------------------------------------------------------------
template <int E, int NBOX>
__global__ void __launch_bounds__(NUM_THREADS, 1)
knn_candidates_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const MmaSched sched, const int nslot, const int64_t nq, const int ntiles, const int tiles_per_split,
                      int32_t* __restrict__ cand_idx,   // [nsplit][nq][32E]
                      float* __restrict__ cand_score,   // [nsplit][nq][32E] (may be null)
                      float* __restrict__ thr_out,      // [nsplit][nq]
                      const int dbg_mode,               // 0 = normal; 1 = epilogue only recycles stages; 2 = + TMEM loads, no filter
                      const int csize,                  // thread-block cluster size (1, 2 or 4): CTAs of a cluster share every B box
                      long long* __restrict__ dbg_ts,   // optional [64][32] clock64 trace of CTA (0,0) (measurement aid)
                      const int trace_start)            // first tile recorded in the trace
{
    // 1024-byte alignment is what SWIZZLE_128B needs; keeping plain pointer arithmetic on the __shared__ array keeps
    // the address space visible to the compiler (LDS/STS instead of generic LD/ST in the epilogue).
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int nbox = NBOX;   // boxes (64 fp16 columns) per operand row: compile-time so the MMA issue path unrolls

    uint8_t* smA = smem;
    uint8_t* smB = smA + (size_t)nbox * A_BOX_BYTES;
    unsigned long long* lists = reinterpret_cast<unsigned long long*>(smB + (size_t)nslot * B_BOX_BYTES);
    float4* stage_all = reinterpret_cast<float4*>(lists + 4 * Cand<E>::WARP_KEYS + 2);   // [4 warps][8][32] float4, 16-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + 4 * 8 * 32);
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
    const bool trace = dbg_ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define B200_TS(tile, k)                                                        \
    do {                                                                        \
        if (trace && lane == 0 && (tile) >= trace_start && (tile) < trace_start + 64) dbg_ts[((tile) - trace_start) * 32 + (k)] = clock64(); \
    } while (0)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        // a slot is free again once the MMA threads of ALL CTAs of the cluster have retired their reads of it
        for (int i = 0; i < nslot; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), (uint32_t)csize); }
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
    if (csize > 1) cluster_sync_all();   // every CTA's barriers are initialised before any peer multicasts into it
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t crank = csize > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(afull), (uint32_t)(nbox * A_BOX_BYTES));
            for (int b = 0; b < nbox; ++b) tma_load_2d(smem_u32(smA + (size_t)b * A_BOX_BYTES), &tmA, smem_u32(afull), b * KBOX, m0);
            int slot = 0;
            uint32_t phase = 0;
            const int rows_per = BN / csize;                       // this CTA fetches 1/csize of every box ...
            const uint32_t part_off = crank * (uint32_t)rows_per * 128u;  // ... and multicasts it to the whole cluster
            for (int t = tile0; t < tile1; ++t) {
                for (int b = 0; b < nbox; ++b) {
                    mbar_wait(smem_u32(&empty[slot]), phase ^ 1);
                    if (dbg_mode == 3) {  // measurement aid: no TMA traffic at all, MMAs run on stale shared memory
                        mbar_arrive(smem_u32(&full[slot]));
                        if (++slot == nslot) { slot = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(smem_u32(&full[slot]), (uint32_t)B_BOX_BYTES);
                    const uint32_t dst = smem_u32(smB + (size_t)slot * B_BOX_BYTES);
                    if (csize == 1) tma_load_2d(dst, &tmB, smem_u32(&full[slot]), b * KBOX, t * BN);
                    else tma_load_2d_mc(dst + part_off, &tmB, smem_u32(&full[slot]), b * KBOX, t * BN + (int)crank * rows_per, cmask);
                    if (++slot == nslot) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == 6) {
        // ===================== MMA issuers =====================
        // tcgen05.mma issue blocks until the tensor pipe accepts the instruction, so every barrier wait / commit of an
        // issuing thread is dead time for the pipe.  Two issuing warps alternate tiles (warp 1: even tiles = accumulator
        // stage 0, warp 6: odd tiles = stage 1): while one sits in its waits the other keeps the pipe fed.
        // A single thread has no latency hiding, so the per-tile issue path must be straight-line code: every
        // descriptor of the schedule is precomputed into registers (static indices after unrolling over NBOX x 6).
        // The whole warp runs the loop (warp-uniform values stay in uniform registers); one elected lane issues.
        {
            constexpr uint32_t idesc = make_idesc();
            uint64_t dA[NBOX][MAX_MMA_PER_BOX];
            uint32_t oB[NBOX][MAX_MMA_PER_BOX];
            int nm[NBOX];
            const uint32_t a_base = smem_u32(smA);
#pragma unroll
            for (int b = 0; b < NBOX; ++b) {
                nm[b] = sched.nmma[b];
#pragma unroll
                for (int m = 0; m < MAX_MMA_PER_BOX; ++m) {
                    const int as = sched.a_slice[b][m];
                    dA[b][m] = make_sw128_desc(a_base + (uint32_t)(as >> 2) * A_BOX_BYTES + (uint32_t)(as & 3) * 32u);
                    oB[b][m] = (uint32_t)sched.b_sub[b][m] * 2u;   // 32 bytes per K slice, in the descriptor's 16-byte units
                }
            }
            const uint64_t db_base = make_sw128_desc(smem_u32(smB));
            mbar_wait_u(smem_u32(afull), 0);
            tc_fence_after();
            const int mma_id = (warp == 1) ? 0 : 1;
            for (int tl = mma_id; tl < my_tiles; tl += 2) {
                const int stage = tl & 1;
                const uint32_t use = (uint32_t)(tl >> 1);
                B200_TS(tl, 0);
                mbar_wait_u(smem_u32(&tempty[stage]), (use & 1) ^ 1);
                tc_fence_after();
                B200_TS(tl, 1);
                const uint32_t tmem_d = tmem_base + (uint32_t)(stage * BN);
                const int box0 = tl * NBOX;                 // global box counter of this tile's first box
                int slot = box0 % nslot;
                uint32_t phase = (uint32_t)((box0 / nslot) & 1);
#pragma unroll
                for (int b = 0; b < NBOX; ++b) {
                    mbar_wait_u(smem_u32(&full[slot]), phase);
                    tc_fence_after();
                    B200_TS(tl, 2 + 2 * (b < 2 ? b : 1));
                    const uint64_t db0 = db_base + (uint64_t)((uint32_t)slot * (uint32_t)(B_BOX_BYTES >> 4));
                    if (elect_one()) {
                        if (dbg_mode != 4) {  // mode 4: TMA pipeline alone (no MMAs issued)
#pragma unroll
                            for (int m = 0; m < MAX_MMA_PER_BOX; ++m) {
                                if (m < nm[b]) {
                                    if (b == 0 && m == 0) umma_f16<false>(tmem_d, dA[b][m], db0 + oB[b][m], idesc);
                                    else umma_f16<true>(tmem_d, dA[b][m], db0 + oB[b][m], idesc);
                                }
                            }
                        }
                        // frees the smem slot (in every CTA of the cluster) once these MMAs retire
                        if (csize == 1) umma_commit(smem_u32(&empty[slot]));
                        else umma_commit_mc(smem_u32(&empty[slot]), cmask);
                        if (b == NBOX - 1) umma_commit(smem_u32(&tfull[stage]));   // accumulator stage complete
                    }
                    __syncwarp();
                    B200_TS(tl, 3 + 2 * (b < 2 ? b : 1));
                    if (++slot == nslot) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> threshold filter -> sorted lists =====================
        const int q4 = warp & 3;  // TMEM lane quarter this warp may access
        unsigned long long* mybuf = lists + (size_t)(warp - 2) * Cand<E>::WARP_KEYS;   // this warp's candidate buffers
        float4* stg = stage_all + (size_t)(warp - 2) * 8 * 32;                       // this warp's staging tile
        int cnt = 0;          // keys currently in this lane's row buffer
        bool sorted = false;  // retained part sorted (true after the first compaction)
        bool dirty = false;   // this row ever dropped a hit (buffer full) -> certificate must fail -> exact rescue
        float thr = __int_as_float(0x7f800000);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        constexpr int BOOT_TILES = 6;                       // tiles scanned in bootstrap mode (compaction after every chunk)
        for (int tl = 0; tl < my_tiles; ++tl) {
            const int stage = tl & 1;
            const uint32_t use = (uint32_t)(tl >> 1);
            if (warp == 2) B200_TS(tl, 8);
            mbar_wait_u(smem_u32(&tfull[stage]), use & 1);
            tc_fence_after();
            if (warp == 2) B200_TS(tl, 9);
            const int colbase = (tile0 + tl) * BN;
            const uint32_t tbase = lane_base + (uint32_t)(stage * BN);
            if (dbg_mode == 1 || dbg_mode == 3 || dbg_mode == 4) {  // measurement aid: pure TMA + MMA pipeline
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tempty[stage]));
                if (warp == 2) B200_TS(tl, 10);
                continue;
            }
            if (trace && warp == 2 && tl >= trace_start && tl < trace_start + 64) {
                const int tot = __reduce_add_sync(0xffffffffu, cnt);
                if (lane == 0) dbg_ts[(tl - trace_start) * 32 + 24] = tot;
            }
            uint32_t v0[32], v1[32], v2[32], v3[32];
            if (tl < BOOT_TILES) {
                // Bootstrap: thresholds are still loose (every column of the very first tile is a hit), so rows are
                // compacted after every chunk; a chunk adds at most 32 = PEND keys, hence no row can overflow here.
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    tmem_ld32(tbase + (uint32_t)(c * 32), v0);
                    tmem_ld_wait();
                    if (c == BN / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tempty[stage]));
                        if (warp == 2) B200_TS(tl, 10);
                    }
                    scan_chunk<E>(v0, colbase + c * 32, thr, cnt, dirty, mybuf, stg, lane);
                    if (__any_sync(0xffffffffu, cnt > Cand<E>::KEEP)) compact_rows<E>(mybuf, lane, Cand<E>::KEEP, thr, cnt, sorted);
                }
            } else {
                // Steady state.  Two 32-column loads are always in flight while two others are scanned: reading freshly
                // written accumulators out of TMEM runs at ~64 B/cycle/SM, which makes the TMEM read port the limiter of
                // this kernel (128 KB per tile = ~2k cycles vs ~1.3k cycles of MMA); it must never idle.
                tmem_ld32(tbase, v0);
                tmem_ld32(tbase + 32u, v1);
                tmem_ld_wait();
#pragma unroll 1
                for (int c = 0; c < BN / 32; c += 4) {
                    tmem_ld32(tbase + (uint32_t)((c + 2) * 32), v2);
                    tmem_ld32(tbase + (uint32_t)((c + 3) * 32), v3);
                    if (dbg_mode != 2) {
                        scan_chunk<E>(v0, colbase + c * 32, thr, cnt, dirty, mybuf, stg, lane);
                        scan_chunk<E>(v1, colbase + (c + 1) * 32, thr, cnt, dirty, mybuf, stg, lane);
                    } else if (__uint_as_float(v0[0]) == 1.2345e-30f && __uint_as_float(v1[31]) == 1.2345e-30f) thr = 0.f;
                    if (warp == 2) B200_TS(tl, 16 + c);
                    tmem_ld_wait();
                    if (c + 4 < BN / 32) {
                        tmem_ld32(tbase + (uint32_t)((c + 4) * 32), v0);
                        tmem_ld32(tbase + (uint32_t)((c + 5) * 32), v1);
                    } else {  // every column of this stage is now in registers: hand the stage back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tempty[stage]));
                        if (warp == 2) B200_TS(tl, 10);
                    }
                    if (dbg_mode != 2) {
                        scan_chunk<E>(v2, colbase + (c + 2) * 32, thr, cnt, dirty, mybuf, stg, lane);
                        scan_chunk<E>(v3, colbase + (c + 3) * 32, thr, cnt, dirty, mybuf, stg, lane);
                    } else if (__uint_as_float(v2[0]) == 1.2345e-30f && __uint_as_float(v3[31]) == 1.2345e-30f) thr = 0.f;
                    if (warp == 2) B200_TS(tl, 18 + c);
                    tmem_ld_wait();
                }
                // one compaction site per tile (the last tile compacts every row: final sorted lists + thresholds)
                // (hits per row per tile fall like 48/tl: leave more head-room while they are still frequent)
                const int limit = (tl == my_tiles - 1) ? -1 : (tl < 32 ? Cand<E>::KEEP + 8 : Cand<E>::KEEP + 16);
                if (__any_sync(0xffffffffu, cnt > limit)) compact_rows<E>(mybuf, lane, limit, thr, cnt, sorted);
            }
            if (warp == 2) B200_TS(tl, 11);
            if (trace && warp == 2 && tl >= trace_start && tl < trace_start + 64) {
                const int tot = __reduce_add_sync(0xffffffffu, cnt);
                if (lane == 0) dbg_ts[(tl - trace_start) * 32 + 25] = tot;
            }
        }
        // rows that never saw a steady-state tile still need their final compaction
        if (my_tiles <= BOOT_TILES) compact_rows<E>(mybuf, lane, -1, thr, cnt, sorted);
        if (dirty) thr = __int_as_float(0xff800000);   // -inf: the re-rank certificate cannot hold -> exact rescue
        const int64_t rowbase = (int64_t)m0 + q4 * 32;
        const int64_t sbase = (int64_t)blockIdx.y * nq;
#pragma unroll 1
        for (int r = 0; r < 32; ++r) {
            const int64_t row = rowbase + r;
            if (row < nq) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long key = mybuf[(lane + 32 * e) * ROWPITCH + r];
                    const int64_t o = (sbase + row) * (32 * E) + lane + 32 * e;
                    cand_idx[o] = (int32_t)(uint32_t)key;       // EMPTY_KEY -> -1
                    if (cand_score) cand_score[o] = key_score(key);
                }
            }
        }
        if (rowbase + lane < nq) thr_out[sbase + rowbase + lane] = thr;
        tc_fence_before();
    }
    __syncthreads();
    if (csize > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it or arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// K2 (TS variant): the query operand lives in TMEM, eight epilogue warps
// ------------------------------------------------------------------------------------------------
// Same algorithm as knn_candidates_kernel, different data flow, motivated by what bounds that kernel on B200:
//  * SS-mode UMMA fetches A and B from shared memory (96 B/cycle) while TMA writes the next boxes (~47 B/cycle): the
//    128 B/cycle shared-memory port is oversubscribed.  Here every query row is written into TMEM once (tcgen05.st,
//    row = lane, two fp16 per 32-bit column) and the MMAs take A from TMEM (TS mode): the tensor core reads only B
//    from shared memory.
//  * Tiles are 128 references wide and THREE accumulator stages (3 x 128 columns) sit next to the A columns.
//  * Measured: with the tensor pipe busy, a lone epilogue warp per scheduler issues one ALU instruction every ~5
//    cycles (fixed-latency stalls, nothing to switch to), which made the epilogue -- not the MMAs -- the bound.  So
//    the epilogue runs EIGHT warps, two per scheduler: warps w and w+4 read the same TMEM lane quarter and split its
//    ROWS (16 query rows each, all 128 columns of a tile: tcgen05.ld 16x32bx2).  A stage is released as soon as the
//    tile is in registers (TMEM loads take ~30 cycles), so the MMA pipeline always has ~3 tiles of slack.
//  * Every row has ONE candidate list (its KEEP best scores, unsorted, in shared memory) and one owning warp: a hit
//    replaces the current maximum (one vote to find its holder, one REDUX for the new maximum = the new threshold).
//    No pending buffers, no locks, no capacity that could overflow, nothing shared between warps.  With cluster
//    pruning nearly every scored tile belongs to the queries' own component, so hits are frequent and their cost --
//    not the quiet path -- decides the kernel time; the former sorted-list-with-pending-segment scheme spent 5 of 9 k
//    cycles per tile in bitonic merges and lock waits there.
constexpr int TS_BN = 128;                   // references per tile
constexpr int TS_STAGES = 3;                 // accumulator stages in TMEM
constexpr int TS_B_BOX_BYTES = TS_BN * 128;  // 16 KB
constexpr int TS_MAX_NBOX = 4;               // A columns: 32 per box, 3*128 accumulator columns -> at most 4 boxes
constexpr int TS_THREADS = 352;              // warp 0: TMA, warps 1 and 10: MMA issuers (even / odd tiles), warps 2..9: epilogue
constexpr int TS_EPI_WARPS = 8;
constexpr int TS_RING = 32;                  // tile-id ring between the producer and its consumers (>= tiles in flight + 2)
// Pruned search (knn_cluster.cuh): per query tile the reference clusters in ascending order of their lower bound.
// cl_list == nullptr: dense scan of tiles [tile0, tile1).
struct PruneArgs {
    const int2* cl_list;     // [query tile][C] (cluster, float bits of S^2 LB^2)
    const int* cl_tile0;     // [C + 1] first reference tile of every cluster
    const float* qoff;       // [slot] S^2 ||q||^2 rounded up, -inf for padding slots
    int C;
    unsigned long long* visited;   // optional counter: tiles scored by this launch
    unsigned long long* mma_count; // optional counter (profiling): tcgen05.mma instructions issued = tiles x MMAs per tile
    int mma_per_tile;
};
__device__ __forceinline__ void umma_f16_ts_impl(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :
                 : "r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Order-preserving map float -> uint32 (the high word of make_key) and back.
__device__ __forceinline__ uint32_t ord_bits(float s) {
    uint32_t u = __float_as_uint(s);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t u) { return __uint_as_float(u ^ (((u >> 31) - 1u) | 0x80000000u)); }
constexpr uint32_t ORD_INF = 0xFF800000u;   // ord_bits(+inf)

// Select that the compiler cannot turn back into a branch (the insertion rounds must stay straight-line code so that the
// dependent chains of several rows interleave).
__device__ __forceinline__ uint32_t sel_u32(const bool c, const uint32_t a, const uint32_t b) {
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((uint32_t)c));
    return r;
}

// Conditional replacement of a list entry (score bits, id) under ONE predicate, as straight-line code.
__device__ __forceinline__ void sel_entry(const bool c, const uint32_t score, const uint32_t id, uint2& e) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\tselp.b32 %0, %2, %0, p;\n\tselp.b32 %1, %3, %1, p;\n\t}"
        : "+r"(e.x), "+r"(e.y)
        : "r"(score), "r"(id), "r"((uint32_t)c));
}

// Stages one 32x32 chunk of scores (registers, lane = row) in shared memory, XOR-swizzled so that both these row-wise
// writes and the column-wise reads of ts_chunk_score are bank-conflict free.
__device__ __forceinline__ void ts_stage_chunk(const uint32_t (&v)[32], float4* __restrict__ stg, const int lane) {
#define VF(i) __uint_as_float(v[i])
#pragma unroll
    for (int j = 0; j < 8; ++j) stg[j * 32 + (lane ^ j)] = make_float4(VF(4 * j), VF(4 * j + 1), VF(4 * j + 2), VF(4 * j + 3));
#undef VF
    __syncwarp();
}
// score of (row L, column `lane`) of the staged chunk
__device__ __forceinline__ float ts_chunk_score(const float4* __restrict__ stg, const int L, const int lane) {
    const int g = lane >> 2;
    return reinterpret_cast<const float*>(stg)[((g * 32 + (L ^ g)) << 2) + (lane & 3)];
}
// First chunks of a warp: copied straight into segment `seg` of every row's list (no selection needed yet).
// First load of a warp: the staged values go straight into the lists (no selection needed yet).  Holder lane h < 16
// carries columns [col0, col0+32) of row h, holder lane h + 16 columns [col0+32, col0+64) of the same row: the former
// fill segment 0 of every list, the latter segment 1 when the lists have one (E == 2).
template <int E>
__device__ __forceinline__ void ts_fill_staged(const int col0, uint2* __restrict__ list, const float4* __restrict__ stg, const int lane) {
    constexpr int KEEP = 32 * E;
#pragma unroll 8
    for (int H = 0; H < 16 * E; ++H) {
        const float sc = ts_chunk_score(stg, H, lane);
        const uint32_t o = ord_bits(sc);
        list[(H & 15) * KEEP + (H >> 4) * 32 + lane] = make_uint2(o, o >= ORD_INF ? 0xFFFFFFFFu : (uint32_t)(col0 + (H >> 4) * 32 + lane));   // padding rows score +inf
    }
    __syncwarp();
}
// Hits of one staged load: a holder lane L carries 32 columns of row L & 15.  For every hitting holder the 32 lanes each
// look at ONE of its scores; every score below the row's threshold replaces the current maximum of the row's list.
// Invariant: thr (register of both holder lanes of a row) == max of the row's list.  The caller passes the holders of
// the left and of the right 32 columns in separate calls, so no two slots of a group work on the same row.
// One insertion is a chain of dependent warp collectives (SHFL -> VOTE -> REDUX, ~100+ cycles of latency and nothing else
// to issue), so INS_ROWS rows are processed side by side in branch-free lock step: their chains interleave.
#ifndef B200_INS_ROWS
#define B200_INS_ROWS 3
#endif
template <int E>
__device__ __forceinline__ void ts_insert_staged(unsigned hit, const int col0, float& thr, uint2* __restrict__ list,
                                                 volatile float* __restrict__ thr_pub, const float4* __restrict__ stg, const int lane, long long* stat) {
    constexpr int KEEP = 32 * E;
    constexpr int NR = B200_INS_ROWS;
    long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
    (void)tc0;
    if (stat) { stat[0] += 1; stat[1] += __popc(hit); if (__activemask() != 0xffffffffu) stat[8] += 1; tc0 = clock64(); }
    do {
        if (stat) tc1 = clock64();
        int L[NR];
        unsigned m[NR];
        uint32_t so[NR], curmax[NR];
        uint2 e0[NR], e1[NR];
        unsigned any = 0;
        float thrL[NR], sc[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            L[i] = 31 - __clz(hit);                      // highest hitting row first (FLO; ffs would need BREV + FLO); -1: slot unused,
            hit &= ~((L[i] >= 0 ? 1u : 0u) << (L[i] & 31));   // rides along on row L[0], never stored
        }
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int LL = L[i] < 0 ? L[0] : L[i];
            thrL[i] = __shfl_sync(0xffffffffu, thr, LL);
            sc[i] = ts_chunk_score(stg, LL, lane);
            e0[i] = list[(LL & 15) * KEEP + lane];      // holder lane LL carries 32 columns of row LL & 15
            e1[i] = make_uint2(0u, 0u);
            if (E == 2) e1[i] = list[(LL & 15) * KEEP + 32 + lane];
        }
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const unsigned mm = __ballot_sync(0xffffffffu, sc[i] < thrL[i]);
            m[i] = L[i] < 0 ? 0u : mm;
            so[i] = ord_bits(sc[i]);
            curmax[i] = ord_bits(thrL[i]);
            any |= m[i];
            if (stat) stat[2] += __popc(m[i]);
        }
        if (stat) { tc2 = clock64(); stat[5] += tc2 - tc1; }
        while (any) {   // warp-uniform: one candidate of every row per round, in phases so that the collectives of the rows overlap
            uint32_t sj[NR], idj[NR];
            unsigned b0[NR], b1[NR];
            bool act[NR];
            any = 0;
            if (stat) { stat[3] += 1; if (__activemask() != 0xffffffffu) stat[9] += 1; }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                act[i] = m[i] != 0u;
                const int j = act[i] ? 31 - __clz(m[i]) : 0;   // any order of the row's hits will do
                m[i] &= ~((act[i] ? 1u : 0u) << j);
                any |= m[i];
                idj[i] = (uint32_t)(col0 + ((L[i] >> 4) & 1) * 32 + j);
                sj[i] = __shfl_sync(0xffffffffu, so[i], j);
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                b0[i] = __ballot_sync(0xffffffffu, e0[i].x == curmax[i]);
                b1[i] = (E == 2) ? __ballot_sync(0xffffffffu, e1[i].x == curmax[i]) : 0u;
            }
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const bool doit = act[i] && sj[i] < curmax[i];
                // any holder of the maximum will do: the highest such lane.  sel_entry is inline PTX (setp + 2 selp): the
                // compiler cannot turn it back into a branch, so the rounds stay straight-line code and the rows' chains interleave.
                const bool in0 = doit && (E == 1 || b0[i] != 0u) && lane == 31 - __clz(b0[i]);
                sel_entry(in0, sj[i], idj[i], e0[i]);
                if (E == 2) {
                    const bool in1 = doit && b0[i] == 0u && lane == 31 - __clz(b1[i]);
                    sel_entry(in1, sj[i], idj[i], e1[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NR; ++i)   // unchanged when nothing was replaced
                curmax[i] = __reduce_max_sync(0xffffffffu, E == 1 ? e0[i].x : max(e0[i].x, e1[i].x));
        }
        if (stat) { tc3 = clock64(); stat[6] += tc3 - tc2; }
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (L[i] >= 0) {   // warp-uniform
                list[(L[i] & 15) * KEEP + lane] = e0[i];
                if (E == 2) list[(L[i] & 15) * KEEP + 32 + lane] = e1[i];
                if ((lane & 15) == (L[i] & 15)) thr = ord_float(curmax[i]);   // both holder lanes of the row
                if (lane == L[i]) thr_pub[L[i] & 15] = ord_float(curmax[i]);
            }
        }
        if (stat) stat[7] += clock64() - tc3;
    } while (hit);
    __syncwarp();   // the staging area is reused by the next chunk
}
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
    float m = __uint_as_float(v[0]);
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
    return m;
}

// ------------------------------------------------------------------------------------------------
// Per-thread epilogue of the TS kernel (EPI == 1; lists of 32 candidates per row, i.e. E == 1)
// ------------------------------------------------------------------------------------------------
// The replace-the-maximum lists above spend ~25 issue slots per insertion in chains of dependent warp collectives, and
// with cluster pruning nearly every scored tile belongs to the queries' own component (one insertion per row and tile
// or so).  Here nothing on the hot path is cooperative and the lists live in REGISTERS:
//  * every THREAD owns the scores it reads from TMEM (16x32bx2: lanes r and r + 16 hold 64 columns each of row r) and
//    an ascending list of its SL_KEEP = 24 best scores in as many registers (two threads per row: 48 kept for 32
//    candidates, enough for every k of E == 1).  A list entry is the order-preserving integer image of the
//    score with its low 5 bits replaced by a SLOT number; the reference id sits in that slot of a small per-thread table
//    in shared memory and never moves.  Inserting v is  l[i] = max(l[i-1], min(l[i], v))  for all i -- 46 integer
//    min/max instructions, no loads, no branches, no dependent chain (a lane with nothing to insert passes v = ~0);
//    the entry that drops out hands its slot to the new one.  (Measured alternatives: a binary heap in shared memory
//    needs ~800 cycles per insertion round at two warps per scheduler -- four dependent load-compare-store levels;
//    a sorted list in shared memory moves 16 KB per round and warp and made the kernel LSU-bound.)
//  * hot path, per 4 scores: two MINs, one compare, two address computations and three PREDICATED instructions that
//    push the whole quad as a record (4 scores + id of the first) onto the thread's record stack when its minimum is
//    below the thread's threshold -- no vote, no branch, no staging.  Store addresses are computed from the record
//    count into fresh registers: updating one pointer register in place made every store wait for the previous one to
//    leave the LSU queue (write-after-read on the address register, ~11 cycles per store, measured);
//  * records are popped by ALL lanes side by side (one record per lane and step); the scores of the record that are
//    still below the thread's threshold are inserted one per round, smallest first.  A step is taken after a tile when
//    at least half of the lanes have a record (SIMT efficiency), and before 32 scores are pushed whenever some lane
//    could run out of room.  Spreading the steps over the tiles matters: all eight warps must have read a tile before
//    its accumulator stage is recycled, so one warp's long drain stalls the whole pipeline;
//  * thresholds.  T = l[31] with the slot bits cleared is a lower bound (in the integer image) of every score this
//    thread ever dropped or refused, and it only decreases.  The thread's filter is min(own T, partner's T) as a float,
//    refreshed after every step (a stale threshold only admits more records);
//    (Tried: a tighter per-row filter, the R-th best score of both lists together.  R = 32 doubles the queries the
//    one-term tier cannot certify, R = 28 leaves 5 % uncertified: the certificate needs the margin the two per-list
//    thresholds give, about the 45th best score of the row.)
//  * at the end of the stream the two lists of a row are merged (one min/max step across the lanes): the 32 smallest
//    are the row's candidates and thr = min(both T, smallest dropped entry), so that every scanned reference that is
//    not a candidate has score >= thr -- the same certificate the re-rank expects.
#ifndef B200_SL_KEEP
#define B200_SL_KEEP 24
#endif
constexpr int SL_KEEP = B200_SL_KEEP;                // list entries (registers) and id slots per thread (>= the largest k of E == 1)
#ifndef B200_SL_ONECHECK
#define B200_SL_ONECHECK 2
#endif
constexpr int SL_PREC = B200_SL_ONECHECK ? 24 : 16;  // record stack entries per thread
#ifndef B200_SL_POP
#define B200_SL_POP 16
#endif
constexpr int SL_POP_LANES = B200_SL_POP;            // a step after a tile is worth it when this many lanes have a record
constexpr int SL_ID_BYTES = SL_KEEP * 32 * 4;        // per warp: id of slot s of lane l at [s][l]
constexpr int SL_RS_BYTES = SL_PREC * 32 * 16;       // record scores (float4 per record and lane)
constexpr int SL_RI_BYTES = SL_PREC * 32 * 4;        // record ids
constexpr int SL_WARP_BYTES = SL_ID_BYTES + SL_RS_BYTES + SL_RI_BYTES;
constexpr uint32_t SL_SLOT_MASK = 31u;
// Lists of 64 candidates per row (E == 2, k = 25 .. SL_KMAX2): 32 entries per thread, and the 2 x 32 entries of a row ARE its
// candidates (no merge at the end).  Larger k keeps the replace-the-maximum lists.
constexpr int SL_KEEP2 = 32;
constexpr int SL_KMAX2 = 30;
__host__ __device__ constexpr int sl_keep(int E) { return E == 1 ? SL_KEEP : SL_KEEP2; }
__host__ __device__ constexpr int sl_warp_bytes(int E) { return sl_keep(E) * 32 * 4 + SL_RS_BYTES + SL_RI_BYTES; }

// Pushes the quad (a, b, c, d) = scores of references id0 .. id0 + 3 when its minimum is below thr.
// cnt: this lane's record count; rs_addr / ri_addr: shared-memory addresses of this lane's record 0 (scores / id).
__device__ __forceinline__ void sl_push_quad(const float a, const float b, const float c, const float d, const float thr, const uint32_t id0,
                                             uint32_t& cnt, const uint32_t rs_addr, const uint32_t ri_addr) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f32 m;\n\t.reg .b32 as, ai;\n\t"
        "min.f32 m, %1, %2, %3;\n\t"
        "min.f32 m, m, %4;\n\t"
        "setp.lt.f32 p, m, %5;\n\t"
        "mad.lo.u32 as, %0, 512, %7;\n\t"
        "mad.lo.u32 ai, %0, 128, %8;\n\t"
        "@p st.shared.v4.f32 [as], {%1, %2, %3, %4};\n\t"
        "@p st.shared.u32 [ai], %6;\n\t"
        "@p add.u32 %0, %0, 1;\n\t}"
        : "+r"(cnt)
        : "f"(a), "f"(b), "f"(c), "f"(d), "f"(thr), "r"(id0), "r"(rs_addr), "r"(ri_addr)
        : "memory");
}
// The 32 scores v[] of references idb ...
__device__ __forceinline__ void sl_push32(const uint32_t (&v)[32], const float thr, const uint32_t idb, uint32_t& cnt, const uint32_t rs_addr,
                                          const uint32_t ri_addr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4)
        sl_push_quad(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]), thr, idb + i, cnt,
                     rs_addr, ri_addr);
}
// Inserts key v into the ascending register list (v = ~0: nothing happens); the largest entry drops out.
template <int LK>
__device__ __forceinline__ void sl_insert(uint32_t (&l)[LK], const uint32_t v) {
#pragma unroll
    for (int i = LK - 1; i >= 1; --i) l[i] = max(l[i - 1], min(l[i], v));
    l[0] = min(l[0], v);
}
// One step: every lane with a record pops its top record and inserts the scores that are below its threshold.
// cnt: this lane's record count; ids: this lane's slot table (slot s at ids[s * 32]); all lanes of the warp take part.
template <int LK>
__device__ __forceinline__ void sl_step(uint32_t (&l)[LK], uint32_t* ids, const float4* rs, const uint32_t* ri, uint32_t& cnt, const float thr,
                                        long long* stat) {
    const bool have = cnt > 0u;
    const uint32_t j = have ? cnt - 1u : 0u;
    float4 s = rs[j * 32];
    const uint32_t id0 = ri[j * 32];
    const float inf = __int_as_float(0x7f800000);
    if (!have) s = make_float4(inf, inf, inf, inf);
    cnt = j;
    float lim = thr;   // <= own T at all times
    if (stat) { stat[0] += 1; stat[1] += __popc(__ballot_sync(0xffffffffu, have)); }
#pragma unroll 1
    while (true) {
        const float cm = fminf(fmin3(s.x, s.y, s.z), s.w);
        const bool act = cm < lim;
        if (!__any_sync(0xffffffffu, act)) break;
        if (stat) { stat[3] += 1; stat[4] += __popc(__ballot_sync(0xffffffffu, act)); }
        const uint32_t q = (cm == s.x) ? 0u : ((cm == s.y) ? 1u : ((cm == s.z) ? 2u : 3u));
        s.x = (q == 0u) ? inf : s.x;
        s.y = (q == 1u) ? inf : s.y;
        s.z = (q == 2u) ? inf : s.z;
        s.w = (q == 3u) ? inf : s.w;
        const uint32_t slot = l[LK - 1] & SL_SLOT_MASK;               // the entry that drops out hands over its slot
        const uint32_t key = (ord_bits(cm) & ~SL_SLOT_MASK) | slot;    // < own T (multiple of 32) whenever act
        if (act) ids[slot * 32] = id0 + q;
        sl_insert(l, act ? key : 0xFFFFFFFFu);
        lim = fminf(lim, ord_float(l[LK - 1] & ~SL_SLOT_MASK));
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Split epilogue (EPI == 2): the pushers of EPI == 1 and their lists in different warps
// ------------------------------------------------------------------------------------------------
// EPI == 1 leaves every scheduler with two epilogue warps whose instructions each wait ~4 cycles on the one before
// (issue slots half used, ncu).  Here warps 2..9 only READ and PUSH (TMEM -> registers -> quad records), and eight more
// warps (11..18) only POP and INSERT: lane l of inserter warp w owns the register list of lane l of pusher warp w and is
// fed through a single-producer single-consumer ring of SP_RING records in shared memory (head written by the pusher
// after a block fence, tail by the inserter).  Four worker warps per scheduler instead of two; the work per tile is the
// same plus the ring bookkeeping.  Flow control: a pusher needs 8 free slots before 32 scores; an inserter pops one
// record per lane when at least half of its lanes have one, when some ring is more than half full (its pusher may be
// waiting), or when the stream has ended.
constexpr int SP_RING = 16;                               // records per ring (power of two)
constexpr int SP_THREADS = 608;                           // 19 warps
constexpr int SP_RS_BYTES = SP_RING * 32 * 16;
constexpr int SP_RI_BYTES = SP_RING * 32 * 4;
constexpr int SP_CTL_BYTES = 3 * 32 * 4 + 32;             // head[32], tail[32], thr[32], done
constexpr int SP_WARP_BYTES = SL_ID_BYTES + SP_RS_BYTES + SP_RI_BYTES + SP_CTL_BYTES;

// Pushes the quad when its minimum is below thr: record slot head & (SP_RING - 1).
__device__ __forceinline__ void sp_push_quad(const float a, const float b, const float c, const float d, const float thr, const uint32_t id0,
                                             uint32_t& head, const uint32_t rs_addr, const uint32_t ri_addr) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f32 m;\n\t.reg .b32 as, ai, t;\n\t"
        "min.f32 m, %1, %2, %3;\n\t"
        "min.f32 m, m, %4;\n\t"
        "setp.lt.f32 p, m, %5;\n\t"
        "and.b32 t, %0, 15;\n\t"
        "mad.lo.u32 as, t, 512, %7;\n\t"
        "mad.lo.u32 ai, t, 128, %8;\n\t"
        "@p st.shared.v4.f32 [as], {%1, %2, %3, %4};\n\t"
        "@p st.shared.u32 [ai], %6;\n\t"
        "@p add.u32 %0, %0, 1;\n\t}"
        : "+r"(head)
        : "f"(a), "f"(b), "f"(c), "f"(d), "f"(thr), "r"(id0), "r"(rs_addr), "r"(ri_addr)
        : "memory");
}
__device__ __forceinline__ void sp_push32(const uint32_t (&v)[32], const float thr, const uint32_t idb, uint32_t& head, const uint32_t rs_addr,
                                          const uint32_t ri_addr) {
    static_assert(SP_RING == 16, "sp_push_quad masks the record count with 15");
#pragma unroll
    for (int i = 0; i < 32; i += 4)
        sp_push_quad(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]), thr, idb + i, head,
                     rs_addr, ri_addr);
}

template <int E, int NBOX, int EPI>
__global__ void __launch_bounds__(EPI == 2 ? SP_THREADS : TS_THREADS, 1)
knn_candidates_ts_kernel(const __half* __restrict__ opA,   // [nq_pad][a_pitch] query operand rows (global); the first NBOX*64 columns are used
                         const int a_pitch,
                         const int32_t* __restrict__ qmap,  // optional: CTA row j serves query qmap[j] (second tier: the uncertified queries)
                         const int* __restrict__ qcount,    // with qmap: number of valid entries (device side, no host sync)
                         const __grid_constant__ CUtensorMap tmB, const MmaSched sched, const int nslot, const int64_t nq,
                         const int ntiles, const int tiles_per_split,
                         int32_t* __restrict__ cand_idx,   // [nsplit][nq][32E]
                         float* __restrict__ cand_score,   // [nsplit][nq][32E] (may be null)
                         float* __restrict__ thr_out,      // [nsplit][nq]
                         const int dbg_mode,               // measurement aid: 1 = stages recycled unread, 3 = no TMA, 4 = no MMAs
                         long long* __restrict__ dbg_ts,   // optional: per-warp cycle accounting of CTA (0,0)
                         const int trace_start, const PruneArgs P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    const int64_t nq_eff = qcount ? (int64_t)*qcount : nq;   // rows that exist; nq stays the stride of the output arrays
    if ((int64_t)blockIdx.x * BM >= nq_eff) return;          // whole CTA, before any barrier or TMEM allocation
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int A_COLS = NBOX * 32;                      // TMEM columns holding the query operand
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TS_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    constexpr int KEEP = 32 * E;

    uint8_t* smB = smem;
    // EPI == 0: [8 warps][16 rows][KEEP] (score bits, id) + [8 warps][8][32] float4 staging; EPI == 1: [8 warps][SL_WARP_BYTES]
    uint2* lists = reinterpret_cast<uint2*>(smB + (size_t)nslot * TS_B_BOX_BYTES);
    float4* stage_all = reinterpret_cast<float4*>(lists + (size_t)TS_EPI_WARPS * 16 * KEEP);
    float* thr_s = EPI == 2 ? reinterpret_cast<float*>(smB + (size_t)nslot * TS_B_BOX_BYTES + (size_t)TS_EPI_WARPS * SP_WARP_BYTES)
                 : EPI == 1 ? reinterpret_cast<float*>(smB + (size_t)nslot * TS_B_BOX_BYTES + (size_t)TS_EPI_WARPS * sl_warp_bytes(E))
                            : reinterpret_cast<float*>(stage_all + TS_EPI_WARPS * 8 * 32);                     // [128] row thresholds
    volatile int* ring = reinterpret_cast<int*>(thr_s + BM);                                                   // [TS_RING] tile ids, -1 = end
    float* qoff_s = reinterpret_cast<float*>(thr_s + BM) + TS_RING;                                            // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(qoff_s + BM);
    uint64_t* full = bars;                     // [MAX_SLOTS]
    uint64_t* empty = bars + MAX_SLOTS;        // [MAX_SLOTS]
    uint64_t* aready = bars + 2 * MAX_SLOTS;   // [1] query operand is in TMEM (4 warp arrivals)
    uint64_t* tfull = aready + 1;              // [TS_STAGES]
    uint64_t* tempty = tfull + TS_STAGES;      // [TS_STAGES] (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + TS_STAGES);

    const int tile0 = blockIdx.y * tiles_per_split;
    const int tile1 = min(ntiles, tile0 + tiles_per_split);
    const int m0 = blockIdx.x * BM;
#ifdef B200_TRACE   // per-warp cycle accounting of CTA (0,0): compiled in only for measurement builds (tools/build_variant.sh)
    const bool trace = dbg_ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#else
    constexpr bool trace = false;
    (void)dbg_ts;
#endif
    (void)trace_start;
    if (trace && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg_ts[192] = (long long)gt;
        dbg_ts[193] = clock64();
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < nslot; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        mbar_init(smem_u32(aready), 4);
        for (int i = 0; i < TS_STAGES; ++i) { mbar_init(smem_u32(&tfull[i]), 1); mbar_init(smem_u32(&tempty[i]), TS_EPI_WARPS); }
        fence_barrier_init();
    }
    if (threadIdx.x < BM) {
        thr_s[threadIdx.x] = __int_as_float(0x7f800000);
        qoff_s[threadIdx.x] = (P.cl_list && (int64_t)m0 + threadIdx.x < nq_eff) ? P.qoff[m0 + threadIdx.x] : __int_as_float(0xff800000);
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_base = tmem_base + (uint32_t)A_COLS;   // accumulator stage s at columns A_COLS + s*TS_BN

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The producer decides which tiles are scored and tells its consumers through the tile-id ring: entry seq % TS_RING
        // is written before the tile's first box is armed (the mbarrier arrive releases it), -1 ends the stream.  Every
        // entry -- the two end markers (one per MMA issuer) included -- occupies NBOX slots of the box ring.
        int slot = 0;
        uint32_t phase = 0;
        int seq = 0;
        auto emit = [&](const int tile) {
            if (lane == 0) {
                ring[seq & (TS_RING - 1)] = tile;
#pragma unroll
                for (int b = 0; b < NBOX; ++b) {
                    if (tile < 0 && b > 0) {   // an end marker arms only its first box (nobody consumes, hence frees, the others)
                        if (++slot == nslot) { slot = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_wait_parked(smem_u32(&empty[slot]), phase ^ 1);   // lane 0 only
                    if (tile < 0 || dbg_mode == 3) {
                        mbar_arrive(smem_u32(&full[slot]));
                    } else {
                        mbar_arrive_expect_tx(smem_u32(&full[slot]), (uint32_t)TS_B_BOX_BYTES);
                        tma_load_2d(smem_u32(smB + (size_t)slot * TS_B_BOX_BYTES), &tmB, smem_u32(&full[slot]), b * KBOX, tile * TS_BN);
                    }
                    if (++slot == nslot) { slot = 0; phase ^= 1; }
                }
            }
            ++seq;
            __syncwarp();
        };
        if (P.cl_list) {
            // Pruned search: clusters in ascending order of their lower bound; the stream ends at the first cluster whose
            // bound S^2 LB^2 is no smaller than S^2 d^2 of the threshold of EVERY row (thresholds only decrease, so a stale
            // read is merely lenient).  Unseen references then satisfy score >= threshold like every rejected one.
            const int2* list = P.cl_list + (size_t)blockIdx.x * P.C;
            const volatile float* thr_v = thr_s;
            bool stop = false;
            for (int ci = 0; ci < P.C && !stop; ++ci) {
                const int2 e = list[ci];
                const float lb = __int_as_float(e.y);
                const int t0 = P.cl_tile0[e.x], t1 = P.cl_tile0[e.x + 1];
                for (int t = t0; t < t1; ++t) {
                    if (lb > 0.f) {
                        float m = __int_as_float(0xff800000);
#pragma unroll
                        for (int r = lane; r < BM; r += 32) {
                            const float qo = qoff_s[r];
                            if (qo > __int_as_float(0xff800000)) m = fmaxf(m, __fadd_ru(thr_v[r], qo));
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                        if (lb >= m) { stop = true; break; }
                    }
                    emit(t);
                }
            }
        } else {
            for (int t = tile0; t < tile1; ++t) emit(t);
        }
        if (P.visited && lane == 0) atomicAdd(P.visited, (unsigned long long)seq);
        if (P.mma_count && lane == 0) atomicAdd(P.mma_count, (unsigned long long)seq * (unsigned long long)P.mma_per_tile);
        emit(-1);
        emit(-1);
    } else if (warp == 1 || warp == 10) {
        // ===================== MMA issuers (even / odd tiles) =====================
        uint32_t aT[NBOX][MAX_MMA_PER_BOX];   // TMEM address of the A slice of every scheduled MMA
        uint32_t oB[NBOX][MAX_MMA_PER_BOX];
        int nm[NBOX];
#pragma unroll
        for (int b = 0; b < NBOX; ++b) {
            nm[b] = sched.nmma[b];
#pragma unroll
            for (int m = 0; m < MAX_MMA_PER_BOX; ++m) {
                aT[b][m] = tmem_base + (uint32_t)sched.a_slice[b][m] * 8u;   // 16 fp16 of K = 8 columns per slice
                oB[b][m] = (uint32_t)sched.b_sub[b][m] * 2u;
            }
        }
        const uint64_t db_base = make_sw128_desc(smem_u32(smB));
        mbar_wait_u(smem_u32(aready), 0);
        tc_fence_after();
        const int mma_id = (warp == 1) ? 0 : 1;
        bool more = true;
        for (int tl = mma_id; more; tl += 2) {
            const int stage = tl % TS_STAGES;
            const uint32_t use = (uint32_t)(tl / TS_STAGES);
            mbar_wait_parked_u(smem_u32(&tempty[stage]), (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = acc_base + (uint32_t)(stage * TS_BN);
            const int box0 = tl * NBOX;
            int slot = box0 % nslot;
            uint32_t phase = (uint32_t)((box0 / nslot) & 1);
#pragma unroll
            for (int b = 0; b < NBOX; ++b) {
                mbar_wait_parked_u(smem_u32(&full[slot]), phase);
                tc_fence_after();
                if (b == 0 && ring[tl & (TS_RING - 1)] < 0) {   // end marker: pass it on to the epilogue, done
                    if (elect_one()) mbar_arrive(smem_u32(&tfull[stage]));
                    __syncwarp();
                    more = false;
                    break;
                }
                const uint64_t db0 = db_base + (uint64_t)((uint32_t)slot * (uint32_t)(TS_B_BOX_BYTES >> 4));
                if (elect_one()) {
                    if (dbg_mode != 4) {
                        const int nmb = (dbg_mode == 6) ? min(nm[b], 2) : nm[b];   // 6: measurement aid, fewer MMAs per tile
#pragma unroll
                        for (int m = 0; m < MAX_MMA_PER_BOX; ++m)
                            if (m < nmb) umma_f16_ts_impl(tmem_d, aT[b][m], db0 + oB[b][m], idesc, (b == 0 && m == 0) ? 0u : 1u);
                    }
                    umma_commit(smem_u32(&empty[slot]));
                    if (b == NBOX - 1) umma_commit(smem_u32(&tfull[stage]));
                }
                __syncwarp();
                if (++slot == nslot) { slot = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: warps 2..9 =====================
        // Two warps per TMEM lane quarter, 16 query rows each, all 128 columns of every tile: every row has ONE candidate
        // list and one owner, so nothing is shared between warps.  tcgen05.ld 16x32bx2 hands thread t < 16 the columns
        // [c, c+32) of row t and thread t + 16 the columns [c+32, c+64) of the same row; two such loads cover a tile.
        const bool is_ins = EPI == 2 && warp >= 11;       // EPI == 2: inserter warp 11 + w serves pusher warp 2 + w
        const int pw = is_ins ? warp - 9 : warp;          // the pusher warp of this pair
        const int grp = (pw - 2) >> 2;            // 0: rows 0..15 of the quarter, 1: rows 16..31
        const int q4 = pw & 3;                    // TMEM lane quarter the pusher may access
        const int row0 = q4 * 32 + grp * 16;      // first of this warp's 16 rows (CTA-relative)
        uint2* mylist = lists + (size_t)(pw - 2) * 16 * KEEP;                        // this warp's candidate lists
        float4* stg = stage_all + (size_t)(pw - 2) * 8 * 32;
        volatile float* thr_pub = thr_s + row0;                                      // read by the producer's stop test
        const uint32_t lane_acc = acc_base + ((uint32_t)row0 << 16);
        long long acc_t[4] = {0, 0, 0, 0};   // measurement aid (trace only): tfull wait, TMEM load, scan + hits
        long long stat[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // measurement aid (trace only), see ts_insert_staged
        uint32_t v0[32], v1[32];
        if constexpr (EPI == 0) {
#pragma unroll
            for (int i = 0; i < KEEP / 2; ++i) mylist[i * 32 + lane] = make_uint2(ORD_INF, 0xFFFFFFFFu);   // empty lists: score +inf, id -1
        }

        if (grp == 0 && !is_ins) {
            // this thread's query row -> TMEM (A operand of every MMA of this CTA); the warp covers the whole lane quarter
            const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
            const int64_t j = (int64_t)m0 + q4 * 32 + lane;
            int64_t src = j;                                       // rows past nq are zero padding of opA
            if (qmap) src = (j < nq_eff) ? (int64_t)qmap[j] : -1;
            const uint4* arow = reinterpret_cast<const uint4*>(opA + (size_t)(src < 0 ? 0 : src) * (size_t)a_pitch);
#pragma unroll
            for (int sl = 0; sl < NBOX * 4; ++sl) {
                uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
                if (src >= 0) { lo = __ldg(arow + 2 * sl); hi = __ldg(arow + 2 * sl + 1); }
                tmem_st8(tmem_base + lane_sel + (uint32_t)(sl * 8), lo, hi);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(aready));
        }
        __syncwarp();

        if constexpr (EPI == 2) {
            // ---------- split epilogue (see sp_* above) ----------
            uint8_t* wbase = reinterpret_cast<uint8_t*>(lists) + (size_t)(pw - 2) * SP_WARP_BYTES;
            uint32_t* ids = reinterpret_cast<uint32_t*>(wbase) + lane;                            // slot s at ids[s * 32]
            const float4* rs = reinterpret_cast<const float4*>(wbase + SL_ID_BYTES) + lane;       // record j at rs[j * 32]
            const uint32_t* ri = reinterpret_cast<const uint32_t*>(wbase + SL_ID_BYTES + SP_RS_BYTES) + lane;
            volatile uint32_t* headp = reinterpret_cast<volatile uint32_t*>(wbase + SL_ID_BYTES + SP_RS_BYTES + SP_RI_BYTES) + lane;
            volatile uint32_t* tailp = headp + 32;
            volatile float* thrp = reinterpret_cast<volatile float*>(headp + 64);
            volatile uint32_t* donep = reinterpret_cast<volatile uint32_t*>(wbase + SL_ID_BYTES + SP_RS_BYTES + SP_RI_BYTES) + 96;   // one word per pair
            const float inf = __int_as_float(0x7f800000);
            if (!is_ins) {
                *headp = 0u;
                *tailp = 0u;
                *thrp = inf;
                if (lane == 0) *donep = 0u;
            }
            // pair barrier: the control words are initialised before either side looks at them (named barrier 1 + pair)
            __syncwarp();
            asm volatile("bar.sync %0, 64;" ::"r"(1 + (pw - 2)) : "memory");
            if (!is_ins) {
                // ===== pusher =====
                const uint32_t rs_addr = smem_u32(rs), ri_addr = smem_u32(ri);
                uint32_t head = 0u;
                int seq = 0, stage = 0;
                uint32_t par = 0;
                while (true) {
                    mbar_wait_u(smem_u32(&tfull[stage]), par);
                    tc_fence_after();
                    const int tile = ring[seq & (TS_RING - 1)];
                    if (tile < 0) break;
                    const uint32_t tbase = lane_acc + (uint32_t)(stage * TS_BN);
                    tmem_ld16x2(tbase, v0);
                    tmem_ld16x2(tbase + 64u, v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tempty[stage]));
                    if (++stage == TS_STAGES) { stage = 0; par ^= 1u; }
                    ++seq;
                    if (dbg_mode == 1) continue;
                    const float thr = *thrp;
                    if (P.cl_list == nullptr && !__any_sync(0xffffffffu, fminf(chunk_min(v0), chunk_min(v1)) < thr)) continue;   // quiet tile of a dense scan
                    const uint32_t idb = (uint32_t)tile * (uint32_t)TS_BN + (uint32_t)(lane >> 4) * 32u;
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        // 32 scores = 8 quads need 8 free slots of the ring
                        if (__any_sync(0xffffffffu, head - *tailp > (uint32_t)(SP_RING - 8))) {
                            const long long t0 = clock64();
                            while (__any_sync(0xffffffffu, head - *tailp > (uint32_t)(SP_RING - 8))) {
                                __nanosleep(40);
                                if (clock64() - t0 > 20000000000LL) {
                                    if (lane == 0) printf("b200mnn: record ring watchdog fired (block %d warp %d)\n", blockIdx.x, warp);
                                    __trap();
                                }
                            }
                        }
                        if (h == 0) sp_push32(v0, thr, idb, head, rs_addr, ri_addr);
                        else sp_push32(v1, thr, idb + 64u, head, rs_addr, ri_addr);
                        __threadfence_block();   // the records before the count
                        *headp = head;
                    }
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) *donep = 1u;
                tc_fence_before();
            } else {
                // ===== inserter =====
                uint32_t l[SL_KEEP];   // ascending keys: (order image of the score & ~31) | slot; empty = image of +inf
#pragma unroll
                for (int i = 0; i < SL_KEEP; ++i) {
                    l[i] = ORD_INF | (uint32_t)i;
                    ids[i * 32] = 0xFFFFFFFFu;   // id -1
                }
                __syncwarp();
                uint32_t tail = 0u;
                float thr = inf;
                const long long t_start = clock64();
                while (true) {
                    const bool done = __any_sync(0xffffffffu, *donep != 0u);   // read BEFORE the counts: a set flag means the counts below are final
                    __threadfence_block();
                    const uint32_t avail = *headp - tail;
                    const unsigned have = __ballot_sync(0xffffffffu, avail != 0u);
                    const bool urgent = __any_sync(0xffffffffu, avail > (uint32_t)(SP_RING - 8));
                    if (have == 0u) {
                        if (done) break;
                        __nanosleep(20);
                        if (clock64() - t_start > 40000000000LL) __trap();
                        continue;
                    }
                    if (!(urgent || done || __popc(have) >= SL_POP_LANES)) { __nanosleep(20); continue; }
                    __threadfence_block();   // the counts before the records
                    // one record per lane that has one
                    const bool mine = avail != 0u;
                    const uint32_t j = tail & (uint32_t)(SP_RING - 1);
                    float4 sc = rs[j * 32];
                    const uint32_t id0 = ri[j * 32];
                    if (!mine) sc = make_float4(inf, inf, inf, inf);
                    tail += mine ? 1u : 0u;
                    float lim = thr;
#pragma unroll 1
                    while (true) {
                        const float cm = fminf(fmin3(sc.x, sc.y, sc.z), sc.w);
                        const bool act = cm < lim;
                        if (!__any_sync(0xffffffffu, act)) break;
                        const uint32_t q = (cm == sc.x) ? 0u : ((cm == sc.y) ? 1u : ((cm == sc.z) ? 2u : 3u));
                        sc.x = (q == 0u) ? inf : sc.x;
                        sc.y = (q == 1u) ? inf : sc.y;
                        sc.z = (q == 2u) ? inf : sc.z;
                        sc.w = (q == 3u) ? inf : sc.w;
                        const uint32_t slot = l[SL_KEEP - 1] & SL_SLOT_MASK;
                        const uint32_t key = (ord_bits(cm) & ~SL_SLOT_MASK) | slot;
                        if (act) ids[slot * 32] = id0 + q;
                        sl_insert(l, act ? key : 0xFFFFFFFFu);
                        lim = fminf(lim, ord_float(l[SL_KEEP - 1] & ~SL_SLOT_MASK));
                    }
                    __syncwarp();
                    *tailp = tail;   // the record slot may be reused
                    const float own = ord_float(l[SL_KEEP - 1] & ~SL_SLOT_MASK);
                    thr = fminf(own, __shfl_xor_sync(0xffffffffu, own, 16));
                    *thrp = thr;
                    if (lane < 16) thr_pub[lane] = thr;
                }
                __syncwarp();
                // Output: as EPI == 1, by the inserter (the keys go over the record ring, which is empty now)
                uint32_t* keys = reinterpret_cast<uint32_t*>(wbase + SL_ID_BYTES);   // [SL_KEEP entries][32 lanes]
#pragma unroll
                for (int i = 0; i < SL_KEEP; ++i) keys[i * 32 + lane] = l[i];
                __syncwarp();
                const uint32_t* idw = reinterpret_cast<const uint32_t*>(wbase);      // [32 slots][32 lanes]
                const int64_t rowbase = (int64_t)m0 + row0;
                const int64_t sbase = (int64_t)blockIdx.y * nq;
#pragma unroll 1
                for (int r = 0; r < 16; ++r) {
                    const int64_t row = rowbase + r;
                    if (row >= nq_eff) break;   // warp-uniform
                    const uint32_t ka = lane < SL_KEEP ? keys[lane * 32 + r] : 0xFFFFFFFFu;
                    const uint32_t kb = 31 - lane < SL_KEEP ? keys[(31 - lane) * 32 + r + 16] : 0xFFFFFFFFu;
                    const bool alo = ka <= kb;
                    const uint32_t lo = alo ? ka : kb, hi = alo ? kb : ka;
                    const uint32_t id = idw[(lo & SL_SLOT_MASK) * 32 + (alo ? r : r + 16)];
                    const uint32_t dropped = __reduce_min_sync(0xffffffffu, hi & ~SL_SLOT_MASK);
                    const uint32_t ta = __shfl_sync(0xffffffffu, l[SL_KEEP - 1], r) & ~SL_SLOT_MASK;
                    const uint32_t tb = __shfl_sync(0xffffffffu, l[SL_KEEP - 1], r + 16) & ~SL_SLOT_MASK;
                    const int64_t o = (sbase + row) * KEEP + lane;
                    cand_idx[o] = (int32_t)id;
                    if (cand_score) cand_score[o] = ord_float(lo & ~SL_SLOT_MASK);
                    if (lane == 0) thr_out[sbase + row] = ord_float(min(min(ta, tb), dropped));
                }
            }
        } else if constexpr (EPI == 1) {
            // ---------- per-thread epilogue (see sl_* above) ----------
            uint8_t* wbase = reinterpret_cast<uint8_t*>(lists) + (size_t)(warp - 2) * sl_warp_bytes(E);
            constexpr int LK = sl_keep(E);              // list entries per thread
            constexpr int ID_BYTES = LK * 32 * 4;
            uint32_t* ids = reinterpret_cast<uint32_t*>(wbase) + lane;                            // slot s at ids[s * 32]
            const float4* rs = reinterpret_cast<const float4*>(wbase + ID_BYTES) + lane;       // record j at rs[j * 32]
            const uint32_t* ri = reinterpret_cast<const uint32_t*>(wbase + ID_BYTES + SL_RS_BYTES) + lane;
            const uint32_t rs_addr = smem_u32(rs), ri_addr = smem_u32(ri);
            const float inf = __int_as_float(0x7f800000);
            uint32_t l[LK];   // ascending keys: (order image of the score & ~31) | slot; empty = image of +inf
#pragma unroll
            for (int i = 0; i < LK; ++i) {
                l[i] = ORD_INF | (uint32_t)i;
                ids[i * 32] = 0xFFFFFFFFu;   // id -1
            }
            __syncwarp();
            uint32_t cnt = 0;    // records on this lane's stack
            float thr = inf;     // what a score must beat to be recorded: min(own T, partner's T) as a float
            int seq = 0, stage = 0;
            uint32_t par = 0;
            auto step = [&]() {
                long long td = 0;
                if (trace) td = clock64();
                sl_step(l, ids, rs, ri, cnt, thr, trace ? stat : nullptr);
                const float own = ord_float(l[LK - 1] & ~SL_SLOT_MASK);
                thr = fminf(own, __shfl_xor_sync(0xffffffffu, own, 16));
                if (lane < 16) thr_pub[lane] = thr;
                if (trace) acc_t[3] += clock64() - td;
            };
            while (true) {
                long long c0 = 0, c1 = 0, c2 = 0;
                if (trace) c0 = clock64();
                mbar_wait_u(smem_u32(&tfull[stage]), par);
                tc_fence_after();
                const int tile = ring[seq & (TS_RING - 1)];
                if (tile < 0) break;
                if (trace) c1 = clock64();
                const uint32_t tbase = lane_acc + (uint32_t)(stage * TS_BN);
                tmem_ld16x2(tbase, v0);
                tmem_ld16x2(tbase + 64u, v1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tempty[stage]));
                if (++stage == TS_STAGES) { stage = 0; par ^= 1u; }
                ++seq;
                if (trace) { c2 = clock64(); acc_t[0] += c1 - c0; acc_t[1] += c2 - c1; }
                if (dbg_mode == 1) continue;
                // Dense scan (no cluster plan): most tiles hold no score below any threshold of the warp -- two min trees and
                // one vote skip them.  With a plan nearly every tile is one of the queries' own component and has hits, so
                // the test is not made.  (Tried: testing every 8th tile and staying in a tested mode while tiles come out
                // quiet, for weakly separated data.  One Gaussian blob gained 5 %; the mixture lost 7-10 %, because the
                // mode -- as a loop-carried flag or as two copies of the loop -- cost ptxas its proof that the votes of the
                // loop are convergent.)
                if (P.cl_list == nullptr && !__any_sync(0xffffffffu, fminf(chunk_min(v0), chunk_min(v1)) < thr)) continue;
                const uint32_t idb = (uint32_t)tile * (uint32_t)TS_BN + (uint32_t)(lane >> 4) * 32u;   // reference of v0[0]; v1[0] is 64 further
                // ONE copy of the step code (the loop is not unrolled); 32 scores = 8 quads need room for 8 records
#if B200_SL_ONECHECK == 2
                while (__any_sync(0xffffffffu, cnt > (uint32_t)(SL_PREC - 16))) step();
                sl_push32(v0, thr, idb, cnt, rs_addr, ri_addr);
                sl_push32(v1, thr, idb + 64u, cnt, rs_addr, ri_addr);
                if (__popc(__ballot_sync(0xffffffffu, cnt != 0u)) >= SL_POP_LANES) step();
#elif B200_SL_ONECHECK
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    // before the tile's 64 scores (16 quads need room for 16 records): steps until every lane has room;
                    // after the tile: one step if at least half of the lanes have a record (worth it)
                    bool need = (h == 0) ? __any_sync(0xffffffffu, cnt > (uint32_t)(SL_PREC - 16))
                                         : (__popc(__ballot_sync(0xffffffffu, cnt != 0u)) >= SL_POP_LANES);
                    while (need) {
                        step();
                        need = (h == 0) && __any_sync(0xffffffffu, cnt > (uint32_t)(SL_PREC - 16));
                    }
                    if (h == 0) {
                        sl_push32(v0, thr, idb, cnt, rs_addr, ri_addr);
                        sl_push32(v1, thr, idb + 64u, cnt, rs_addr, ri_addr);
                    }
                }
#else
#pragma unroll 1
                for (int h = 0; h < 3; ++h) {
                    // before a push: steps until every lane has room; after the tile: one step if at least half of the
                    // lanes have a record (worth it)
                    bool need = (h < 2) ? __any_sync(0xffffffffu, cnt > (uint32_t)(SL_PREC - 8))
                                        : (__popc(__ballot_sync(0xffffffffu, cnt != 0u)) >= SL_POP_LANES);
                    while (need) {
                        step();
                        need = (h < 2) && __any_sync(0xffffffffu, cnt > (uint32_t)(SL_PREC - 8));
                    }
                    if (h == 0) sl_push32(v0, thr, idb, cnt, rs_addr, ri_addr);
                    else if (h == 1) sl_push32(v1, thr, idb + 64u, cnt, rs_addr, ri_addr);
                }
#endif
                if (trace) acc_t[2] += clock64() - c2;   // includes the steps (acc_t[3])
            }
            __syncwarp();
            while (__any_sync(0xffffffffu, cnt != 0u)) step();
            if (trace && lane == 0) {
                for (int i = 0; i < 4; ++i) dbg_ts[(warp - 2) * 8 + i] = acc_t[i];
                for (int i = 0; i < 3; ++i) dbg_ts[(warp - 2) * 8 + 4 + i] = stat[i];
                dbg_ts[(warp - 2) * 8 + 7] = seq;
                for (int i = 0; i < 7; ++i) dbg_ts[64 + (warp - 2) * 8 + i] = stat[3 + i];
            }
            // Output.  The register lists go to shared memory (over the record stacks, which are empty now); row r of this
            // warp then merges the ascending lists of lanes r and r + 16: the 32 smallest of the 64 keys.
            static_assert(2 * SL_KEEP >= 32 && SL_KEEP <= 32 && SL_KEEP2 == 32, "the output below fills 32 / 64 slots from two lists");
            uint32_t* keys = reinterpret_cast<uint32_t*>(wbase + ID_BYTES);   // [LK entries][32 lanes]
#pragma unroll
            for (int i = 0; i < LK; ++i) keys[i * 32 + lane] = l[i];
            __syncwarp();
            const uint32_t* idw = reinterpret_cast<const uint32_t*>(wbase);      // [LK slots][32 lanes]
            const int64_t rowbase = (int64_t)m0 + row0;
            const int64_t sbase = (int64_t)blockIdx.y * nq;
#pragma unroll 1
            for (int r = 0; r < 16; ++r) {
                const int64_t row = rowbase + r;
                if (row >= nq_eff) break;   // warp-uniform
                const uint32_t ta = __shfl_sync(0xffffffffu, l[LK - 1], r) & ~SL_SLOT_MASK;
                const uint32_t tb = __shfl_sync(0xffffffffu, l[LK - 1], r + 16) & ~SL_SLOT_MASK;
                if constexpr (E == 2) {
                    // 64 candidates per row: both lists as they are
                    const uint32_t ka = keys[lane * 32 + r], kb = keys[lane * 32 + r + 16];
                    const int64_t o = (sbase + row) * KEEP + lane;
                    cand_idx[o] = (int32_t)idw[(ka & SL_SLOT_MASK) * 32 + r];
                    cand_idx[o + 32] = (int32_t)idw[(kb & SL_SLOT_MASK) * 32 + r + 16];
                    if (cand_score) { cand_score[o] = ord_float(ka & ~SL_SLOT_MASK); cand_score[o + 32] = ord_float(kb & ~SL_SLOT_MASK); }
                    if (lane == 0) thr_out[sbase + row] = ord_float(min(ta, tb));
                } else {
                    // a ascending in lanes 0 .. SL_KEEP-1, b descending in lanes 32-SL_KEEP .. 31, ~0 elsewhere: bitonic split
                    const uint32_t ka = lane < LK ? keys[lane * 32 + r] : 0xFFFFFFFFu;
                    const uint32_t kb = 31 - lane < LK ? keys[(31 - lane) * 32 + r + 16] : 0xFFFFFFFFu;
                    const bool alo = ka <= kb;
                    const uint32_t lo = alo ? ka : kb, hi = alo ? kb : ka;
                    const uint32_t id = idw[(lo & SL_SLOT_MASK) * 32 + (alo ? r : r + 16)];
                    const uint32_t dropped = __reduce_min_sync(0xffffffffu, hi & ~SL_SLOT_MASK);     // lower bound of the dropped scores
                    const int64_t o = (sbase + row) * KEEP + lane;
                    cand_idx[o] = (int32_t)id;                            // empty entries carry id -1
                    if (cand_score) cand_score[o] = ord_float(lo & ~SL_SLOT_MASK);
                    if (lane == 0) thr_out[sbase + row] = ord_float(min(min(ta, tb), dropped));
                }
            }
            tc_fence_before();
        } else {
        // Tiles arrive in the producer's order; entry seq of the tile-id ring names the tile in accumulator stage
        // seq % TS_STAGES (-1: end of the stream).  Quiet tiles (no score below any threshold of the warp's rows -- the
        // common case of a dense scan) stay a short dependent chain: wait, two TMEM loads, release, two min trees, ONE vote.
        int seq = 0, stage = 0;
        uint32_t par = 0;
        bool filled = false;                            // the first load of the stream went straight into the lists
        float thr = __int_as_float(0x7f800000);         // threshold of row (lane & 15) == max of its list once that is full
        const bool skip_read = dbg_mode == 1;
        while (true) {
            long long c0 = 0, c1 = 0, c2 = 0;
            if (trace) c0 = clock64();
            mbar_wait_u(smem_u32(&tfull[stage]), par);
            tc_fence_after();
            const int tile = ring[seq & (TS_RING - 1)];
            if (tile < 0) break;
            if (trace) c1 = clock64();
            if (!skip_read) {
                const uint32_t tbase = lane_acc + (uint32_t)(stage * TS_BN);
                tmem_ld16x2(tbase, v0);
                tmem_ld16x2(tbase + 64u, v1);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty[stage]));
            if (++stage == TS_STAGES) { stage = 0; par ^= 1u; }
            ++seq;
            if (trace) c2 = clock64();
            if (skip_read) continue;
            const int col0 = tile * TS_BN;
            const float ma = chunk_min(v0), mb = chunk_min(v1);
            if (filled && !__any_sync(0xffffffffu, fminf(ma, mb) < thr)) continue;   // quiet tile
            // ONE copy of the fill / insertion code for both loads (the loop is not unrolled): the hit path is long, and
            // with a copy per call site the eight epilogue warps thrashed the instruction cache.
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                unsigned h;
                if (!filled) {   // warp-uniform: the very first load of this warp
                    ts_stage_chunk(v0, stg, lane);
                    ts_fill_staged<E>(col0, mylist, stg, lane);
                    // E == 1: the lists hold the first 32 columns; the other 32 are ordinary hits.  E == 2: all 64 are in.
                    float mx = chunk_max(v0);
                    if (E == 2) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
                    else mx = __shfl_sync(0xffffffffu, mx, lane & 15);
                    thr = (dbg_mode == 8) ? __int_as_float(0xff800000) : mx;
                    if (lane < 16) thr_pub[lane] = thr;
                    filled = true;
                    if (E == 2) continue;
                    h = __ballot_sync(0xffffffffu, ma < thr) & 0xffff0000u;
                    if (h) ts_insert_staged<E>(h, col0, thr, mylist, thr_pub, stg, lane, trace ? stat : nullptr);
                    continue;
                }
                h = __ballot_sync(0xffffffffu, (c == 0 ? ma : mb) < thr);   // thr may just have tightened
                if (h == 0u) continue;
                if (c == 0) ts_stage_chunk(v0, stg, lane);
                else ts_stage_chunk(v1, stg, lane);
                // the two holder lanes of a row must not be merged side by side: first the left 32 columns, then the right
                if (h & 0x0000ffffu) ts_insert_staged<E>(h & 0x0000ffffu, col0 + 64 * c, thr, mylist, thr_pub, stg, lane, trace ? stat : nullptr);
                h = __ballot_sync(0xffffffffu, (c == 0 ? ma : mb) < thr) & 0xffff0000u;
                if (h) ts_insert_staged<E>(h, col0 + 64 * c, thr, mylist, thr_pub, stg, lane, trace ? stat : nullptr);
            }
            if (trace) { acc_t[0] += c1 - c0; acc_t[1] += c2 - c1; acc_t[2] += clock64() - c2; }
        }
        __syncwarp();
        if (trace && lane == 0) {
            for (int i = 0; i < 4; ++i) dbg_ts[(warp - 2) * 8 + i] = acc_t[i];
            for (int i = 0; i < 3; ++i) dbg_ts[(warp - 2) * 8 + 4 + i] = stat[i];
            dbg_ts[(warp - 2) * 8 + 7] = seq;
            for (int i = 0; i < 7; ++i) dbg_ts[64 + (warp - 2) * 8 + i] = stat[3 + i];
        }
        // Output: every warp writes its 16 lists as they are (the re-rank orders the candidates by their exact distances
        // anyway).  The list maximum -- this lane's thr, +inf while fewer than KEEP references were seen -- is the row's
        // threshold: a rejected or evicted score was >= the maximum at that time, which only ever decreases.
        const int64_t rowbase = (int64_t)m0 + row0;
        const int64_t sbase = (int64_t)blockIdx.y * nq;
#pragma unroll 1
        for (int r = 0; r < 16; ++r) {
            const int64_t row = rowbase + r;
            if (row >= nq_eff) break;   // warp-uniform
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const uint2 t = mylist[r * KEEP + 32 * e + lane];
                const int64_t o = (sbase + row) * KEEP + 32 * e + lane;
                cand_idx[o] = (int32_t)t.y;                       // empty entries carry id -1
                if (cand_score) cand_score[o] = ord_float(t.x);
            }
        }
        if (lane < 16 && rowbase + lane < nq_eff) thr_out[sbase + rowbase + lane] = thr;
        tc_fence_before();
        }   // EPI == 0
    }
    __syncthreads();
    if (trace && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        dbg_ts[194] = (long long)gt;
        dbg_ts[195] = clock64();
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// K0: row statistics.  One thread per row: squared norm in double (dimension order, unfused -- the same value the
// reference arithmetic would produce), plus the global max |x| and max ||x||^2 (non-negative IEEE values order like
// their bit patterns, so integer atomicMax does it).
// ------------------------------------------------------------------------------------------------
__global__ void rowstat_kernel(const double* __restrict__ X, int64_t n, int d, double* __restrict__ norm_f64,
                               unsigned int* __restrict__ absmax_bits, unsigned long long* __restrict__ maxnorm_bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float am = 0.f;
    double nrm = 0.0;
    if (i < n) {
        const double* x = X + i * d;
        for (int t = 0; t < d; ++t) {
            const double xv = x[t];
            nrm = __dadd_rn(nrm, __dmul_rn(xv, xv));
            const float a = fabsf((float)xv);
            am = (a > am) ? a : am;  // NaN never wins
        }
        norm_f64[i] = nrm;
    }
    unsigned long long nb = (nrm >= 0.0) ? (unsigned long long)__double_as_longlong(nrm) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, nb, o);
        nb = ob > nb ? ob : nb;
    }
    if ((threadIdx.x & 31) == 0) {
        if (am > 0.f) atomicMax(absmax_bits, __float_as_uint(am) + 1u);  // +1 ulp: never below the double value
        if (maxnorm_bits) atomicMax(maxnorm_bits, nb);
    }
}

// Scale exponent e (S = 2^e) such that |2 S x| < 2^13 for every element (fp16 range with headroom, low parts stay
// normal) and S^2 ||x||^2 < 2^30 for every reference row (so N / 2^15 fits fp16).
__global__ void scale_kernel(const unsigned int* __restrict__ absmax_bits, const unsigned long long* __restrict__ maxnorm_bits,
                             int* __restrict__ scale_exp) {
    const float m = __uint_as_float(*absmax_bits);
    int e = 0;
    if (m > 0.f && isfinite(m)) {
        int ex;
        frexpf(m, &ex);  // m < 2^ex
        e = 12 - ex;
        const double mn = sqrt(__longlong_as_double((long long)*maxnorm_bits)) * 1.0000001;
        if (mn > 0.0 && isfinite(mn)) {
            int en;
            frexp(mn, &en);  // ||x|| < 2^en
            e = min(e, 15 - en);
        }
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
                                    const int* __restrict__ scale_exp, __half* __restrict__ op, const double* __restrict__ norm_f64,
                                    float2* __restrict__ qerr,            // queries: (|a - ah|, |ah|) over the full-group dims, rounded up
                                    unsigned int* __restrict__ bmax_bits, // references: max |bh| and max |b - bh| (float bits, rounded up)
                                    const int32_t* __restrict__ rowmap    // optional: operand row i holds X[rowmap[i]] (-1: padding), i < n_pad
                                    ) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const int64_t srow = rowmap ? (int64_t)rowmap[i] : (i < n ? i : -1);
    const int KS = L.nbox * KBOX;
    uint4* row = reinterpret_cast<uint4*>(op + i * KS);
    __align__(16) __half hbuf[16];
    __align__(16) __half lbuf[16];
    if (srow < 0) {
        // padding: zero operand; a padded REFERENCE row gets an infinite norm so that it never becomes a candidate
        for (int c = 0; c < KS / 8; ++c) row[c] = make_uint4(0, 0, 0, 0);
        if (!IS_QUERY) {
#pragma unroll
            for (int j = 0; j < 16; ++j) hbuf[j] = __float2half(0.f);
            hbuf[L.norm_col] = __ushort_as_half((unsigned short)0x7C00);  // +inf
            const uint4* hp = reinterpret_cast<const uint4*>(hbuf);
            row[L.norm_slice * 2 + 0] = hp[0];
            row[L.norm_slice * 2 + 1] = hp[1];
        }
        return;
    }
    const double S = scalbn(1.0, *scale_exp);
    const double mul = IS_QUERY ? -2.0 * S : S;
    const double* x = X + srow * d;
    double hi2 = 0.0, lo2 = 0.0;   // squared norms of the hi parts and of the exact residuals (what the one-term schedule drops)
    for (int g = 0; g < L.groups; ++g) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int t = g * 16 + j;
            const double xv = (t < d) ? x[t] : 0.0;
            split_half(xv * mul, hbuf[j], lbuf[j]);
            const double hv = (double)__half2float(hbuf[j]);
            const double rv = xv * mul - hv;
            hi2 += hv * hv;
            lo2 += rv * rv;
        }
        const uint4* hp = reinterpret_cast<const uint4*>(hbuf);
        const uint4* lp = reinterpret_cast<const uint4*>(lbuf);
        row[g * 2 + 0] = hp[0];
        row[g * 2 + 1] = hp[1];
        row[(L.nprim + g) * 2 + 0] = lp[0];
        row[(L.nprim + g) * 2 + 1] = lp[1];
    }
    {
        const float fh = __double2float_ru(sqrt(hi2) * 1.0000001), fl = __double2float_ru(sqrt(lo2) * 1.0000001);
        if (IS_QUERY) {
            if (qerr) qerr[srow] = make_float2(fl, fh);
        } else if (bmax_bits) {   // non-negative floats order like their bit patterns
            atomicMax(bmax_bits + 0, __float_as_uint(fh));
            atomicMax(bmax_bits + 1, __float_as_uint(fl));
        }
    }
    // norm columns: constants on the query side, the split scaled norm on the reference side
    __half nc[3];
    if (IS_QUERY) {
        nc[0] = __float2half(32768.f);      // 2^15
        nc[1] = __float2half(16.f);         // 2^4
        nc[2] = __float2half(0.0078125f);   // 2^-7
    } else {
        const double N = norm_f64[srow] * S * S;
        nc[0] = __double2half(N * 0.000030517578125);                           // N / 2^15
        const double r1 = N - (double)__half2float(nc[0]) * 32768.0;
        nc[1] = __double2half(r1 * 0.0625);                                     // r1 / 2^4
        const double r2 = r1 - (double)__half2float(nc[1]) * 16.0;
        nc[2] = __double2half(r2 * 128.0);                                      // r2 / 2^-7
    }
    int written = L.groups;
    if (L.rem_slice >= 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) hbuf[j] = __float2half(0.f);
        const int r = L.rem;
        for (int j = 0; j < r; ++j) {
            __half h, l;
            split_half(x[L.groups * 16 + j] * mul, h, l);
            hbuf[j] = h;
            if (IS_QUERY) { hbuf[r + j] = h; hbuf[2 * r + j] = l; }
            else          { hbuf[r + j] = l; hbuf[2 * r + j] = h; }
        }
        if (L.norm_slice == L.rem_slice) { hbuf[L.norm_col] = nc[0]; hbuf[L.norm_col + 1] = nc[1]; hbuf[L.norm_col + 2] = nc[2]; }
        const uint4* hp = reinterpret_cast<const uint4*>(hbuf);
        row[written * 2 + 0] = hp[0];
        row[written * 2 + 1] = hp[1];
        ++written;
    }
    if (L.norm_slice != L.rem_slice) {
#pragma unroll
        for (int j = 0; j < 16; ++j) hbuf[j] = __float2half(0.f);
        hbuf[0] = nc[0]; hbuf[1] = nc[1]; hbuf[2] = nc[2];
        const uint4* hp = reinterpret_cast<const uint4*>(hbuf);
        row[written * 2 + 0] = hp[0];
        row[written * 2 + 1] = hp[1];
        ++written;
    }
    // written == L.nprim here; the lo slices follow, the rest of the last box is zero
    for (int sl = L.nslices; sl < L.nbox * 4; ++sl) { row[sl * 2 + 0] = make_uint4(0, 0, 0, 0); row[sl * 2 + 1] = make_uint4(0, 0, 0, 0); }
}

// ------------------------------------------------------------------------------------------------
// K3: exact fp64 re-rank + certificate.  One warp per query, one candidate per lane per round.
// ------------------------------------------------------------------------------------------------
constexpr int RR_WARPS = 4;
constexpr int RR_DCH = 16;  // dims staged per chunk (small staging tile + few registers: the gather is latency-bound, occupancy pays)
constexpr int RR_MAXROUNDS = MAX_SPLIT * 2;   // most candidate lists a query can have (splits x E)

__device__ __forceinline__ bool pair_less(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

template <int MAXR /* candidate lists of 32 per query this instance can hold */>
__global__ void __launch_bounds__(RR_WARPS * 32)
rerank_kernel(const double* __restrict__ X, const double* __restrict__ Q, int64_t nq, int d, int k,
              const int32_t* __restrict__ cand_idx, const float* __restrict__ thr, int nsplit, int ncand_per_split,
              const int* __restrict__ scale_exp, const double* __restrict__ qnorm, const unsigned long long* __restrict__ maxnorm_bits,
              int32_t* __restrict__ out_idx, double* __restrict__ out_dist, int* __restrict__ flag_count, int32_t* __restrict__ flag_list,
              double* __restrict__ dbg_d2 /* [nq][ncand] or null */,
              const int32_t* __restrict__ qmap, const int* __restrict__ qcount,   // optional: list slot j holds query qmap[j], j < *qcount
              const int fast /* candidates came from the one-term schedule: wider error bound */,
              const float2* __restrict__ qerr, const unsigned int* __restrict__ bmax_bits,
              const int32_t* __restrict__ refmap /* optional: candidate ids are rows of the grouped operand -> original rows */,
              double* __restrict__ flag_dk2 /* optional, parallel to flag_list: exact squared distance of the k-th candidate */) {
    __shared__ double stage[RR_WARPS][32][RR_DCH + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t jq = (int64_t)blockIdx.x * RR_WARPS + warp;   // slot in the candidate / threshold arrays
    if (jq >= (qcount ? (int64_t)*qcount : nq)) return;
    const int64_t q = qmap ? (int64_t)qmap[jq] : jq;            // the query itself
    if (q < 0) return;                                          // padding slot of the grouped query order
    const int ncand = nsplit * ncand_per_split;
    const int rounds = ncand / 32;
    double cd[MAXR];
    int ci[MAXR];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) { cd[r] = INFINITY; ci[r] = -1; }
    const double* qv = Q + q * d;

#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        if (r < rounds) {
            const int c = r * 32 + lane;
            const int sp = c / ncand_per_split, within = c % ncand_per_split;
            int id = cand_idx[((int64_t)sp * nq + jq) * ncand_per_split + within];
            if (refmap && id >= 0) id = refmap[id];
            ci[r] = id;
            double acc = 0.0;
            // Cooperative staging of the 32 candidate rows through shared memory, 16 dims (one 128-byte line per row) at a
            // time: each half-warp fetches one row per step.  The gather is latency-bound, so all 16 loads of a chunk are issued
            // before the first store, and the loads of the NEXT chunk are in flight while this chunk's distances accumulate (a
            // serial chain: dimension order is part of the contract).  (Tried: every lane reading the row of its own candidate
            // with 13 independent 16-byte loads in flight -- 4.8 instead of 3.5 ms per 1M queries: 32 scattered sectors per load
            // instruction cost more in L1 than the staging.)
            const int sub = lane & 15, hw = lane >> 4;
            double vv[16];
            auto fetch = [&](const int t0) {
                const int len = min(RR_DCH, d - t0);
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) {
                    const int rid = __shfl_sync(0xffffffffu, id, 2 * rr + hw);
                    vv[rr] = (rid >= 0 && sub < len) ? __ldg(X + (int64_t)rid * d + t0 + sub) : 0.0;
                }
            };
            fetch(0);
            for (int t0 = 0; t0 < d; t0 += RR_DCH) {
                const int len = min(RR_DCH, d - t0);
#pragma unroll
                for (int rr = 0; rr < 16; ++rr)
                    if (sub < len) stage[warp][2 * rr + hw][sub] = vv[rr];
                __syncwarp();
                if (t0 + RR_DCH < d) fetch(t0 + RR_DCH);   // warp-uniform
                for (int t = 0; t < len; ++t) {
                    const double df = __dsub_rn(qv[t0 + t], stage[warp][lane][t]);
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
                __syncwarp();
            }
            cd[r] = (id >= 0) ? acc : INFINITY;
            if (dbg_d2) dbg_d2[jq * ncand + c] = cd[r];
        }
    }
    double dk = INFINITY;
    if (rounds == 1 && k <= 32) {
        // One candidate per lane (the usual case): bitonic sort of the 32 (distance, index) pairs across the lanes, then
        // lane j < k holds the j-th neighbour.
        double sd = cd[0];
        int si = (ci[0] >= 0) ? ci[0] : 0x7fffffff;   // empty slots sort last
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, sd, j);
                const int oi = __shfl_xor_sync(0xffffffffu, si, j);
                const bool up = (lane & kk) == 0, low = (lane & j) == 0;
                const bool mine_first = pair_less(sd, si, od, oi);
                const bool keep = (mine_first == (up == low));
                sd = keep ? sd : od;
                si = keep ? si : oi;
            }
        }
        if (lane < k) {
            out_idx[q * k + lane] = (si == 0x7fffffff) ? -1 : si;
            if (out_dist) out_dist[q * k + lane] = sqrt(sd);
        }
        dk = __shfl_sync(0xffffffffu, sd, k - 1);
    } else {
    // k rounds of warp arg-min under (distance, index)
    for (int j = 0; j < k; ++j) {
        double bd = INFINITY;
        int bi = 0x7fffffff, br = -1;
#pragma unroll
        for (int r = 0; r < MAXR; ++r)
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
            for (int r = 0; r < MAXR; ++r)
                if (r == br) ci[r] = -1;
        }
        if (lane == 0) {
            out_idx[q * k + j] = (wi == 0x7fffffff) ? -1 : wi;
            if (out_dist) out_dist[q * k + j] = sqrt(wd);
        }
        dk = wd;
    }
    }
    // certificate: every non-candidate j has score_j >= min_split thr, and |score_j/S^2 + ||q||^2 - d2_j| <= eps
    if (lane == 0) {
        float tmin = __int_as_float(0x7f800000);
        for (int sp = 0; sp < nsplit; ++sp) tmin = fminf(tmin, thr[(int64_t)sp * nq + jq]);
        bool ok = true;
        if (tmin < __int_as_float(0x7f800000)) {
            const double inv = scalbn(1.0, -2 * (*scale_exp));
            const double qn = qnorm[q];
            const double M2 = __longlong_as_double((long long)*maxnorm_bits);
            double eps = 1.52587890625e-05 * (sqrt(qn * M2) + M2);  // 2^-16 (|q| M + M^2)
            // one-term schedule: with a = ah + ra, b = bh + rb over the full-group dims (ra, rb the exact residuals) the score
            // lacks ra.bh + ah.rb + ra.rb, so |error| <= |ra||bh| + |ah||rb| + |ra||rb| (Cauchy-Schwarz) with |ra|, |ah| of
            // THIS query and the maxima of |bh|, |rb| over the references -- all measured by the operand kernels.
            if (fast) {
                const float2 qe = qerr[q];
                const double BH = (double)__uint_as_float(bmax_bits[0]), BL = (double)__uint_as_float(bmax_bits[1]);
                eps += ((double)qe.x * BH + (double)qe.y * BL + (double)qe.x * BL) * inv * 1.001;
            }
            const double bound = (double)tmin * inv + qn - eps;
            ok = dk < bound;
        }
        if (!(ok)) {
            const int slot = atomicAdd(flag_count, 1);
            flag_list[slot] = (int32_t)q;
            if (flag_dk2) flag_dk2[slot] = dk;   // +inf when the query has fewer than k candidates
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4: exact rescue / generic path.  One block per flagged query, k rounds of block arg-min over the
// pairs strictly greater than the previous pick.  O(k n d) per query -- only for rare queries.
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_CAP = 1024;   // references a block can collect in its single-pass mode

// flag_dk2 (optional, parallel to flag_list): the exact squared distance of the query's k-th candidate -- an upper bound
// of its true k-th neighbour distance.  With it a block needs ONE pass over the references: it collects every reference
// within that bound (normally a few dozen: the neighbours plus the ties that defeated the certificate) and selects among
// those; only if more than RS_CAP qualify does it fall back to k full passes.
__global__ void __launch_bounds__(RS_THREADS)
rescue_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ Q, int d, int k,
              const int* __restrict__ flag_count, const int32_t* __restrict__ flag_list, const double* __restrict__ flag_dk2, int64_t nq_all,
              int32_t* __restrict__ out_idx, double* __restrict__ out_dist, const int use_smem,
              const int coop_limit, const int* __restrict__ coop_left /* optional [coop_limit]: 1 = not answered by the sliced path */) {
    __shared__ double sd[RS_THREADS / 32];
    __shared__ int si[RS_THREADS / 32];
    __shared__ double pick_d;
    __shared__ int pick_i;
    __shared__ double col_d[RS_CAP];
    __shared__ int col_i[RS_CAP];
    __shared__ int col_n;
    extern __shared__ double qs_sm[];  // [d] when use_smem
    const int count = flag_list ? *flag_count : (int)nq_all;
    for (int f = blockIdx.x; f < count; f += gridDim.x) {
        if (coop_left && f < coop_limit && coop_left[f] == 0) continue;   // answered by rescue_collect / rescue_select (block-uniform)
        const int64_t q = flag_list ? flag_list[f] : f;
        __syncthreads();
        const double* qs = use_smem ? qs_sm : Q + q * d;   // very wide rows are read through L1 instead
        if (use_smem)
            for (int t = threadIdx.x; t < d; t += blockDim.x) qs_sm[t] = Q[q * d + t];
        if (threadIdx.x == 0) { pick_d = -1.0; pick_i = -1; col_n = 0; }
        __syncthreads();
        // single-pass collection of everything within the bound
        bool collected = false;
        const double bound = (flag_list && flag_dk2) ? flag_dk2[f] : INFINITY;
        if (bound < INFINITY) {
            for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
                const double* xv = X + r * d;
                double acc = 0.0;
                for (int t = 0; t < d; ++t) {
                    const double df = __dsub_rn(qs[t], xv[t]);
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
                if (acc <= bound) {
                    const int pos = atomicAdd(&col_n, 1);
                    if (pos < RS_CAP) { col_d[pos] = acc; col_i[pos] = (int)r; }
                }
            }
            __syncthreads();
            collected = col_n <= RS_CAP;   // (col_n >= k always: the k candidates themselves qualify)
        }
        const int64_t nscan = collected ? (int64_t)col_n : n;
        for (int j = 0; j < k; ++j) {
            const double pd = pick_d;
            const int pi = pick_i;
            double bd = INFINITY;
            int bi = 0x7fffffff;
            for (int64_t r = threadIdx.x; r < nscan; r += blockDim.x) {
                double acc;
                int ri;
                if (collected) {
                    acc = col_d[r];
                    ri = col_i[r];
                } else {
                    const double* xv = X + r * d;
                    acc = 0.0;
                    for (int t = 0; t < d; ++t) {
                        const double df = __dsub_rn(qs[t], xv[t]);
                        acc = __dadd_rn(acc, __dmul_rn(df, df));
                    }
                    ri = (int)r;
                }
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

// A handful of flagged queries (a few exact ties in 1M cells) would leave rescue_kernel with one block per query, each
// streaming every reference on its own: ~11 ms per query at 1M x 50.  The first RSC_FMAX flagged queries are therefore
// answered by a SLICED scan: block (s, y) computes the exact distances of slice s of the references to the queries
// f = y, y + gridDim.y, ... and appends those within the query's bound to a global buffer; rescue_select_kernel then picks
// the k best per query under (distance, index).  A query without a bound, or with more than RS_CAP references inside it,
// is left to rescue_kernel (coop_left).
constexpr int RSC_FMAX = 1024;   // beyond this many flagged queries rescue_kernel has enough blocks of its own
constexpr int RSC_SLICES = 148;
constexpr int RSC_QROWS = 16;
__global__ void __launch_bounds__(RS_THREADS)
rescue_collect_kernel(const double* __restrict__ X, int64_t n, const double* __restrict__ Q, int d, const int* __restrict__ flag_count,
                      const int32_t* __restrict__ flag_list, const double* __restrict__ flag_dk2, int* __restrict__ gcount,
                      double* __restrict__ gd, int* __restrict__ gi) {
    extern __shared__ double qrow[];   // [d]
    const int count = min(*flag_count, RSC_FMAX);
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(n, r0 + chunk);
    for (int f = blockIdx.y; f < count; f += gridDim.y) {
        const double bound = flag_dk2[f];
        if (!(bound < INFINITY)) continue;   // block-uniform
        const int64_t q = flag_list[f];
        __syncthreads();
        for (int t = threadIdx.x; t < d; t += blockDim.x) qrow[t] = Q[q * d + t];
        __syncthreads();
        for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            const double* xv = X + r * d;
            double acc = 0.0;
            for (int t = 0; t < d; ++t) {
                const double df = __dsub_rn(qrow[t], xv[t]);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
            if (acc <= bound) {
                const int pos = atomicAdd(gcount + f, 1);
                if (pos < RS_CAP) { gd[(size_t)f * RS_CAP + pos] = acc; gi[(size_t)f * RS_CAP + pos] = (int)r; }
            }
        }
    }
}
__global__ void __launch_bounds__(RS_THREADS)
rescue_select_kernel(int k, const int* __restrict__ flag_count, const int32_t* __restrict__ flag_list, const double* __restrict__ flag_dk2,
                     const int* __restrict__ gcount, const double* __restrict__ gd, const int* __restrict__ gi, int* __restrict__ coop_left,
                     int32_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    __shared__ double sd[RS_THREADS / 32];
    __shared__ int si[RS_THREADS / 32];
    __shared__ double pick_d;
    __shared__ int pick_i;
    const int count = min(*flag_count, RSC_FMAX);
    for (int f = blockIdx.x; f < count; f += gridDim.x) {
        const int cn = gcount[f];
        const bool ok = flag_dk2[f] < INFINITY && cn <= RS_CAP && cn >= k;   // block-uniform
        if (threadIdx.x == 0) coop_left[f] = ok ? 0 : 1;
        if (!ok) continue;
        const int64_t q = flag_list[f];
        __syncthreads();
        if (threadIdx.x == 0) { pick_d = -1.0; pick_i = -1; }
        __syncthreads();
        for (int j = 0; j < k; ++j) {
            const double pd = pick_d;
            const int pi = pick_i;
            double bd = INFINITY;
            int bi = 0x7fffffff;
            for (int r = threadIdx.x; r < cn; r += blockDim.x) {
                const double acc = gd[(size_t)f * RS_CAP + r];
                const int ri = gi[(size_t)f * RS_CAP + r];
                const bool after = acc > pd || (acc == pd && ri > pi);   // strictly after the previous pick
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

__global__ void write_stats_kernel(const int* __restrict__ flag_count, int64_t* __restrict__ stats, int64_t lists, int64_t path,
                                   const int* __restrict__ tier2_count = nullptr, const unsigned long long* __restrict__ visited = nullptr,
                                   int64_t dense_tiles = 0) {
    stats[0] = flag_count ? *flag_count : 0;      // queries answered by the exact rescue kernel
    stats[1] = lists;
    stats[2] = path;                              // 1: tensor path, 2: tensor path with cluster pruning
    stats[3] = tier2_count ? *tier2_count : 0;    // queries the one-term tier could not certify (re-scored with three terms)
    stats[4] = visited ? (int64_t)visited[0] : 0; // 128x128 score tiles computed by the first tier ...
    stats[5] = visited ? (int64_t)visited[1] : 0; // ... and by the second
    stats[6] = dense_tiles;                       // tiles of a dense scan of the same problem
    stats[7] = 0;
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

static size_t ts_smem_bytes(int nslot, int E, int epi = 0) {
    const size_t lists = epi == 2 ? (size_t)TS_EPI_WARPS * SP_WARP_BYTES
                       : epi == 1 ? (size_t)TS_EPI_WARPS * sl_warp_bytes(E)
                                  : (size_t)TS_EPI_WARPS * 16 * (32 * E) * 8 + (size_t)TS_EPI_WARPS * 8 * 32 * 16;
    return (size_t)1024 /* alignment slack */ + (size_t)nslot * TS_B_BOX_BYTES + lists + (size_t)BM * 4 + (size_t)TS_RING * 4 + (size_t)BM * 4 +
           (size_t)(2 * MAX_SLOTS + 1 + 2 * TS_STAGES) * 8 + 16;
}
// TS variant: the operand must fit the TMEM columns next to the accumulators and at least nbox+1 reference boxes must
// fit in shared memory next to the candidate rows.
static bool ts_variant_fits(int nbox, int E, int epi = 0) { return nbox <= TS_MAX_NBOX && ts_smem_bytes(nbox + 1, E, epi) <= (size_t)232448; }

static size_t candidates_smem_bytes(int nbox, int nslot, int E) {
    return 1024 /* alignment slack */ + (size_t)nbox * A_BOX_BYTES + (size_t)nslot * B_BOX_BYTES + (size_t)(4 * (32 * E + 32) * ROWPITCH + 2) * 8 + (size_t)4 * 8 * 32 * 16 +
           (2 * MAX_SLOTS + 5) * 8 + 16;
}

bool tensor_path_supported(int64_t n, int64_t nq, int d, int k) {
    if (d < 1 || k < 1 || n < 1 || nq < 1) return false;
    if (k > MAX_K_TENSOR) return false;
    if (n > (int64_t)INT32_MAX - 512 || nq > (int64_t)INT32_MAX - 512) return false;
    const KLayout L = make_layout(d);
    if (L.nbox > MAX_NBOX) return false;
    if (ts_variant_fits(L.nbox, k <= 24 ? 1 : 2)) return true;
    // SS variant: the resident query operand, the candidate buffers and at least three reference boxes must fit in 227 KB
    return candidates_smem_bytes(L.nbox, 3, k <= 24 ? 1 : 2) <= (size_t)232448;
}

// Optional timing of the dominant kernel (bench.py's roofline figure): CUDA events recorded on the launch stream
// around every knn_candidates_kernel launch while enabled; collected (and synchronised) on demand.
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static double g_prof_flops = 0.0;
static unsigned long long* g_prof_mma = nullptr;   // device counter of issued MMAs (TS kernel), allocated on first use

int profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (auto& e : g_prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    g_prof_events.clear();
    g_prof_flops = 0.0;
    g_prof_on = on != 0;
    if (g_prof_on) {
        if (!g_prof_mma) B200_CUDA(cudaMalloc(&g_prof_mma, sizeof(unsigned long long)));
        B200_CUDA(cudaMemset(g_prof_mma, 0, sizeof(unsigned long long)));
    }
    return 0;
}

// Executed tensor-core work of the profiled launches: every tcgen05.mma of the TS kernel is M128 x N128 x K16.
int profile_collect_executed(double* executed_flops) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    unsigned long long h = 0;
    if (g_prof_mma) B200_CUDA(cudaMemcpy(&h, g_prof_mma, sizeof(h), cudaMemcpyDeviceToHost));
    if (executed_flops) *executed_flops = (double)h * 2.0 * BM * TS_BN * SLICE;
    return 0;
}

int profile_collect(double* total_ms, int64_t* launches, double* flops) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    double ms = 0.0;
    for (auto& e : g_prof_events) {
        B200_CUDA(cudaEventSynchronize(e.second));
        float t = 0.f;
        B200_CUDA(cudaEventElapsedTime(&t, e.first, e.second));
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)g_prof_events.size();
    if (flops) *flops = g_prof_flops;
    return 0;
}

struct DebugOut {
    int32_t* cand_idx = nullptr;  // [nq x ncand]
    double* cand_d2 = nullptr;    // approximate squared distances
    double* thr = nullptr;        // [nq]
    int64_t capacity = 0;         // entries available per query in the arrays above
    int64_t* ncand_host = nullptr;
};

// The query row is staged in shared memory when it fits next to the kernel's static buffers (opt-in up to the device
// limit); wider rows are read from global memory.
static int rescue_smem(int d, int* use_smem, size_t* bytes) {
    int dev = 0, max_smem = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t need = (size_t)std::max(d, 1) * sizeof(double);
    const size_t room = (size_t)max_smem > (size_t)16 * 1024 ? (size_t)max_smem - 16 * 1024 : 0;   // static: col_d, col_i, reductions
    *use_smem = need <= room ? 1 : 0;
    *bytes = *use_smem ? need : 0;
    if (*use_smem && need > 32 * 1024)
        B200_CUDA(cudaFuncSetAttribute(rescue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    return 0;
}

int launch_rescue(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, const int* flag_count, const int32_t* flag_list,
                  const double* flag_dk2, int32_t* d_idx, double* d_dist, cudaStream_t stream) {
    int use_smem = 0;
    size_t bytes = 0;
    B200_TRY(rescue_smem(d, &use_smem, &bytes));
    Scratch ws(stream);
    int* coop_left = nullptr;
    const int fmax = (int)std::min<int64_t>(nq, RSC_FMAX);
    if (flag_dk2 && flag_list && (size_t)d * sizeof(double) <= (size_t)40 * 1024 && !getenv("B200MNN_RESCUE_SERIAL")) {
        // sliced scan for the first RSC_FMAX flagged queries (see rescue_collect_kernel); no host synchronisation: the grids
        // are sized for the worst case and every block reads the flag count on the device
        int* gcount = ws.get<int>((size_t)fmax);
        coop_left = ws.get<int>((size_t)fmax);
        double* gd = ws.get<double>((size_t)fmax * RS_CAP);
        int* gi = ws.get<int>((size_t)fmax * RS_CAP);
        if (!ws.ok()) return B200MNN_ENOMEM;
        B200_CUDA(cudaMemsetAsync(gcount, 0, sizeof(int) * (size_t)fmax, stream));
        B200_CUDA(cudaMemsetAsync(coop_left, 0xFF, sizeof(int) * (size_t)fmax, stream));   // non-zero: left to rescue_kernel unless answered
        const dim3 grid((unsigned)std::min<int64_t>(RSC_SLICES, std::max<int64_t>(1, n / 1024)), (unsigned)std::min(fmax, RSC_QROWS), 1);
        rescue_collect_kernel<<<grid, RS_THREADS, (size_t)d * sizeof(double), stream>>>(dX, n, dQ, d, flag_count, flag_list, flag_dk2, gcount, gd, gi);
        B200_LAUNCH_CHECK();
        rescue_select_kernel<<<(unsigned)std::min(fmax, 4 * sm_count()), RS_THREADS, 0, stream>>>(k, flag_count, flag_list, flag_dk2, gcount, gd, gi,
                                                                                                 coop_left, d_idx, d_dist);
        B200_LAUNCH_CHECK();
    }
    rescue_kernel<<<sm_count() * 4, RS_THREADS, bytes, stream>>>(dX, n, dQ, d, k, flag_count, flag_list, flag_dk2, nq, d_idx, d_dist, use_smem,
                                                                coop_left ? fmax : 0, coop_left);
    B200_LAUNCH_CHECK();
    return 0;
}

int launch_rescue_all(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                      int64_t* d_stats, cudaStream_t stream) {
    int use_smem = 0;
    size_t bytes = 0;
    B200_TRY(rescue_smem(d, &use_smem, &bytes));
    const int grid = (int)std::min<int64_t>(nq, (int64_t)sm_count() * 8);
    rescue_kernel<<<grid, RS_THREADS, bytes, stream>>>(dX, n, dQ, d, k, nullptr, nullptr, nullptr, nq, d_idx, d_dist, use_smem, 0, nullptr);
    B200_LAUNCH_CHECK();
    if (d_stats) { write_stats_kernel<<<1, 1, 0, stream>>>(nullptr, d_stats, 0, 0); B200_LAUNCH_CHECK(); }
    return 0;
}

int write_stats(const int* flag_count, int64_t* d_stats, int64_t lists, int64_t path, cudaStream_t stream) {
    write_stats_kernel<<<1, 1, 0, stream>>>(flag_count, d_stats, lists, path);
    B200_LAUNCH_CHECK();
    return 0;
}

int query_knn_device(const double* dX, int64_t n, const double* dQ, int64_t nq, int d, int k, int32_t* d_idx, double* d_dist,
                     int64_t* d_stats, cudaStream_t stream, const DebugOut* dbg, RefCache* cache) {
    B200_TRY(ensure_device());
    if (n < 0 || nq < 0 || d < 0) return fail(B200MNN_EINVAL, "negative dimension");
    if (k < 0 || k > n) return fail(B200MNN_EINVAL, "'k' must be positive and no larger than the number of points in 'X'");
    if ((nq == 0 && !cache) || k == 0) return 0;   // with a cache, nq == 0 prepares the reference side only
    if (n > (int64_t)INT32_MAX - 512 || nq > (int64_t)INT32_MAX - 512) return fail(B200MNN_EINVAL, "more than 2^31 points are not supported");
    Scratch ws(stream);
    const bool ref_ready = cache && cache->ready;
    if (nq == 0 && ref_ready) return 0;
    if (ref_ready && (cache->dX != dX || cache->n != n || cache->d != d || cache->k != k))
        return fail(B200MNN_EINVAL, "internal: reference cache used with a different reference set");
    Scratch& rws = cache ? cache->ws : ws;   // owner of the reference-side buffers

    if (!tensor_path_supported(n, std::max<int64_t>(nq, 1), d, k) || d == 0) {
        if (nq == 0) return 0;   // nothing to prepare outside the tensor path
        if (dbg) return fail(B200MNN_EINVAL, "debug candidates requested for a shape outside the tensor path");
        // wide rows (gene space) or large k: K-streamed tensor-core scoring + selection (knn_wide.cu); B200MNN_WIDE=0 forces
        // the generic exact scan, which is also what remains for k beyond the candidate capacity
        const char* wenv = getenv("B200MNN_WIDE");
        if (d > 0 && wide_path_supported(n, nq, d, k) && !(wenv && atoi(wenv) == 0)) return query_knn_wide(dX, n, dQ, nq, d, k, d_idx, d_dist, d_stats, stream);
        return launch_rescue_all(dX, n, dQ, nq, d, k, d_idx, d_dist, d_stats, stream);
    }

    const KLayout L = make_layout(d);
    const MmaSched sched = make_sched(L, false);
    const int KS = L.nbox * KBOX;
    const int E = (k <= 24) ? 1 : 2;
    const int per = 32 * E;
    // TS variant (query operand in TMEM, 128-reference tiles, three accumulator stages) unless the operand is too wide
    // for the TMEM columns left next to the accumulators, or B200MNN_KERNEL=ss asks for the shared-memory variant.
    const char* kenv = getenv("B200MNN_KERNEL");
    const bool ss_fits = candidates_smem_bytes(L.nbox, 3, E) <= (size_t)232448;
    const bool use_ts = ts_variant_fits(L.nbox, E) && !(ss_fits && kenv && strcmp(kenv, "ss") == 0);
    const int bn = use_ts ? TS_BN : BN;
    // Epilogue of the TS kernel: per-thread heaps (lists of 32, the default) or the replace-the-maximum lists
    // (lists of 64, B200MNN_EPI=0, or when the heaps do not fit next to the reference boxes).
    const char* eenv = getenv("B200MNN_EPI");
    int epi = (use_ts && (E == 1 || k <= SL_KMAX2) && ts_variant_fits(L.nbox, E, 1) && !(eenv && atoi(eenv) == 0)) ? 1 : 0;
    if (epi == 1 && E == 1 && eenv && atoi(eenv) == 2) epi = 2;   // split epilogue (pushers and inserters in different warps): experimental
    // Pruned search (knn_cluster.cuh): reference rows grouped by a coarse k-means, query rows grouped by nearest centroid,
    // whole clusters skipped by a rigorous lower bound.  B200MNN_PRUNE=0 never, =1 whenever the shape allows, unset: for
    // searches large enough to pay for the clustering.  B200MNN_CLUSTERS sets the number of clusters (power of two).
    int nclusters = 64;
    if (const char* ce = getenv("B200MNN_CLUSTERS")) {
        const int c = atoi(ce);
        if (c >= 16 && c <= CL_MAXC && (c & (c - 1)) == 0) nclusters = c;
    }
    const char* penv = getenv("B200MNN_PRUNE");
    const bool prune_ok = use_ts && !dbg && n >= 8 * (int64_t)nclusters && ((size_t)CL_TILE * (d | 1) + (size_t)nclusters * d) * sizeof(double) <= (size_t)200 * 1024;
    const int64_t nq_plan = (cache && nq == 0) ? cache->nq_hint : nq;   // rows a prepare-only call expects per search
    const bool use_prune = ref_ready ? cache->use_prune : (prune_ok && (penv ? atoi(penv) == 1 : (n >= 65536 && nq_plan >= 16384)));
    const int64_t n_pad = use_prune ? round_up(n, CL_TILE) + (int64_t)nclusters * CL_TILE : round_up(n, bn);
    const int64_t nq_pad = round_up(nq, BM);
    const int64_t nslots = use_prune ? round_up(nq, CL_TILE) + (int64_t)nclusters * CL_TILE : nq;   // rows of the candidate / threshold arrays
    const int ntiles = (int)(n_pad / bn);
    const int mtiles = use_prune ? (int)(nslots / BM) : (int)(nq_pad / BM);
    int nsplit = 1;
    if (!use_prune && mtiles > 0 && mtiles < 2 * sm_count()) nsplit = (int)std::min<int64_t>(std::min<int64_t>(MAX_SPLIT, ntiles), ceil_div(2 * sm_count(), mtiles));
    const int tiles_per_split = (int)ceil_div(ntiles, nsplit);
    nsplit = (int)ceil_div(ntiles, tiles_per_split);

    // ---- reference side: computed once per reference set when a cache is given ----
    __half* opB = ref_ready ? static_cast<__half*>(cache->opB) : rws.get<__half>((size_t)n_pad * KS);
    double* xnorm = ref_ready ? cache->xnorm : rws.get<double>((size_t)n_pad);
    unsigned char* rscalars = ref_ready ? cache->scalars : rws.get<unsigned char>(64);
    if (!rws.ok()) return B200MNN_ENOMEM;
    unsigned int* absmax_bits = reinterpret_cast<unsigned int*>(rscalars);
    int* scale_exp = reinterpret_cast<int*>(rscalars + 8);
    unsigned long long* maxnorm_bits = reinterpret_cast<unsigned long long*>(rscalars + 16);
    unsigned int* bmax_bits = reinterpret_cast<unsigned int*>(rscalars + 32);   // [2]
    double* qnorm = nq > 0 ? ws.get<double>((size_t)nq_pad) : nullptr;
    if (!ws.ok()) return B200MNN_ENOMEM;
    ClusterPlan plan;
    if (ref_ready) plan = cache->plan;
    if (!ref_ready) {
        B200_CUDA(cudaMemsetAsync(rscalars, 0, 64, stream));
        rowstat_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(dX, n, d, xnorm, absmax_bits, maxnorm_bits);
        B200_LAUNCH_CHECK();
    }
    if (!cache) {   // one scale for both sides; with a cache the queries do not enter the scale (see RefCache)
        rowstat_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, stream>>>(dQ, nq, d, qnorm, absmax_bits, nullptr);
        B200_LAUNCH_CHECK();
    }
    if (!ref_ready) {
        scale_kernel<<<1, 1, 0, stream>>>(absmax_bits, maxnorm_bits, scale_exp);
        B200_LAUNCH_CHECK();
        if (use_prune) B200_TRY(build_ref_plan(dX, n, d, nclusters, xnorm, maxnorm_bits, rws, stream, &plan));
        prep_operand_kernel<false><<<(unsigned)ceil_div(n_pad, 128), 128, 0, stream>>>(dX, n, n_pad, d, L, scale_exp, opB, xnorm, nullptr, bmax_bits,
                                                                                      use_prune ? plan.refmap : nullptr);
        B200_LAUNCH_CHECK();
        if (cache) {
            cache->ready = true;
            cache->dX = dX;
            cache->n = n;
            cache->d = d;
            cache->k = k;
            cache->use_prune = use_prune;
            cache->opB = opB;
            cache->xnorm = xnorm;
            cache->scalars = rscalars;
            cache->plan = plan;
        }
    }
    if (nq == 0) return 0;   // reference side only (cache)

    // ---- query side ----
    __half* opA = ws.get<__half>((size_t)nq_pad * KS);
    int32_t* cand_idx = ws.get<int32_t>((size_t)nsplit * nslots * per);
    float* cand_score = dbg ? ws.get<float>((size_t)nsplit * nslots * per) : nullptr;
    float* thr = ws.get<float>((size_t)nsplit * nslots);
    int32_t* flag_list = ws.get<int32_t>((size_t)nq);
    int32_t* flag_list2 = ws.get<int32_t>((size_t)nq);
    double* flag_dk2 = ws.get<double>((size_t)nq);   // bound for the exact rescue (of whichever re-rank flags last)
    float2* qerr = ws.get<float2>((size_t)nq_pad);
    unsigned char* scalars = ws.get<unsigned char>(64);
    unsigned long long* visited = ws.get<unsigned long long>(2);
    if (!ws.ok()) return B200MNN_ENOMEM;
    int* flag_count = reinterpret_cast<int*>(scalars + 24);
    int* flag_count2 = reinterpret_cast<int*>(scalars + 28);
    B200_CUDA(cudaMemsetAsync(scalars, 0, 64, stream));
    B200_CUDA(cudaMemsetAsync(visited, 0, 16, stream));
    if (cache) {
        rowstat_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, stream>>>(dQ, nq, d, qnorm, reinterpret_cast<unsigned int*>(scalars), nullptr);
        B200_LAUNCH_CHECK();
    }
    int2* lists2 = nullptr;
    float* qoff2 = nullptr;
    int32_t* qmap2 = nullptr;   // second tier: the uncertified queries regrouped by cluster
    int* work2 = nullptr;
    if (use_prune) {
        B200_TRY(build_query_plan(&plan, dQ, nq, d, qnorm, scale_exp, maxnorm_bits, ws, stream));
        lists2 = ws.get<int2>((size_t)(nslots / CL_TILE) * plan.C);
        qoff2 = ws.get<float>((size_t)nslots);
        qmap2 = ws.get<int32_t>((size_t)nslots);
        work2 = ws.get<int>((size_t)3 * CL_MAXC + 4);
        if (!ws.ok()) return B200MNN_ENOMEM;
    }
    prep_operand_kernel<true><<<(unsigned)ceil_div(nq_pad, 128), 128, 0, stream>>>(dQ, nq, nq_pad, d, L, scale_exp, opA, nullptr, qerr, nullptr, nullptr);
    B200_LAUNCH_CHECK();

    // Thread-block clusters (optional, B200MNN_CLUSTER=2|4): the CTAs of a cluster work on different query tiles against
    // the SAME stream of reference boxes; each CTA fetches 1/csize of every box and TMA-multicasts it, dividing the
    // L2->SM traffic by csize.  Measured on B200 the kernel is NOT L2-bound (TMEM readout and the two-stage accumulator
    // hand-off are), and lock-stepping the CTAs of a cluster costs 20-25 %, so the default is no cluster.
    int csize = 1;
    if (const char* ce = getenv("B200MNN_CLUSTER")) {
        const int c = atoi(ce);
        if (c == 1 || c == 2 || c == 4) csize = c;
    }
    CUtensorMap tmA, tmB;
    B200_TRY(make_operand_map(&tmA, opA, nq_pad, KS, BM));
    if (use_ts) csize = 1;
    B200_TRY(make_operand_map(&tmB, opB, n_pad, KS, bn / csize));

    int dev = 0, max_smem = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const char* dbg_env = getenv("B200MNN_DEBUG_MODE");   // measurement aid only (results are void when non-zero)
    const int dbg_mode = dbg_env ? atoi(dbg_env) : 0;
    long long* dbg_ts = nullptr;
    const int trace_start = getenv("B200MNN_TRACE") ? atoi(getenv("B200MNN_TRACE")) : 0;
    if (getenv("B200MNN_TRACE")) {  // measurement aid: cycle accounting of CTA (0,0), printed after the launch
        dbg_ts = ws.get<long long>(64 * 32 + 64);
        if (!dbg_ts) return B200MNN_ENOMEM;
    }
    // Two tiers (TS variant only).  Tier 1 scores every query with the ONE-term schedule (2.5x fewer MMAs, half the
    // operand traffic at d = 50); its re-rank certifies with the correspondingly wider error bound.  Queries it cannot
    // certify are scored again by the three-term schedule (tier 2, reading its query rows through the flag list --
    // no host synchronisation, surplus CTAs exit at once), and what even that cannot certify goes to the exact
    // rescue kernel.  B200MNN_TIERS=1 (or a layout with no full 16-dim group) runs the three-term schedule alone.
    const char* tenv = getenv("B200MNN_TIERS");
    // The one-term tier certifies a query when its k-th distance clears the KEEP-th best score by the error bound:
    // worthwhile while k leaves a third of the kept candidates as margin.
    const bool two_tier = use_ts && !dbg && L.groups > 0 && 3 * k <= 2 * per && !(tenv && atoi(tenv) == 1);

    auto launch_candidates = [&](bool fast, const int32_t* qmap, const int* qcount, const PruneArgs& pargs, bool count_flops) -> int {
        const MmaSched sch = make_sched(L, fast);
        const int nb = fast ? L.nbox_fast : L.nbox;
        int nslot = MAX_SLOTS;
        size_t smem = 0;
        if (use_ts) {
            while (nslot > nb + 1 && ts_smem_bytes(nslot, E, epi) > (size_t)max_smem) --nslot;
            smem = ts_smem_bytes(nslot, E, epi);
        } else {
            while (nslot > 3 && candidates_smem_bytes(nb, nslot, E) > (size_t)max_smem) --nslot;
            smem = candidates_smem_bytes(nb, nslot, E);
        }
        if (smem > (size_t)max_smem) return fail(B200MNN_ECUDA, "device does not offer enough shared memory per block for the kNN kernel");
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_prof_mu);
            if (g_prof_on) {
                B200_CUDA(cudaEventCreate(&ev0));
                B200_CUDA(cudaEventCreate(&ev1));
                B200_CUDA(cudaEventRecord(ev0, stream));
            }
        }
        if (dbg_ts) B200_CUDA(cudaMemsetAsync(dbg_ts, 0, sizeof(long long) * (64 * 32 + 64), stream));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)round_up(mtiles, csize), (unsigned)nsplit, 1);   // padding CTAs see only zero-filled query rows
        cfg.blockDim = dim3(use_ts ? (epi == 2 ? SP_THREADS : TS_THREADS) : NUM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)csize;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const int64_t nq_c = nslots;
        const __half* opA_c = opA;
        const int a_pitch = KS;
#define B200_LAUNCH_CAND(EE, NB)                                                                                              \
    do {                                                                                                                       \
        B200_CUDA(cudaFuncSetAttribute(knn_candidates_kernel<EE, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        B200_CUDA(cudaLaunchKernelEx(&cfg, knn_candidates_kernel<EE, NB>, tmA, tmB, sch, nslot, nq_c, ntiles,                   \
                                     tiles_per_split, cand_idx, cand_score, thr, dbg_mode, csize, dbg_ts, trace_start));       \
    } while (0)
#define B200_LAUNCH_CAND_E(NB)               \
    do {                                     \
        if (E == 1) B200_LAUNCH_CAND(1, NB); \
        else B200_LAUNCH_CAND(2, NB);        \
    } while (0)
#define B200_LAUNCH_TS(EE, NB, EP)                                                                                                   \
    do {                                                                                                                              \
        B200_CUDA(cudaFuncSetAttribute(knn_candidates_ts_kernel<EE, NB, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        B200_CUDA(cudaLaunchKernelEx(&cfg, knn_candidates_ts_kernel<EE, NB, EP>, opA_c, a_pitch, qmap, qcount, tmB, sch, nslot, nq_c,  \
                                     ntiles, tiles_per_split, cand_idx, cand_score, thr, dbg_mode, dbg_ts, trace_start, pargs));      \
    } while (0)
#define B200_LAUNCH_TS_E(NB)                       \
    do {                                           \
        if (epi == 2) B200_LAUNCH_TS(1, NB, 2);    \
        else if (epi == 1 && E == 1) B200_LAUNCH_TS(1, NB, 1);    \
        else if (epi == 1) B200_LAUNCH_TS(2, NB, 1);    \
        else if (E == 1) B200_LAUNCH_TS(1, NB, 0); \
        else B200_LAUNCH_TS(2, NB, 0);             \
    } while (0)
        if (use_ts) {
            switch (nb) {
                case 1: B200_LAUNCH_TS_E(1); break;
                case 2: B200_LAUNCH_TS_E(2); break;
                case 3: B200_LAUNCH_TS_E(3); break;
                default: B200_LAUNCH_TS_E(4); break;
            }
        } else {
            switch (nb) {
                case 1: B200_LAUNCH_CAND_E(1); break;
                case 2: B200_LAUNCH_CAND_E(2); break;
                case 3: B200_LAUNCH_CAND_E(3); break;
                case 4: B200_LAUNCH_CAND_E(4); break;
                case 5: B200_LAUNCH_CAND_E(5); break;
                default: B200_LAUNCH_CAND_E(6); break;
            }
        }
#undef B200_LAUNCH_CAND_E
#undef B200_LAUNCH_CAND
#undef B200_LAUNCH_TS_E
#undef B200_LAUNCH_TS
        B200_LAUNCH_CHECK();
        if (ev0) {
            B200_CUDA(cudaEventRecord(ev1, stream));
            std::lock_guard<std::mutex> lock(g_prof_mu);
            g_prof_events.emplace_back(ev0, ev1);
            if (count_flops) g_prof_flops += 2.0 * (double)nq * (double)n * (double)d;   // algorithmic flops of the call, counted once
        }
        if (dbg_ts) {
            static long long h[64 * 32 + 64];
            B200_CUDA(cudaMemcpyAsync(h, dbg_ts, sizeof(h), cudaMemcpyDeviceToHost, stream));
            B200_CUDA(cudaStreamSynchronize(stream));
            fprintf(stderr, "--- %s schedule, %d box(es) per tile, %d slots ---\n", fast ? "one-term" : "three-term", nb, nslot);
            if (use_ts) {
                if (h[194] > h[192])
                    fprintf(stderr, "CTA (0,0): %lld cycles in %.1f us -> SM clock %.0f MHz while this kernel ran\n", h[195] - h[193], (h[194] - h[192]) / 1e3,
                            (double)(h[195] - h[193]) / ((h[194] - h[192]) / 1e3));
                for (int w = 0; w < TS_EPI_WARPS; ++w) {
                    const long long* a = h + w * 8;
                    const double nt = a[7] ? (double)a[7] : 1.0;
                    fprintf(stderr, "epilogue warp %d (%s half), cycles/tile: tfull wait %.0f, TMEM load %.0f, scan %.0f, merge %.0f; %lld tiles, %lld hit chunks, %lld hit rows, %lld keys\n",
                            w, w < 4 ? "left" : "right", a[0] / nt, a[1] / nt, a[2] / nt, a[3] / nt, a[7], a[4], a[5], a[6]);
                    const long long* m = h + 64 + w * 8;
                    fprintf(stderr, "    insertion: %lld rounds; cycles: staging %lld, row set-up %lld, rounds %lld, stores %lld; entered diverged %lld times, rounds diverged %lld\n",
                            m[0], m[1], m[2], m[3], m[4], m[5], m[6]);
                    const long long* ps = h + 128 + w * 4;
                    if (ps[3]) fprintf(stderr, "    in-situ dependent-chain latency: SHFL+IADD %.1f, LDS %.1f cycles/op (%lld probes)\n",
                                       (double)ps[0] / (32.0 * ps[3]), (double)ps[2] / (16.0 * ps[3]), ps[3]);
                }
            } else {
                const long long t00 = h[0];
                fprintf(stderr, "tile: mma[wait_tempty got_tempty got_full0 issued0 got_full1 issued1] epi[wait_tfull got_tfull released done] (cycles rel.)\n");
                for (int t = 0; t < 24; ++t) {
                    fprintf(stderr, "%5d:", t + trace_start);
                    for (int k2 = 0; k2 < 12; ++k2) if (k2 < 6 || k2 >= 8) fprintf(stderr, " %7lld", h[t * 32 + k2] ? h[t * 32 + k2] - t00 : -1LL);
                    fprintf(stderr, "  chunks:");
                    for (int k2 = 16; k2 < 24; ++k2) fprintf(stderr, " %5lld", h[t * 32 + k2] ? h[t * 32 + k2] - h[t * 32 + 9] : -1LL);
                    fprintf(stderr, " cnt0=%lld cnt1=%lld\n", h[t * 32 + 24], h[t * 32 + 25]);
                }
            }
        }
        return 0;
    };

    const bool fast_only_env = tenv && atoi(tenv) == 3;   // measurement aid: tier 1 only (uncertified queries go straight to the rescue)
    unsigned long long* mma_counter = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_prof_mu);
        if (g_prof_on) mma_counter = g_prof_mma;
    }
    auto mma_per_tile = [&](bool fast) { const MmaSched sc = make_sched(L, fast); int t = 0; for (int b = 0; b < sc.nbox; ++b) t += sc.nmma[b]; return t; };
    PruneArgs prune1 = {nullptr, nullptr, nullptr, 0, visited, mma_counter, mma_per_tile(two_tier)};
    PruneArgs prune2 = {nullptr, nullptr, nullptr, 0, visited + 1, mma_counter, mma_per_tile(false)};
    if (use_prune) {
        prune1 = PruneArgs{plan.cl_list, plan.cl_tile0, plan.qoff, plan.C, visited, mma_counter, mma_per_tile(two_tier)};
        prune2 = PruneArgs{lists2, plan.cl_tile0, qoff2, plan.C, visited + 1, mma_counter, mma_per_tile(false)};
    }
    const int32_t* qmap1 = use_prune ? plan.qmap : nullptr;
    const int* qcount1 = use_prune ? plan.nslots : nullptr;
    const int32_t* refmap = use_prune ? plan.refmap : nullptr;
    B200_TRY(launch_candidates(two_tier, qmap1, qcount1, prune1, true));

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

    if (dbg_mode != 0) {   // measurement modes leave garbage candidates: do not re-rank (or rescue) them, results are void
        B200_CUDA(cudaMemsetAsync(d_idx, 0, sizeof(int32_t) * (size_t)nq * k, stream));
        B200_CUDA(cudaMemsetAsync(d_dist, 0, sizeof(double) * (size_t)nq * k, stream));
        return 0;
    }
    const unsigned rr_grid = (unsigned)ceil_div(nslots, RR_WARPS);
    const int rr_lists = nsplit * per / 32;
    auto rr_kernel = rr_lists == 1 ? rerank_kernel<1> : (rr_lists == 2 ? rerank_kernel<2> : rerank_kernel<RR_MAXROUNDS>);
    rr_kernel<<<rr_grid, RR_WARPS * 32, 0, stream>>>(dX, dQ, nslots, d, k, cand_idx, thr, nsplit, per, scale_exp, qnorm, maxnorm_bits, d_idx,
                                                         d_dist, flag_count, flag_list, nullptr, qmap1, qcount1, two_tier ? 1 : 0, qerr, bmax_bits, refmap, flag_dk2);
    B200_LAUNCH_CHECK();
    const int* rescue_count = flag_count;
    const int32_t* rescue_list = flag_list;
    if (two_tier && !fast_only_env) {
        // tier 2: the flagged queries again, three-term schedule, same grid (CTAs past the flag count exit immediately)
        const int32_t* qm2 = flag_list;
        const int* qc2 = flag_count;
        if (use_prune) {
            // cluster-pure tiles again: a tile that mixed clusters would meet each row's own cluster late, with loose thresholds
            int* nslots2 = work2 + 3 * CL_MAXC;
            B200_TRY(regroup_query_list(plan, flag_list, flag_count, nq, qmap2, nslots, nslots2, work2, stream));
            B200_TRY(build_tile_lists(plan, dQ, d, qmap2, nslots2, nslots, qnorm, scale_exp, maxnorm_bits, lists2, qoff2, stream));
            qm2 = qmap2;
            qc2 = nslots2;
        }
        B200_TRY(launch_candidates(false, qm2, qc2, prune2, false));
        rr_kernel<<<rr_grid, RR_WARPS * 32, 0, stream>>>(dX, dQ, nslots, d, k, cand_idx, thr, nsplit, per, scale_exp, qnorm, maxnorm_bits, d_idx,
                                                             d_dist, flag_count2, flag_list2, nullptr, qm2, qc2, 0, qerr, bmax_bits, refmap, flag_dk2);
        B200_LAUNCH_CHECK();
        rescue_count = flag_count2;
        rescue_list = flag_list2;
    }
    B200_TRY(launch_rescue(dX, n, dQ, nq, d, k, rescue_count, rescue_list, flag_dk2, d_idx, d_dist, stream));
    if (d_stats) {
        write_stats_kernel<<<1, 1, 0, stream>>>(rescue_count, d_stats, nsplit, use_prune ? 2 : 1, two_tier ? flag_count : nullptr, use_ts ? visited : nullptr,
                                                ceil_div(nq, BM) * ceil_div(n, TS_BN));
        B200_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace knn
}  // namespace b200

extern "C" {

int b200mnn_profile_enable(int on) { return b200::knn::profile_enable(on); }

int b200mnn_profile_collect(double* total_ms, int64_t* launches, double* algorithmic_flops) {
    return b200::knn::profile_collect(total_ms, launches, algorithmic_flops);
}

int b200mnn_profile_collect_executed(double* executed_flops) { return b200::knn::profile_collect_executed(executed_flops); }

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
