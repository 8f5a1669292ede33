// K-streamed split-fp16 GEMM for sm_100a: TMA-fed shared-memory ring -> tcgen05.mma (kind::f16, fp32 accumulators in
// TMEM, M128 x N256 x K16) -> tcgen05.ld -> register re-basing -> fused epilogue.  See gemm_tc.cuh for the contract.
//
// One CTA per 128 x 256 output tile, 320 threads:
//   warp 0       TMA producer: per K box (64 fp16 columns) the A-hi, A-lo (128 rows) and B-hi, B-lo (256 rows) boxes of
//                the tile into one ring stage (96 KB; two stages).  One-term schedule: A-hi and B-hi only (48 KB; four).
//   warp 1       MMA issuer (one elected lane): per stage 4 K16 slices x {Ah.Bh, Al.Bh, Ah.Bl}; tcgen05.commit frees the
//                stage; after `chunk_boxes` stages commits the accumulator buffer to the epilogue and switches to the other
//                (2 x 256 TMEM columns).
//   warps 2..9   epilogue: warps w and w+4 share a TMEM lane quarter and split the 256 columns; every thread keeps the
//                128 running sums of its (row, column half) in registers and adds each accumulator chunk to them in fp32
//                (round to nearest), so no tensor-core accumulation chain is longer than 3 * 4 * chunk_boxes MMAs; after the
//                last chunk it applies the epilogue and writes its 512 contiguous bytes of the output row.
// Tiles are rasterised in 8 x 16 super-tiles so that a wave of CTAs re-uses ~40 MB of operand boxes out of L2; the
// bound of the three-term schedule is the L2->SM path (64 B/clk/SM needed at tensor peak vs ~42 B/clk/SM available),
// DESIGN.md section 4.3.
#include "gemm_tc.cuh"

#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "tc_ptx.cuh"

namespace b200 {
namespace gemm {

using namespace tc;

constexpr int A_BOX = BM * 128;   // 16 KB
constexpr int B_BOX = BN * 128;   // 32 KB
constexpr int THREADS = 320;
constexpr int SUPER_M = 8, SUPER_N = 16;

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

template <int NTERMS>
struct Cfg {
    static constexpr int STAGE_BYTES = (NTERMS == 3) ? 2 * A_BOX + 2 * B_BOX : A_BOX + B_BOX;
    static constexpr int NSTAGE = (NTERMS == 3) ? 2 : 4;
    static constexpr int OFF_AL = A_BOX;
    static constexpr int OFF_BH = (NTERMS == 3) ? 2 * A_BOX : A_BOX;
    static constexpr int OFF_BL = 2 * A_BOX + B_BOX;
    static constexpr size_t SMEM = 1024 + (size_t)NSTAGE * STAGE_BYTES + 256;
};

template <int NTERMS, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const int nkbox,
                  const int chunk_boxes, const int ntm, const int ntn, const EpiArgs ep) {
    using C = Cfg<NTERMS>;
    // tile of this CTA (super-tile rasterisation); surplus CTAs of ragged super-tiles leave before touching anything
    const int stn = (ntn + SUPER_N - 1) / SUPER_N;
    const int st = blockIdx.x / (SUPER_M * SUPER_N), wi = blockIdx.x % (SUPER_M * SUPER_N);
    const int tm = (st / stn) * SUPER_M + (wi % SUPER_M), tn = (st % stn) * SUPER_N + (wi / SUPER_M);
    if (tm >= ntm || tn >= ntn) return;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::NSTAGE * C::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::NSTAGE;
    uint64_t* tfull = bars + 2 * C::NSTAGE;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmAh);
        tma_prefetch_desc(&tmBh);
        if (NTERMS == 3) { tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBl); }
        for (int i = 0; i < C::NSTAGE; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull[i]), 1); mbar_init(smem_u32(&tempty[i]), 8); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int m0 = tm * BM, n0 = tn * BN;
    const int nchunks = (nkbox + chunk_boxes - 1) / chunk_boxes;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkbox; ++kb) {
                mbar_wait(smem_u32(&empty[slot]), phase ^ 1);
                const uint32_t bar = smem_u32(&full[slot]);
                mbar_arrive_expect_tx(bar, (uint32_t)C::STAGE_BYTES);
                const uint32_t base = smem_u32(smem + (size_t)slot * C::STAGE_BYTES);
                tma_load_2d(base, &tmAh, bar, kb * KBOX, m0);
                tma_load_2d(base + C::OFF_BH, &tmBh, bar, kb * KBOX, n0);
                if (NTERMS == 3) {
                    tma_load_2d(base + C::OFF_AL, &tmAl, bar, kb * KBOX, m0);
                    tma_load_2d(base + C::OFF_BL, &tmBl, bar, kb * KBOX, n0);
                }
                if (++slot == C::NSTAGE) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_f16(BM, BN);
        const uint64_t d0 = make_sw128_desc(smem_u32(smem));
        int slot = 0, kb = 0;
        uint32_t phase = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1;
            const uint32_t use = (uint32_t)(c >> 1);
            mbar_wait_u(smem_u32(&tempty[buf]), (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
            const int nb = min(chunk_boxes, nkbox - kb);
            for (int b = 0; b < nb; ++b) {
                mbar_wait_u(smem_u32(&full[slot]), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ds = d0 + (uint64_t)((uint32_t)slot * (uint32_t)(C::STAGE_BYTES >> 4));
                    const uint64_t dAh = ds, dAl = ds + (uint64_t)(C::OFF_AL >> 4), dBh = ds + (uint64_t)(C::OFF_BH >> 4),
                                   dBl = ds + (uint64_t)(C::OFF_BL >> 4);
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint64_t o = (uint64_t)(s * 2);   // 32 bytes per K16 slice, in 16-byte units
                        umma_f16_ss(tmem_d, dAh + o, dBh + o, idesc, (b | s) ? 1u : 0u);
                        if (NTERMS == 3) {
                            umma_f16_ss(tmem_d, dAl + o, dBh + o, idesc, 1u);
                            umma_f16_ss(tmem_d, dAh + o, dBl + o, idesc, 1u);
                        }
                    }
                    umma_commit(smem_u32(&empty[slot]));
                    if (b == nb - 1) umma_commit(smem_u32(&tfull[buf]));
                }
                __syncwarp();
                if (++slot == C::NSTAGE) { slot = 0; phase ^= 1; }
            }
            kb += nb;
        }
    } else {
        // ===================== epilogue =====================
        const int q4 = warp & 3;              // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;     // column half
        float acc[128];
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1;
            const uint32_t use = (uint32_t)(c >> 1);
            mbar_wait_u(smem_u32(&tfull[buf]), use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * BN + half * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)(j * 16), v);
                tmem_ld_wait();
                if (c == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[j * 16 + i] = __uint_as_float(v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[j * 16 + i] += __uint_as_float(v[i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty[buf]));
        }
        const int64_t row = (int64_t)m0 + q4 * 32 + lane;
        if (row < ep.M) {
            const int64_t cbase = (int64_t)n0 + half * 128;
            float4* orow = reinterpret_cast<float4*>(ep.out + row * ep.ldo + cbase);
            if (EPI == EPI_PLAIN) {
                const float al = (float)ep.alpha;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    orow[j] = make_float4(acc[4 * j] * al, acc[4 * j + 1] * al, acc[4 * j + 2] * al, acc[4 * j + 3] * al);
            } else if (EPI == EPI_SCORE) {
                const float al = (float)ep.alpha;   // a power of two: the product is exact, one rounding in the add
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float cv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const int64_t cc = cbase + 4 * j + i; cv[i] = (cc < ep.N) ? ep.colf[cc] : __int_as_float(0x7f800000); }
                    orow[j] = make_float4(fmaf(al, acc[4 * j], cv[0]), fmaf(al, acc[4 * j + 1], cv[1]), fmaf(al, acc[4 * j + 2], cv[2]),
                                          fmaf(al, acc[4 * j + 3], cv[3]));
                }
            } else {
                const double rd = ep.rowd[row];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t cc = cbase + 4 * j + i;
                        if (cc < ep.N) {
                            double d2 = (rd + ep.cold[cc]) - ep.beta * (double)acc[4 * j + i];
                            d2 = (ep.diag >= 0 && cc == ep.diag + row) ? 0.0 : fmax(d2, 0.0);
                            o[i] = (float)(-d2 * ep.inv_sigma - (ep.dens ? ep.dens[cc] : 0.0));
                        } else o[i] = __int_as_float(0xff800000);
                    }
                    orow[j] = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_half(double xs, __half& hi, __half& lo) {
    hi = __double2half(xs);
    lo = __double2half(xs - (double)__half2float(hi));
}

// one warp per output row
__global__ void __launch_bounds__(256)
split_rows_kernel(const double* __restrict__ X, int64_t ldx, const int32_t* __restrict__ gather, int64_t rows, int64_t rows_pad, int64_t K,
                  int64_t Kp, const double* __restrict__ centre, double mul, __half* __restrict__ hi, __half* __restrict__ lo,
                  double* __restrict__ norm2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + warp;
    if (r >= rows_pad) return;
    const int64_t src = (r < rows) ? (gather ? (int64_t)gather[r] : r) : -1;
    double s = 0.0;
    for (int64_t k = lane; k < Kp; k += 32) {
        double v = 0.0;
        if (src >= 0 && k < K) {
            v = X[src * ldx + k];
            if (centre) v -= centre[k];
        }
        s = fma(v, v, s);
        __half h, l;
        split_half(v * mul, h, l);
        hi[r * Kp + k] = h;
        lo[r * Kp + k] = l;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && norm2 && r < rows) norm2[r] = s;
}

// out row r (< rows), column k (< K) = X[k * ldx + r]; 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
split_transposed_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int64_t rows_pad, int64_t K, int64_t Kp, double mul,
                        __half* __restrict__ hi, __half* __restrict__ lo) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int64_t r0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 32;
    for (int j = ty; j < 32; j += 8) {
        const int64_t k = k0 + j, r = r0 + tx;
        tile[j][tx] = (k < K && r < rows) ? X[k * ldx + r] : 0.0;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int64_t r = r0 + j, k = k0 + tx;
        if (r < rows_pad && k < Kp) {
            __half h, l;
            split_half(tile[tx][j] * mul, h, l);
            hi[r * Kp + k] = h;
            lo[r * Kp + k] = l;
        }
    }
}

__global__ void __launch_bounds__(256)
col_sum_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int64_t K, double inv_rows, double* __restrict__ mean) {
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const int64_t r0 = (int64_t)blockIdx.y * 512, r1 = min(rows, r0 + 512);
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += X[r * ldx + k];
    atomicAdd(&mean[k], s * inv_rows);
}

__global__ void __launch_bounds__(256)
row_stats_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int64_t K, const double* __restrict__ centre,
                 unsigned int* __restrict__ absmax_bits, unsigned long long* __restrict__ maxnorm2_bits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + warp;
    if (r >= rows) return;
    double s = 0.0, m = 0.0;
    for (int64_t k = lane; k < K; k += 32) {
        double v = X[r * ldx + k];
        if (centre) v -= centre[k];
        s = fma(v, v, s);
        m = fmax(m, fabs(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if (lane == 0) {
        atomicMax(absmax_bits, __float_as_uint(__double2float_ru(m)));            // non-negative floats order like their bits
        atomicMax(maxnorm2_bits, (unsigned long long)__double_as_longlong(s));
    }
}

int col_mean(const double* X, int64_t ldx, int64_t rows, int64_t K, double* mean, cudaStream_t stream) {
    B200_CUDA(cudaMemsetAsync(mean, 0, sizeof(double) * (size_t)std::max<int64_t>(K, 1), stream));
    if (rows <= 0 || K <= 0) return 0;
    dim3 grid((unsigned)ceil_div(K, 256), (unsigned)ceil_div(rows, 512));
    col_sum_kernel<<<grid, 256, 0, stream>>>(X, ldx, rows, K, 1.0 / (double)rows, mean);
    B200_LAUNCH_CHECK();
    return 0;
}

int row_stats(const double* X, int64_t ldx, int64_t rows, int64_t K, const double* centre, unsigned int* absmax_bits,
              unsigned long long* maxnorm2_bits, cudaStream_t stream) {
    if (rows <= 0) return 0;
    row_stats_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, stream>>>(X, ldx, rows, K, centre, absmax_bits, maxnorm2_bits);
    B200_LAUNCH_CHECK();
    return 0;
}

int pick_scale_exp(float absmax, double maxnorm2) {
    int e = 0;
    if (absmax > 0.f && std::isfinite(absmax)) {
        int ex;
        frexpf(absmax, &ex);   // absmax < 2^ex
        e = 12 - ex;
        const double mn = sqrt(maxnorm2) * 1.0000001;
        if (mn > 0.0 && std::isfinite(mn)) {
            int en;
            frexp(mn, &en);    // ||x|| < 2^en
            e = std::min(e, 15 - en);
        }
    }
    return e;
}

int alloc_split(Scratch& ws, int64_t rows, int64_t K, SplitMat* out) {
    out->rows = rows;
    out->rows_pad = pad_rows(rows);
    out->K = K;
    out->Kp = pad_k(K);
    out->hi = ws.get<__half>((size_t)out->rows_pad * out->Kp);
    out->lo = ws.get<__half>((size_t)out->rows_pad * out->Kp);
    return ws.ok() ? 0 : B200MNN_ENOMEM;
}

int split_rows(const double* X, int64_t ldx, const int32_t* gather, int64_t rows, int64_t K, const double* centre, int scale_exp,
               const SplitMat& out, double* norm2, cudaStream_t stream) {
    split_rows_kernel<<<(unsigned)ceil_div(out.rows_pad, 8), 256, 0, stream>>>(X, ldx, gather, rows, out.rows_pad, K, out.Kp, centre,
                                                                               scalbn(1.0, scale_exp), out.hi, out.lo, norm2);
    B200_LAUNCH_CHECK();
    return 0;
}

int split_transposed(const double* X, int64_t ldx, int64_t rows, int64_t K, int scale_exp, const SplitMat& out, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(out.rows_pad, 32), (unsigned)ceil_div(out.Kp, 32));
    split_transposed_kernel<<<grid, 256, 0, stream>>>(X, ldx, rows, out.rows_pad, K, out.Kp, scalbn(1.0, scale_exp), out.hi, out.lo);
    B200_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

static int make_map(CUtensorMap* map, const __half* base, int64_t rows_pad, int64_t Kp, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(B200MNN_ECUDA, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows_pad};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * sizeof(__half)};
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

static std::mutex g_mu;
static bool g_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;
static double g_flops = 0.0;

int profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& e : g_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    g_events.clear();
    g_flops = 0.0;
    g_on = on != 0;
    return 0;
}

int profile_collect(double* total_ms, int64_t* launches, double* executed_flops) {
    std::lock_guard<std::mutex> lock(g_mu);
    double tot = 0.0;
    for (auto& e : g_events) {
        B200_CUDA(cudaEventSynchronize(e.second));
        float ms = 0.f;
        B200_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = (int64_t)g_events.size();
    if (executed_flops) *executed_flops = g_flops;
    return 0;
}

template <int NTERMS, int EPI>
static int launch(const CUtensorMap& tAh, const CUtensorMap& tAl, const CUtensorMap& tBh, const CUtensorMap& tBl, int nkbox, int chunk_boxes,
                  int ntm, int ntn, const EpiArgs& ep, cudaStream_t stream) {
    const size_t smem = Cfg<NTERMS>::SMEM;
    B200_CUDA(cudaFuncSetAttribute(gemm_split_kernel<NTERMS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t stm = ceil_div(ntm, SUPER_M), stn = ceil_div(ntn, SUPER_N);
    const int64_t grid = stm * stn * SUPER_M * SUPER_N;
    if (grid > (int64_t)INT32_MAX) return fail(B200MNN_EINVAL, "GEMM grid too large");
    gemm_split_kernel<NTERMS, EPI><<<(unsigned)grid, THREADS, smem, stream>>>(tAh, tAl, tBh, tBl, nkbox, chunk_boxes, ntm, ntn, ep);
    B200_LAUNCH_CHECK();
    return 0;
}

int gemm_split(const SplitMat& A, const SplitMat& B, int terms, int epilogue, const EpiArgs& ep, int chunk_boxes, cudaStream_t stream) {
    if (A.Kp != B.Kp || A.Kp % KBOX) return fail(B200MNN_EINVAL, "internal: GEMM operands disagree on K");
    if (ep.M <= 0 || ep.N <= 0) return 0;
    if (ep.M > A.rows_pad || ep.N > B.rows_pad) return fail(B200MNN_EINVAL, "internal: GEMM operand rows");
    const int ntm = (int)ceil_div(ep.M, BM), ntn = (int)ceil_div(ep.N, BN);
    if (ep.ldo < (int64_t)ntn * BN || (ep.ldo & 3)) return fail(B200MNN_EINVAL, "internal: GEMM output pitch");
    const int nkbox = (int)(A.Kp / KBOX);
    if (chunk_boxes < 1) chunk_boxes = nkbox;
    CUtensorMap tAh, tAl, tBh, tBl;
    B200_TRY(make_map(&tAh, A.hi, A.rows_pad, A.Kp, BM));
    B200_TRY(make_map(&tAl, A.lo ? A.lo : A.hi, A.rows_pad, A.Kp, BM));
    B200_TRY(make_map(&tBh, B.hi, B.rows_pad, B.Kp, BN));
    B200_TRY(make_map(&tBl, B.lo ? B.lo : B.hi, B.rows_pad, B.Kp, BN));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (g_on) {
            B200_CUDA(cudaEventCreate(&ev0));
            B200_CUDA(cudaEventCreate(&ev1));
            B200_CUDA(cudaEventRecord(ev0, stream));
        }
    }
    int rc;
#define B200_GEMM_CASE(T, E) rc = launch<T, E>(tAh, tAl, tBh, tBl, nkbox, chunk_boxes, ntm, ntn, ep, stream)
    if (terms == 3) {
        if (epilogue == EPI_PLAIN) B200_GEMM_CASE(3, EPI_PLAIN);
        else if (epilogue == EPI_SCORE) B200_GEMM_CASE(3, EPI_SCORE);
        else B200_GEMM_CASE(3, EPI_LOGIT);
    } else {
        if (epilogue == EPI_PLAIN) B200_GEMM_CASE(1, EPI_PLAIN);
        else if (epilogue == EPI_SCORE) B200_GEMM_CASE(1, EPI_SCORE);
        else B200_GEMM_CASE(1, EPI_LOGIT);
    }
#undef B200_GEMM_CASE
    if (rc) return rc;
    if (ev0) {
        B200_CUDA(cudaEventRecord(ev1, stream));
        std::lock_guard<std::mutex> lock(g_mu);
        g_events.emplace_back(ev0, ev1);
        g_flops += 2.0 * (double)ntm * BM * (double)ntn * BN * (double)A.Kp * (terms == 3 ? 3.0 : 1.0);
    }
    return 0;
}

}  // namespace gemm
}  // namespace b200

// Debug/validation hook: out[M x ldo] (fp32) = A[M x K] . B[N x K]^T through the split-fp16 tensor-core GEMM.
extern "C" int b200mnn_dev_debug_gemm(const double* dA, int64_t M, const double* dB, int64_t N, int64_t K, int terms, int chunk_boxes,
                                      float* d_out, int64_t ldo, void* stream_) {
    using namespace b200;
    using namespace b200::gemm;
    B200_TRY(ensure_device());
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (M <= 0 || N <= 0 || K <= 0) return fail(B200MNN_EINVAL, "empty GEMM");
    Scratch ws(stream);
    SplitMat A, B;
    B200_TRY(alloc_split(ws, M, K, &A));
    B200_TRY(alloc_split(ws, N, K, &B));
    unsigned char* sc = ws.get<unsigned char>(32);
    if (!ws.ok()) return B200MNN_ENOMEM;
    B200_CUDA(cudaMemsetAsync(sc, 0, 32, stream));
    unsigned int* amax = reinterpret_cast<unsigned int*>(sc);
    unsigned long long* nmax = reinterpret_cast<unsigned long long*>(sc + 8);
    B200_TRY(row_stats(dA, K, M, K, nullptr, amax, nmax, stream));
    B200_TRY(row_stats(dB, K, N, K, nullptr, amax + 1, nmax + 1, stream));
    unsigned char h[32];
    B200_CUDA(cudaMemcpyAsync(h, sc, 32, cudaMemcpyDeviceToHost, stream));
    B200_CUDA(cudaStreamSynchronize(stream));
    float am[2];
    double nm[2];
    memcpy(am, h, 8);
    memcpy(nm, h + 8, 16);
    const int ea = pick_scale_exp(am[0], nm[0]), eb = pick_scale_exp(am[1], nm[1]);
    B200_TRY(split_rows(dA, K, nullptr, M, K, nullptr, ea, A, nullptr, stream));
    B200_TRY(split_rows(dB, K, nullptr, N, K, nullptr, eb, B, nullptr, stream));
    EpiArgs ep;
    ep.out = d_out; ep.ldo = ldo; ep.M = M; ep.N = N;
    ep.alpha = scalbn(1.0, -(ea + eb));
    return gemm_split(A, B, terms, EPI_PLAIN, ep, chunk_boxes, stream);
}

extern "C" int b200mnn_gemm_profile_enable(int on) { return b200::gemm::profile_enable(on); }
extern "C" int b200mnn_gemm_profile_collect(double* total_ms, int64_t* launches, double* executed_flops) {
    return b200::gemm::profile_collect(total_ms, launches, executed_flops);
}
