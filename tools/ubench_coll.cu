// Latency / throughput of warp collectives on sm_100a: dependent chains of SHFL, VOTE, REDUX (CREDUX) -- 1 warp per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int iters) {
    const int lane = threadIdx.x & 31;
    unsigned x = lane * 2654435761u, y = x ^ 0x1234u, z = x + 77u, w = x * 3u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + lane;
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) x = __ballot_sync(0xffffffffu, x & 1) + lane;
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) x = __reduce_max_sync(0xffffffffu, x) + lane;
    long long t3 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = __reduce_max_sync(0xffffffffu, x) + lane; y = __reduce_max_sync(0xffffffffu, y) + lane;
        z = __reduce_max_sync(0xffffffffu, z) + lane; w = __reduce_max_sync(0xffffffffu, w) + lane;
    }
    long long t4 = clock64();
    for (int i = 0; i < iters; ++i) {
        x = __shfl_xor_sync(0xffffffffu, x, 1) + lane; y = __shfl_xor_sync(0xffffffffu, y, 2) + lane;
        z = __shfl_xor_sync(0xffffffffu, z, 4) + lane; w = __shfl_xor_sync(0xffffffffu, w, 8) + lane;
    }
    long long t5 = clock64();
    for (int i = 0; i < iters; ++i) {   // max via 5 butterfly steps
        unsigned m = x;
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        x = m + lane;
    }
    long long t6 = clock64();
    for (int i = 0; i < iters; ++i) x = __match_any_sync(0xffffffffu, x & 3) + lane;
    long long t7 = clock64();
    if (lane == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5; out[6] = t7 - t6;
    }
    if (x + y + z + w == 0x7fffffffu) out[7] = x;
}
int main() {
    long long* d; cudaMalloc(&d, 64); const int iters = 4096;
    for (int warps = 1; warps <= 8; warps *= 2) {
        k<<<1, 32 * warps>>>(d, iters); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("%d warp(s): dependent SHFL %.1f, VOTE %.1f, REDUX.MAX %.1f cycles/op; 4 independent REDUX %.1f cycles/4, 4 independent SHFL %.1f cycles/4; 5-step butterfly max %.1f; MATCH %.1f\n",
               warps, (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, (double)h[5] / iters, (double)h[6] / iters);
    }
    return 0;
}
