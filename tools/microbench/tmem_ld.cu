// Microbenchmark (not part of the product): tcgen05.ld throughput TMEM -> registers per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../cmt-cooperative-perception_b200/csrc/common.cuh"
namespace cmt { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -2; } }
using namespace cmt;

template <int MODE>  // 0: 4x ld32 then one wait ; 1: ld32+wait each
__global__ void k(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t a[32], b[32], c[32], d[32];
        tmem_ld32(base + 0, a);
        if (MODE == 1) tc_wait_ld();
        tmem_ld32(base + 32, b);
        if (MODE == 1) tc_wait_ld();
        tmem_ld32(base + 64, c);
        if (MODE == 1) tc_wait_ld();
        tmem_ld32(base + 96, d);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= a[i] ^ b[i] ^ c[i] ^ d[i];
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
    long long* dc; uint32_t* ds; cudaMalloc(&dc, 148 * 8); cudaMalloc(&ds, 4);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(iters, dc, ds); else k<1><<<148, warps * 32>>>(iters, dc, ds);
                cudaDeviceSynchronize();
            }
            long long h[148]; cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
            double cyc = double(h[0]) / iters;
            double bytes = double(warps) * 32 * 128 * 4;
            printf("mode %d warps %2d: %8.1f cycles per 128-column row-block load per warp-set, %6.1f B/clk/SM (%s)\n", mode, warps, cyc,
                   bytes / cyc, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
