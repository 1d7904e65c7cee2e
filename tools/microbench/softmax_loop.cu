// Microbenchmark: the per-tile softmax work of the attention kernel in isolation (no MMA, no TMA):
// each warp repeatedly loads a 32-lane x 128-column fp32 score block from TMEM, takes the row max,
// exponentiates, packs to bf16 and stores the 64 packed columns back to TMEM.
// Ideal (MUFU bound): warps * 128 * 32 exps / 16 per clk = 2048 cycles per iteration with 8 warps.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../cmt-cooperative-perception_b200/csrc/common.cuh"
namespace cmt { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -2; } }
using namespace cmt;
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;
    const float n = t - 12582912.0f;
    const float f = x - n;
    float p = fmaf(f, 0.0555054f, 0.2402265f);
    p = fmaf(p, f, 0.6931472f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE, int PACK = 0, int POLY = 0>  // POLY n>0: one exponential in 2n on the FMA pipes (cubic)  // PACK 0: cvt.rn.bf16x2 (F2FP); 1: integer round + PRMT; 2: PRMT truncate
// MODE 0: full; 1: no TMEM ld/st (registers only); 2: no exps (ld/max/st only); 3: chunked (max from previous tile)
__global__ void __launch_bounds__(384, 1) k(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = slot + lane_base + (warp >> 2) * 128;          // S region of "my" query tile
    const uint32_t t_p = slot + lane_base + 256 + (warp >> 2) * 64;     // P region
    float m = 0.25f;
    uint32_t acc = 0;
    uint32_t s[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) s[c][i] = __float_as_uint(1e-3f * (threadIdx.x + c * 32 + i));
    if (MODE != 1) { tmem_st32(t_s, s[0]); tmem_st32(t_s + 32, s[1]); tmem_st32(t_s + 64, s[2]); tmem_st32(t_s + 96, s[3]); tc_wait_st(); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE != 1) {
            tmem_ld32(t_s + 0, s[0]); tmem_ld32(t_s + 32, s[1]); tmem_ld32(t_s + 64, s[2]); tmem_ld32(t_s + 96, s[3]);
            tc_wait_ld();
        }
        float mx0 = -1e30f, mx1 = -1e30f, mx2 = -1e30f, mx3 = -1e30f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            mx0 = fmaxf(mx0, __uint_as_float(s[0][i])); mx1 = fmaxf(mx1, __uint_as_float(s[1][i]));
            mx2 = fmaxf(mx2, __uint_as_float(s[2][i])); mx3 = fmaxf(mx3, __uint_as_float(s[3][i]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (MODE == 3) { /* use stale m for this tile, update after */ } else if (mx - m > 8.0f) m = mx;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float e0, e1;
                if (MODE == 2) { e0 = __uint_as_float(s[c][2 * i]) - m; e1 = __uint_as_float(s[c][2 * i + 1]) - m; }
                else { e0 = ex2_approx(__uint_as_float(s[c][2 * i]) - m); const float x1 = __uint_as_float(s[c][2 * i + 1]) - m; e1 = (POLY > 0 && (i % (POLY > 0 ? POLY : 1)) == POLY - 1) ? ex2_poly(x1) : ex2_approx(x1); }
                if (PACK == 0) pk[i] = pack_bf16x2(e0, e1);
                else if (PACK == 1) pk[i] = __byte_perm(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u, 0x7632);
                else pk[i] = __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);
            }
            if (MODE != 1) tmem_st16(t_p + c * 16, pk);
            else { for (int i = 0; i < 16; ++i) acc ^= pk[i]; }
        }
        if (MODE != 1) tc_wait_st();
        if (MODE == 3 && mx - m > 8.0f) m = mx;
        if (MODE == 1) { for (int c = 0; c < 4; ++c) for (int i = 0; i < 32; ++i) s[c][i] += acc & 1; }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345 || m == 123.f) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
    long long* dc; uint32_t* ds; cudaMalloc(&dc, 148 * 8); cudaMalloc(&ds, 4);
    const int iters = 2000;
    const char* names[8] = {"full (ld, max, exp, pack, st)", "registers only (no TMEM)", "no exp (ld, max, pack, st)", "full, stale max",
                            "full, poly 1/2", "full, poly 1/4", "full, poly 1/8", "full, poly 1/16"};
    for (int mode = 0; mode < 8; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            const int th = getenv("THREADS") ? atoi(getenv("THREADS")) : 256;
            if (mode == 0) k<0><<<148, th>>>(iters, dc, ds);
            if (mode == 1) k<1><<<148, th>>>(iters, dc, ds);
            if (mode == 2) k<2><<<148, th>>>(iters, dc, ds);
            if (mode == 3) k<3><<<148, th>>>(iters, dc, ds);
            if (mode == 4) k<0, 0, 1><<<148, th>>>(iters, dc, ds);
            if (mode == 5) k<0, 0, 2><<<148, th>>>(iters, dc, ds);
            if (mode == 6) k<0, 0, 4><<<148, th>>>(iters, dc, ds);
            if (mode == 7) k<0, 0, 8><<<148, th>>>(iters, dc, ds);
            cudaDeviceSynchronize();
        }
        long long h[148]; cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-34s %8.1f cycles per tile-step (8 warps x 32 rows x 128 cols; MUFU bound = 2048)  [%s]\n", names[mode],
               double(h[0]) / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
