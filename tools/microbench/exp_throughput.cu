// Microbenchmark (not part of the product): throughput of the exp2 variants the attention softmax
// could use on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_tp exp_throughput.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define ITERS 4096
#define ILP 8

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void k(float* out, float seed) {
    float acc = 0.f;
    float v[ILP];
    uint32_t u[ILP];
    for (int i = 0; i < ILP; ++i) { v[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f; u[i] = 0xB800B800u + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { v[i] = ex2f(v[i]) - 1.0f; }                       // 1 MUFU + 1 FADD per element
            if (MODE == 1) { u[i] = ex2h2(u[i]) ^ 0x80008000u; }               // f16x2: 2 elements per op
            if (MODE == 2) { u[i] = ex2b2(u[i]) ^ 0x80008000u; }               // bf16x2
            if (MODE == 3) {                                                   // polynomial exp2 on FMA/ALU pipes
                float x = fmaxf(v[i], -126.0f);
                float t = x + 12582912.0f;                                     // round to nearest integer (magic 1.5*2^23)
                float n = t - 12582912.0f;
                float f = x - n;                                               // in [-0.5, 0.5]
                float p = fmaf(f, 0.0555054f, 0.2402265f);
                p = fmaf(p, f, 0.6931472f);
                p = fmaf(p, f, 1.0f);
                v[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) - 1.0f;
            }
        }
    }
    for (int i = 0; i < ILP; ++i) acc += v[i] + __uint_as_float(u[i]);
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int per_op) {
    float* d; cudaMalloc(&d, 4);
    int blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(d, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double elems = double(blocks) * threads * ITERS * ILP * per_op;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-34s %8.3f ms  %8.1f Gelem/s  = %6.2f elem/clk/SM @%d MHz (nominal)\n", name, ms, elems / ms * 1e-6,
           elems / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
}

int main() {
    run<0>("ex2.approx.ftz.f32 (+FADD)", 1);
    run<1>("ex2.approx.f16x2 (+LOP)", 2);
    run<2>("ex2.approx.ftz.bf16x2 (+LOP)", 2);
    run<3>("poly3 exp2 on FMA/ALU", 1);
    return 0;
}
