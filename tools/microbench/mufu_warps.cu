// Microbenchmark: MUFU.EX2 throughput per SM as a function of resident warps per scheduler and ILP,
// for the softmax inner loop shape (FADD -> MUFU -> pack pairs).  One block per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
template <int ILP>
__global__ void k(int iters, float m, uint32_t* sink, long long* cyc) {
    float s[ILP];
    for (int i = 0; i < ILP; ++i) s[i] = (threadIdx.x * 37 + i) * 1e-4f;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; i += 2) {
            float e0 = ex2f(s[i] - m), e1 = ex2f(s[i + 1] - m);
            acc ^= pack(e0, e1);
            s[i] += 1e-3f; s[i + 1] -= 1e-3f;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x1234567) sink[0] = acc;
}
template <int ILP> void run(int warps) {
    uint32_t* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 148 * 8);
    int iters = 4096 / ILP * 8;
    k<ILP><<<148, warps * 32>>>(iters, 0.5f, d, c); cudaDeviceSynchronize();
    k<ILP><<<148, warps * 32>>>(iters, 0.5f, d, c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double exps = double(warps) * 32 * iters * ILP;
    printf("ILP %3d warps/SM %2d (%d per scheduler): %6.2f ex2/clk/SM\n", ILP, warps, warps / 4, exps / double(h));
    cudaFree(d); cudaFree(c);
}
int main() {
    for (int w : {4, 8, 12, 16, 32}) { run<8>(w); run<32>(w); run<128>(w); }
    return 0;
}
