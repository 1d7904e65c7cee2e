// Fused residual-add + LayerNorm for the decoder's small ops (mmcv BaseTransformerLayer order
// self_attn, norm, cross_attn, norm, ffn, norm + PETRTransformerDecoder's shared post_norm,
// projects/mmdet3d_plugin/models/utils/petr_transformer.py:347-371).  One warp per 256-wide row:
//   y   = LN(x + r; gamma, beta)                               fp32
//   y2  = LN(y; gamma2, beta2)            (optional: the stacked post-normed intermediate)
//   ylp = cast(y), yadd = cast(y + add)   (optional: bf16|fp32 A operands of the next projections,
//                                          `add` = query_pos)
// HBM/launch bound; replaces ~6 eager torch launches per norm.
#include "kernels.cuh"

namespace cmt {

template <int kPerLane, bool kLpBf16>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, int M,
                                                            float* __restrict__ y, const float* __restrict__ gamma2,
                                                            const float* __restrict__ beta2, float* __restrict__ y2,
                                                            const float* __restrict__ add, void* __restrict__ ylp,
                                                            void* __restrict__ yadd) {
    constexpr int C = kPerLane * 32;
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
        const long long base = static_cast<long long>(row) * C;
        float v[kPerLane];
        float s = 0.f;
        // lane owns 4-element groups: element index = (g * 32 + lane) * 4 + e  (coalesced 16 B accesses)
#pragma unroll
        for (int g = 0; g < kPerLane / 4; ++g) {
            const int idx = (g * 32 + lane) * 4;
            float4 a = *reinterpret_cast<const float4*>(x + base + idx);
            if (r != nullptr) {
                const float4 b = *reinterpret_cast<const float4*>(r + base + idx);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            v[4 * g] = a.x; v[4 * g + 1] = a.y; v[4 * g + 2] = a.z; v[4 * g + 3] = a.w;
            s += (a.x + a.y) + (a.z + a.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / C) + eps);
        float s2 = 0.f;
#pragma unroll
        for (int g = 0; g < kPerLane / 4; ++g) {
            const int idx = (g * 32 + lane) * 4;
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + idx));
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta + idx));
            v[4 * g] = (v[4 * g] - mean) * rstd * ga.x + be.x;
            v[4 * g + 1] = (v[4 * g + 1] - mean) * rstd * ga.y + be.y;
            v[4 * g + 2] = (v[4 * g + 2] - mean) * rstd * ga.z + be.z;
            v[4 * g + 3] = (v[4 * g + 3] - mean) * rstd * ga.w + be.w;
            *reinterpret_cast<float4*>(y + base + idx) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            s2 += (v[4 * g] + v[4 * g + 1]) + (v[4 * g + 2] + v[4 * g + 3]);
            if (ylp != nullptr) {
                if (kLpBf16) {
                    uint2 w;
                    w.x = pack_bf16x2(v[4 * g], v[4 * g + 1]);
                    w.y = pack_bf16x2(v[4 * g + 2], v[4 * g + 3]);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ylp) + base + idx) = w;
                } else {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ylp) + base + idx) =
                        make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                }
            }
            if (yadd != nullptr) {
                const float4 ad = *reinterpret_cast<const float4*>(add + base + idx);
                const float a0 = v[4 * g] + ad.x, a1 = v[4 * g + 1] + ad.y, a2 = v[4 * g + 2] + ad.z,
                            a3 = v[4 * g + 3] + ad.w;
                if (kLpBf16) {
                    uint2 w;
                    w.x = pack_bf16x2(a0, a1);
                    w.y = pack_bf16x2(a2, a3);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(yadd) + base + idx) = w;
                } else {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(yadd) + base + idx) = make_float4(a0, a1, a2, a3);
                }
            }
        }
        if (y2 != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            const float mean2 = s2 * (1.0f / C);
            float q2 = 0.f;
#pragma unroll
            for (int i = 0; i < kPerLane; ++i) { const float d = v[i] - mean2; q2 = fmaf(d, d, q2); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
            const float rstd2 = rsqrtf(q2 * (1.0f / C) + eps);
#pragma unroll
            for (int g = 0; g < kPerLane / 4; ++g) {
                const int idx = (g * 32 + lane) * 4;
                const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma2 + idx));
                const float4 be = __ldg(reinterpret_cast<const float4*>(beta2 + idx));
                *reinterpret_cast<float4*>(y2 + base + idx) =
                    make_float4((v[4 * g] - mean2) * rstd2 * ga.x + be.x, (v[4 * g + 1] - mean2) * rstd2 * ga.y + be.y,
                                (v[4 * g + 2] - mean2) * rstd2 * ga.z + be.z, (v[4 * g + 3] - mean2) * rstd2 * ga.w + be.w);
            }
        }
    }
}

int launch_add_layernorm(const float* x, const float* r, const float* gamma, const float* beta, float eps, int M, int C,
                         float* y, const float* gamma2, const float* beta2, float* y2, const float* add, void* ylp,
                         void* yadd, int lp_dtype, cudaStream_t stream) {
    CMT_CHECK_ARG(x && gamma && beta && y && M > 0, "cmt_add_layernorm: bad arguments");
    CMT_CHECK_ARG(C == 256, "cmt_add_layernorm: embed dim 256 only (got %d)", C);
    CMT_CHECK_ARG(y2 == nullptr || (gamma2 && beta2), "cmt_add_layernorm: second norm needs gamma2/beta2");
    CMT_CHECK_ARG(yadd == nullptr || add != nullptr, "cmt_add_layernorm: yadd needs add");
    CMT_CHECK_ARG(lp_dtype == CMT_F32 || lp_dtype == CMT_BF16, "cmt_add_layernorm: bad lp dtype");
    const int rows_per_block = 8;
    long long blocks = (M + rows_per_block - 1) / rows_per_block;
    const long long cap = static_cast<long long>(device_sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    if (lp_dtype == CMT_BF16)
        launch_pdl(add_layernorm_kernel<8, true>, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, x, r, gamma, beta, eps, M, y, gamma2,
                                                                                   beta2, y2, add, ylp, yadd);
    else
        launch_pdl(add_layernorm_kernel<8, false>, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, x, r, gamma, beta, eps, M, y, gamma2,
                                                                                    beta2, y2, add, ylp, yadd);
    CMT_LAUNCH_CHECK("cmt_add_layernorm");
    return CMT_OK;
}

// ---------------------------------------------------------------------------------------------
// Task-head tail (SeparateTaskHead, cmt_head.py:116-150 + GroupLayerNorm1d :53-94, final_kernel = 1): after the first
// grouped 1x1 conv (a per-decoder-layer GEMM, h = x W1^T), every (layer, query row, output head) needs
//   y = ReLU(LN_64(h) * gamma + beta),   out[o] = y . w2[o] + b2[o]        (o < c_out <= CMAX)
// Eager torch runs this as ~10 elementwise / reduction passes over the 66 MB h tensor plus an einsum (0.45 ms per
// forward at B = 8).  Here one THREAD owns a (layer, row): per output head it pulls the 64 hidden values into
// registers (sixteen 16-byte loads of its own 256 contiguous bytes), takes the statistics and the affine + ReLU in
// registers, and runs the CMAX 64-long dot products against w2 held in shared memory (all lanes of a warp read the same
// weight: broadcast).  No shuffles (a warp-per-row version spent its time in 70 shuffles per head: 146 us), fp32
// throughout (these outputs feed the top-k).
constexpr int TH_HC = 64;   // hidden channels per head (head_conv = 64 in every reference config)

__global__ void __launch_bounds__(128) task_head_tail_kernel(const float* __restrict__ h, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ w2,
                                                             const float* __restrict__ b2, float* __restrict__ out, int M,
                                                             int NH, int CMAX, float eps) {
    extern __shared__ float th_smem[];
    const int l = blockIdx.y;
    float* w2s = th_smem;                       // [NH][CMAX][64]
    float* gs = w2s + NH * CMAX * TH_HC;        // [NH][64]
    float* bs = gs + NH * TH_HC;                // [NH][64]
    float* b2s = bs + NH * TH_HC;               // [NH][CMAX]
    for (int i = threadIdx.x; i < NH * CMAX * TH_HC; i += blockDim.x) w2s[i] = w2[static_cast<long long>(l) * NH * CMAX * TH_HC + i];
    for (int i = threadIdx.x; i < NH * TH_HC; i += blockDim.x) {
        gs[i] = gamma[l * NH * TH_HC + i];
        bs[i] = beta[l * NH * TH_HC + i];
    }
    for (int i = threadIdx.x; i < NH * CMAX; i += blockDim.x) b2s[i] = b2[l * NH * CMAX + i];
    __syncthreads();
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        const float4* hrow = reinterpret_cast<const float4*>(h + (static_cast<long long>(l) * M + m) * NH * TH_HC);
        float* orow = out + (static_cast<long long>(l) * M + m) * NH * CMAX;
        for (int hd = 0; hd < NH; ++hd) {
            float y[TH_HC];
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < TH_HC / 4; ++i) {
                const float4 v = __ldg(hrow + hd * (TH_HC / 4) + i);
                y[4 * i] = v.x; y[4 * i + 1] = v.y; y[4 * i + 2] = v.z; y[4 * i + 3] = v.w;
                sum += (v.x + v.y) + (v.z + v.w);
            }
            const float mu = sum * (1.0f / TH_HC);
            float sq = 0.0f;
#pragma unroll
            for (int c = 0; c < TH_HC; ++c) {
                y[c] -= mu;
                sq = fmaf(y[c], y[c], sq);
            }
            const float sd = sqrtf(sq * (1.0f / TH_HC) + eps);
            const float4* g4 = reinterpret_cast<const float4*>(gs + hd * TH_HC);
            const float4* b4 = reinterpret_cast<const float4*>(bs + hd * TH_HC);
#pragma unroll
            for (int i = 0; i < TH_HC / 4; ++i) {
                const float4 g = g4[i], b = b4[i];
                y[4 * i] = fmaxf(y[4 * i] / sd * g.x + b.x, 0.0f);
                y[4 * i + 1] = fmaxf(y[4 * i + 1] / sd * g.y + b.y, 0.0f);
                y[4 * i + 2] = fmaxf(y[4 * i + 2] / sd * g.z + b.z, 0.0f);
                y[4 * i + 3] = fmaxf(y[4 * i + 3] / sd * g.w + b.w, 0.0f);
            }
            for (int o = 0; o < CMAX; ++o) {
                const float4* w4 = reinterpret_cast<const float4*>(w2s + (hd * CMAX + o) * TH_HC);
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
                for (int i = 0; i < TH_HC / 4; ++i) {
                    const float4 w = w4[i];
                    a0 = fmaf(y[4 * i], w.x, a0);
                    a1 = fmaf(y[4 * i + 1], w.y, a1);
                    a2 = fmaf(y[4 * i + 2], w.z, a2);
                    a3 = fmaf(y[4 * i + 3], w.w, a3);
                }
                orow[hd * CMAX + o] = (a0 + a1) + (a2 + a3) + b2s[hd * CMAX + o];
            }
        }
    }
}

int launch_task_head_tail(const float* h, const float* gamma, const float* beta, const float* w2, const float* b2, float* out,
                          int L, int M, int NH, int HC, int CMAX, float eps, cudaStream_t stream) {
    CMT_CHECK_ARG(HC == TH_HC, "cmt_task_head_tail: head_conv must be %d (got %d)", TH_HC, HC);
    CMT_CHECK_ARG(L > 0 && M > 0 && NH > 0 && CMAX > 0 && CMAX <= 32, "cmt_task_head_tail: bad shape L=%d M=%d NH=%d CMAX=%d", L, M,
                  NH, CMAX);
    const size_t smem = static_cast<size_t>(NH) * (CMAX * TH_HC + 2 * TH_HC + CMAX) * sizeof(float);
    CMT_CHECK_ARG(smem <= 200 * 1024, "cmt_task_head_tail: weights do not fit in shared memory");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(task_head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(task_head_tail)");
    }
    int bx = (M + 127) / 128;
    task_head_tail_kernel<<<dim3(bx, L), 128, smem, stream>>>(h, gamma, beta, w2, b2, out, M, NH, CMAX, eps);
    CMT_LAUNCH_CHECK("cmt_task_head_tail");
    return CMT_OK;
}

}  // namespace cmt
