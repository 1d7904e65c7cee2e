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
        add_layernorm_kernel<8, true><<<static_cast<int>(blocks), 256, 0, stream>>>(x, r, gamma, beta, eps, M, y, gamma2,
                                                                                   beta2, y2, add, ylp, yadd);
    else
        add_layernorm_kernel<8, false><<<static_cast<int>(blocks), 256, 0, stream>>>(x, r, gamma, beta, eps, M, y, gamma2,
                                                                                    beta2, y2, add, ylp, yadd);
    CMT_LAUNCH_CHECK("cmt_add_layernorm");
    return CMT_OK;
}

}  // namespace cmt
