// K2: projection / MLP GEMM on the 5th-gen tensor cores.
//
//   C = act((A * B^T + bias) * alpha),  A:[M,K] bf16, B:[N,K] bf16 (both K-major), fp32 accumulate.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer  : 3-D tensor maps (K, rows, batch), 128B swizzle, 4-stage smem ring
//   warp 1      MMA issuer    : one elected thread, tcgen05.mma cta_group::1 kind::f16,
//                               M=128 x N=256 x K=16 per instruction, accumulator in TMEM
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, ping-pong)
//   warps 4..11 epilogue      : tcgen05.ld 32x32b -> bias / alpha / ReLU -> bf16|fp32 -> global,
//                               overlapped with the next tile's main loop
//
// Replaces F.linear in models/utils/attention.py:21-27,138 and the PE MLPs of
// models/dense_heads/cmt_head.py:292-301 (reference runs them as fp32 cuBLAS SGEMMs).
#include <cstdlib>
#include "kernels.cuh"

namespace cmt {

namespace gemm {
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STORE_STAGING = 8 * 4096;  // per epilogue warp: two 32x32 bf16 tiles or one 32x32 fp32 tile
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_STAGING + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int THREADS = 384;  // 4 control warps + 8 epilogue warps
}  // namespace gemm

struct TcGemmParams {
    const float* bias;
    void* C;
    int M, N, K;
    long long ldc, cb, cb_stride, strideC;
    float alpha;
    int relu, bias_per_row, out_bf16;
    int a_batched, b_batched;
    int tma_store;  // epilogue through shared memory + TMA tensor store (full-sector, coalesced writes)
    int cb32;       // column block size as int (for the store coordinates)
    int m_tiles, n_tiles, total_tiles, num_kb;
};

__global__ void __launch_bounds__(gemm::THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const TcGemmParams p) {
    using namespace gemm;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STORE_STAGING);
    uint64_t* full_bar = bars;                   // [STAGES]
    uint64_t* empty_bar = bars + STAGES;         // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;     // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    // warp index / TMEM base through shuffles: provably warp-uniform, so the producer and issuer warps keep
    // their descriptors in uniform registers and TMA / tcgen05.mma instructions issue back to back (a divergent
    // `lane == 0` role costs a ~12-instruction ELECT/R2UR waterfall per instruction).
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 8);  // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const int tiles_per_batch = p.m_tiles * p.n_tiles;

    if (warp == 0) {
        // ----------------------------- TMA producer -----------------------------
        {
            const bool leader = elect_one();
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int z = tile / tiles_per_batch;
                const int r = tile - z * tiles_per_batch;
                const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (leader) {
                        uint8_t* sa = smem + stage * STAGE_BYTES;
                        uint8_t* sb = sa + A_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_3d(sa, &tma_a, &full_bar[stage], kb * BK, mt * BM, p.a_batched ? z : 0);
                        tma_load_3d(sb, &tma_b, &full_bar[stage], kb * BK, nt * BN, p.b_batched ? z : 0);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t adesc = make_kmajor_desc(sa, 128);
                        const uint64_t bdesc = make_kmajor_desc(sa + A_BYTES, 128);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // +32 bytes per K=16 step inside the 128B swizzle atom (>>4 -> +2)
                            tc_mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (leader) tc_commit(&tmem_full[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------- epilogue -------------------------------
        // 8 warps: warp w reads TMEM lane quadrant (w & 3) and column half ((w - 4) >> 2), so every
        // scheduler has two epilogue warps to overlap TMEM / global latencies.  All option tests
        // (bias kind, alpha, relu, dtype) are hoisted out of the per-element loops.
        const int quad = warp & 3;
        const int half = (warp - 4) >> 2;
        const bool has_bias = p.bias != nullptr;
        const bool col_bias = has_bias && !p.bias_per_row;
        const bool scale = p.alpha != 1.0f;
        uint32_t store_cnt = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int z = tile / tiles_per_batch;
            const int r = tile - z * tiles_per_batch;
            const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
            const int m = mt * BM + quad * 32 + lane;
            const bool m_ok = m < p.M;
            const float row_bias = (has_bias && p.bias_per_row && m_ok) ? __ldg(p.bias + m) : 0.0f;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
            // column-block addressing without a 64-bit division per chunk: one division per tile, then the
            // (block, offset-in-block) pair advances by 32 columns per chunk
            const int c_first = half * (BN / 64);
            const long long n_first = static_cast<long long>(nt) * BN + c_first * 32;
            long long blk = n_first / p.cb;
            long long rem = n_first - blk * p.cb;
            const long long row_off = z * p.strideC + static_cast<long long>(m) * p.ldc;
            // bias for one 32-column chunk (independent of the accumulator: issued before the TMEM wait)
            auto load_bias = [&](int n0, float (&f)[32]) {
                if (col_bias) {
                    if (n0 + 32 <= p.N) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);  // n0 % 32 == 0
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b = __ldg(b4 + i);
                            f[4 * i] = b.x; f[4 * i + 1] = b.y; f[4 * i + 2] = b.z; f[4 * i + 3] = b.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.0f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = row_bias;
                }
            };
            auto finish = [&](int n0, long long off, const uint32_t (&v)[32], float (&f)[32]) {
                const bool full = (n0 + 32 <= p.N);
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] += __uint_as_float(v[i]);
                if (scale) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] *= p.alpha;
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
                }
                if (p.tma_store) {
                    // registers -> this warp's staging tile [32 rows][32 cols] -> one TMA tensor store.
                    // A thread's direct 16-byte stores land in 64-byte-strided rows: half-filled sectors and
                    // 32 sectors per request; the TMA store writes whole lines and clips the M / N tails itself.
                    uint8_t* tile = staging + (warp - 4) * 4096 + (p.out_bf16 ? (store_cnt & 1) * 2048 : 0);
                    if (lane == 0) {
                        if (p.out_bf16) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
                    }
                    __syncwarp();
                    if (p.out_bf16) {
                        uint4* row = reinterpret_cast<uint4*>(tile + lane * 64);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 w;
                            w.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
                            w.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
                            w.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
                            w.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
                            row[i] = w;
                        }
                    } else {
                        float4* row = reinterpret_cast<float4*>(tile + lane * 128);
#pragma unroll
                        for (int i = 0; i < 8; ++i) row[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const int nb = n0 / p.cb32;
                        tma_store_4d(&tma_c, tile, n0 - nb * p.cb32, mt * BM + quad * 32, nb, z);
                        tma_store_commit();
                    }
                    ++store_cnt;
                    return;
                }
                if (!m_ok) return;
                if (p.out_bf16) {
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + off;
                    if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 w;
                            w.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
                            w.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
                            w.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
                            w.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
                            reinterpret_cast<uint4*>(dst)[i] = w;
                        }
                    } else {
                        for (int i = 0; i < 32; ++i)
                            if (n0 + i < p.N) dst[i] = __float2bfloat16_rn(f[i]);
                    }
                } else {
                    float* dst = reinterpret_cast<float*>(p.C) + off;
                    if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                    } else {
                        for (int i = 0; i < 32; ++i)
                            if (n0 + i < p.N) dst[i] = f[i];
                    }
                }
            };
            // two 32-column chunks per round: both TMEM loads and both bias fetches are in flight together
#pragma unroll 1
            for (int c = c_first; c < c_first + BN / 64; c += 2) {
                const int n0a = nt * BN + c * 32, n0b = n0a + 32;
                if (n0a >= p.N) break;  // warp-uniform
                const bool has_b = n0b < p.N;
                const long long offa = row_off + blk * p.cb_stride + rem;
                rem += 32;
                if (rem >= p.cb) { rem -= p.cb; ++blk; }
                const long long offb = row_off + blk * p.cb_stride + rem;
                rem += 32;
                if (rem >= p.cb) { rem -= p.cb; ++blk; }
                uint32_t va[32], vb[32];
                float fa[32], fb[32];
                tmem_ld32(t_row + c * 32, va);
                if (has_b) tmem_ld32(t_row + (c + 1) * 32, vb);
                load_bias(n0a, fa);
                if (has_b) load_bias(n0b, fb);
                tc_wait_ld();
                finish(n0a, offa, va, fa);
                if (has_b) finish(n0b, offb, vb, fb);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.tma_store && lane == 0) tma_store_wait_all<0>();  // all tensor stores of this warp have landed
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int launch_tc_gemm(const GemmArgs& g, int batch, cudaStream_t stream) {
    using namespace gemm;
    CMT_CHECK_ARG(g.K % 8 == 0 && g.lda % 8 == 0 && g.ldb % 8 == 0,
                  "cmt_gemm_bias_act(bf16): K, lda, ldb must be multiples of 8 (K=%d lda=%lld ldb=%lld)",
                  g.K, g.lda, g.ldb);
    CMT_CHECK_ARG(((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.B)) & 15) == 0,
                  "cmt_gemm_bias_act(bf16): operands must be 16-byte aligned");
    CMT_CHECK_ARG(g.cb >= g.N || g.cb % 32 == 0,
                  "cmt_gemm_bias_act(bf16): column block must be >= N or a multiple of 32");
    CMT_CHECK_ARG(g.strideA % 8 == 0 && g.strideB % 8 == 0, "cmt_gemm_bias_act(bf16): batch strides must be multiples of 8");

    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gemm)");
        attr_done = true;
    }

    CUtensorMap ta, tb;
    {
        const bool batched = g.strideA != 0 && batch > 1;
        uint64_t dims[3] = {static_cast<uint64_t>(g.K), static_cast<uint64_t>(g.M),
                            static_cast<uint64_t>(batched ? batch : 1)};
        uint64_t strides[2] = {static_cast<uint64_t>(g.lda) * 2,
                               static_cast<uint64_t>(batched ? g.strideA : (long long)g.M * g.lda) * 2};
        uint32_t box[3] = {BK, BM, 1};
        int rc = encode_tma_bf16(&ta, g.A, 3, dims, strides, box, 128);
        if (rc) return rc;
    }
    {
        const bool batched = g.strideB != 0 && batch > 1;
        uint64_t dims[3] = {static_cast<uint64_t>(g.K), static_cast<uint64_t>(g.N),
                            static_cast<uint64_t>(batched ? batch : 1)};
        uint64_t strides[2] = {static_cast<uint64_t>(g.ldb) * 2,
                               static_cast<uint64_t>(batched ? g.strideB : (long long)g.N * g.ldb) * 2};
        uint32_t box[3] = {BK, BN, 1};
        int rc = encode_tma_bf16(&tb, g.B, 3, dims, strides, box, 128);
        if (rc) return rc;
    }

    // C tensor map in "column block" coordinates (n % cb, m, n / cb, batch); used when the layout is TMA-able
    CUtensorMap tc;
    const int esz = g.out_bf16 ? 2 : 4;
    const bool plain = g.cb >= g.N;
    bool tma_store = (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && (g.ldc * esz) % 16 == 0 &&
                     (plain || ((g.cb_stride * esz) % 16 == 0 && g.cb % 32 == 0)) &&
                     (batch == 1 || (g.strideC * esz) % 16 == 0) && g.cb < (1ll << 31);
    if (getenv("CMT_GEMM_NO_TMA_STORE")) tma_store = false;
    if (tma_store) {
        const uint64_t d0 = plain ? static_cast<uint64_t>(g.N) : static_cast<uint64_t>(g.cb);
        const uint64_t d2 = plain ? 1 : static_cast<uint64_t>((g.N + g.cb - 1) / g.cb);
        uint64_t dims[4] = {d0, static_cast<uint64_t>(g.M), d2, static_cast<uint64_t>(batch)};
        uint64_t strides[3] = {static_cast<uint64_t>(g.ldc) * esz,
                               static_cast<uint64_t>(plain ? g.M * g.ldc : g.cb_stride) * esz,
                               static_cast<uint64_t>(batch > 1 ? g.strideC : (plain ? g.M * g.ldc : g.cb_stride * d2)) * esz};
        uint32_t box[4] = {32, 32, 1, 1};
        if (d0 < 32) box[0] = static_cast<uint32_t>(d0);
        int rc = encode_tma(&tc, g.C, g.out_bf16 ? CMT_BF16 : CMT_F32, 4, dims, strides, box, 0);
        if (rc) tma_store = false;  // fall back to direct stores (e.g. stride not encodable)
        if (d0 < 32) tma_store = false;
    }
    if (!tma_store) tc = ta;  // unused placeholder

    TcGemmParams p{};
    p.tma_store = tma_store ? 1 : 0;
    p.cb32 = static_cast<int>(plain ? (1ll << 30) : g.cb);
    p.bias = g.bias;
    p.C = g.C;
    p.M = g.M;
    p.N = g.N;
    p.K = g.K;
    p.ldc = g.ldc;
    p.cb = g.cb;
    p.cb_stride = g.cb_stride;
    p.strideC = g.strideC;
    p.alpha = g.alpha;
    p.relu = g.relu;
    p.bias_per_row = g.bias_per_row;
    p.out_bf16 = g.out_bf16;
    p.a_batched = (g.strideA != 0 && batch > 1) ? 1 : 0;
    p.b_batched = (g.strideB != 0 && batch > 1) ? 1 : 0;
    p.m_tiles = (g.M + BM - 1) / BM;
    p.n_tiles = (g.N + BN - 1) / BN;
    const long long total = static_cast<long long>(p.m_tiles) * p.n_tiles * batch;
    CMT_CHECK_ARG(total < (1ll << 31), "cmt_gemm_bias_act(bf16): too many tiles");
    p.total_tiles = static_cast<int>(total);
    p.num_kb = (g.K + BK - 1) / BK;
    int grid = device_sm_count();
    if (grid > p.total_tiles) grid = p.total_tiles;
    tc_gemm_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(ta, tb, tc, p);
    CMT_LAUNCH_CHECK("cmt_gemm_bias_act(tcgen05)");
    return CMT_OK;
}

}  // namespace cmt
