// K3: flash-style cross-attention of the object queries over the BEV ++ image tokens on the
// 5th-gen tensor cores (head dim 32, bf16 operands, fp32 accumulate / softmax statistics).
//
// Replaces flash_attn_unpadded_kvpacked_func as called from
// projects/mmdet3d_plugin/models/utils/attention.py:46-92 (softmax(QK^T/sqrt(d))V, non-causal).
//
// Work decomposition.  An "item" is (frame b, head h, block of 256 queries); it needs
// T = ceil(n_tokens/128) KV tile-steps.  The flat space items x T is cut into gridDim.x equal
// contiguous ranges (stream-K), one persistent CTA per SM, so that any batch size fills all 148
// SMs; every (item, CTA) overlap ("segment") writes a normalised fp32 partial + log2-sum-exp into
// the workspace and a second kernel merges the segments of each item.  The same partial/LSE
// algebra serves the multi-GPU KV-token split (cmt_lse_merge).
//
// CTA layout (384 threads):
//   warps 0-3   softmax warpgroup 0 : owns query rows   0..127 of the block (TMEM lanes = rows)
//   warps 4-7   softmax warpgroup 1 : owns query rows 128..255
//   warp  8     TMA producer        : Q (64B swizzle), K tiles [128 tok x 32] (64B swizzle),
//                                     V^T tiles [32 x 128 tok] (two 128B-swizzle boxes), 4-stage rings
//   warp  9     MMA issuer          : S_i = Q_i K^T  (tcgen05.mma SS, M128 N128 K16 x2)
//                                     O_i += P_i [V | 1] (tcgen05.mma TS, A = P in TMEM, M128 N48 K16 x8)
//   warp 10     TMEM allocator
// TMEM columns: S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,432) O1 [432,480).
//
// Softmax is the online form with exp2 (Q arrives pre-multiplied by log2(e)/sqrt(d)) and a lazy
// rescale: the running maximum is only raised (and O rescaled in TMEM) when it grows by more
// than 2^8, so the common tile does no accumulator traffic at all.
//
// Row sums on the tensor pipe: every V^T stage carries 16 extra constant rows (row 32 = ones, rows
// 33..47 = zeros), so the PV MMA (N = 48) also produces sum_j P_ij in accumulator column 32 -- from
// exactly the bf16-rounded P the numerator uses, and without one FADD per score on the CUDA cores.
//
// Latency hiding: S_i(j+1) is issued as soon as warpgroup i has pulled S_i(j) into registers
// (s_empty), not after P_i(j) -- so the next score tile is computed while the current softmax runs and
// the MUFU pipe (the real bound at d_head = 32) is not left idle across the
// softmax -> MMA -> softmax round trip.  pv_done_i orders P_i / O_i reuse.
#include "kernels.cuh"

namespace cmt {

namespace attn {
constexpr int QBLK = 256;
constexpr int KT = 128;
constexpr int NK = 4, NV = 4;
constexpr int TILE_BYTES = 128 * 32 * 2;       // 8 KB: one Q tile or one K tile
constexpr int V_BOX_BYTES = 48 * 128;          // 32 TMA rows (d) + 16 constant rows, 64 tokens (128 B) wide
constexpr int V_STAGE_BYTES = 2 * V_BOX_BYTES; // two 64-token boxes per 128-token tile
constexpr int V_TMA_BYTES = 2 * 32 * 128;      // bytes TMA delivers per stage
constexpr int ON = 48;                         // PV accumulator columns: 32 (O) + 1 (row sum) + 15 (zero)
constexpr int THREADS = 384;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;
constexpr int OFF_V = OFF_K + NK * TILE_BYTES;
constexpr int OFF_BAR = OFF_V + NV * V_STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr uint32_t COL_S = 0, COL_P = 256, COL_O = 384;
constexpr float RESCALE_THRESHOLD = 8.0f;
}  // namespace attn

struct TcAttnParams {
    int B, H, Nq;
    int kv_begin, kv_end;
    int T;            // KV tile-steps per item
    int qblocks;      // ceil(Nq / 256)
    long long W;      // B * H * T : tile-steps of one query-block column
    int groups;       // CTA groups; group g walks range g of [0, W), its qblocks CTAs take one query block each
    int S_max;        // partial slots per item
    float* part_o;    // [items*S_max][256][32]
    float* part_lse;  // [items*S_max][256]   (log2 domain)
};

// Debug-only phase timing (tools/attn_timing.py): when set, CTA 0 accumulates clock64 deltas per wait site.
__device__ long long* g_attn_timing = nullptr;
#define TWAIT(slot, bar, parity)                                   \
    do {                                                           \
        if (tim) {                                                 \
            const long long t0__ = clock64();                      \
            mbar_wait(bar, parity);                                \
            atomicAdd(reinterpret_cast<unsigned long long*>(g_attn_timing + (slot)), static_cast<unsigned long long>(clock64() - t0__)); \
        } else {                                                   \
            mbar_wait(bar, parity);                                \
        }                                                          \
    } while (0)
#define TMARK(slot, since)                                         \
    do {                                                           \
        if (tim) {                                                 \
            const long long now__ = clock64();                     \
            atomicAdd(reinterpret_cast<unsigned long long*>(g_attn_timing + (slot)), static_cast<unsigned long long>(now__ - (since))); \
            (since) = now__;                                       \
        }                                                          \
    } while (0)

__device__ __forceinline__ long long range_start(long long c, long long W, long long G) {
    return (c * W) / G;
}
__device__ __forceinline__ int cta_of(long long x, long long W, long long G) {
    return static_cast<int>(((x + 1) * G + W - 1) / W - 1);
}

__global__ void __launch_bounds__(attn::THREADS, 1)
tc_attn_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
               const __grid_constant__ CUtensorMap tma_v, const TcAttnParams p) {
    using namespace attn;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;             // [NK]
    uint64_t* k_empty = bars + 2 + NK;       // [NK]
    uint64_t* v_full = bars + 2 + 2 * NK;    // [NV]
    uint64_t* v_empty = v_full + NV;         // [NV]
    uint64_t* s_full = v_empty + NV;         // [2]  MMA -> softmax : S_i(t) is in TMEM
    uint64_t* s_empty = s_full + 2;          // [2]  softmax -> MMA : S_i(t) is in registers (128 arrivals)
    uint64_t* p_full = s_empty + 2;          // [2]  softmax -> MMA : P_i(t) is in TMEM      (128 arrivals)
    uint64_t* pv_done = p_full + 2;          // [2]  MMA -> softmax : O_i += P_i(t) V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_v);
    }
    if (warp == 9 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < NK; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < NV; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&pv_done[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 10) tmem_alloc(tmem_slot, 512);
    // constant rows of every V^T stage: row 32 = 1.0 (bf16 0x3F80), rows 33..47 = 0.  A row of identical
    // 16-byte chunks is invariant under the 128B swizzle, so plain stores are enough.
    for (int idx = threadIdx.x; idx < NV * 2 * 16 * 8; idx += THREADS) {
        const int chunk = idx & 7, row = (idx >> 3) & 15, box = (idx >> 7) & 1, stage = idx >> 8;
        const uint32_t v = (row == 0) ? 0x3F803F80u : 0u;
        uint4* dst = reinterpret_cast<uint4*>(smem + OFF_V + stage * V_STAGE_BYTES + box * V_BOX_BYTES +
                                              (32 + row) * 128 + chunk * 16);
        *dst = make_uint4(v, v, v, v);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The qblocks CTAs of a group stream the SAME K / V^T tiles at the same time, one 256-query block
    // each, so a tile is fetched from HBM once and served to the other CTAs of the group from L2.
    const bool tim = (g_attn_timing != nullptr) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp >= 8);
    const long long G = p.groups;
    const int qb = static_cast<int>(blockIdx.x) % p.qblocks;
    const long long grp = static_cast<long long>(blockIdx.x) / p.qblocks;
    const long long pos_begin = grp < G ? range_start(grp, p.W, G) : 0;
    const long long pos_end = grp < G ? range_start(grp + 1, p.W, G) : 0;

    if (warp >= 8) {
        setmaxnreg_dec<80>();
        if (warp == 8 && lane == 0) {
            // ----------------------------- TMA producer -----------------------------
            uint32_t kc = 0, vc = 0, seg = 0;
            for (long long pos = pos_begin; pos < pos_end;) {
                const int item = static_cast<int>(pos / p.T);
                const int j0 = static_cast<int>(pos - static_cast<long long>(item) * p.T);
                const int n = static_cast<int>(min(static_cast<long long>(p.T - j0), pos_end - pos));
                const int h = item % p.H;
                const int b = item / p.H;
                TWAIT(18, q_empty, (seg & 1) ^ 1);
                mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
                tma_load_4d(smem + OFF_Q, &tma_q, q_full, 0, qb * QBLK, h, b);
                tma_load_4d(smem + OFF_Q + TILE_BYTES, &tma_q, q_full, 0, qb * QBLK + 128, h, b);
                for (int jj = 0; jj < n; ++jj) {
                    const int tok0 = p.kv_begin + (j0 + jj) * KT;
                    const uint32_t ks = kc % NK, vs = vc % NV;
                    TWAIT(16, &k_empty[ks], ((kc / NK) & 1) ^ 1);
                    mbar_arrive_expect_tx(&k_full[ks], TILE_BYTES);
                    tma_load_4d(smem + OFF_K + ks * TILE_BYTES, &tma_k, &k_full[ks], 0, tok0, h, b);
                    ++kc;
                    TWAIT(17, &v_empty[vs], ((vc / NV) & 1) ^ 1);
                    mbar_arrive_expect_tx(&v_full[vs], V_TMA_BYTES);
                    uint8_t* sv = smem + OFF_V + vs * V_STAGE_BYTES;
                    tma_load_4d(sv, &tma_v, &v_full[vs], tok0, 0, h, b);
                    tma_load_4d(sv + V_BOX_BYTES, &tma_v, &v_full[vs], tok0 + 64, 0, h, b);
                    ++vc;
                }
                pos += n;
                ++seg;
            }
        } else if (warp == 9 && lane == 0) {
            // ------------------------------ MMA issuer ------------------------------
            constexpr uint32_t idesc_s = make_idesc_bf16(128, KT);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, ON);
            const uint32_t sq = smem_u32(smem + OFF_Q);
            uint32_t kc = 0, vc = 0, seg = 0;
            uint32_t p_cnt[2] = {0, 0};   // P_i tiles consumed
            uint32_t s_cnt[2] = {0, 0};   // S_i tiles issued
            auto issue_s = [&](int i, uint64_t kdesc) {
                // S_i may only be overwritten once warpgroup i holds the previous tile in registers
                if (s_cnt[i] > 0) {
                    TWAIT(9 + i, &s_empty[i], (s_cnt[i] - 1) & 1);
                    tc_fence_after();
                }
                const uint64_t qdesc = make_kmajor_desc(sq + i * TILE_BYTES, 64);
                tc_mma_ss(tmem_base + COL_S + i * 128, qdesc, kdesc, idesc_s, 0);
                tc_mma_ss(tmem_base + COL_S + i * 128, qdesc + 2, kdesc + 2, idesc_s, 1);
                tc_commit(&s_full[i]);
                ++s_cnt[i];
            };
            for (long long pos = pos_begin; pos < pos_end;) {
                const int item = static_cast<int>(pos / p.T);
                const int j0 = static_cast<int>(pos - static_cast<long long>(item) * p.T);
                const int n = static_cast<int>(min(static_cast<long long>(p.T - j0), pos_end - pos));
                TWAIT(15, q_full, seg & 1);
                {
                    const uint32_t ks = kc % NK;
                    TWAIT(8, &k_full[ks], (kc / NK) & 1);
                    tc_fence_after();
                    const uint64_t kdesc = make_kmajor_desc(smem_u32(smem + OFF_K + ks * TILE_BYTES), 64);
                    issue_s(0, kdesc);
                    issue_s(1, kdesc);
                    tc_commit(&k_empty[ks]);
                    ++kc;
                    if (n == 1) tc_commit(q_empty);
                }
                for (int jj = 0; jj < n; ++jj) {
                    if (jj + 1 < n) {  // next score tiles first: they overlap the running softmax
                        const uint32_t ks = kc % NK;
                        TWAIT(8, &k_full[ks], (kc / NK) & 1);
                        tc_fence_after();
                        const uint64_t kdesc = make_kmajor_desc(smem_u32(smem + OFF_K + ks * TILE_BYTES), 64);
                        issue_s(0, kdesc);
                        issue_s(1, kdesc);
                        tc_commit(&k_empty[ks]);
                        ++kc;
                        if (jj + 2 == n) tc_commit(q_empty);
                    }
                    const uint32_t vs = vc % NV;
                    TWAIT(11, &v_full[vs], (vc / NV) & 1);
                    tc_fence_after();
                    const uint64_t vdesc = make_kmajor_desc(smem_u32(smem + OFF_V + vs * V_STAGE_BYTES), 128);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        TWAIT(12 + i, &p_full[i], p_cnt[i] & 1);
                        ++p_cnt[i];
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {
                            const uint64_t vd = vdesc + (((kk >> 2) * V_BOX_BYTES + (kk & 3) * 32) >> 4);
                            tc_mma_ts(tmem_base + COL_O + i * ON, tmem_base + COL_P + i * 64 + kk * 8, vd, idesc_o,
                                      (jj > 0 || kk > 0) ? 1u : 0u);
                        }
                        tc_commit(&pv_done[i]);
                    }
                    tc_commit(&v_empty[vs]);
                    ++vc;
                }
                pos += n;
                ++seg;
            }
        }
    } else {
        // --------------------------- softmax warpgroups ---------------------------
        setmaxnreg_inc<200>();
        const int wg = warp >> 2;                       // 0 or 1 -> Q tile
        const int r = (warp & 3) * 32 + lane;           // row inside the 128-row tile == TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t t_s = tmem_base + lane_base + COL_S + wg * 128;
        const uint32_t t_p = tmem_base + lane_base + COL_P + wg * 64;
        const uint32_t t_o = tmem_base + lane_base + COL_O + wg * ON;
        uint32_t t = 0;  // tiles processed by this warpgroup == phase index of its four barriers
        for (long long pos = pos_begin; pos < pos_end;) {
            const int item = static_cast<int>(pos / p.T);
            const int j0 = static_cast<int>(pos - static_cast<long long>(item) * p.T);
            const int n = static_cast<int>(min(static_cast<long long>(p.T - j0), pos_end - pos));
            float m = -INFINITY;
            for (int jj = 0; jj < n; ++jj, ++t) {
                long long tphase = tim ? clock64() : 0;
                TWAIT(0, &s_full[wg], t & 1);
                if (tim) tphase = clock64();
                tc_fence_after();
                uint32_t s[4][32];
                tmem_ld32(t_s + 0, s[0]);
                tmem_ld32(t_s + 32, s[1]);
                tmem_ld32(t_s + 64, s[2]);
                tmem_ld32(t_s + 96, s[3]);
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(&s_empty[wg]);  // the MMA warp may now overwrite S with the next tile
                TMARK(1, tphase);
                const int valid = p.kv_end - (p.kv_begin + (j0 + jj) * KT);
                if (valid < KT) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
                }
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    mx0 = fmaxf(mx0, __uint_as_float(s[0][i]));
                    mx1 = fmaxf(mx1, __uint_as_float(s[1][i]));
                    mx2 = fmaxf(mx2, __uint_as_float(s[2][i]));
                    mx3 = fmaxf(mx3, __uint_as_float(s[3][i]));
                }
                const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                TMARK(2, tphase);
                bool pv_waited = (jj == 0);  // first tile of a segment: the previous epilogue already waited
                if (jj == 0) {
                    m = mx;  // O_i is overwritten by the first PV of the segment: nothing to rescale
                } else {
                    const bool need = (mx - m) > RESCALE_THRESHOLD;
                    if (__any_sync(0xffffffffu, need)) {
                        // rare: O_i is ours again only once the previous PV retired
                        mbar_wait(&pv_done[wg], (t - 1) & 1);
                        tc_fence_after();
                        pv_waited = true;
                        const float m_new = need ? mx : m;
                        const float alpha = ex2_approx(m - m_new);
                        uint32_t o[32];
                        tmem_ld32(t_o, o);
                        uint32_t rs = tmem_ld1(t_o + 32);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        rs = __float_as_uint(__uint_as_float(rs) * alpha);
                        tmem_st32(t_o, o);
                        tmem_st1(t_o + 32, rs);
                        m = m_new;
                    }
                }
                // all 128 exponentials first (the MUFU phase overlaps the previous PV's round trip) ...
                uint32_t pk[4][16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e0 = ex2_approx(__uint_as_float(s[c][2 * i]) - m);
                        const float e1 = ex2_approx(__uint_as_float(s[c][2 * i + 1]) - m);
                        pk[c][i] = pack_bf16x2(e0, e1);
                    }
                }
                // ... then P_i(t-1) must have been consumed before P_i(t) overwrites it
                TMARK(3, tphase);
                if (!pv_waited) {
                    mbar_wait(&pv_done[wg], (t - 1) & 1);
                    tc_fence_after();
                }
                TMARK(4, tphase);
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_st16(t_p + c * 16, pk[c]);
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(&p_full[wg]);
                TMARK(5, tphase);
                if (tim) atomicAdd(reinterpret_cast<unsigned long long*>(g_attn_timing + 6), 1ull);
            }
            // segment epilogue: normalised partial + log2-sum-exp into the workspace
            mbar_wait(&pv_done[wg], (t - 1) & 1);
            tc_fence_after();
            uint32_t o[32];
            tmem_ld32(t_o, o);
            const float l = __uint_as_float(tmem_ld1(t_o + 32));  // row sum from the ones column
            tc_wait_ld();
            const int slot = (item * p.qblocks + qb) * p.S_max +
                             (static_cast<int>(grp) - cta_of(static_cast<long long>(item) * p.T, p.W, G));
            const long long prow = static_cast<long long>(slot) * QBLK + wg * 128 + r;
            const float inv = 1.0f / l;
            float4* dst = reinterpret_cast<float4*>(p.part_o + prow * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                dst[i] = make_float4(__uint_as_float(o[4 * i]) * inv, __uint_as_float(o[4 * i + 1]) * inv,
                                     __uint_as_float(o[4 * i + 2]) * inv, __uint_as_float(o[4 * i + 3]) * inv);
            p.part_lse[prow] = m + log2f(l);
            tc_fence_before();
            pos += n;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 10) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Merge the per-CTA segments of each item.  One thread = (item row, 4 output dims).
template <bool kBf16>
__global__ void __launch_bounds__(256) tc_attn_merge_kernel(TcAttnParams p, long long G, void* o,
                                                            float* lse) {
    const long long items = static_cast<long long>(p.B) * p.H * p.qblocks;
    const long long total = items * attn::QBLK * 8;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int q4 = static_cast<int>(t & 7);
        const int rr = static_cast<int>((t >> 3) % attn::QBLK);
        const int item = static_cast<int>((t >> 3) / attn::QBLK);
        const int qb = item % p.qblocks;
        const int h = (item / p.qblocks) % p.H;
        const int b = item / (p.qblocks * p.H);
        const int row = qb * attn::QBLK + rr;
        if (row >= p.Nq) continue;
        const long long x0 = static_cast<long long>(item / p.qblocks) * p.T;
        const int nseg = cta_of(x0 + p.T - 1, p.W, G) - cta_of(x0, p.W, G) + 1;
        float mx = -INFINITY;
        for (int s = 0; s < nseg; ++s)
            mx = fmaxf(mx, p.part_lse[(static_cast<long long>(item) * p.S_max + s) * attn::QBLK + rr]);
        float den = 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < nseg; ++s) {
            const long long prow = (static_cast<long long>(item) * p.S_max + s) * attn::QBLK + rr;
            const float w = exp2f(p.part_lse[prow] - mx);
            den += w;
            const float4 x = reinterpret_cast<const float4*>(p.part_o + prow * 32)[q4];
            acc.x = fmaf(w, x.x, acc.x);
            acc.y = fmaf(w, x.y, acc.y);
            acc.z = fmaf(w, x.z, acc.z);
            acc.w = fmaf(w, x.w, acc.w);
        }
        const float inv = 1.0f / den;
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        const long long oidx = ((static_cast<long long>(b) * p.Nq + row) * p.H + h) * 8 + q4;  // float4 units
        if (kBf16) {
            uint2 w;
            w.x = pack_bf16x2(acc.x, acc.y);
            w.y = pack_bf16x2(acc.z, acc.w);
            reinterpret_cast<uint2*>(o)[oidx] = w;
        } else {
            reinterpret_cast<float4*>(o)[oidx] = acc;
        }
        if (lse != nullptr && q4 == 0)
            lse[(static_cast<long long>(b) * p.H + h) * p.Nq + row] = (mx + log2f(den)) * 0.6931471805599453f;
    }
}

static void attn_plan(int B, int H, int Nq, int n_tok, int sms, TcAttnParams* p, int* grid) {
    p->qblocks = (Nq + attn::QBLK - 1) / attn::QBLK;
    p->T = (n_tok + attn::KT - 1) / attn::KT;
    p->W = static_cast<long long>(B) * H * p->T;
    long long G = sms / p->qblocks;
    if (G > p->W) G = p->W;
    if (G < 1) G = 1;
    const long long chunk_min = p->W / G;  // >= 1
    p->S_max = static_cast<int>((p->T - 1) / chunk_min + 2);
    p->groups = static_cast<int>(G);
    *grid = static_cast<int>(G) * p->qblocks;
}

int tc_attn_set_timing_buffer(long long* dev_buf) {
    cudaError_t e = cudaMemcpyToSymbol(g_attn_timing, &dev_buf, sizeof(dev_buf));
    return e == cudaSuccess ? CMT_OK : cuda_fail(e, "tc_attn_set_timing_buffer");
}

size_t tc_attn_workspace_bytes(int B, int H, int Nq, int n_kv_tokens) {
    if (B <= 0 || H <= 0 || Nq <= 0 || n_kv_tokens <= 0) return 0;
    TcAttnParams p{};
    int grid;
    attn_plan(B, H, Nq, n_kv_tokens, device_sm_count(), &p, &grid);
    const size_t slots = static_cast<size_t>(B) * H * p.qblocks * p.S_max;
    return slots * attn::QBLK * 33 * sizeof(float) + 256;
}

int launch_tc_attn(const AttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    using namespace attn;
    const int n_tok = a.kv_end - a.kv_begin;
    CMT_CHECK_ARG(n_tok > 0, "cmt_cross_attn_fwd: empty token range");
    CMT_CHECK_ARG(a.q_ld % 8 == 0 && a.v_ld % 8 == 0 && a.k_bstride % 8 == 0 && a.k_hstride % 8 == 0 &&
                      a.v_bstride % 8 == 0 && a.v_hstride % 8 == 0,
                  "cmt_cross_attn_fwd(bf16): strides must be multiples of 8 elements");
    CMT_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) |
                    reinterpret_cast<uintptr_t>(a.vt) | reinterpret_cast<uintptr_t>(a.o)) & 15) == 0,
                  "cmt_cross_attn_fwd(bf16): pointers must be 16-byte aligned");
    TcAttnParams p{};
    int grid;
    attn_plan(a.B, a.H, a.Nq, n_tok, device_sm_count(), &p, &grid);
    p.B = a.B;
    p.H = a.H;
    p.Nq = a.Nq;
    p.kv_begin = a.kv_begin;
    p.kv_end = a.kv_end;
    const size_t need = tc_attn_workspace_bytes(a.B, a.H, a.Nq, n_tok);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("cmt_cross_attn_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
        return CMT_ERR_WORKSPACE;
    }
    const size_t slots = static_cast<size_t>(a.B) * a.H * p.qblocks * p.S_max;
    uintptr_t wsp = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    p.part_o = reinterpret_cast<float*>(wsp);
    p.part_lse = p.part_o + slots * QBLK * 32;

    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_attn)");
        attr_done = true;
    }
    CUtensorMap tq, tk, tv;
    {
        uint64_t dims[4] = {32, static_cast<uint64_t>(a.Nq), static_cast<uint64_t>(a.H), static_cast<uint64_t>(a.B)};
        uint64_t strides[3] = {static_cast<uint64_t>(a.q_ld) * 2, 64, static_cast<uint64_t>(a.Nq) * a.q_ld * 2};
        uint32_t box[4] = {32, 128, 1, 1};
        int rc = encode_tma_bf16(&tq, a.q, 4, dims, strides, box, 64);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {32, static_cast<uint64_t>(a.kv_end), static_cast<uint64_t>(a.H), static_cast<uint64_t>(a.B)};
        uint64_t strides[3] = {64, static_cast<uint64_t>(a.k_hstride) * 2, static_cast<uint64_t>(a.k_bstride) * 2};
        uint32_t box[4] = {32, KT, 1, 1};
        int rc = encode_tma_bf16(&tk, a.k, 4, dims, strides, box, 64);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {static_cast<uint64_t>(a.kv_end), 32, static_cast<uint64_t>(a.H), static_cast<uint64_t>(a.B)};
        uint64_t strides[3] = {static_cast<uint64_t>(a.v_ld) * 2, static_cast<uint64_t>(a.v_hstride) * 2,
                               static_cast<uint64_t>(a.v_bstride) * 2};
        uint32_t box[4] = {64, 32, 1, 1};
        int rc = encode_tma_bf16(&tv, a.vt, 4, dims, strides, box, 128);
        if (rc) return rc;
    }
    tc_attn_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
    CMT_LAUNCH_CHECK("cmt_cross_attn_fwd(tcgen05)");
    const long long total = static_cast<long long>(a.B) * a.H * p.qblocks * QBLK * 8;
    long long mblocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(device_sm_count()) * 8;
    if (mblocks > cap) mblocks = cap;
    if (a.o_bf16)
        tc_attn_merge_kernel<true><<<static_cast<int>(mblocks), 256, 0, stream>>>(p, p.groups, a.o, a.lse);
    else
        tc_attn_merge_kernel<false><<<static_cast<int>(mblocks), 256, 0, stream>>>(p, p.groups, a.o, a.lse);
    CMT_LAUNCH_CHECK("cmt_cross_attn_fwd(merge)");
    return CMT_OK;
}

}  // namespace cmt
