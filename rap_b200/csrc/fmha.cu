// Varlen (segment-packed) non-causal softmax attention, head_dim 64, fp16 in / fp32 accumulate, on
// tcgen05 + TMEM + TMA.  Replaces flash_attn.flash_attn_varlen_qkvpacked_func as called by the
// reference at flow_model/layer.py:106-111 (part-wise: segments = views) and :123-128 (global:
// segments = scenes).  The two calls differ only in the segment table, so this is one kernel.
//
// One CTA = one 256-row query block (two 128-row tiles) of one head, one CTA per SM (193 KB smem, 512 TMEM columns).
//   warp 0   : TMA producer  (Q tiles once; K and V blocks through separate 3-deep rings)
//   warp 1   : UMMA issuer   (S_t = Q_t K^T into TMEM cols [128t,128t+128); O_t += P_t V into cols [256+64t, ...))
//   warps 2-5: softmax of tile 0, warps 6-9: softmax of tile 1 (thread = query row: tcgen05.ld S, online max/sum in the
//              log2 domain, P -> smem as fp16 in the K-major SWIZZLE_128B layout UMMA reads, lazy O rescale in TMEM)
#include <cstdlib>

#include "fmha.cuh"
#include "sm100.cuh"

namespace rap {
namespace {

#define FMHA_TS(slot) do { if (dbg_on) p.dbg[(size_t)j * 16 + (slot)] = clock64(); } while (0)

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;         // 16 KB: 128 rows x 128 bytes
// NT = query tiles per CTA.  NT=2: one 256-row CTA per SM (K/V fetched once for both tiles, one UMMA warp ping-pongs them).
// NT=1: one 128-row CTA, two CTAs per SM (fully independent pipelines).  Selected per launch (RAP_FMHA_NT, default 2).
template <int NT> struct FCfg {
    static constexpr int KV_STAGES = NT == 2 ? 4 : 3;
    static constexpr int OFF_Q = 0, OFF_K = NT * TILE_BYTES, OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
    static constexpr int OFF_BAR = OFF_V + KV_STAGES * TILE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr int kThreads = 64 + NT * 128;   // TMA warp, UMMA warp, NT x 4 softmax warps
    static constexpr uint32_t TMEM_COLS = NT * 256;
    static_assert((3 - NT) * (SMEM_BYTES + 1024) <= 233472, "shared memory budget");
};
// TMEM columns of query tile t: S (fp32 scores) [256t, +128), P (fp16 pairs) [256t+128, +64), O (fp32) [256t+192, +64)
constexpr uint32_t TM_TILE = 256, TM_S = 0, TM_P = 128, TM_O = 192;
// Handing S back before the exponentials (possible whenever the gamma bound excludes overflow) was MEASURED slower: both query
// tiles then run in lockstep and fight for the MUFU pipe, while the release-after-exp order staggers them naturally.
constexpr bool kEarlyRelease = false;
constexpr float kRowSumLimit = 32768.0f;         // a block row sum >= 2^15 means some P may leave the fp16 range: re-reference that row

// One CTA = one 256-row query block (two 128-row tiles t=0,1) of one head; one CTA per SM.
// The UMMA warp ping-pongs between the two tiles -- QK0, PV0, QK1, PV1 per K/V block -- so while softmax warpgroup 0
// exponentiates block j of tile 0, the tensor core produces tile 1's scores and vice versa: the exp (MUFU) pipe, which
// bounds head_dim-64 attention, stays busy and each K/V block is fetched once for 256 query rows.
template <int NT>
__global__ void __launch_bounds__(FCfg<NT>::kThreads, 3 - NT)
fmha_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FmhaArgs p) {
    using F = FCfg<NT>;
    constexpr int KV_STAGES = F::KV_STAGES, OFF_Q = F::OFF_Q, OFF_K = F::OFF_K, OFF_V = F::OFF_V, OFF_BAR = F::OFF_BAR;
    constexpr uint32_t TMEM_COLS = F::TMEM_COLS;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int4 w = p.tiles[blockIdx.x];            // {q_row0, seg_start, seg_end, -}  (plan data, not produced by the previous kernel)
    const int q0 = w.x, kv_begin = w.y, kv_end = w.z;
    const int n_kv = (kv_end - kv_begin + BKV - 1) / BKV;
    const bool two = NT == 2 && q0 + BQ < kv_end;  // second 128-row tile has rows inside the segment

    const uint32_t bar = sbase + OFF_BAR;
    const uint32_t q_full = bar;
    auto k_full = [&](int s) { return bar + 8 + 8 * s; };
    auto k_empty = [&](int s) { return bar + 40 + 8 * s; };
    auto v_full = [&](int s) { return bar + 72 + 8 * s; };
    auto v_empty = [&](int s) { return bar + 104 + 8 * s; };
    auto s_full = [&](int t) { return bar + 136 + 8 * t; };
    auto p_full = [&](int t) { return bar + 152 + 8 * t; };
    auto pv_done = [&](int t) { return bar + 168 + 8 * t; };
    auto s_free = [&](int t) { return bar + 184 + 8 * t; };
    const uint32_t tmem_slot = bar + 200;

    if (threadIdx.x == 0) {
        if (sbase & 1023u) { atomicCAS(&g_dev_error, 0u, 0xA1u); __trap(); }
        tma_prefetch_desc(&tm_qkv);
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(s_full(t), 1); mbar_init(p_full(t), 4); mbar_init(pv_done(t), 1); mbar_init(s_free(t), 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 200);

    const int col_q = head * HD, col_k = p.d_model + head * HD, col_v = 2 * p.d_model + head * HD;
    pdl_wait();                                    // qkv comes from the previous kernel

    if (warp == 0) {
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(q_full, two ? 2 * TILE_BYTES : TILE_BYTES);
            tma_load_2d(sbase + OFF_Q, &tm_qkv, q_full, col_q, q0);
            if (two) tma_load_2d(sbase + OFF_Q + TILE_BYTES, &tm_qkv, q_full, col_q, q0 + BQ);
        }
        int s = 0; uint32_t ph = 0;
        for (int j = 0; j < n_kv; ++j) {
            const int row = kv_begin + j * BKV;
            mbar_wait(k_empty(s), ph ^ 1, 0x41);
            if (leader) {
                mbar_expect_tx(k_full(s), TILE_BYTES);
                tma_load_2d(sbase + OFF_K + s * TILE_BYTES, &tm_qkv, k_full(s), col_k, row);
            }
            mbar_wait(v_empty(s), ph ^ 1, 0x42);
            if (leader) {
                mbar_expect_tx(v_full(s), TILE_BYTES);
                tma_load_2d(sbase + OFF_V + s * TILE_BYTES, &tm_qkv, v_full(s), col_v, row);
            }
            if (++s == KV_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // Whole warp converged; only the UMMA / commit instructions are predicated on the elected lane.
        const bool leader = elect_one();
        const bool dbg_on = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && leader;
        constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV, false, false);   // A=Q K-major, B=K K-major, N=128
        constexpr uint32_t idesc_pv = umma_idesc_f16(BQ, HD, false, true);     // A=P K-major, B=V MN-major, N=64
        const int nt = two ? 2 : 1;
        auto issue_qk = [&](int t, int s) {
            const uint64_t dq = umma_desc_sw128(sbase + OFF_Q + t * TILE_BYTES, 1024);
            const uint64_t dk = umma_desc_sw128(sbase + OFF_K + s * TILE_BYTES, 1024);
            if (leader) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem + t * TM_TILE + TM_S, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
                umma_commit(s_full(t));
            }
        };
        mbar_wait(q_full, 0, 0x50);
        mbar_wait(k_full(0), 0, 0x51);
        tc_fence_after();
        for (int t = 0; t < nt; ++t) issue_qk(t, 0);
        if (leader) umma_commit(k_empty(0));
        int s = 0; uint32_t ph = 0;                      // stage / phase of K/V block j
        for (int j = 0; j < n_kv; ++j) {
            const int s1 = (s + 1 == KV_STAGES) ? 0 : s + 1;
            const uint32_t ph1 = (s + 1 == KV_STAGES) ? ph ^ 1 : ph;
            const bool more = j + 1 < n_kv;
            const uint64_t dv0 = umma_desc_sw128(sbase + OFF_V + s * TILE_BYTES, 1024);
            for (int t = 0; t < nt; ++t) {
                if (more) {                              // next scores of this tile as soon as its softmax is done with S_t(j)
                    if (t == 0) mbar_wait(k_full(s1), ph1, 0x53);
                    mbar_wait(s_free(t), j & 1, 0x55);
                    tc_fence_after();
                    FMHA_TS(8 + 4 * t);
                    issue_qk(t, s1);
                    if (t == nt - 1 && leader) umma_commit(k_empty(s1));
                }
                FMHA_TS(9 + 4 * t);
                if (t == 0) mbar_wait(v_full(s), ph, 0x54);
                mbar_wait(p_full(t), j & 1, 0x52);       // P_t(j) in TMEM, O_t rescaled if needed
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)   // A = P from TMEM: 8 columns per K=16 step; V: +2 KB per 16 rows (encoded >> 4)
                        umma_f16_ts(tmem + t * TM_TILE + TM_O, tmem + t * TM_TILE + TM_P + k * 8, dv0 + (uint64_t)(k * 128), idesc_pv, (j | k) != 0);
                    umma_commit(pv_done(t));
                    if (t == nt - 1) umma_commit(v_empty(s));
                }
                FMHA_TS(10 + 4 * t);
            }
            s = s1; ph = ph1;
        }
    } else {
        const int t = (warp - 2) >> 2;                   // query tile of this warpgroup
        const int qd = warp & 3;                         // TMEM lane quarter of this warp
        const int r = qd * 32 + lane;                    // query row inside the tile
        if (t == 0 || two) {
            const uint32_t t_lane = tmem + (uint32_t(qd * 32) << 16);
            const uint32_t tm_s = t_lane + t * TM_TILE + TM_S, tm_p = t_lane + t * TM_TILE + TM_P, tm_o = t_lane + t * TM_TILE + TM_O;
            const float sl2 = p.scale_log2;
            const bool dbg_on = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && warp == 2 && lane == 0;
            const float bound = p.score_bound ? p.score_bound[head] : INFINITY;
            float m_run = -INFINITY, l_run = 0.f;      // m_run: reference exponent (log2 domain), only raised lazily
            uint32_t sreg[4][32];
            for (int j = 0; j < n_kv; ++j) {
                FMHA_TS(0);
                mbar_wait(s_full(t), j & 1, 0x61);
                tc_fence_after();
                FMHA_TS(1);
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld32(tm_s + c * 32, sreg[c]);
                tmem_ld_wait();
                // If no score of this head can exceed the current reference by 2^15 (bound from the qk-norm gammas), the stale-reference
                // pass below cannot overflow, the slow path cannot trigger, and S_t can be handed back to the tensor core right now.
                const bool early = kEarlyRelease && j > 0 && __all_sync(0xffffffffu, bound - m_run < 14.0f);
                if (early) {
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(s_free(t));
                }
                FMHA_TS(2);
                const int valid = kv_end - (kv_begin + j * BKV);     // columns >= valid belong to another segment
                if (valid < BKV) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i >= valid) sreg[c][i] = 0xff800000u;   // -inf
                }
                if (j == 0) {                                        // first block: exact row max as the reference exponent
                    float mxa[8];
#pragma unroll
                    for (int a = 0; a < 8; ++a) mxa[a] = -INFINITY;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(sreg[c][i]));
                    m_run = sl2 * fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])), fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
                }
                const bool pv_ready = j == 0 || mbar_try_wait(pv_done(t), (j - 1) & 1);   // polled here, its latency hides under the exps
                FMHA_TS(3);
                // One pass: exponentials against the (possibly stale) reference exponent.  No per-element max tracking: the only thing
                // a stale reference can break is fp16 overflow of P, and any element >= 2^15 shows up in the block's row sum.
                f32x2 ls[4] = {0ull, 0ull, 0ull, 0ull};                // packed pairs: FFMA2 / FADD2 halve the issue slots around the MUFUs
                const f32x2 sl2x2 = pack2(sl2, sl2), nmx2 = pack2(-m_run, -m_run);
                uint32_t preg[64];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float ua, ub;
                        unpack2(ffma2(pack2(__uint_as_float(sreg[c][i]), __uint_as_float(sreg[c][i + 1])), sl2x2, nmx2), ua, ub);
                        const float a = ex2f(ua), b = ex2f(ub);
                        ls[(i >> 1) & 3] = fadd2(ls[(i >> 1) & 3], pack2(a, b));
                        preg[c * 16 + i / 2] = pack_half2(a, b);
                    }
                float lsum;
                {
                    float s0, s1;
                    unpack2(fadd2(fadd2(ls[0], ls[1]), fadd2(ls[2], ls[3])), s0, s1);
                    lsum = s0 + s1;
                }
                const bool grow = !(lsum < kRowSumLimit);            // also catches inf / NaN
                const bool slow = __any_sync(0xffffffffu, grow);
                if (!slow && !early) {                               // common case: the scores are consumed, let the tensor core
                    tc_fence_before();                               // overwrite S_t with block j+1 while P_t(j) is being stored
                    if (lane == 0) mbar_arrive(s_free(t));
                }
                FMHA_TS(4);
                if (!__all_sync(0xffffffffu, pv_ready)) mbar_wait(pv_done(t), (j - 1) & 1, 0x62);   // PV_t(j-1) retired: P columns free, O stable
                tc_fence_after();
                if (slow) {
                    // Rare slow path (some score exceeded the reference by ~2^15): take this block's exact row max as the new
                    // reference, redo the block from the scores still in TMEM, and rescale O / l.  (Rows that did not overflow
                    // keep their reference and recompute identical values.)
                    float bmax = -INFINITY;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        tmem_ld32(tm_s + c * 32, sreg[0]);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) bmax = fmaxf(bmax, (c * 32 + i < valid) ? __uint_as_float(sreg[0][i]) : -INFINITY);
                    }
                    const float m_new = grow ? bmax * sl2 : m_run;
                    const float alpha = ex2f(m_run - m_new);
                    m_run = m_new;
                    l_run *= alpha;
                    lsum = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        tmem_ld32(tm_s + c * 32, sreg[0]);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const bool v0 = c * 32 + i < valid, v1 = c * 32 + i + 1 < valid;
                            const float a = v0 ? ex2f(fmaf(__uint_as_float(sreg[0][i]), sl2, -m_run)) : 0.f;
                            const float b = v1 ? ex2f(fmaf(__uint_as_float(sreg[0][i + 1]), sl2, -m_run)) : 0.f;
                            lsum += a + b;
                            preg[c * 16 + i / 2] = pack_half2(a, b);
                        }
                    }
                    if (j > 0) {
                        uint32_t o[8];
#pragma unroll 1
                        for (int c = 0; c < HD / 8; ++c) {
                            tmem_ld8(tm_o + c * 8, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st8(tm_o + c * 8, o);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0 && !early) mbar_arrive(s_free(t));   // (early && slow cannot happen: the bound excludes overflow)
                }
                l_run += lsum;
                FMHA_TS(5);
                // P row -> TMEM (fp16 pairs, 64 columns): the A operand of the P.V UMMA
                {
                    uint32_t (&plo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&preg[0]);
                    uint32_t (&phi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&preg[32]);
                    tmem_st32(tm_p, plo);
                    tmem_st32(tm_p + 32, phi);
                    tmem_st_wait();
                }
                FMHA_TS(6);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(t));
                FMHA_TS(7);
            }
            // epilogue: O / l -> fp16
            mbar_wait(pv_done(t), (n_kv - 1) & 1, 0x63);
            tc_fence_after();
            const float inv_l = 1.0f / l_run;
            const int row = q0 + t * BQ + r;
            uint32_t o[32], o2[32];
            tmem_ld32(tm_o, o);
            tmem_ld32(tm_o + 32, o2);
            tmem_ld_wait();
            if (row < kv_end) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + (size_t)row * p.d_model + head * HD);
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    dst[i / 8] = make_uint4(pack_half2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l),
                                            pack_half2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l),
                                            pack_half2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l),
                                            pack_half2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l));
                    dst[4 + i / 8] = make_uint4(pack_half2(__uint_as_float(o2[i]) * inv_l, __uint_as_float(o2[i + 1]) * inv_l),
                                                pack_half2(__uint_as_float(o2[i + 2]) * inv_l, __uint_as_float(o2[i + 3]) * inv_l),
                                                pack_half2(__uint_as_float(o2[i + 4]) * inv_l, __uint_as_float(o2[i + 5]) * inv_l),
                                                pack_half2(__uint_as_float(o2[i + 6]) * inv_l, __uint_as_float(o2[i + 7]) * inv_l));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace

template <int NT>
static int launch_t(const CUtensorMap& tm_qkv, const FmhaArgs& a, int n_tiles, int heads, cudaStream_t st) {
    using F = FCfg<NT>;
    static bool configured = false;
    if (!configured) {
        RAP_CUDA(cudaFuncSetAttribute(fmha_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES));
        configured = true;
    }
    ProfScope _ps(KC_FMHA, st);
    RAP_CUDA(launch_pdl(fmha_kernel<NT>, dim3(n_tiles, heads), dim3(F::kThreads), F::SMEM_BYTES, st, tm_qkv, a));
    note_launch();
    return RAP_OK;
}

int fmha_rows_per_cta() {
    static int rows = 0;
    if (!rows) { const char* e = getenv("RAP_FMHA_NT"); rows = (e && atoi(e) == 1) ? 128 : 256; }
    return rows;
}

int launch_fmha(const CUtensorMap& tm_qkv, const FmhaArgs& a, int n_tiles, int heads, cudaStream_t st) {
    RAP_REQUIRE(n_tiles > 0 && heads > 0 && a.d_model == heads * HD, "fmha: bad args (tiles=%d heads=%d d=%d)", n_tiles, heads, a.d_model);
    return fmha_rows_per_cta() == 128 ? launch_t<1>(tm_qkv, a, n_tiles, heads, st) : launch_t<2>(tm_qkv, a, n_tiles, heads, st);
}

}  // namespace rap
