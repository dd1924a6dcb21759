// Varlen (segment-packed) non-causal softmax attention, head_dim 64, fp16 in / fp32 accumulate, on
// tcgen05 + TMEM + TMA.  Replaces flash_attn.flash_attn_varlen_qkvpacked_func as called by the
// reference at flow_model/layer.py:106-111 (part-wise: segments = views) and :123-128 (global:
// segments = scenes).  The two calls differ only in the segment table, so this is one kernel.
//
// One CTA = one work item: a 256-row query block (two 128-row tiles) of one head against a key range of its segment (the whole
// segment, or a piece of it when the unit belongs to the last, partially filled wave -- fmha.cuh), one CTA per SM (161 KB smem, 512 TMEM
// columns).
//   warp 0   : TMA producer  (Q tiles once; K and V blocks through separate 4-deep rings)
//   warp 1   : UMMA issuer   (per tile t: S_t = Q_t K^T -> TMEM cols [256t, +128);  O_t += P_t V -> cols [256t+192, +64), P_t read
//              straight from TMEM cols [256t+128, +64) as the A operand (TS mode), V as an MN-major shared-memory B operand)
//   warps 2-5: softmax of tile 0, warps 6-9: softmax of tile 1 (thread = query row: tcgen05.ld of the 128 scores, exponentials in the
//              log2 domain against a stale reference exponent on packed f32x2 math, P written back to TMEM as fp16 pairs with tcgen05.st;
//              O is only rescaled on the rare overflow path)
#include <cstdlib>

#include "fmha.cuh"
#include <algorithm>
#include <cstdlib>
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
constexpr int kWsRow = 68;                       // floats per partial row: 64 outputs, reference exponent, row sum, 2 pad (16-byte rows)
constexpr float kRowSumLimit = 32768.0f;         // a block row sum >= 2^15 means some P may leave the fp16 range: re-reference that row

// One CTA = one 256-row query block (two 128-row tiles t=0,1) of one head; one CTA per SM.
// The UMMA warp ping-pongs between the two tiles -- QK0, PV0, QK1, PV1 per K/V block -- so while softmax warpgroup 0
// exponentiates block j of tile 0, the tensor core produces tile 1's scores and vice versa: the exp (MUFU) pipe, which
// bounds head_dim-64 attention, stays busy and each K/V block is fetched once for 256 query rows.
template <int NT>
__global__ void __launch_bounds__(FCfg<NT>::kThreads + (NT == 2 ? 32 : 0), 3 - NT)
fmha_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FmhaArgs p) {
    using F = FCfg<NT>;
    constexpr int KV_STAGES = F::KV_STAGES, OFF_Q = F::OFF_Q, OFF_K = F::OFF_K, OFF_V = F::OFF_V, OFF_BAR = F::OFF_BAR;
    constexpr uint32_t TMEM_COLS = F::TMEM_COLS;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) p.dbg[15] = clock64();                 // CTA entry
    const FmhaItem item = p.items[blockIdx.x];     // plan data, not produced by the previous kernel
    const int head = item.head;
    const int q0 = item.q_row0, seg_end = item.seg_end;
    const int kv_begin = item.kv_lo, kv_end = item.kv_hi;            // this piece's keys (the whole segment unless the unit was split)
    const int n_kv = (kv_end - kv_begin + BKV - 1) / BKV;
    const bool two = NT == 2 && q0 + BQ < seg_end; // second 128-row tile has rows inside the segment
    // mma2: each query tile has its OWN UMMA-issuing warp (warp 1: tile 0, warp 10: tile 1).  One warp issuing for both tiles blocks on the
    // in-order tensor queue (issuing tile 0's P.V, 8 UMMAs behind its QK^T) while tile 1's scores could already be requested, and vice versa.
    const bool mma2 = NT == 2 && p.mma2 && two && blockDim.x > (unsigned)FCfg<NT>::kThreads;

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
        for (int s = 0; s < KV_STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), mma2 ? 2 : 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), mma2 ? 2 : 1); }
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
    } else if (warp == 1 || warp == 10) {
        // Whole warp converged; only the UMMA / commit instructions are predicated on the elected lane.  Tiles [t_lo, t_hi) of this warp:
        // both for warp 1 alone, one each with mma2.
        const int nt = two ? 2 : 1;
        const int t_lo = (warp == 10) ? 1 : 0, t_hi = mma2 ? t_lo + 1 : (warp == 10 ? 0 : nt);     // (warp 10 idles unless mma2)
        const bool leader = elect_one();
        const bool dbg_on = p.dbg && blockIdx.x == 0 && leader && warp == 1;
        constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV, false, false);   // A=Q K-major, B=K K-major, N=128
        constexpr uint32_t idesc_pv = umma_idesc_f16(BQ, HD, false, true);     // A=P K-major, B=V MN-major, N=64
        auto issue_qk = [&](int t, int s) {
            const uint64_t dq = umma_desc_sw128(sbase + OFF_Q + t * TILE_BYTES, 1024);
            const uint64_t dk = umma_desc_sw128(sbase + OFF_K + s * TILE_BYTES, 1024);
            if (leader) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem + t * TM_TILE + TM_S, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
                umma_commit(s_full(t));
            }
        };
        if (t_lo < t_hi) {
            mbar_wait(q_full, 0, 0x50);
            mbar_wait(k_full(0), 0, 0x51);
            tc_fence_after();
            for (int t = t_lo; t < t_hi; ++t) issue_qk(t, 0);
            if (leader) umma_commit(k_empty(0));             // (with mma2 the stage is free when both warps' commits have arrived)
            int s = 0; uint32_t ph = 0;                      // stage / phase of K/V block j
            for (int j = 0; j < n_kv; ++j) {
                const int s1 = (s + 1 == KV_STAGES) ? 0 : s + 1;
                const uint32_t ph1 = (s + 1 == KV_STAGES) ? ph ^ 1 : ph;
                const bool more = j + 1 < n_kv;
                const uint64_t dv0 = umma_desc_sw128(sbase + OFF_V + s * TILE_BYTES, 1024);
                for (int t = t_lo; t < t_hi; ++t) {
                    if (more) {                              // next scores of this tile as soon as its softmax is done with S_t(j)
                        if (t == t_lo) mbar_wait(k_full(s1), ph1, 0x53);
                        mbar_wait(s_free(t), j & 1, 0x55);
                        tc_fence_after();
                        FMHA_TS(8 + 4 * t);
                        issue_qk(t, s1);
                        if (t == t_hi - 1 && leader) umma_commit(k_empty(s1));
                    }
                    FMHA_TS(9 + 4 * t);
                    if (t == t_lo) mbar_wait(v_full(s), ph, 0x54);
                    mbar_wait(p_full(t), j & 1, 0x52);       // P_t(j) in TMEM, O_t rescaled if needed
                    tc_fence_after();
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)   // A = P from TMEM: 8 columns per K=16 step; V: +2 KB per 16 rows (encoded >> 4)
                            umma_f16_ts(tmem + t * TM_TILE + TM_O, tmem + t * TM_TILE + TM_P + k * 8, dv0 + (uint64_t)(k * 128), idesc_pv, (j | k) != 0);
                        umma_commit(pv_done(t));
                        if (t == t_hi - 1) umma_commit(v_empty(s));
                    }
                    FMHA_TS(10 + 4 * t);
                }
                s = s1; ph = ph1;
            }
        }
    } else {
        const int t = (warp - 2) >> 2;                   // query tile of this warpgroup
        const int qd = warp & 3;                         // TMEM lane quarter of this warp
        const int r = qd * 32 + lane;                    // query row inside the tile
        if (t == 0 || two) {
            const uint32_t t_lane = tmem + (uint32_t(qd * 32) << 16);
            const uint32_t tm_s = t_lane + t * TM_TILE + TM_S, tm_p = t_lane + t * TM_TILE + TM_P, tm_o = t_lane + t * TM_TILE + TM_O;
            const float sl2 = p.scale_log2;
            const bool dbg_on = p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0;
            const float bound = p.score_bound ? p.score_bound[head] : INFINITY;
            float m_run = -INFINITY, l_run = 0.f;      // m_run: reference exponent (log2 domain), only raised lazily
            uint32_t sreg[4][32];
            for (int j = 0; j < n_kv; ++j) {
                FMHA_TS(0);
                mbar_wait(s_full(t), j & 1, 0x61);
                tc_fence_after();
                FMHA_TS(1);
                // (Loading the scores in two halves, the second still in flight under the first 64 exponentials, was measured 7 % SLOWER:
                // the extra tcgen05.wait::ld splits the exp stream and the two tiles lose their stagger.)
                const int valid = kv_end - (kv_begin + j * BKV);     // columns >= valid belong to another segment
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld32(tm_s + c * 32, sreg[c]);
                tmem_ld_wait();
                // If no score of this head can exceed the current reference by 2^15 (bound from the qk-norm gammas), the stale-reference
                // pass below cannot overflow, the slow path cannot trigger, and S_t can be handed back to the tensor core right now.
                const bool early = p.early_release && j > 0 && __all_sync(0xffffffffu, bound - m_run < 14.0f);
                if (early) {
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(s_free(t));
                }
                FMHA_TS(2);
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
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float ua, ub;
                        unpack2(ffma2(pack2(__uint_as_float(sreg[c][i]), __uint_as_float(sreg[c][i + 1])), sl2x2, nmx2), ua, ub);
                        const float a = ex2f(ua), b = ex2f(ub);
                        ls[(i >> 1) & 3] = fadd2(ls[(i >> 1) & 3], pack2(a, b));
                        preg[c * 16 + i / 2] = pack_half2(a, b);
                    }
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
            // epilogue
            mbar_wait(pv_done(t), (n_kv - 1) & 1, 0x63);
            tc_fence_after();
            const int row = q0 + t * BQ + r;
            uint32_t o[32], o2[32];
            tmem_ld32(tm_o, o);
            tmem_ld32(tm_o + 32, o2);
            tmem_ld_wait();
            bool write_out = true;
            if (item.slot >= 0) {
                // Split unit: publish this piece's un-normalised row (O, reference exponent, row sum); the piece that arrives last
                // re-references all of them to the largest exponent, adds them up and writes the output.
                const int piece = item.piece & 0xff, n_pieces = item.piece >> 8;
                const int nthreads = two ? 256 : 128;
                float* unit_ws = p.ws + (size_t)item.slot * p.ws_pieces * (NT * BQ) * kWsRow;
                float4* mine = reinterpret_cast<float4*>(unit_ws + ((size_t)piece * (NT * BQ) + t * BQ + r) * kWsRow);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    mine[i / 4] = make_float4(__uint_as_float(o[i]), __uint_as_float(o[i + 1]), __uint_as_float(o[i + 2]), __uint_as_float(o[i + 3]));
                    mine[8 + i / 4] = make_float4(__uint_as_float(o2[i]), __uint_as_float(o2[i + 1]), __uint_as_float(o2[i + 2]), __uint_as_float(o2[i + 3]));
                }
                mine[16] = make_float4(m_run, l_run, 0.f, 0.f);
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
                if (warp == 2 && lane == 0) {
                    const int arrived = atomicAdd(p.counters + item.slot, 1);
                    if (arrived == n_pieces - 1) p.counters[item.slot] = 0;      // every piece has arrived: ready for the next launch
                    *reinterpret_cast<volatile int*>(smem + OFF_BAR + 208) = arrived == n_pieces - 1;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
                write_out = *reinterpret_cast<volatile int*>(smem + OFF_BAR + 208) != 0;
                if (write_out) {
                    // Pieces are added in piece order from the workspace (this CTA's own rows included), so the result does not depend
                    // on which piece happened to arrive last.
                    __threadfence();
                    const float* rows = unit_ws + (size_t)(t * BQ + r) * kWsRow;
                    const size_t piece_stride = (size_t)(NT * BQ) * kWsRow;
                    float m_all = -INFINITY;
                    for (int q = 0; q < n_pieces; ++q) m_all = fmaxf(m_all, __ldcg(rows + q * piece_stride + 64));
                    l_run = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) { o[i] = 0u; o2[i] = 0u; }
                    for (int q = 0; q < n_pieces; ++q) {
                        const float4* src = reinterpret_cast<const float4*>(rows + q * piece_stride);
                        const float4 ml = __ldcg(src + 16);
                        const float wq = ex2f(ml.x - m_all);
                        l_run = fmaf(ml.y, wq, l_run);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 a = __ldcg(src + i / 4), b = __ldcg(src + 8 + i / 4);
                            o[i] = __float_as_uint(fmaf(a.x, wq, __uint_as_float(o[i])));
                            o[i + 1] = __float_as_uint(fmaf(a.y, wq, __uint_as_float(o[i + 1])));
                            o[i + 2] = __float_as_uint(fmaf(a.z, wq, __uint_as_float(o[i + 2])));
                            o[i + 3] = __float_as_uint(fmaf(a.w, wq, __uint_as_float(o[i + 3])));
                            o2[i] = __float_as_uint(fmaf(b.x, wq, __uint_as_float(o2[i])));
                            o2[i + 1] = __float_as_uint(fmaf(b.y, wq, __uint_as_float(o2[i + 1])));
                            o2[i + 2] = __float_as_uint(fmaf(b.z, wq, __uint_as_float(o2[i + 2])));
                            o2[i + 3] = __float_as_uint(fmaf(b.w, wq, __uint_as_float(o2[i + 3])));
                        }
                    }
                }
            }
            const float inv_l = 1.0f / l_run;
            if (write_out && row < seg_end) {
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
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) p.dbg[16 + 15] = clock64();            // softmax warp 2 done (epilogue written)
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}


// ======================================================================================= three query tiles, 64-key blocks (RAP_FMHA_NT=3)
// The two-tile kernel keeps TWO softmax warps on every scheduler; whenever one of them is loading scores / storing P / waiting, the other
// streams its exponentials alone at ~12.5 cycles per MUFU instruction instead of the pipe's 8, which is where its 18 % to the MUFU bound go
// (tools/fmha_timeline.py; removing the waits with a second UMMA warp only moved the time into MUFU contention).  Here a CTA owns THREE
// 128-row query tiles (384 rows) of one head and walks 64-key blocks, so three softmax warps share each scheduler: TMEM per tile = S 64 + P 32
// + O 64 columns (3 x 160 = 480 of 512), 14 warps (TMA, UMMA, 3 x 4 softmax).  Everything else -- stale reference exponent, overflow path,
// TS-mode P.V, key-split units with the handshake combine -- is the two-tile kernel's.
namespace k3 {
constexpr int NT = 3, BK3 = 64, STAGES = 6;
constexpr int KV_TILE = BK3 * HD * 2;                  // 8 KB: 64 keys x 128 bytes
constexpr int OFF_Q = 0, OFF_K = NT * TILE_BYTES, OFF_V = OFF_K + STAGES * KV_TILE, OFF_BAR = OFF_V + STAGES * KV_TILE;
constexpr int SMEM_BYTES = OFF_BAR + 512;
constexpr int kThreads = 64 + NT * 128;
constexpr uint32_t TM_TILE3 = 160, TM_S3 = 0, TM_P3 = 64, TM_O3 = 96;
}  // namespace k3

__global__ void __launch_bounds__(k3::kThreads + 64, 1)
fmha3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const FmhaArgs p) {
    using namespace k3;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FmhaItem item = p.items[blockIdx.x];
    const int head = item.head, q0 = item.q_row0, seg_end = item.seg_end;
    const int kv_begin = item.kv_lo, kv_end = item.kv_hi;
    const int n_kv = (kv_end - kv_begin + BK3 - 1) / BK3;
    const int nt = min(NT, (seg_end - q0 + BQ - 1) / BQ);           // query tiles with rows inside the segment
    // mma3 (launch with 512 threads): one UMMA-issuing warp per query tile (warps 1, 14, 15) instead of one warp walking the tiles in a fixed
    // order -- with three tiles the fixed order keeps them in lockstep.
    const bool mma3 = blockDim.x > (unsigned)kThreads;

    const uint32_t bar = sbase + OFF_BAR;
    const uint32_t q_full = bar;
    auto k_full = [&](int s) { return bar + 8 + 8 * s; };
    auto k_empty = [&](int s) { return bar + 56 + 8 * s; };
    auto v_full = [&](int s) { return bar + 104 + 8 * s; };
    auto v_empty = [&](int s) { return bar + 152 + 8 * s; };
    auto s_full = [&](int t) { return bar + 200 + 8 * t; };
    auto p_full = [&](int t) { return bar + 224 + 8 * t; };
    auto pv_done = [&](int t) { return bar + 248 + 8 * t; };
    auto s_free = [&](int t) { return bar + 272 + 8 * t; };
    const uint32_t tmem_slot = bar + 296;

    if (threadIdx.x == 0) {
        if (sbase & 1023u) { atomicCAS(&g_dev_error, 0u, 0xA1u); __trap(); }
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), mma3 ? nt : 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), mma3 ? nt : 1); }
        for (int t = 0; t < NT; ++t) { mbar_init(s_full(t), 1); mbar_init(p_full(t), 4); mbar_init(pv_done(t), 1); mbar_init(s_free(t), 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 296);
    const int col_q = head * HD, col_k = p.d_model + head * HD, col_v = 2 * p.d_model + head * HD;
    pdl_wait();                                    // qkv comes from the previous kernel

    if (warp == 0) {
        const bool leader = elect_one();
        if (leader) {
            mbar_expect_tx(q_full, nt * TILE_BYTES);
            for (int t = 0; t < nt; ++t) tma_load_2d(sbase + OFF_Q + t * TILE_BYTES, &tm_q, q_full, col_q, q0 + t * BQ);
        }
        int s = 0; uint32_t ph = 0;
        for (int j = 0; j < n_kv; ++j) {
            const int row = kv_begin + j * BK3;
            mbar_wait(k_empty(s), ph ^ 1, 0x41);
            if (leader) { mbar_expect_tx(k_full(s), KV_TILE); tma_load_2d(sbase + OFF_K + s * KV_TILE, &tm_kv, k_full(s), col_k, row); }
            mbar_wait(v_empty(s), ph ^ 1, 0x42);
            if (leader) { mbar_expect_tx(v_full(s), KV_TILE); tma_load_2d(sbase + OFF_V + s * KV_TILE, &tm_kv, v_full(s), col_v, row); }
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1 || warp >= 14) {
        const bool leader = elect_one();
        const int t_lo = mma3 ? (warp == 1 ? 0 : warp - 13) : 0, t_hi = mma3 ? min(t_lo + 1, nt) : (warp == 1 ? nt : 0);
        constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BK3, false, false);   // A = Q (K-major), B = K (K-major), N = 64 keys
        constexpr uint32_t idesc_pv = umma_idesc_f16(BQ, HD, false, true);     // A = P (TMEM), B = V (MN-major), N = 64 dims
        auto issue_qk = [&](int t, int s) {
            const uint64_t dq = umma_desc_sw128(sbase + OFF_Q + t * TILE_BYTES, 1024);
            const uint64_t dk = umma_desc_sw128(sbase + OFF_K + s * KV_TILE, 1024);
            if (leader) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem + t * TM_TILE3 + TM_S3, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
                umma_commit(s_full(t));
            }
        };
        if (t_lo < t_hi) {
        mbar_wait(q_full, 0, 0x50);
        mbar_wait(k_full(0), 0, 0x51);
        tc_fence_after();
        for (int t = t_lo; t < t_hi; ++t) issue_qk(t, 0);
        if (leader) umma_commit(k_empty(0));
        int s = 0; uint32_t ph = 0;
        for (int j = 0; j < n_kv; ++j) {
            const int s1 = (s + 1 == STAGES) ? 0 : s + 1;
            const uint32_t ph1 = (s + 1 == STAGES) ? ph ^ 1 : ph;
            const bool more = j + 1 < n_kv;
            const uint64_t dv0 = umma_desc_sw128(sbase + OFF_V + s * KV_TILE, 1024);
            for (int t = t_lo; t < t_hi; ++t) {
                if (more) {
                    if (t == t_lo) mbar_wait(k_full(s1), ph1, 0x53);
                    mbar_wait(s_free(t), j & 1, 0x55);
                    tc_fence_after();
                    issue_qk(t, s1);
                    if (t == t_hi - 1 && leader) umma_commit(k_empty(s1));
                }
                if (t == t_lo) mbar_wait(v_full(s), ph, 0x54);
                mbar_wait(p_full(t), j & 1, 0x52);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int k = 0; k < BK3 / 16; ++k)       // A = P from TMEM: 8 columns per K=16 step; V: +2 KB per 16 keys (encoded >> 4)
                        umma_f16_ts(tmem + t * TM_TILE3 + TM_O3, tmem + t * TM_TILE3 + TM_P3 + k * 8, dv0 + (uint64_t)(k * 128), idesc_pv, (j | k) != 0);
                    umma_commit(pv_done(t));
                    if (t == t_hi - 1) umma_commit(v_empty(s));
                }
            }
            s = s1; ph = ph1;
        }
        }
    } else {
        const int t = (warp - 2) >> 2;                   // query tile of this warpgroup
        const int qd = warp & 3;                         // TMEM lane quarter of this warp
        const int r = qd * 32 + lane;                    // query row inside the tile
        if (t < nt) {
            const uint32_t t_lane = tmem + (uint32_t(qd * 32) << 16);
            const uint32_t tm_s = t_lane + t * TM_TILE3 + TM_S3, tm_p = t_lane + t * TM_TILE3 + TM_P3, tm_o = t_lane + t * TM_TILE3 + TM_O3;
            const float sl2 = p.scale_log2;
            float m_run = -INFINITY, l_run = 0.f;
            uint32_t sreg[2][32];
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(s_full(t), j & 1, 0x61);
                tc_fence_after();
                const int valid = kv_end - (kv_begin + j * BK3);
                tmem_ld32(tm_s, sreg[0]);
                tmem_ld32(tm_s + 32, sreg[1]);
                tmem_ld_wait();
                if (valid < BK3) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i >= valid) sreg[c][i] = 0xff800000u;   // -inf
                }
                if (j == 0) {
                    float mxa[8];
#pragma unroll
                    for (int a = 0; a < 8; ++a) mxa[a] = -INFINITY;
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(sreg[c][i]));
                    m_run = sl2 * fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])), fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
                }
                const bool pv_ready = j == 0 || mbar_try_wait(pv_done(t), (j - 1) & 1);
                f32x2 ls[4] = {0ull, 0ull, 0ull, 0ull};
                const f32x2 sl2x2 = pack2(sl2, sl2), nmx2 = pack2(-m_run, -m_run);
                uint32_t preg[32];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float ua, ub;
                        unpack2(ffma2(pack2(__uint_as_float(sreg[c][i]), __uint_as_float(sreg[c][i + 1])), sl2x2, nmx2), ua, ub);
                        const float a = ex2f(ua), b = ex2f(ub);
                        ls[(i >> 1) & 3] = fadd2(ls[(i >> 1) & 3], pack2(a, b));
                        preg[c * 16 + i / 2] = pack_half2(a, b);
                    }
                }
                float lsum;
                {
                    float s0, s1;
                    unpack2(fadd2(fadd2(ls[0], ls[1]), fadd2(ls[2], ls[3])), s0, s1);
                    lsum = s0 + s1;
                }
                const bool grow = !(lsum < kRowSumLimit);
                const bool slow = __any_sync(0xffffffffu, grow);
                if (!slow) {
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(s_free(t));
                }
                if (!__all_sync(0xffffffffu, pv_ready)) mbar_wait(pv_done(t), (j - 1) & 1, 0x62);
                tc_fence_after();
                if (slow) {
                    float bmax = -INFINITY;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
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
                    for (int c = 0; c < 2; ++c) {
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
                    if (lane == 0) mbar_arrive(s_free(t));
                }
                l_run += lsum;
                tmem_st32(tm_p, preg);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(t));
            }
            // epilogue
            mbar_wait(pv_done(t), (n_kv - 1) & 1, 0x63);
            tc_fence_after();
            const int row = q0 + t * BQ + r;
            uint32_t o[32], o2[32];
            tmem_ld32(tm_o, o);
            tmem_ld32(tm_o + 32, o2);
            tmem_ld_wait();
            bool write_out = true;
            if (item.slot >= 0) {
                const int piece = item.piece & 0xff, n_pieces = item.piece >> 8;
                const int nthreads = nt * 128;
                float* unit_ws = p.ws + (size_t)item.slot * p.ws_pieces * (NT * BQ) * kWsRow;
                float4* mine = reinterpret_cast<float4*>(unit_ws + ((size_t)piece * (NT * BQ) + t * BQ + r) * kWsRow);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    mine[i / 4] = make_float4(__uint_as_float(o[i]), __uint_as_float(o[i + 1]), __uint_as_float(o[i + 2]), __uint_as_float(o[i + 3]));
                    mine[8 + i / 4] = make_float4(__uint_as_float(o2[i]), __uint_as_float(o2[i + 1]), __uint_as_float(o2[i + 2]), __uint_as_float(o2[i + 3]));
                }
                mine[16] = make_float4(m_run, l_run, 0.f, 0.f);
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
                if (warp == 2 && lane == 0) {
                    const int arrived = atomicAdd(p.counters + item.slot, 1);
                    if (arrived == n_pieces - 1) p.counters[item.slot] = 0;
                    *reinterpret_cast<volatile int*>(smem + OFF_BAR + 304) = arrived == n_pieces - 1;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
                write_out = *reinterpret_cast<volatile int*>(smem + OFF_BAR + 304) != 0;
                if (write_out) {
                    __threadfence();
                    const float* rows = unit_ws + (size_t)(t * BQ + r) * kWsRow;
                    const size_t piece_stride = (size_t)(NT * BQ) * kWsRow;
                    float m_all = -INFINITY;
                    for (int q = 0; q < n_pieces; ++q) m_all = fmaxf(m_all, __ldcg(rows + q * piece_stride + 64));
                    l_run = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) { o[i] = 0u; o2[i] = 0u; }
                    for (int q = 0; q < n_pieces; ++q) {
                        const float4* src = reinterpret_cast<const float4*>(rows + q * piece_stride);
                        const float4 ml = __ldcg(src + 16);
                        const float wq = ex2f(ml.x - m_all);
                        l_run = fmaf(ml.y, wq, l_run);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 a = __ldcg(src + i / 4), b = __ldcg(src + 8 + i / 4);
                            o[i] = __float_as_uint(fmaf(a.x, wq, __uint_as_float(o[i])));
                            o[i + 1] = __float_as_uint(fmaf(a.y, wq, __uint_as_float(o[i + 1])));
                            o[i + 2] = __float_as_uint(fmaf(a.z, wq, __uint_as_float(o[i + 2])));
                            o[i + 3] = __float_as_uint(fmaf(a.w, wq, __uint_as_float(o[i + 3])));
                            o2[i] = __float_as_uint(fmaf(b.x, wq, __uint_as_float(o2[i])));
                            o2[i + 1] = __float_as_uint(fmaf(b.y, wq, __uint_as_float(o2[i + 1])));
                            o2[i + 2] = __float_as_uint(fmaf(b.z, wq, __uint_as_float(o2[i + 2])));
                            o2[i + 3] = __float_as_uint(fmaf(b.w, wq, __uint_as_float(o2[i + 3])));
                        }
                    }
                }
            }
            const float inv_l = 1.0f / l_run;
            if (write_out && row < seg_end) {
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
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// ======================================================================================= persistent (stream-K) variant
// Same per-block pipeline as fmha_kernel<2>, but ONE CTA per SM walks a list of items (FmhaSchedule::cta_ptr): barriers, TMEM and the
// K/V ring are set up once and keep running across items.  Phase bookkeeping is by running counters instead of the block index:
//   cq[t] QK^T issues / cp[t] P.V issues on tile t (UMMA warp), cs blocks consumed (softmax warps of tile t), n items (Q buffer).
// Hand-over rules at an item boundary (everything else is the steady-state protocol):
//   * Q tiles of item n+1 are loaded once the last QK^T of item n has retired (q_empty, committed behind it);
//   * QK^T(0) of item n+1 waits, like any QK^T, for s_free of the tile's previous block (the last block of item n);
//   * P.V(0) of item n+1 overwrites O (accumulate = 0): it waits for p_full, which the softmax warps only signal after their epilogue of
//     item n has read O out of TMEM (program order in those warps), so no extra barrier is needed.
// No CTA ever waits for another CTA: split units only publish partial rows, the piece that arrives last combines them.
constexpr int kPersistThreads = FCfg<2>::kThreads + 32;     // + warp 10: publishes the "partial rows written" flag of a split unit's later piece
__global__ void __launch_bounds__(kPersistThreads, 1)
fmha_persist_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FmhaArgs p) {
    using F = FCfg<2>;
    constexpr int NT = 2;
    constexpr int KV_STAGES = F::KV_STAGES, OFF_Q = F::OFF_Q, OFF_K = F::OFF_K, OFF_V = F::OFF_V, OFF_BAR = F::OFF_BAR;
    constexpr uint32_t TMEM_COLS = F::TMEM_COLS;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int it0 = p.cta_ptr[blockIdx.x], it1 = p.cta_ptr[blockIdx.x + 1];      // plan data, not produced by the previous kernel
    long long dbg_t0 = 0, dbg_c0 = 0;
    if (p.dbg && threadIdx.x == 64) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0)); dbg_c0 = clock64(); }

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
    const uint32_t q_empty = bar + 216;
    auto part_ready = [&](int t) { return bar + 224 + 8 * t; };

    if (threadIdx.x == 0) {
        if (sbase & 1023u) { atomicCAS(&g_dev_error, 0u, 0xA1u); __trap(); }
        tma_prefetch_desc(&tm_qkv);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        mbar_init(part_ready(0), 128);
        mbar_init(part_ready(1), 128);
        for (int s = 0; s < KV_STAGES; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(s_full(t), 1); mbar_init(p_full(t), 4); mbar_init(pv_done(t), 1); mbar_init(s_free(t), 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 200);
    pdl_wait();                                    // qkv comes from the previous kernel

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        const bool leader = elect_one();
        int s = 0; uint32_t ph = 0;
        for (int it = it0, n = 0; it < it1; ++it, ++n) {
            const FmhaItem item = p.items[it];
            const int q0 = item.q_row0, kv_begin = item.kv_lo;
            const int n_kv = (item.kv_hi - kv_begin + BKV - 1) / BKV;
            const bool two = q0 + BQ < item.seg_end;
            const int col_q = item.head * HD, col_k = p.d_model + item.head * HD, col_v = 2 * p.d_model + item.head * HD;
            if (n > 0) mbar_wait(q_empty, (n - 1) & 1, 0x40);          // every QK^T of the previous item has read its Q tiles
            if (leader) {
                mbar_expect_tx(q_full, two ? 2 * TILE_BYTES : TILE_BYTES);
                tma_load_2d(sbase + OFF_Q, &tm_qkv, q_full, col_q, q0);
                if (two) tma_load_2d(sbase + OFF_Q + TILE_BYTES, &tm_qkv, q_full, col_q, q0 + BQ);
            }
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
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ UMMA issuer (warp converged, one elected lane issues)
        const bool leader = elect_one();
        constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV, false, false);   // A=Q K-major, B=K K-major, N=128
        constexpr uint32_t idesc_pv = umma_idesc_f16(BQ, HD, false, true);     // A=P K-major, B=V MN-major, N=64
        auto issue_qk = [&](int t, int s) {
            const uint64_t dq = umma_desc_sw128(sbase + OFF_Q + t * TILE_BYTES, 1024);
            const uint64_t dk = umma_desc_sw128(sbase + OFF_K + s * TILE_BYTES, 1024);
            if (leader) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem + t * TM_TILE + TM_S, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
                umma_commit(s_full(t));
            }
        };
        int s = 0; uint32_t ph = 0;                      // K/V ring position of the current block
        uint32_t cq[2] = {0, 0}, cp[2] = {0, 0};
        for (int it = it0, n = 0; it < it1; ++it, ++n) {
            const FmhaItem item = p.items[it];
            const int n_kv = (item.kv_hi - item.kv_lo + BKV - 1) / BKV;
            const int nt = (item.q_row0 + BQ < item.seg_end) ? 2 : 1;
            mbar_wait(q_full, n & 1, 0x50);
            mbar_wait(k_full(s), ph, 0x51);
            for (int t = 0; t < nt; ++t) {
                if (cq[t] > 0) mbar_wait(s_free(t), (cq[t] - 1) & 1, 0x56);     // the tile's previous block (of the previous item) is consumed
                tc_fence_after();
                issue_qk(t, s);
                ++cq[t];
            }
            if (leader) { umma_commit(k_empty(s)); if (n_kv == 1) umma_commit(q_empty); }
            for (int j = 0; j < n_kv; ++j) {
                const int s1 = (s + 1 == KV_STAGES) ? 0 : s + 1;
                const uint32_t ph1 = (s + 1 == KV_STAGES) ? ph ^ 1 : ph;
                const bool more = j + 1 < n_kv;
                const uint64_t dv0 = umma_desc_sw128(sbase + OFF_V + s * TILE_BYTES, 1024);
                for (int t = 0; t < nt; ++t) {
                    if (more) {                              // next scores of this tile as soon as its softmax is done with S_t(j)
                        if (t == 0) mbar_wait(k_full(s1), ph1, 0x53);
                        mbar_wait(s_free(t), (cq[t] - 1) & 1, 0x55);
                        tc_fence_after();
                        issue_qk(t, s1);
                        ++cq[t];
                        if (t == nt - 1 && leader) { umma_commit(k_empty(s1)); if (j + 2 == n_kv) umma_commit(q_empty); }
                    }
                    if (t == 0) mbar_wait(v_full(s), ph, 0x54);
                    mbar_wait(p_full(t), cp[t] & 1, 0x52);   // P_t(j) in TMEM, O_t rescaled if needed (j = 0: O_t of the previous item read out)
                    tc_fence_after();
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)   // A = P from TMEM: 8 columns per K=16 step; V: +2 KB per 16 rows (encoded >> 4)
                            umma_f16_ts(tmem + t * TM_TILE + TM_O, tmem + t * TM_TILE + TM_P + k * 8, dv0 + (uint64_t)(k * 128), idesc_pv, (j | k) != 0);
                        umma_commit(pv_done(t));
                        if (t == nt - 1) umma_commit(v_empty(s));
                    }
                    ++cp[t];
                }
                s = s1; ph = ph1;
            }
        }
    } else if (warp == 10) {
        // ------------------------------------------------------------ publisher: a later piece (index >= 1) of a split unit is by construction the
        // FIRST item of its CTA.  Its softmax threads store their partial rows and arrive on part_ready (cta-scope release); this warp then
        // issues the gpu-scope fence (cumulative over what it synchronised with) and raises the piece's flag, so the fence latency is off the
        // softmax warps' critical path.  The unit's head piece (last item of the preceding CTA) polls the flag.
        if (it0 < it1) {
            const FmhaItem item = p.items[it0];
            if (item.slot >= 0 && (item.piece & 0xff) != 0) {
                const bool two = item.q_row0 + BQ < item.seg_end;
                mbar_wait(part_ready(0), 0, 0x72);
                if (two) mbar_wait(part_ready(1), 0, 0x73);
                __threadfence();
                if (lane == 0) {
                    float* flag = p.ws + ((size_t)item.slot * p.ws_pieces + (item.piece & 0xff)) * (NT * BQ) * kWsRow + 66;
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"((uint32_t)p.gen) : "memory");
                }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax + epilogue warps
        const int t = (warp - 2) >> 2;                   // query tile of this warpgroup
        const int qd = warp & 3;                         // TMEM lane quarter of this warp
        const int r = qd * 32 + lane;                    // query row inside the tile
        const uint32_t t_lane = tmem + (uint32_t(qd * 32) << 16);
        const uint32_t tm_s = t_lane + t * TM_TILE + TM_S, tm_p = t_lane + t * TM_TILE + TM_P, tm_o = t_lane + t * TM_TILE + TM_O;
        const float sl2 = p.scale_log2;
        uint32_t cs = 0;                                 // blocks of this tile consumed so far (all items)
        uint32_t sreg[4][32];
        // optional timeline of CTA 5 (tools/fmha_cta_times.py): per item [8] = {start, block 0 done, last block done, last P.V retired,
        // partial rows published / combined, output written, key blocks, split?}
        const bool stamp = p.dbg && blockIdx.x == 5 && warp == 2 && lane == 0;
#define PSTAMP(k) do { if (stamp) p.dbg[8192 + (size_t)(it - it0) * 8 + (k)] = clock64(); } while (0)
        for (int it = it0; it < it1; ++it) {
            const FmhaItem item = p.items[it];
            const int head = item.head, q0 = item.q_row0, seg_end = item.seg_end;
            const int kv_begin = item.kv_lo, kv_end = item.kv_hi;
            const int n_kv = (kv_end - kv_begin + BKV - 1) / BKV;
            const bool two = q0 + BQ < seg_end;
            if (t == 1 && !two) continue;                // this item has no second tile: its barriers do not move
            float m_run = -INFINITY, l_run = 0.f;        // m_run: reference exponent (log2 domain), only raised lazily
            PSTAMP(0);
            if (stamp) { p.dbg[8192 + (size_t)(it - it0) * 8 + 6] = n_kv; p.dbg[8192 + (size_t)(it - it0) * 8 + 7] = item.slot >= 0 ? (item.piece & 0xff) + 1 : 0; }
            for (int j = 0; j < n_kv; ++j, ++cs) {
                mbar_wait(s_full(t), cs & 1, 0x61);
                tc_fence_after();
                const int valid = kv_end - (kv_begin + j * BKV);     // columns >= valid belong to another segment
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld32(tm_s + c * 32, sreg[c]);
                tmem_ld_wait();
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
                const bool pv_ready = j == 0 || mbar_try_wait(pv_done(t), (cs - 1) & 1);   // polled here, its latency hides under the exps
                f32x2 ls[4] = {0ull, 0ull, 0ull, 0ull};                // packed pairs: FFMA2 / FADD2 halve the issue slots around the MUFUs
                const f32x2 sl2x2 = pack2(sl2, sl2), nmx2 = pack2(-m_run, -m_run);
                uint32_t preg[64];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float ua, ub;
                        unpack2(ffma2(pack2(__uint_as_float(sreg[c][i]), __uint_as_float(sreg[c][i + 1])), sl2x2, nmx2), ua, ub);
                        const float a = ex2f(ua), b = ex2f(ub);
                        ls[(i >> 1) & 3] = fadd2(ls[(i >> 1) & 3], pack2(a, b));
                        preg[c * 16 + i / 2] = pack_half2(a, b);
                    }
                }
                float lsum;
                {
                    float s0, s1;
                    unpack2(fadd2(fadd2(ls[0], ls[1]), fadd2(ls[2], ls[3])), s0, s1);
                    lsum = s0 + s1;
                }
                const bool grow = !(lsum < kRowSumLimit);            // also catches inf / NaN
                const bool slow = __any_sync(0xffffffffu, grow);
                if (!slow) {                                         // common case: the scores are consumed, let the tensor core
                    tc_fence_before();                               // overwrite S_t with the next block while P_t(j) is being stored
                    if (lane == 0) mbar_arrive(s_free(t));
                }
                if (!__all_sync(0xffffffffu, pv_ready)) mbar_wait(pv_done(t), (cs - 1) & 1, 0x62);   // PV_t(j-1) retired: P columns free, O stable
                tc_fence_after();
                if (slow) {
                    // Rare slow path (some score exceeded the reference by ~2^15): take this block's exact row max as the new
                    // reference, redo the block from the scores still in TMEM, and rescale O / l.
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
                    if (lane == 0) mbar_arrive(s_free(t));
                }
                l_run += lsum;
                {   // P row -> TMEM (fp16 pairs, 64 columns): the A operand of the P.V UMMA
                    uint32_t (&plo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&preg[0]);
                    uint32_t (&phi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&preg[32]);
                    tmem_st32(tm_p, plo);
                    tmem_st32(tm_p + 32, phi);
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(t));
                if (j == 0) PSTAMP(1);
            }
            // ---- item epilogue
            PSTAMP(2);
            mbar_wait(pv_done(t), (cs - 1) & 1, 0x63);
            tc_fence_after();
            PSTAMP(3);
            const int row = q0 + t * BQ + r;
            uint32_t o[32], o2[32];
            tmem_ld32(tm_o, o);
            tmem_ld32(tm_o + 32, o2);
            tmem_ld_wait();
            bool write_out = true;
            if (item.slot >= 0) {
                // Split unit.  Pieces are numbered in key order and the persistent schedule hands them out so that piece 0 (the head) is
                // always the LAST item of its CTA while every later piece is the FIRST (or only) item of a following CTA.  So a later
                // piece stores its un-normalised rows (O, reference exponent, row sum), hands the publication (gpu-scope fence + a flag
                // carrying this launch's generation number) to warp 10 and moves on at once; the head piece, which finishes last by
                // construction, re-references all pieces to the largest exponent and adds them up in piece order (deterministic), its own
                // row straight from TMEM.  No atomics, no CTA-wide handshake; the head's wait is bounded and in practice already
                // satisfied.  Deadlock-free under any dispatch order: a waited-for piece depends on nothing.
                const int piece = item.piece & 0xff, n_pieces = item.piece >> 8;
                float* unit_ws = p.ws + (size_t)item.slot * p.ws_pieces * (NT * BQ) * kWsRow;
                if (piece != 0) {
                    float4* mine = reinterpret_cast<float4*>(unit_ws + ((size_t)piece * (NT * BQ) + t * BQ + r) * kWsRow);
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        __stcg(mine + i / 4, make_float4(__uint_as_float(o[i]), __uint_as_float(o[i + 1]), __uint_as_float(o[i + 2]), __uint_as_float(o[i + 3])));
                        __stcg(mine + 8 + i / 4, make_float4(__uint_as_float(o2[i]), __uint_as_float(o2[i + 1]), __uint_as_float(o2[i + 2]), __uint_as_float(o2[i + 3])));
                    }
                    __stcg(reinterpret_cast<float2*>(mine + 16), make_float2(m_run, l_run));
                    mbar_arrive(part_ready(t));                                    // warp 10 fences and raises the piece's flag
                    write_out = false;
                } else {
                    const float* rows = unit_ws + (size_t)(t * BQ + r) * kWsRow;
                    const size_t piece_stride = (size_t)(NT * BQ) * kWsRow;
                    const long long t_spin = clock64();
                    for (int q = 1; q < n_pieces; ++q) {                          // bounded wait for the row flags of the later pieces
                        const uint32_t* flag = reinterpret_cast<const uint32_t*>(unit_ws + q * piece_stride + 66);
                        uint32_t seen;
                        uint32_t spins = 0;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");   // pairs with warp 10's fence + store
                            if (seen != (uint32_t)p.gen && (++spins & 0x3ff) == 0 && clock64() - t_spin > 8000000000ll) {
                                atomicCAS(&g_dev_error, 0u, 0x71u | (blockIdx.x << 8));
                                __threadfence_system();
                                __trap();
                            }
                        } while (seen != (uint32_t)p.gen);
                    }
                    float m_all = m_run;
                    for (int q = 1; q < n_pieces; ++q) m_all = fmaxf(m_all, __ldcg(rows + q * piece_stride + 64));
                    {
                        const float w0 = ex2f(m_run - m_all);                      // own row first (piece order)
                        l_run *= w0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) { o[i] = __float_as_uint(__uint_as_float(o[i]) * w0); o2[i] = __float_as_uint(__uint_as_float(o2[i]) * w0); }
                    }
                    for (int q = 1; q < n_pieces; ++q) {
                        const float4* src = reinterpret_cast<const float4*>(rows + q * piece_stride);
                        const float2 ml = __ldcg(reinterpret_cast<const float2*>(src + 16));
                        const float wq = ex2f(ml.x - m_all);
                        l_run = fmaf(ml.y, wq, l_run);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 a = __ldcg(src + i / 4), b = __ldcg(src + 8 + i / 4);
                            o[i] = __float_as_uint(fmaf(a.x, wq, __uint_as_float(o[i])));
                            o[i + 1] = __float_as_uint(fmaf(a.y, wq, __uint_as_float(o[i + 1])));
                            o[i + 2] = __float_as_uint(fmaf(a.z, wq, __uint_as_float(o[i + 2])));
                            o[i + 3] = __float_as_uint(fmaf(a.w, wq, __uint_as_float(o[i + 3])));
                            o2[i] = __float_as_uint(fmaf(b.x, wq, __uint_as_float(o2[i])));
                            o2[i + 1] = __float_as_uint(fmaf(b.y, wq, __uint_as_float(o2[i + 1])));
                            o2[i + 2] = __float_as_uint(fmaf(b.z, wq, __uint_as_float(o2[i + 2])));
                            o2[i + 3] = __float_as_uint(fmaf(b.w, wq, __uint_as_float(o2[i + 3])));
                        }
                    }
                    // Consumed: clear the flags (this CTA is their only reader), so a workspace never carries a flag into a later launch -- the
                    // generation number is then a second line of defence only (e.g. replays of a captured launch reuse their numbers).
                    asm volatile("bar.sync 2, %0;" ::"r"(two ? 256 : 128) : "memory");
                    if (warp == 2 && lane == 0)
                        for (int q = 1; q < n_pieces; ++q) *reinterpret_cast<volatile uint32_t*>(unit_ws + q * piece_stride + 66) = 0u;
                }
            }
            PSTAMP(4);
            const float inv_l = 1.0f / l_run;
            if (write_out && row < seg_end) {
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
            PSTAMP(5);
        }
#undef PSTAMP
    }
    if (p.dbg && threadIdx.x == 64) {            // softmax warp 2, lane 0: [cta][4] = {start ns, end ns, cycles, SM id} (tools/fmha_cta_times.py)
        long long t1; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.dbg[blockIdx.x * 4 + 0] = dbg_t0; p.dbg[blockIdx.x * 4 + 1] = t1; p.dbg[blockIdx.x * 4 + 2] = clock64() - dbg_c0; p.dbg[blockIdx.x * 4 + 3] = smid;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace

template <int NT>
static int launch_t(const CUtensorMap& tm_qkv, const FmhaArgs& a, int n_items, cudaStream_t st) {
    using F = FCfg<NT>;
    static bool configured = false;
    if (!configured) {
        RAP_CUDA(cudaFuncSetAttribute(fmha_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES));
        configured = true;
    }
    ProfScope _ps(KC_FMHA, st);
    RAP_CUDA(launch_pdl(fmha_kernel<NT>, dim3(n_items), dim3(F::kThreads + ((NT == 2 && a.mma2) ? 32 : 0)), F::SMEM_BYTES, st, tm_qkv, a));
    note_launch();
    return RAP_OK;
}

int g_fmha_rows_override = 0;                         // micro entry points / tests: force 128 / 256 / 384 query rows per CTA
int fmha_rows_per_cta() {
    static int rows = 0;
    if (!rows) { const char* e = getenv("RAP_FMHA_NT"); const int nt = e ? atoi(e) : 2; rows = nt == 1 ? 128 : nt == 3 ? 384 : 256; }
    return g_fmha_rows_override > 0 ? g_fmha_rows_override : rows;
}

size_t FmhaSchedule::workspace_floats() const { return (size_t)n_slots * max_pieces * fmha_rows_per_cta() * kWsRow; }

// Every (query block, head) unit costs its number of 128-key blocks.  Units are issued longest first; the hardware hands the next
// CTA to whichever SM frees up, so whole waves balance by themselves and only the last, partially filled wave leaves SMs idle
// (cfg2: 256 units on 148 SMs = 13.5 % of the kernel).  The R units of that wave are therefore cut along the key axis:
//   * R <= SMs/2: each into k = SMs/R equal pieces;
//   * otherwise: each into a big piece (its own SM) and a small one, q = ceil(R / idle SMs) small pieces per idle SM, sized 1/(q+1).
FmhaSchedule build_fmha_schedule(const std::vector<int>& cu, int heads, bool force_split) {
    static int split = -1;
    if (split < 0) { const char* e = getenv("RAP_FMHA_SPLIT"); split = (e && atoi(e) == 0) ? 0 : 1; }
    const int rows = fmha_rows_per_cta();
    const int slots_per_wave = num_sms() * (rows == 128 ? 2 : 1);
    constexpr int kMaxPieces = 8;
    struct Unit { int q, s0, s1, head, blocks; };
    std::vector<Unit> units;
    for (int h = 0; h < heads; ++h)
        for (size_t s = 0; s + 1 < cu.size(); ++s)
            for (int q = cu[s]; q < cu[s + 1]; q += rows) units.push_back({q, cu[s], cu[s + 1], h, (cu[s + 1] - cu[s] + BKV - 1) / BKV});
    std::stable_sort(units.begin(), units.end(), [](const Unit& a, const Unit& b) { return a.blocks > b.blocks; });
    FmhaSchedule sch;
    const int U = (int)units.size();
    // A piece costs about kPieceOverhead key blocks on top of its own (CTA start-up, pipeline fill, partial-row round trip through L2;
    // measured with tools/fmha_split_sweep.py: ~7 us against 1.55 us per block).  Splitting is skipped when it would not pay.
    constexpr double kPieceOverhead = 4.5;
    int R = (split || force_split) ? U % slots_per_wave : 0;
    int k_equal = 0, q_small = 0;
    static double frac_env = -1.0;                        // RAP_FMHA_SMALL_FRAC: override of the small piece's share (tuning aid)
    if (frac_env < 0.0) { const char* e = getenv("RAP_FMHA_SMALL_FRAC"); frac_env = e ? atof(e) : 0.0; }
    double small_frac = 0.0;
    if (R > 0) {
        const double o_rel = kPieceOverhead / std::max(1, units[U - 1].blocks);
        if (2 * R <= slots_per_wave) {
            k_equal = std::min(slots_per_wave / R, kMaxPieces);
            if (!force_split && 1.0 / k_equal + o_rel > 0.92) R = 0;
        } else {
            q_small = (R + (slots_per_wave - R) - 1) / (slots_per_wave - R);
            small_frac = frac_env > 0.0 ? frac_env : (1.0 - 2.0 * o_rel) / (q_small + 1);      // q (x + o) = 1 - x + o
            if (force_split) small_frac = std::max(small_frac, 1.0 / (q_small + 1));
            else if (frac_env <= 0.0 && small_frac - o_rel < 0.08) R = 0;                           // the last wave would shrink by < 8 %
        }
    }
    auto whole = [&](const Unit& u) { sch.items.push_back({u.q, u.s0, u.s1, u.head, u.s0, u.s1, -1, 1 << 8}); };
    for (int i = 0; i < U - R; ++i) whole(units[i]);
    std::vector<FmhaItem> later;
    for (int i = U - R; i < U; ++i) {
        const Unit& u = units[i];
        std::vector<int> cuts;                            // block counts of the pieces, first = biggest
        if (k_equal >= 2) {
            const int k = std::min(k_equal, u.blocks);
            for (int p = 0; p < k; ++p) cuts.push_back(u.blocks / k + (p < u.blocks % k ? 1 : 0));
        } else if (q_small > 0) {
            const int small = std::max(1, (int)(u.blocks * small_frac + 0.5));
            if (u.blocks - small >= 1) { cuts.push_back(u.blocks - small); cuts.push_back(small); }
        }
        if (cuts.size() < 2) { whole(u); continue; }
        const int slot = sch.n_slots++;
        sch.max_pieces = std::max(sch.max_pieces, (int)cuts.size());
        int lo = u.s0;
        for (size_t p = 0; p < cuts.size(); ++p) {
            const int hi = std::min(u.s1, lo + cuts[p] * BKV);
            FmhaItem it{u.q, u.s0, u.s1, u.head, lo, hi, slot, (int)p | ((int)cuts.size() << 8)};
            if (p == 0) sch.items.push_back(it); else later.push_back(it);
            lo = hi;
        }
    }
    // big pieces were appended in unit order; the smaller pieces follow, piece index by piece index
    std::stable_sort(later.begin(), later.end(), [](const FmhaItem& a, const FmhaItem& b) { return (a.piece & 0xff) < (b.piece & 0xff); });
    sch.items.insert(sch.items.end(), later.begin(), later.end());
    return sch;
}

// RAP_FMHA_PERSIST: 0 / unset = one item per CTA, hardware-scheduled, last wave key-split (default); 1 = persistent stream-K kernel, whole waves
// dealt round-robin; 2 = persistent, one contiguous range per CTA.  The persistent kernel is 6-9 % faster as an isolated launch (r02: 94.6 /
// 171.8 us against 104.7 / 186.8) but NOT inside the sampler (same-box alternating A/B, ms per pair at 1 pair | ms per step at 8 pairs:
// default 98.52 / 98.13 | 760.8 / 760.6; =1: 99.97 / 100.21 | 776.9 / 776.8; =2: 98.87 / 98.93 | 834.4 / 835.9): under programmatic dependent
// launch the ragged end of a hardware-scheduled grid lets the next kernel's CTAs start (and run their prologues) while the last ones
// finish, whereas 148 persistent CTAs end together and the step is power-capped anyway.
int fmha_persist_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_FMHA_PERSIST"); v = e ? atoi(e) : 0; if (v < 0 || v > 2) v = 0; }
    return fmha_rows_per_cta() == 256 ? v : 0;
}
bool fmha_persistent_enabled() { return fmha_persist_mode() != 0; }

int g_persist_contiguous_override = -1;               // rap_bench_fmha_variant / tests: force the contiguous (1) or wave-major (0) persistent schedule

// Cost model of the persistent schedule, in units of one 128-key block of a 256-row unit (~2.5 k cycles): a unit whose second query tile
// is empty keeps only one softmax warpgroup busy; every item pays a fixed start/drain cost (first scores, row max, epilogue, Q reload).
namespace {
constexpr double kOneTileBlock = 0.55, kItemOverhead = 2.0;
constexpr int kMinPieceBlocks = 2;
}

FmhaSchedule build_fmha_schedule_persistent(const std::vector<int>& cu, int heads, int n_ctas, bool force_split) {
    const int rows = 256;
    static int env_ctas = -1;                             // RAP_FMHA_CTAS: number of persistent CTAs (measurement aid; default one per SM)
    if (env_ctas < 0) { const char* e = getenv("RAP_FMHA_CTAS"); env_ctas = e ? atoi(e) : 0; }
    int G = n_ctas > 0 ? n_ctas : (env_ctas > 0 ? env_ctas : num_sms());
    struct Unit { int q, s0, s1, head, blocks; double w; };
    std::vector<Unit> units;
    // segment-major, then head, then query block: CTAs that run side by side read the same K/V stream out of L2
    for (size_t s = 0; s + 1 < cu.size(); ++s)
        for (int h = 0; h < heads; ++h)
            for (int q = cu[s]; q < cu[s + 1]; q += rows)
                units.push_back({q, cu[s], cu[s + 1], h, (cu[s + 1] - cu[s] + BKV - 1) / BKV, (q + BQ < cu[s + 1]) ? 1.0 : kOneTileBlock});
    FmhaSchedule sch;
    sch.cta_ptr.push_back(0);
    if (units.empty()) return sch;
    double work = 0.0;
    for (const Unit& u : units) work += u.blocks * u.w;
    struct Piece { int unit, b0, b1; };
    std::vector<std::vector<Piece>> per_cta, best_cta;
    std::vector<int> pieces_of_unit, best_pieces;
    double best_max = 1e300;
    const double floor_target = (force_split ? 3.0 : 6.0) + kItemOverhead;   // a CTA is not worth starting for less than a few blocks
    size_t n_items = units.size() + (size_t)G;                               // first guess: one boundary (= one extra item) per CTA
    if (!force_split && work + kItemOverhead * (double)units.size() <= floor_target * G) {
        // Small launch (fewer blocks than the SMs could take a minimum share of): no cutting, whole units, longest first onto the least
        // loaded CTA -- at most one CTA per unit.
        std::vector<int> order(units.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return units[a].blocks * units[a].w > units[b].blocks * units[b].w; });
        const int n = (int)std::min<size_t>((size_t)G, units.size());
        std::vector<double> load(n, 0.0);
        std::vector<std::vector<int>> mine(n);
        for (int ui : order) {
            const int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            mine[c].push_back(ui);
            load[c] += units[ui].blocks * units[ui].w + kItemOverhead;
        }
        for (const auto& l : mine) {
            for (int ui : l) sch.items.push_back({units[ui].q, units[ui].s0, units[ui].s1, units[ui].head, units[ui].s0, units[ui].s1, -1, 1 << 8});
            sch.cta_ptr.push_back((int)sch.items.size());
        }
        return sch;
    }
    // Whole waves first: unit k * G + c goes to CTA c, so that at any moment all CTAs work inside the same window of ~G consecutive units
    // (a handful of (segment, head) K/V streams -> L2-resident; contiguous per-CTA ranges of a large batch would spread the CTAs over every
    // (scene, head) stream at once -- 128 MB of K/V at 8 pairs -- and thrash L2).  Only the LAST, partial wave is cut stream-K style, each
    // CTA receiving what it is short of the common target.
    const size_t U = units.size();
    const bool contiguous = g_persist_contiguous_override >= 0 ? g_persist_contiguous_override == 1 : fmha_persist_mode() == 2;
    const size_t n_whole = contiguous ? 0 : (U / (size_t)G) * (size_t)G;
    std::vector<double> load((size_t)G, 0.0);
    for (size_t ui = 0; ui < n_whole; ++ui) load[ui % (size_t)G] += units[ui].blocks * units[ui].w + kItemOverhead;
    double rem_work = 0.0;
    for (size_t ui = n_whole; ui < U; ++ui) rem_work += units[ui].blocks * units[ui].w;
    double load_sum = 0.0;
    for (double l : load) load_sum += l;
    n_items = (U - n_whole) + (size_t)G;
    for (int pass = 0; pass < 8; ++pass) {                                   // the item count feeds back into the target: keep the best pass
        // capacity of CTA c = target - load[c] (what it is short of); CTA c's share of the remainder ends where the running cost reaches the
        // running capacity: rounding to whole blocks does not accumulate towards the last CTA
        double target = std::max((load_sum + rem_work + kItemOverhead * (double)n_items) / G, floor_target);
        for (int it = 0; it < 4; ++it) {                                     // CTAs already above the target take nothing: re-level the others
            double cap = 0.0;
            int open = 0;
            for (double l : load) if (l < target) { cap += target - l; ++open; }
            const double need = rem_work + kItemOverhead * (double)n_items;
            if (open == 0 || std::abs(cap - need) < 1e-6 * (need + 1.0)) break;
            target += (need - cap) / open;
        }
        std::vector<double> cum((size_t)G, 0.0);
        double run = 0.0;
        for (int c = 0; c < G; ++c) { run += std::max(0.0, target - load[c]); cum[c] = run; }
        per_cta.assign((size_t)G, {});
        pieces_of_unit.assign(units.size(), 0);
        double assigned = 0.0;
        int c = 0;
        for (size_t ui = n_whole; ui < U; ++ui) {
            const Unit& u = units[ui];
            int b = 0;
            while (b < u.blocks) {
                const bool last_cta = c >= G - 1;
                const double budget = last_cta ? 1e30 : cum[c] - assigned;
                const int left = u.blocks - b;
                int take = left;
                if (left * u.w + kItemOverhead > budget + 0.5 * u.w) {
                    take = (int)((budget - kItemOverhead) / u.w + 0.5);
                    if (take > left - kMinPieceBlocks) take = left - kMinPieceBlocks;   // never leave a sliver behind ...
                    if (take < kMinPieceBlocks) take = 0;                               // ... nor start one
                    if (take == 0 && per_cta[c].empty() && n_whole == 0) take = left <= 2 * kMinPieceBlocks ? left : kMinPieceBlocks;
                }
                if (take > 0) {
                    per_cta[c].push_back({(int)ui, b, b + take});
                    ++pieces_of_unit[ui];
                    assigned += take * u.w + kItemOverhead;
                    b += take;
                }
                if (!last_cta && (b < u.blocks || cum[c] - assigned < 0.5 * kItemOverhead + 0.5)) ++c;
            }
        }
        size_t items = 0;
        double worst = 0.0;
        for (int k = 0; k < G; ++k) {
            items += per_cta[k].size();
            double cost = load[k];
            for (const Piece& pc : per_cta[k]) cost += (pc.b1 - pc.b0) * units[pc.unit].w + kItemOverhead;
            worst = std::max(worst, cost);
        }
        if (worst < best_max) { best_max = worst; best_cta = per_cta; best_pieces = pieces_of_unit; }
        if (items == n_items) break;
        n_items = (items + n_items + 1) / 2;                                 // damped: plain substitution can oscillate
    }
    per_cta.swap(best_cta);
    pieces_of_unit.swap(best_pieces);
    // Per-CTA order: [later piece (index >= 1) of a cut unit] [whole units, wave by wave, then whole units of the remainder] [head piece (index 0)
    // of a cut unit].  The kernel's combine protocol needs exactly this: a waited-for piece is the first thing its CTA does, the waiting
    // head piece the last.
    for (int k = 0; k < G; ++k) {
        std::vector<Piece> first, mid, last;
        for (const Piece& pc : per_cta[k]) {
            const bool cut = pieces_of_unit[pc.unit] > 1;
            (cut && pc.b0 > 0 ? first : cut ? last : mid).push_back(pc);
        }
        std::vector<Piece> order(first);
        for (size_t ui = (size_t)k; ui < n_whole; ui += (size_t)G) { order.push_back({(int)ui, 0, units[ui].blocks}); pieces_of_unit[ui] = 1; }
        order.insert(order.end(), mid.begin(), mid.end());
        order.insert(order.end(), last.begin(), last.end());
        per_cta[k].swap(order);
    }
    while (!per_cta.empty() && per_cta.back().empty()) per_cta.pop_back();
    for (auto it = per_cta.begin(); it != per_cta.end();) it = it->empty() ? per_cta.erase(it) : it + 1;
    std::vector<int> slot_of_unit(units.size(), -1), next_piece(units.size(), 0);
    for (size_t ui = 0; ui < units.size(); ++ui)
        if (pieces_of_unit[ui] > 1) { slot_of_unit[ui] = sch.n_slots++; sch.max_pieces = std::max(sch.max_pieces, pieces_of_unit[ui]); }
    // piece index = rank of the piece's first key block inside its unit (key order), whatever order the CTAs run them in
    std::vector<std::vector<int>> starts(units.size());
    for (const auto& list : per_cta)
        for (const Piece& pc : list) if (pieces_of_unit[pc.unit] > 1) starts[pc.unit].push_back(pc.b0);
    for (auto& v : starts) std::sort(v.begin(), v.end());
    (void)next_piece;
    for (const auto& list : per_cta) {
        for (const Piece& pc : list) {
            const Unit& u = units[pc.unit];
            const int lo = u.s0 + pc.b0 * BKV, hi = std::min(u.s1, u.s0 + pc.b1 * BKV);
            const int idx = pieces_of_unit[pc.unit] > 1 ? (int)(std::lower_bound(starts[pc.unit].begin(), starts[pc.unit].end(), pc.b0) - starts[pc.unit].begin()) : 0;
            sch.items.push_back({u.q, u.s0, u.s1, u.head, lo, hi, slot_of_unit[pc.unit], idx | (pieces_of_unit[pc.unit] << 8)});
        }
        sch.cta_ptr.push_back((int)sch.items.size());
    }
    return sch;
}

FmhaSchedule build_fmha_schedule_auto(const std::vector<int>& cu, int heads, bool force_split) {
    return fmha_persistent_enabled() ? build_fmha_schedule_persistent(cu, heads, 0, force_split) : build_fmha_schedule(cu, heads, force_split);
}

int launch_fmha(const CUtensorMap& tm_qkv, const FmhaArgs& a, int n_ctas, cudaStream_t st, const CUtensorMap* tm_kv64) {
    RAP_REQUIRE(n_ctas > 0 && a.items && a.d_model % HD == 0, "fmha: bad args (ctas=%d d=%d)", n_ctas, a.d_model);
    if (!a.cta_ptr && fmha_rows_per_cta() == 384) {        // three query tiles, 64-key blocks
        RAP_REQUIRE(tm_kv64, "fmha: the 3-tile kernel needs the 64-row K/V tensor map");
        static bool configured3 = false;
        if (!configured3) {
            RAP_CUDA(cudaFuncSetAttribute(fmha3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::SMEM_BYTES));
            configured3 = true;
        }
        ProfScope _ps(KC_FMHA, st);
        static int mma3 = -1;                            // RAP_FMHA_MMA3=0: one UMMA warp walks the three tiles in a fixed order
        if (mma3 < 0) { const char* e = getenv("RAP_FMHA_MMA3"); mma3 = (e && atoi(e) == 0) ? 0 : 1; }
        RAP_CUDA(launch_pdl(fmha3_kernel, dim3(n_ctas), dim3(k3::kThreads + (mma3 ? 64 : 0)), k3::SMEM_BYTES, st, tm_qkv, *tm_kv64, a));
        note_launch();
        return RAP_OK;
    }
    if (a.cta_ptr) {
        static bool configured = false;
        if (!configured) {
            RAP_CUDA(cudaFuncSetAttribute(fmha_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FCfg<2>::SMEM_BYTES));
            configured = true;
        }
        ProfScope _ps(KC_FMHA, st);
        RAP_CUDA(launch_pdl(fmha_persist_kernel, dim3(n_ctas), dim3(kPersistThreads), FCfg<2>::SMEM_BYTES, st, tm_qkv, a));
        note_launch();
        return RAP_OK;
    }
    static int early = -1;
    if (early < 0) { const char* e = getenv("RAP_FMHA_EARLY"); early = (e && atoi(e) == 1) ? 1 : 0; }
    static int mma2 = -1;                                // RAP_FMHA_MMA2: one UMMA-issuing warp per query tile (A/B)
    if (mma2 < 0) { const char* e = getenv("RAP_FMHA_MMA2"); mma2 = e ? (atoi(e) != 0 ? 1 : 0) : 0; }
    FmhaArgs b = a;
    b.mma2 = (a.mma2 || mma2) ? 1 : 0;
    b.early_release = early && a.score_bound;
    return fmha_rows_per_cta() == 128 ? launch_t<1>(tm_qkv, b, n_ctas, st) : launch_t<2>(tm_qkv, b, n_ctas, st);
}

}  // namespace rap
