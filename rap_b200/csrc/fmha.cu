// Varlen (segment-packed) non-causal softmax attention, head_dim 64, fp16 in / fp32 accumulate, on
// tcgen05 + TMEM + TMA.  Replaces flash_attn.flash_attn_varlen_qkvpacked_func as called by the
// reference at flow_model/layer.py:106-111 (part-wise: segments = views) and :123-128 (global:
// segments = scenes).  The two calls differ only in the segment table, so this is one kernel.
//
// One CTA = one 128-row query tile of one head; two CTAs are resident per SM (112.1 KB smem,
// 256 TMEM columns each) so one CTA's UMMAs run under the other's softmax.
//   warp 0 : TMA producer  (Q once; K and V tiles through separate 2-deep rings)
//   warp 1 : UMMA issuer   (S = Q K^T into TMEM cols [0,128); O += P V into cols [128,192))
//   warps 2-5: softmax     (thread = query row: tcgen05.ld S, online max/sum, P -> smem as fp16 in
//                           the K-major SWIZZLE_128B layout UMMA reads, lazy O rescale in TMEM)
#include "fmha.cuh"
#include "sm100.cuh"

namespace rap {
namespace {

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;         // 16 KB: 128 rows x 128 bytes
constexpr int P_BYTES = BQ * BKV * 2;            // 32 KB: two 64-column K-major slabs
constexpr int KV_STAGES = 2;
constexpr int OFF_Q = 0, OFF_K = TILE_BYTES, OFF_V = OFF_K + KV_STAGES * TILE_BYTES, OFF_P = OFF_V + KV_STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int kThreads = 192;
constexpr uint32_t TMEM_COLS = 256, TM_S = 0, TM_O = 128;
static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs per SM");

__global__ void __launch_bounds__(kThreads, 2)
fmha_kernel(const __grid_constant__ CUtensorMap tm_qkv, const FmhaArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int4 w = p.tiles[blockIdx.x];            // {q_row0, seg_start, seg_end, -}
    const int q0 = w.x, kv_begin = w.y, kv_end = w.z;
    const int n_kv = (kv_end - kv_begin + BKV - 1) / BKV;

    const uint32_t bar = sbase + OFF_BAR;
    const uint32_t q_full = bar, k_full = bar + 8, k_empty = bar + 24, v_full = bar + 40, v_empty = bar + 56;
    const uint32_t s_full = bar + 72, p_full = bar + 80, pv_done = bar + 88, tmem_slot = bar + 96;

    if (threadIdx.x == 0) {
        if (sbase & 1023u) { atomicCAS(&g_dev_error, 0u, 0xA1u); __trap(); }
        tma_prefetch_desc(&tm_qkv);
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1);
            mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, 1);
        }
        mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 96);

    const int col_q = head * HD, col_k = p.d_model + head * HD, col_v = 2 * p.d_model + head * HD;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, TILE_BYTES);
            tma_load_2d(sbase + OFF_Q, &tm_qkv, q_full, col_q, q0);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1; const uint32_t ph = (j >> 1) & 1;
                const int row = kv_begin + j * BKV;
                mbar_wait(k_empty + 8 * s, ph ^ 1, 0x41);
                mbar_expect_tx(k_full + 8 * s, TILE_BYTES);
                tma_load_2d(sbase + OFF_K + s * TILE_BYTES, &tm_qkv, k_full + 8 * s, col_k, row);
                mbar_wait(v_empty + 8 * s, ph ^ 1, 0x42);
                mbar_expect_tx(v_full + 8 * s, TILE_BYTES);
                tma_load_2d(sbase + OFF_V + s * TILE_BYTES, &tm_qkv, v_full + 8 * s, col_v, row);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV, false, false);   // A=Q K-major, B=K K-major, N=128
            constexpr uint32_t idesc_pv = umma_idesc_f16(BQ, HD, false, true);     // A=P K-major, B=V MN-major, N=64
            const uint64_t dq = umma_desc_sw128(sbase + OFF_Q, 1024);
            auto issue_qk = [&](int j) {
                const int s = j & 1;
                mbar_wait(k_full + 8 * s, (j >> 1) & 1, 0x51);
                tc_fence_after();
                const uint64_t dk = umma_desc_sw128(sbase + OFF_K + s * TILE_BYTES, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem + TM_S, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
                umma_commit(s_full);
                umma_commit(k_empty + 8 * s);
            };
            mbar_wait(q_full, 0, 0x50);
            issue_qk(0);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                mbar_wait(p_full, j & 1, 0x52);          // P_j in smem, S_j consumed, O rescaled
                tc_fence_after();
                if (j + 1 < n_kv) issue_qk(j + 1);       // next scores first: softmax j+1 can start while PV_j runs
                mbar_wait(v_full + 8 * s, (j >> 1) & 1, 0x53);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {
                    const uint64_t dp = umma_desc_sw128(sbase + OFF_P + (k >> 2) * (P_BYTES / 2) + (k & 3) * 32, 1024);
                    const uint64_t dv = umma_desc_sw128(sbase + OFF_V + s * TILE_BYTES + k * 2048, 1024);
                    umma_f16_ss(tmem + TM_O, dp, dv, idesc_pv, (j | k) != 0);
                }
                umma_commit(pv_done);
                umma_commit(v_empty + 8 * s);
            }
        }
    } else {
        const int qd = warp & 3;                         // TMEM lane quarter of this warp
        const int r = qd * 32 + lane;                    // query row inside the tile
        const uint32_t t_lane = tmem + (uint32_t(qd * 32) << 16);
        const float sl2 = p.scale_log2;
        float m_run = -INFINITY, l_run = 0.f;
        uint32_t sreg[4][32];
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, j & 1, 0x61);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(t_lane + TM_S + c * 32, sreg[c]);
            tmem_ld_wait();
            const int valid = kv_end - (kv_begin + j * BKV);     // columns >= valid belong to another segment
            if (valid < BKV) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i >= valid) sreg[c][i] = 0xff800000u;   // -inf
            }
            // Row max / row sum with 8 independent accumulators: a single dependent FMNMX/FADD chain over 128 columns
            // is latency-bound with only two softmax warps per scheduler.
            float mxa[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) mxa[a] = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(sreg[c][i]));
            const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])), fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
            const float m_new = fmaxf(m_run, mx * sl2);
            const float alpha = ex2f(m_run - m_new);
            const bool grew = m_new > m_run;
            m_run = m_new;
            float ls[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            uint32_t preg[64];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float a = ex2f(fmaf(__uint_as_float(sreg[c][i]), sl2, -m_new));
                    const float b = ex2f(fmaf(__uint_as_float(sreg[c][i + 1]), sl2, -m_new));
                    ls[(i >> 1) & 7] += a + b;
                    preg[c * 16 + i / 2] = pack_half2(a, b);
                }
            const float lsum = ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
            l_run = l_run * alpha + lsum;

            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1, 0x62);           // PV_{j-1} retired: P buffer free, O readable
                tc_fence_after();
                if (__any_sync(0xffffffffu, grew)) {
                    uint32_t o[32];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        tmem_ld32(t_lane + TM_O + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(t_lane + TM_O + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            // P row -> smem, K-major SWIZZLE_128B: 16-byte chunk index XOR (row & 7), two 64-column slabs.
            const uint32_t prow = sbase + OFF_P + r * 128;
#pragma unroll
            for (int ch = 0; ch < 16; ++ch) {
                const uint32_t addr = prow + (ch >> 3) * (P_BYTES / 2) + (((ch & 7) ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(preg[4 * ch]), "r"(preg[4 * ch + 1]),
                             "r"(preg[4 * ch + 2]), "r"(preg[4 * ch + 3]) : "memory");
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // epilogue: O / l -> fp16
        mbar_wait(pv_done, (n_kv - 1) & 1, 0x63);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        const int row = q0 + r;
        uint32_t o[32], o2[32];
        tmem_ld32(t_lane + TM_O, o);
        tmem_ld32(t_lane + TM_O + 32, o2);
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
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace

int launch_fmha(const CUtensorMap& tm_qkv, const FmhaArgs& a, int n_tiles, int heads, cudaStream_t st) {
    RAP_REQUIRE(n_tiles > 0 && heads > 0 && a.d_model == heads * HD, "fmha: bad args (tiles=%d heads=%d d=%d)", n_tiles, heads, a.d_model);
    static bool configured = false;
    if (!configured) {
        RAP_CUDA(cudaFuncSetAttribute(fmha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    ProfScope _ps(KC_FMHA, st);
    fmha_kernel<<<dim3(n_tiles, heads), kThreads, SMEM_BYTES, st>>>(tm_qkv, a);
    RAP_CUDA(cudaGetLastError());
    note_launch();
    return RAP_OK;
}

}  // namespace rap
