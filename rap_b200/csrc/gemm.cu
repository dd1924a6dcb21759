// tcgen05 / TMEM / TMA GEMM with fused DiT epilogues (see gemm.cuh).
//
// Reference call sites this kernel replaces (cuBLAS/cuBLASLt behind nn.Linear + eager elementwise):
//   flow_model/layer.py:100,117 (QKV proj) + norm.py:33 (qk RMS norm)        -> EPI_QKV
//   flow_model/layer.py:113,131 (out proj) + layer.py:155,160 (residual add)  -> EPI_F32 (addend = out = h)
//   diffusers FeedForward/GEGLU used at layer.py:89,164                       -> EPI_GEGLU then EPI_F32
//   flow_model/embedding.py:179 (emb_proj) + point_cloud_dit.py:119-139       -> EPI_F32 (table = anchor emb)
//   flow_model/point_cloud_dit.py:111-117 (final_mlp hidden layers)           -> EPI_F16_SILU
#include "gemm.cuh"
#include "sm100.cuh"

namespace rap {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 fp16 = one 128-byte swizzle row
constexpr int kThreads = 320;        // TMA warp + MMA warp + 8 epilogue warps
constexpr int A_STAGE_BYTES = BM * BK * 2;

template <int BN> struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024 /*align slack*/ + 4096 /*gamma*/;
};

// Exact (erf) GELU.  1+erf(z) through Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the fp16 output grid):
//   z>=0: 2 - poly(t) e^{-z^2},  z<0: poly(t) e^{-z^2},  t = 1/(1 + 0.3275911 |z|)   -- one rcp, one ex2, 7 FMAs.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = x * 0.70710678118654752f, az = fabsf(z);
    const float t = __frcp_rn(fmaf(0.3275911f, az, 1.0f));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    const float e = pl * t * ex2f(-az * az * 1.4426950408889634f);
    return 0.5f * x * (z >= 0.f ? 2.0f - e : e);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));
    float* gamma_s = reinterpret_cast<float*>(smem_gen + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = p.N / BN, k_blocks = p.K / BK;
    const int num_tiles = m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    if (EPI == EPI_QKV) {
        for (int i = threadIdx.x; i < 2 * 8 * 64; i += kThreads) gamma_s[i] = p.gamma[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (warp converged, one elected lane issues)
        const bool leader = elect_one();
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1, 0x11);
                if (leader) {
                    mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m0);
                    tma_load_2d(sa + A_STAGE_BYTES, &tmB, full_bar(stage), kb * BK, n0);
                }
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ UMMA issuer (warp converged: descriptors stay warp-uniform)
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_f16(BM, BN, false, false);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1, 0x21);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase, 0x22);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                const uint64_t da = umma_desc_sw128(sa, 1024), db = umma_desc_sw128(sa + A_STAGE_BYTES, 1024);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)   // +32 bytes (encoded +2) per K=16 step inside the swizzle row
                        umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(empty_bar(stage));       // frees the smem stage when these MMAs retire
                }
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (leader) umma_commit(tfull_bar(acc));     // accumulator ready for the epilogue
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ------------------------------------------------------------ epilogue: 8 warps.  TMEM lane quarter = warp % 4
        // (hardware rule), column half = (warp - 2) / 4, so two warps share a quarter and split the tile's columns.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int HC = BN / 2;                       // columns per warp
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
            uint32_t r[32];

            if constexpr (EPI == EPI_F32) {
                // Prefetch the addend (residual / static embedding) row segment into registers BEFORE waiting for the
                // accumulator: the loads fly under this tile's MMAs instead of serialising with the TMEM reads.
                const int col0 = n0 + half * HC;
                float4 ad[HC / 4];
                if (p.addend && row_ok) {
                    const float4* a4 = reinterpret_cast<const float4*>(p.addend + (size_t)row * p.ld32 + col0);
#pragma unroll
                    for (int i = 0; i < HC / 4; ++i) ad[i] = a4[i];
                } else {
#pragma unroll
                    for (int i = 0; i < HC / 4; ++i) ad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const float* tab = (p.table && row_ok) ? p.table + (size_t)(p.sel[row] ? 1 : 0) * p.N : nullptr;
                mbar_wait(tfull_bar(acc), acc_phase, 0x31);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < HC / 32; ++c) {
                    tmem_ld32(t_row + half * HC + c * 32, r);
                    tmem_ld_wait();
                    if (row_ok) {
                        const int col = col0 + c * 32;
                        float* o = p.out32 + (size_t)row * p.ld32 + col;
                        uint32_t h16[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 a = ad[c * 8 + i / 4];
                            float4 v = make_float4(__uint_as_float(r[i]) + a.x, __uint_as_float(r[i + 1]) + a.y, __uint_as_float(r[i + 2]) + a.z,
                                                   __uint_as_float(r[i + 3]) + a.w);
                            if (p.bias) { float4 b = *reinterpret_cast<const float4*>(p.bias + col + i); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
                            if (tab) { float4 t = *reinterpret_cast<const float4*>(tab + col + i); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
                            *reinterpret_cast<float4*>(o + i) = v;
                            h16[i / 2] = pack_half2(v.x, v.y);
                            h16[i / 2 + 1] = pack_half2(v.z, v.w);
                        }
                        if (p.out16) {
                            uint4* o16 = reinterpret_cast<uint4*>(p.out16 + (size_t)row * p.ld16 + col);
#pragma unroll
                            for (int i = 0; i < 4; ++i) o16[i] = make_uint4(h16[4 * i], h16[4 * i + 1], h16[4 * i + 2], h16[4 * i + 3]);
                        }
                    }
                }
            } else if constexpr (EPI == EPI_F16_SILU) {
                mbar_wait(tfull_bar(acc), acc_phase, 0x31);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < HC / 32; ++c) {
                    tmem_ld32(t_row + half * HC + c * 32, r);
                    tmem_ld_wait();
                    if (row_ok) {
                        const int col = n0 + half * HC + c * 32;
                        uint4* o16 = reinterpret_cast<uint4*>(p.out16 + (size_t)row * p.ld16 + col);
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = silu(__uint_as_float(r[i + j]) + p.bias[col + i + j]);
                            o16[i / 8] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                        }
                    }
                }
            } else if constexpr (EPI == EPI_QKV) {
                // One 64-column head per iteration.  section 0/1 = q/k (RMS-normalised), 2 = v.
                const int dmodel = p.N / 3;
                uint32_t r2[32];
                mbar_wait(tfull_bar(acc), acc_phase, 0x31);
                tc_fence_after();
#pragma unroll 1
                for (int hc = half * (HC / 64); hc < (half + 1) * (HC / 64); ++hc) {
                    tmem_ld32(t_row + hc * 64, r);
                    tmem_ld32(t_row + hc * 64 + 32, r2);
                    tmem_ld_wait();
                    const int col = n0 + hc * 64;
                    const int section = col / dmodel;
                    const int head = (col % dmodel) / 64;
                    float inv = 1.0f;
                    const float* g = gamma_s;
                    if (section < 2) {
                        float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int i = 0; i < 32; ++i) { float a = __uint_as_float(r[i]), b = __uint_as_float(r2[i]); ss[i & 3] += a * a + b * b; }
                        inv = 8.0f / fmaxf(sqrtf((ss[0] + ss[1]) + (ss[2] + ss[3])), 1e-12f);     // F.normalize eps, * sqrt(head_dim)
                        g = gamma_s + (section * 8 + head) * 64;
                    }
                    if (row_ok) {
                        uint4* o16 = reinterpret_cast<uint4*>(p.out16 + (size_t)row * p.ld16 + col);
#pragma unroll
                        for (int i = 0; i < 64; i += 8) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int e = i + j;
                                const float x = __uint_as_float(e < 32 ? r[e] : r2[e - 32]);
                                v[j] = (section < 2) ? x * inv * g[e] : x;
                            }
                            o16[i / 8] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                        }
                    }
                }
            } else {  // EPI_GEGLU
                uint32_t r2[32];
                const int out_col0 = (n0 / BN) * (BN / 2);
                mbar_wait(tfull_bar(acc), acc_phase, 0x31);
                tc_fence_after();
#pragma unroll 1
                for (int c = half * (BN / 128); c < (half + 1) * (BN / 128); ++c) {
                    tmem_ld32(t_row + c * 32, r);              // h part
                    tmem_ld32(t_row + BN / 2 + c * 32, r2);    // g part
                    tmem_ld_wait();
                    if (row_ok) {
                        const float* bh = p.bias + n0 + c * 32;
                        const float* bg = p.bias + n0 + BN / 2 + c * 32;
                        uint4* o16 = reinterpret_cast<uint4*>(p.out16 + (size_t)row * p.ld16 + out_col0 + c * 32);
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                v[j] = (__uint_as_float(r[i + j]) + bh[i + j]) * gelu_erf(__uint_as_float(r2[i + j]) + bg[i + j]);
                            o16[i / 8] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN, int EPI>
int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        RAP_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    const int tiles = ceil_div(a.M, BM) * (a.N / BN);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    ProfScope _ps(EPI == EPI_F32 ? KC_GEMM_F32 : EPI == EPI_F16_SILU ? KC_GEMM_SILU : EPI == EPI_QKV ? KC_GEMM_QKV : KC_GEMM_GEGLU, st);
    gemm_kernel<BN, EPI><<<grid, kThreads, C::SMEM_BYTES, st>>>(tmA, tmB, a);
    RAP_CUDA(cudaGetLastError());
    note_launch();
    return RAP_OK;
}

}  // namespace

int launch_gemm(GemmEpi epi, int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
    RAP_REQUIRE(a.M > 0 && a.N % BN == 0 && a.K % BK == 0 && a.K >= BK, "gemm: bad shape M=%d N=%d K=%d BN=%d", a.M, a.N, a.K, BN);
    if (BN == 256) {
        switch (epi) {
            case EPI_F32: return launch_t<256, EPI_F32>(tmA, tmB, a, st);
            case EPI_QKV: return launch_t<256, EPI_QKV>(tmA, tmB, a, st);
            case EPI_GEGLU: return launch_t<256, EPI_GEGLU>(tmA, tmB, a, st);
            default: break;
        }
    } else if (BN == 128) {
        switch (epi) {
            case EPI_F32: return launch_t<128, EPI_F32>(tmA, tmB, a, st);
            case EPI_F16_SILU: return launch_t<128, EPI_F16_SILU>(tmA, tmB, a, st);
            default: break;
        }
    }
    return fail(RAP_EINVAL, "gemm: unsupported (epi=%d, BN=%d)", (int)epi, BN);
}

}  // namespace rap
