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
// Epilogue staging: each epilogue warp owns 32 rows x 128 B (+16 B pad: conflict-free 16-byte accesses in both directions).
// Accumulator rows leave TMEM one row per thread; going through this buffer turns every global access into 4 rows x 128
// contiguous bytes per warp instruction instead of 32 rows x 16 B (which costs 32 LSU cycles per instruction).
constexpr int STG_ROW_BYTES = 144;
constexpr int STG_WARP_BYTES = 32 * STG_ROW_BYTES;

// TWO (2-CTA UMMA): a CTA holds only half of the 256-row B tile per stage (32 KB stages), so five of them fit where three 48 KB ones did:
// 160 KB instead of 144 KB in flight for 2/3 of the bytes per K step (the K loop is bound by bytes in flight / TMA latency: 3 x 48 KB over
// ~2 k cycles = 72 B/clk per SM, against 96 B/clk that 512-cycle K steps would need).
template <int BN, bool TWO = false> struct Cfg {
    static constexpr int B_STAGE_BYTES = (TWO ? BN / 2 : BN) * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = TWO ? 5 : (BN == 256) ? 3 : 5;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 256;
    static constexpr int GAMMA_BYTES = 4096;
    static constexpr int STG_OFF = STAGES * STAGE_BYTES + BAR_BYTES + GAMMA_BYTES;
    static constexpr int SMEM_BYTES = STG_OFF + 8 * STG_WARP_BYTES + 1024 /*align slack*/;
    // EPI_F32_RED: instead of the transposing staging rows, every epilogue warp owns two 32-row x 128-byte boxes in the TMA's 128-byte swizzle
    // (1024-byte aligned), double-buffered against the bulk reductions that read them.
    static constexpr int RED_BOX_BYTES = 32 * 128;
    static constexpr int RED_OFF = (STG_OFF + 1023) & ~1023;
    static constexpr int RED_SMEM_BYTES = RED_OFF + 8 * 2 * RED_BOX_BYTES + 1024;
};
// EPI_QKV / EPI_GEGLU on plain 256-wide tiles may leave through bulk tensor STORES (GemmArgs::tm_out16): the same two boxes per warp.
template <int BN, int EPI, bool TWO, bool MC = false> constexpr bool tma16_capable() { return (EPI == EPI_QKV || EPI == EPI_GEGLU) && BN == 256 && !TWO && !MC; }
template <int BN, int EPI, bool TWO, bool MC = false> constexpr int smem_bytes() {
    return (EPI == EPI_F32_RED || tma16_capable<BN, EPI, TWO, MC>()) ? Cfg<BN, TWO>::RED_SMEM_BYTES : Cfg<BN, TWO>::SMEM_BYTES;
}

// Exact (erf) GELU.  1+erf(z) through Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the fp16 output grid):
//   z>=0: 2 - poly(t) e^{-z^2},  z<0: poly(t) e^{-z^2},  t = 1/(1 + 0.3275911 |z|)   -- one rcp, one ex2, 7 FMAs.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = x * 0.70710678118654752f, az = fabsf(z);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));   // MUFU.RCP, not the IEEE software path
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    const float e = pl * t * ex2f(-az * az * 1.4426950408889634f);
    return 0.5f * x * (z >= 0.f ? 2.0f - e : e);
}
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// Two GEGLU outputs at once on packed f32x2 arithmetic: (h0, h1) * gelu((g0, g1)), exact-erf GELU through
//   erfc(a) = 2^q(a),  q = degree-7 minimax fit of log2 erfc(|g| / sqrt 2) on |g| <= 5.94 (max |err| 9.5e-6 -> 6.6e-6 relative on erfc, also in
//   the tail where 1 + erf underflows to erfc), so 1 + erf(g / sqrt 2) = E (g < 0), 2 - E (g >= 0), E = 2^q(min(|g|, 5.94)).
// Beyond |g| = 5.94, E < 3e-9: gelu is g or 0 to fp16 precision.  Against the scalar Abramowitz-Stegun form this needs one MUFU per
// output instead of two and 10 issue slots instead of 22 (8 FFMA2 + 2 FADD2 + 2 FMUL2 per PAIR): with two epilogue warps per scheduler the
// scalar version (5.2-6.0 k cycles per 128x256 tile) was longer than the tile's K loop (4.8-5.0 k), i.e. this GEMM was bound by its epilogue.
__device__ __forceinline__ uint32_t geglu2(f32x2 h, f32x2 g) {
    float g0, g1;
    unpack2(g, g0, g1);
    const f32x2 a = pack2(fminf(fabsf(g0), 5.94f), fminf(fabsf(g1), 5.94f));
    f32x2 q = ffma2(pack2(-1.833969122e-06f, -1.833969122e-06f), a, pack2(6.158401811e-05f, 6.158401811e-05f));
    q = ffma2(q, a, pack2(-9.304317646e-04f, -9.304317646e-04f));
    q = ffma2(q, a, pack2(8.507597260e-03f, 8.507597260e-03f));
    q = ffma2(q, a, pack2(-5.395960808e-02f, -5.395960808e-02f));
    q = ffma2(q, a, pack2(-4.584647417e-01f, -4.584647417e-01f));
    q = ffma2(q, a, pack2(-1.151250958e+00f, -1.151250958e+00f));
    q = ffma2(q, a, pack2(4.702056685e-06f, 4.702056685e-06f));
    float q0, q1;
    unpack2(q, q0, q1);
    const float e0 = ex2f(q0), e1 = ex2f(q1);
    // half of (1 + erf): 0.5 + copysign(0.5 - 0.5 E, g)
    float u0, u1;
    unpack2(ffma2(pack2(e0, e1), pack2(-0.5f, -0.5f), pack2(0.5f, 0.5f)), u0, u1);
    u0 = __uint_as_float(__float_as_uint(u0) | (__float_as_uint(g0) & 0x80000000u));
    u1 = __uint_as_float(__float_as_uint(u1) | (__float_as_uint(g1) & 0x80000000u));
    const f32x2 w = fadd2(pack2(u0, u1), pack2(0.5f, 0.5f));
    float o0, o1;
    unpack2(fmul2(fmul2(h, g), w), o0, o1);
    return pack_half2(o0, o1);
}

// MC (multicast): the kernel runs as 2-CTA clusters whose CTAs own vertically adjacent 128-row tiles of the SAME column tile.  Each CTA fetches
// its own A tile and HALF of the shared B tile, multicast into both CTAs' stage (cp.async.bulk.tensor ... .multicast::cluster), so the weight
// operand crosses the L2 -> SM fabric once per pair instead of once per CTA (the premise -- K loops bound by that traffic -- turned out wrong:
// the TMA delivers a 48 KB stage per 214-227 ns on all SMs, the K loops take 375-384 ns per step because of the stage round trip, DESIGN.md
// 5.2).  A stage is free again when BOTH CTAs have consumed it: the empty barriers count two arrivals and every UMMA commit arrives on
// the barrier of both CTAs (tcgen05.commit ... .multicast::cluster).
// TWO (2-CTA UMMA, `cta_group::2`): the same pair of vertically adjacent tiles runs as ONE 256x256 UMMA.  Each CTA keeps its own A tile
// and only ITS half of the B tile in shared memory (32 KB per stage instead of 48); the even CTA of the pair issues the MMAs for both, each
// CTA's accumulator rows land in its own TMEM, and the epilogues are untouched.  Loads of both CTAs are counted on the even CTA's stage barrier, stage / accumulator hand-backs are multicast commits and one
// remote arrive per epilogue warp.
template <int BN, int EPI, bool MC, bool TWO = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const GemmArgs p) {
    using C = Cfg<BN, TWO>;
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

    pdl_launch_dependents();
    // debug timeline (tools/gemm_timeline.py): per CTA, %globaltimer and clock64 at entry and exit in dbg[128 + 4 * blockIdx.x ...]
    if (p.dbg && threadIdx.x == 0) {
        long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.dbg[128 + 4 * blockIdx.x] = t; p.dbg[128 + 4 * blockIdx.x + 1] = clock64();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = p.N / BN, k_blocks = p.K / BK;
    // Work units: tiles (one CTA each), or with MC pairs of vertically adjacent tiles (one cluster each; an odd last row of tiles gets a
    // dummy partner whose loads are zero-filled and whose stores are masked).  first_tile / tile_step walk this CTA's units; unit -> (m0, n0).
    static_assert(!(MC && TWO) && (!TWO || BN == 256), "pair modes: multicast OR 2-CTA UMMA, the latter on 256-wide tiles");
    constexpr bool PAIR = MC || TWO;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const int num_tiles = PAIR ? ((m_tiles + 1) / 2) * n_tiles : m_tiles * n_tiles;
    const int first_tile = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // raster = 0: consecutive tiles walk the columns of one row block (neighbouring CTAs share the A tile); 1: they walk the row blocks of one
    // column tile (neighbours share the weight tile).  Plain tiles only.
    const bool col_major = !PAIR && p.raster == 1;
    auto tile_m0 = [&](int tile) { return (PAIR ? 2 * (tile / n_tiles) + (int)crank : col_major ? tile % m_tiles : tile / n_tiles) * BM; };
    auto tile_n0 = [&](int tile) { return (col_major ? tile / m_tiles : tile % n_tiles) * BN; };
    constexpr int B_BOXES = BN / 128 > 0 ? BN / 128 : 1;      // the B tensor map has 128-row boxes (BN = 256: two per stage)
    constexpr int B_BOX_ROWS = BN < 128 ? BN : 128;
    constexpr int B_BOX_BYTES = B_BOX_ROWS * BK * 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (EPI == EPI_F32_RED || (tma16_capable<BN, EPI, TWO, MC>() && p.tma16)) tma_prefetch_desc(&tmC);
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MC ? 2 : 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TWO ? 16 : 8); }   // TWO: both CTAs' epilogue warps
        mbar_fence_init();
    }
    if (warp == 1) { if constexpr (TWO) tmem_alloc_2cta<C::TMEM_COLS>(tmem_slot); else tmem_alloc<C::TMEM_COLS>(tmem_slot); }
    if (EPI == EPI_QKV) {
        for (int i = threadIdx.x; i < 2 * 8 * 64; i += kThreads) gamma_s[i] = p.gamma[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if constexpr (PAIR) { cluster_arrive(); cluster_wait(); }   // the partner's barriers exist before anything is multicast at them
    const uint32_t tmem_base = *tmem_slot_gen;
    // Programmatic dependent launch: only the A operand, the residual and the outputs depend on the previous kernel.  The weight
    // (B) tiles of the first pipeline stages are requested BEFORE griddepcontrol.wait, so their fetch latency hides under the
    // previous kernel's tail; the UMMA warp never touches global memory and does not wait at all.

    const bool dbg_cta = p.dbg && blockIdx.x == 0;
    int dbg_tile = 0;
#define GEMM_TS(slot) do { if (dbg_cta && lane == 0) p.dbg[(size_t)dbg_tile * 8 + (slot)] = clock64(); } while (0)
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (warp converged, one elected lane issues)
        const bool leader = elect_one();
        int stage = 0; uint32_t phase = 0;
        const int pre = (!TWO && first_tile < num_tiles) ? min(C::STAGES, k_blocks) : 0;   // stages whose B tile is in flight before the wait
        auto load_b = [&](uint32_t sb, uint32_t bar, int kcol, int n0) {            // this CTA's share of the B tile of one stage
            if constexpr (TWO) {
                tma_load_2d_2cta(sb, &tmB, bar, kcol, n0 + (int)crank * B_BOX_ROWS);
            } else if constexpr (MC) {
                tma_load_2d_mc(sb + crank * B_BOX_BYTES, &tmB, bar, kcol, n0 + (int)crank * B_BOX_ROWS, (uint16_t)3);
            } else {
#pragma unroll
                for (int h = 0; h < B_BOXES; ++h) tma_load_2d(sb + h * B_BOX_BYTES, &tmB, bar, kcol, n0 + h * B_BOX_ROWS);
            }
        };
        if (leader) {
            const int n0 = tile_n0(first_tile);
            for (int kb = 0; kb < pre; ++kb) {
                mbar_expect_tx(full_bar(kb), C::STAGE_BYTES);
                load_b(smem_base + kb * C::STAGE_BYTES + A_STAGE_BYTES, full_bar(kb), kb * BK, n0);
            }
        }
        pdl_wait();
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++dbg_tile) {
            const int m0 = tile_m0(tile), n0 = tile_n0(tile);
            const bool first = tile == first_tile;
            GEMM_TS(0);
            for (int kb = 0; kb < k_blocks; ++kb) {
                const bool b_done = first && kb < pre;          // this stage's barrier is armed and its B tile already requested
                if (!b_done) mbar_wait(empty_bar(stage), phase ^ 1, 0x11);
                if (leader) {
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                    if constexpr (TWO) {
                        // the even CTA's barrier counts the bytes of both CTAs: 2 x (A tile + half a B tile)
                        const uint32_t bar0 = dsmem_addr(full_bar(stage), 0u);
                        if (crank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                        tma_load_2d_2cta(sa, &tmA, bar0, kb * BK, m0);
                        load_b(sa + A_STAGE_BYTES, bar0, kb * BK, n0);
                    } else {
                    if (!b_done) mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    int a_col = kb * BK, a_row = m0;
                    if (p.conv_pitch) {                          // implicit 3x3 convolution: shifted box of the halo-padded activation map
                        const int tap = kb / p.conv_chunks;
                        a_col = (kb - tap * p.conv_chunks) * BK;
                        a_row = m0 + (tap / 3 - 1) * p.conv_pitch + (tap % 3 - 1);
                    }
                    tma_load_2d(sa, &tmA, full_bar(stage), a_col, a_row);
                    if (!b_done) load_b(sa + A_STAGE_BYTES, full_bar(stage), kb * BK, n0);
                    }
                }
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            GEMM_TS(1);
        }
        if constexpr (EPI == EPI_F32_LN) { __syncwarp(); cluster_arrive(); cluster_wait(); }   // every thread of the cluster takes part in the exchange barrier
    } else if (warp == 1) {
        // ------------------------------------------------------------ UMMA issuer (warp converged: descriptors stay warp-uniform)
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_f16(TWO ? 2 * BM : BM, BN, false, false);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int tile = (TWO && crank != 0) ? num_tiles : first_tile; tile < num_tiles; tile += tile_step, ++dbg_tile) {   // TWO: the even CTA issues for the pair
            GEMM_TS(2);
            mbar_wait(tempty_bar(acc), acc_phase ^ 1, 0x21);
            tc_fence_after();
            GEMM_TS(3);
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase, 0x22);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                const uint64_t da = umma_desc_sw128(sa, 1024), db = umma_desc_sw128(sa + A_STAGE_BYTES, 1024);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) { // +32 bytes (encoded +2) per K=16 step inside the swizzle row
                        if constexpr (TWO) umma_f16_ss_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        else umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    if constexpr (TWO) umma_commit_2cta(empty_bar(stage), (uint16_t)3);     // the stage is free in BOTH CTAs when the pair's MMAs retire
                    else if constexpr (MC) umma_commit_mc(empty_bar(stage), (uint16_t)3);   // both CTAs of the pair learn that this one is done with the stage
                    else umma_commit(empty_bar(stage));  // frees the smem stage when these MMAs retire
                }
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (leader) { if constexpr (TWO) umma_commit_2cta(tfull_bar(acc), (uint16_t)3); else umma_commit(tfull_bar(acc)); }   // accumulator ready for the epilogue(s)
            GEMM_TS(4);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (EPI == EPI_F32_LN) { __syncwarp(); cluster_arrive(); cluster_wait(); }
    } else {
        // ------------------------------------------------------------ epilogue: 8 warps.  TMEM lane quarter = warp % 4
        // (hardware rule), column half = (warp - 2) / 4, so two warps share a quarter and split the tile's columns.
        pdl_wait();                                      // residual reads and output writes order after the previous kernel
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int HC = BN / 2;                       // accumulator columns per warp
        const uint32_t stg = smem_base + C::STG_OFF + (warp - 2) * STG_WARP_BYTES;
        const uint32_t stg_w = stg + lane * STG_ROW_BYTES;                          // write side: thread = row
        const int rr0 = lane >> 3, cc = lane & 7;                                   // read side: 4 rows x 8 x 16 B per instruction
        const uint32_t stg_r = stg + rr0 * STG_ROW_BYTES + cc * 16;
        auto sts4 = [](uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
        };
        auto lds4 = [](uint32_t addr) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
            return v;
        };
        const bool dbg_w = warp == 2;
        uint32_t box_sel = 0;                                 // bulk-store epilogues: which of this warp's two boxes is next
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++dbg_tile) {
            const int m0 = tile_m0(tile), n0 = tile_n0(tile);
            if (dbg_w) GEMM_TS(5);
            const int row_base = m0 + q * 32;            // first of this warp's 32 rows
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
            uint32_t r[32];

            if constexpr (EPI == EPI_F32_RED) {
                // out32 += acc + bias, no residual read: thread = accumulator row writes its 32 columns of a chunk (+ bias, broadcast reads of
                // this warp's private copy) as eight 16-byte pieces into a 32 x 128 B box, piece j of row r at position j ^ (r & 7) -- the TMA's
                // 128-byte swizzle, conflict-free for the stores -- and one lane hands the box to the TMA as a bulk tensor reduction
                // (UTMAREDG: the fp32 add happens in L2; rows past M are dropped by the tensor map).  Two boxes per warp: chunk c is built
                // while the reduction of chunk c-1 still reads its box.
                const int col0 = n0 + half * HC;
                constexpr int NCH = HC / 32;
                const uint32_t bias_s = smem_base + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES + (warp - 2) * 512;     // 128 floats per warp
                const uint32_t box0 = smem_base + C::RED_OFF + (warp - 2) * 2 * C::RED_BOX_BYTES;
                __syncwarp();                                   // the previous tile's bias reads are done
                {
                    const float4 b = p.bias ? *reinterpret_cast<const float4*>(p.bias + col0 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    sts4(bias_s + lane * 16, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
                }
                __syncwarp();
                mbar_wait(tfull_bar(acc), acc_phase, 0x31);
                tc_fence_after();
                if (dbg_w) GEMM_TS(6);
                const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    tmem_ld32(t_row + half * HC + c * 32, r);
                    const uint32_t box = box0 + (c & 1) * C::RED_BOX_BYTES;
                    if (lane == 0) bulk_wait_group_read<1>();   // the reduction issued two chunks ago has read this box
                    __syncwarp();
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint4 b = lds4(bias_s + (c * 32 + 4 * k) * 4);
                        sts4(box + lane * 128 + ((k ^ sw) << 4), __float_as_uint(__uint_as_float(r[4 * k]) + __uint_as_float(b.x)),
                             __float_as_uint(__uint_as_float(r[4 * k + 1]) + __uint_as_float(b.y)),
                             __float_as_uint(__uint_as_float(r[4 * k + 2]) + __uint_as_float(b.z)),
                             __float_as_uint(__uint_as_float(r[4 * k + 3]) + __uint_as_float(b.w)));
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tmC, box, col0 + c * 32, row_base);
                        bulk_commit_group();
                    }
                }
            } else if constexpr (EPI == EPI_F32 || EPI == EPI_F32_LN) {
                // Residual / static-embedding rows prefetched (coalesced) BEFORE waiting for the accumulator: the loads fly
                // under this tile's MMAs.  Layout of ad[c][it]: chunk c (32 columns), rows rr0 + 4*it, 4 columns at cc*4.
                // All addressing is hoisted: one 64-bit base per warp and tile, the loops only add compile-time multiples of the row
                // pitch.  What remains is the SM's store port: 31-32 B/clk per SM whatever the instruction (tools/microbench/store_rate.cu),
                // 24 + 24 B/clk as load + store, so this 128 KB in + 128 KB out costs >= 5.4 k cycles; measured at speed 6.4 k
                // (tools/gemm_timeline.py).  EPI_F32_RED above avoids the read in the SM and takes the same time.
                const int col0 = n0 + half * HC;
                constexpr int NCH = HC / 32;                 // 32-column chunks per warp
                const int rows_left = p.M - (row_base + rr0);        // rows rr0 + 4*it with 4*it < rows_left exist
                const size_t pitch4 = (size_t)4 * p.ld32;            // 4 rows down
                const size_t base_off = (size_t)(row_base + rr0) * p.ld32 + col0 + cc * 4;
                const float* ad_base = p.addend ? p.addend + base_off : nullptr;
                float* out_base = p.out32 + base_off;
                auto load_addend = [&](float4 (&dst)[8], int c) {
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        dst[it] = (ad_base && 4 * it < rows_left) ? *reinterpret_cast<const float4*>(ad_base + it * pitch4 + c * 32)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                };
                float4 ad[2][8];                             // double-buffered: chunk c+1 streams in while chunk c is finished
                load_addend(ad[0], 0);
                if (NCH > 1) load_addend(ad[1], 1);
                float4 bias4[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    bias4[c] = p.bias ? *reinterpret_cast<const float4*>(p.bias + col0 + c * 32 + cc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                mbar_wait(tfull_bar(acc), acc_phase, 0x31);
                tc_fence_after();
                if (dbg_w) GEMM_TS(6);
                const bool extras = (EPI == EPI_F32) && (p.table || p.out16);      // anchor-row select (embedding GEMM) / fp16 shadow (last FF2): rare launches
                // EPI_F32_LN: per-row moments of the finished rows (rows rr0 + 4*it of this warp, this thread's 4 columns of every chunk) and a
                // private stash of the finished values in the (now idle: one tile per CTA) pipeline stages for the normalisation pass.
                float rs[8], rq[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) { rs[it] = 0.f; rq[it] = 0.f; }
                const uint32_t vstash = smem_base + (warp - 2) * 16384 + lane * 16;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    tmem_ld32(t_row + half * HC + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 8; ++k) sts4(stg_w + k * 16, r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
                    __syncwarp();
                    const float4 b = bias4[c];
                    if (!extras) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const uint4 u = lds4(stg_r + it * 4 * STG_ROW_BYTES);
                            const float4 a = ad[c & 1][it];
                            const float4 v = make_float4(__uint_as_float(u.x) + a.x + b.x, __uint_as_float(u.y) + a.y + b.y,
                                                         __uint_as_float(u.z) + a.z + b.z, __uint_as_float(u.w) + a.w + b.w);
                            if (4 * it < rows_left) *reinterpret_cast<float4*>(out_base + it * pitch4 + c * 32) = v;
                            if constexpr (EPI == EPI_F32_LN) {
                                rs[it] += (v.x + v.y) + (v.z + v.w);
                                rq[it] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, rq[it]))));
                                sts4(vstash + (c * 8 + it) * 512, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
                            }
                        }
                    } else {
                        const int col = col0 + c * 32 + cc * 4;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int row = row_base + it * 4 + rr0;
                            const uint4 u = lds4(stg_r + it * 4 * STG_ROW_BYTES);
                            const float4 a = ad[c & 1][it];
                            float4 v = make_float4(__uint_as_float(u.x) + a.x + b.x, __uint_as_float(u.y) + a.y + b.y, __uint_as_float(u.z) + a.z + b.z,
                                                   __uint_as_float(u.w) + a.w + b.w);
                            if (row < p.M) {
                                if (p.table) {
                                    const float4 t = *reinterpret_cast<const float4*>(p.table + (size_t)(p.sel[row] ? 1 : 0) * p.N + col);
                                    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                                }
                                *reinterpret_cast<float4*>(out_base + it * pitch4 + c * 32) = v;
                                if (p.out16) *reinterpret_cast<uint2*>(p.out16 + (size_t)row * p.ld16 + col) = make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
                            }
                        }
                    }
                    if (c + 2 < NCH) load_addend(ad[c & 1], c + 2);
                    __syncwarp();
                }
                if constexpr (EPI == EPI_F32_LN) {
                    // ---- row moments: reduce over the 8 lanes that share a row (cc = lane & 7), publish to both CTAs of the cluster
                    const uint32_t rank = cluster_ctarank();
                    const uint32_t stats_local = smem_base + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES;      // [4 sources][128 rows] float2 (gamma area)
#pragma unroll
                    for (int it = 0; it < 8; ++it)
#pragma unroll
                        for (int o = 1; o < 8; o <<= 1) { rs[it] += __shfl_xor_sync(0xffffffffu, rs[it], o); rq[it] += __shfl_xor_sync(0xffffffffu, rq[it], o); }
                    if (cc == 0) {
                        const uint32_t peer = dsmem_addr(stats_local, rank ^ 1u), mine = dsmem_addr(stats_local, rank);
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const uint32_t off = ((rank * 2 + half) * 128 + q * 32 + rr0 + 4 * it) * 8;
                            dsmem_st_f32x2(mine + off, rs[it], rq[it]);
                            dsmem_st_f32x2(peer + off, rs[it], rq[it]);
                        }
                    }
                    __syncwarp();
                    cluster_arrive();
                    cluster_wait();
                    // ---- normalise + modulate the stashed values: the next GEMM's fp16 A operand
                    float mean[8], rstd[8];
                    size_t sel[8];
                    bool same = true;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rloc = q * 32 + rr0 + 4 * it;
                        float sm = 0.f, sq = 0.f;
#pragma unroll
                        for (int src = 0; src < 4; ++src) {
                            float a, b2;
                            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b2) : "r"(stats_local + (src * 128 + rloc) * 8) : "memory");
                            sm += a; sq += b2;
                        }
                        mean[it] = sm * (1.0f / 512.0f);
                        rstd[it] = rsqrtf(fmaxf(sq * (1.0f / 512.0f) - mean[it] * mean[it], 0.f) + 1e-5f);
                        const int row = min(m0 + rloc, p.M - 1);
                        sel[it] = (size_t)p.ln_row_of_tok[row] * (size_t)p.ln_row_stride;
                        same = same && sel[it] == sel[0];
                    }
                    same = __all_sync(0xffffffffu, same && sel[0] == __shfl_sync(0xffffffffu, sel[0], 0));
                    __half* ln_base = p.ln_out + (size_t)(row_base + rr0) * p.ln_ld + col0 + cc * 4;
                    const size_t lpitch4 = (size_t)4 * p.ln_ld;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int col = col0 + c * 32 + cc * 4;
                        float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), ad4 = mu;
                        if (same) { mu = *reinterpret_cast<const float4*>(p.ln_mul + sel[0] + col); ad4 = *reinterpret_cast<const float4*>(p.ln_add + sel[0] + col); }
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const uint4 u = lds4(vstash + (c * 8 + it) * 512);
                            if (!same) { mu = *reinterpret_cast<const float4*>(p.ln_mul + sel[it] + col); ad4 = *reinterpret_cast<const float4*>(p.ln_add + sel[it] + col); }
                            const float g = rstd[it], mn = mean[it];
                            const float y0 = (__uint_as_float(u.x) - mn) * g * (p.ln_one_plus + mu.x) + ad4.x;
                            const float y1 = (__uint_as_float(u.y) - mn) * g * (p.ln_one_plus + mu.y) + ad4.y;
                            const float y2 = (__uint_as_float(u.z) - mn) * g * (p.ln_one_plus + mu.z) + ad4.z;
                            const float y3 = (__uint_as_float(u.w) - mn) * g * (p.ln_one_plus + mu.w) + ad4.w;
                            if (4 * it < rows_left) *reinterpret_cast<uint2*>(ln_base + it * lpitch4 + c * 32) = make_uint2(pack_half2(y0, y1), pack_half2(y2, y3));
                        }
                    }
                }
            } else {
                // fp16 outputs: each staging round holds 64 output columns (128 B per row).
                mbar_wait(tfull_bar(acc), acc_phase, 0x31);
                tc_fence_after();
                if (dbg_w) GEMM_TS(6);
                constexpr int ROUNDS = (EPI == EPI_GEGLU) ? 1 : HC / 64;    // GEGLU: 2 x (32 h + 32 g) -> 64 outputs per warp
#pragma unroll 1
                for (int rd = 0; rd < ROUNDS; ++rd) {
                    uint32_t r2[32];
                    uint32_t pk[32];                       // 64 fp16 outputs of this thread's row
                    int out_col;                           // first output column of this round
                    if constexpr (EPI == EPI_F16_SILU) {
                        const int col = n0 + half * HC + rd * 64;
                        out_col = col;
                        tmem_ld32(t_row + half * HC + rd * 64, r);
                        tmem_ld32(t_row + half * HC + rd * 64 + 32, r2);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col + i), b1 = *reinterpret_cast<const float4*>(p.bias + col + 32 + i);
                            float v[8] = {__uint_as_float(r[i]) + b0.x, __uint_as_float(r[i + 1]) + b0.y, __uint_as_float(r[i + 2]) + b0.z, __uint_as_float(r[i + 3]) + b0.w,
                                          __uint_as_float(r2[i]) + b1.x, __uint_as_float(r2[i + 1]) + b1.y, __uint_as_float(r2[i + 2]) + b1.z, __uint_as_float(r2[i + 3]) + b1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = p.act ? fmaxf(v[j], 0.f) : silu(v[j]);
                            pk[i / 2] = pack_half2(v[0], v[1]);
                            pk[i / 2 + 1] = pack_half2(v[2], v[3]);
                            pk[16 + i / 2] = pack_half2(v[4], v[5]);
                            pk[16 + i / 2 + 1] = pack_half2(v[6], v[7]);
                        }
                    } else if constexpr (EPI == EPI_QKV) {
                        // One 64-column head per round.  section 0/1 = q/k (RMS-normalised), 2 = v.
                        const int hc = half * (HC / 64) + rd;
                        const int col = n0 + hc * 64;
                        out_col = col;
                        const int dmodel = p.N / 3, section = col / dmodel, head = (col % dmodel) / 64;
                        tmem_ld32(t_row + hc * 64, r);
                        tmem_ld32(t_row + hc * 64 + 32, r2);
                        tmem_ld_wait();
                        if (section < 2 && p.act == 1) {            // A/B: the scalar form (RAP_QKV_SCALAR=1)
                            float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int i = 0; i < 32; ++i) { const float a = __uint_as_float(r[i]), b = __uint_as_float(r2[i]); ss[i & 3] += a * a + b * b; }
                            const float inv = 8.0f / fmaxf(sqrtf((ss[0] + ss[1]) + (ss[2] + ss[3])), 1e-12f);
                            const float4* g4 = reinterpret_cast<const float4*>(gamma_s + (section * 8 + head) * 64);
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 g0 = g4[i / 4], g1 = g4[8 + i / 4];
                                pk[i / 2] = pack_half2(__uint_as_float(r[i]) * inv * g0.x, __uint_as_float(r[i + 1]) * inv * g0.y);
                                pk[i / 2 + 1] = pack_half2(__uint_as_float(r[i + 2]) * inv * g0.z, __uint_as_float(r[i + 3]) * inv * g0.w);
                                pk[16 + i / 2] = pack_half2(__uint_as_float(r2[i]) * inv * g1.x, __uint_as_float(r2[i + 1]) * inv * g1.y);
                                pk[16 + i / 2 + 1] = pack_half2(__uint_as_float(r2[i + 2]) * inv * g1.z, __uint_as_float(r2[i + 3]) * inv * g1.w);
                            }
                        } else if (section < 2) {
                            // packed f32x2 arithmetic: half the issue slots of the scalar form (the epilogue warps share their schedulers with
                            // the UMMA-issuing and the TMA warp; a lighter epilogue also shortens the K loop it overlaps, tools/gemm_timeline.py)
                            f32x2 ss2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                const f32x2 a = pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), b = pack2(__uint_as_float(r2[i]), __uint_as_float(r2[i + 1]));
                                ss2[(i >> 1) & 1] = ffma2(a, a, ss2[(i >> 1) & 1]);
                                ss2[2 + ((i >> 1) & 1)] = ffma2(b, b, ss2[2 + ((i >> 1) & 1)]);
                            }
                            float s0, s1;
                            unpack2(fadd2(fadd2(ss2[0], ss2[1]), fadd2(ss2[2], ss2[3])), s0, s1);
                            const float inv = 8.0f / fmaxf(sqrtf(s0 + s1), 1e-12f);     // F.normalize eps, * sqrt(head_dim)
                            const f32x2 inv2 = pack2(inv, inv);
                            const float4* g4 = reinterpret_cast<const float4*>(gamma_s + (section * 8 + head) * 64);
                            auto scaled = [&](uint32_t x0, uint32_t x1, float ga, float gb) {
                                float o0, o1;
                                unpack2(fmul2(fmul2(pack2(__uint_as_float(x0), __uint_as_float(x1)), inv2), pack2(ga, gb)), o0, o1);
                                return pack_half2(o0, o1);
                            };
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 g0 = g4[i / 4], g1 = g4[8 + i / 4];
                                pk[i / 2] = scaled(r[i], r[i + 1], g0.x, g0.y);
                                pk[i / 2 + 1] = scaled(r[i + 2], r[i + 3], g0.z, g0.w);
                                pk[16 + i / 2] = scaled(r2[i], r2[i + 1], g1.x, g1.y);
                                pk[16 + i / 2 + 1] = scaled(r2[i + 2], r2[i + 3], g1.z, g1.w);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                pk[i / 2] = pack_half2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                                pk[16 + i / 2] = pack_half2(__uint_as_float(r2[i]), __uint_as_float(r2[i + 1]));
                            }
                        }
                    } else {  // EPI_GEGLU: this warp's two (h,g) chunk pairs -> 64 consecutive output columns
                        out_col = (n0 / BN) * (BN / 2) + half * 64;
#pragma unroll
                        for (int cp = 0; cp < 2; ++cp) {
                            const int c = half * 2 + cp;
                            tmem_ld32(t_row + c * 32, r);              // h part
                            tmem_ld32(t_row + BN / 2 + c * 32, r2);    // g part
                            tmem_ld_wait();
                            const float* bh = p.bias + n0 + c * 32;
                            const float* bg = p.bias + n0 + BN / 2 + c * 32;
                            if (p.act == 1) {                   // A/B: the scalar Abramowitz-Stegun epilogue (RAP_GEGLU_SCALAR=1)
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    const float4 h4 = *reinterpret_cast<const float4*>(bh + i), g4 = *reinterpret_cast<const float4*>(bg + i);
                                    pk[cp * 16 + i / 2] = pack_half2((__uint_as_float(r[i]) + h4.x) * gelu_erf(__uint_as_float(r2[i]) + g4.x),
                                                                     (__uint_as_float(r[i + 1]) + h4.y) * gelu_erf(__uint_as_float(r2[i + 1]) + g4.y));
                                    pk[cp * 16 + i / 2 + 1] = pack_half2((__uint_as_float(r[i + 2]) + h4.z) * gelu_erf(__uint_as_float(r2[i + 2]) + g4.z),
                                                                         (__uint_as_float(r[i + 3]) + h4.w) * gelu_erf(__uint_as_float(r2[i + 3]) + g4.w));
                                }
                                continue;
                            }
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 h4 = *reinterpret_cast<const float4*>(bh + i), g4 = *reinterpret_cast<const float4*>(bg + i);
                                pk[cp * 16 + i / 2] = geglu2(fadd2(pack2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), pack2(h4.x, h4.y)),
                                                             fadd2(pack2(__uint_as_float(r2[i]), __uint_as_float(r2[i + 1])), pack2(g4.x, g4.y)));
                                pk[cp * 16 + i / 2 + 1] = geglu2(fadd2(pack2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), pack2(h4.z, h4.w)),
                                                                 fadd2(pack2(__uint_as_float(r2[i + 2]), __uint_as_float(r2[i + 3])), pack2(g4.z, g4.w)));
                            }
                        }
                    }
                    if (tma16_capable<BN, EPI, TWO, MC>() && p.tma16) {
                        // 64 fp16 outputs of this thread's row = one 128-byte row of a 32-row box in the TMA's 128-byte swizzle (16-byte piece k of
                        // row r at position k ^ (r & 7): conflict-free stores, no transposing read-back), handed to the TMA as one bulk tensor
                        // store (UTMASTG); rows past M are dropped by the tensor map.  Two boxes per warp alternate.
                        const uint32_t box = smem_base + C::RED_OFF + ((warp - 2) * 2 + (int)(box_sel & 1u)) * C::RED_BOX_BYTES;
                        ++box_sel;
                        if (lane == 0) bulk_wait_group_read<1>();      // the store issued two rounds ago has read this box
                        __syncwarp();
                        const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
                        for (int k = 0; k < 8; ++k) sts4(box + lane * 128 + ((k ^ sw) << 4), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmC, box, out_col, row_base);
                            bulk_commit_group();
                        }
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) sts4(stg_w + k * 16, pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
                    __syncwarp();
                    {
                        __half* dst = p.out16 + (size_t)(row_base + rr0) * p.ld16 + out_col + cc * 8;   // hoisted: the loop adds 4-row pitches
                        const size_t pitch4 = (size_t)4 * p.ld16;
                        const int rows_left = p.M - (row_base + rr0);
                        if (!p.conv_pitch) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const uint4 u = lds4(stg_r + it * 4 * STG_ROW_BYTES);
                                if (4 * it < rows_left) *reinterpret_cast<uint4*>(dst + it * pitch4) = u;
                            }
                        } else {                                 // convolution map: real positions only, plus the circular halo copies
                            const ptrdiff_t wrap = (ptrdiff_t)p.conv_w * p.ld16;
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const uint4 u = lds4(stg_r + it * 4 * STG_ROW_BYTES);
                                const int pos = (row_base + rr0 + 4 * it) % p.conv_rows, e = pos / p.conv_pitch, a = pos - e * p.conv_pitch;
                                if (4 * it < rows_left && e < p.conv_h && a >= 1 && a <= p.conv_w) {
                                    *reinterpret_cast<uint4*>(dst + it * pitch4) = u;
                                    if (a == p.conv_w) *reinterpret_cast<uint4*>(dst + it * pitch4 - wrap) = u;      // column 0 = last real column
                                    if (a == 1) *reinterpret_cast<uint4*>(dst + it * pitch4 + wrap) = u;             // column w+1 = first real column
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (dbg_w) GEMM_TS(7);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (TWO && crank != 0) mbar_arrive_cluster(dsmem_addr(tempty_bar(acc), 0u));   // the even CTA's UMMA warp waits for both epilogues
                else mbar_arrive(tempty_bar(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (EPI == EPI_F32_RED || (tma16_capable<BN, EPI, TWO, MC>() && p.tma16)) {   // every bulk reduction / store of this warp has completed before the CTA retires
            // .read: the boxes have been read out of shared memory (the CTA may retire); the global writes themselves complete with the grid,
            // like per-thread stores, and waiting for them here only lengthens the tail of latency-bound launches
            if (lane == 0) bulk_wait_group_read<0>();
            if (dbg_cta && dbg_w && lane == 0) p.dbg[15 * 8] = clock64();       // timeline: slot 7 = last reduction issued, this = all of them complete
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) { cluster_arrive(); cluster_wait(); }   // no CTA leaves while its partner may still multicast into it / arrive on its barriers
    if (warp == 1) { if constexpr (TWO) tmem_dealloc_2cta<C::TMEM_COLS>(tmem_base); else tmem_dealloc<C::TMEM_COLS>(tmem_base); }
    if (p.dbg && threadIdx.x == 0) {
        long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.dbg[128 + 4 * blockIdx.x + 2] = t; p.dbg[128 + 4 * blockIdx.x + 3] = clock64();
    }
}

template <int BN, int EPI, bool MC = false, bool TWO = false>
int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
    constexpr int SMEM = smem_bytes<BN, EPI, TWO, MC>();
    static bool configured = false;
    if (!configured) {
        RAP_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI, MC, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    if constexpr (EPI == EPI_F32_RED)
        RAP_REQUIRE(a.tm_out32 && a.addend == a.out32 && !a.table && !a.out16 && !a.conv_pitch,
                    "gemm: the reduction epilogue is the in-place residual form (addend == out32, tm_out32 set, no table / fp16 shadow / convolution)");
    GemmArgs a2 = a;
    a2.tma16 = (tma16_capable<BN, EPI, TWO, MC>() && a.tm_out16 && !a.conv_pitch) ? 1 : 0;
    const CUtensorMap& tmC = (EPI == EPI_F32_RED) ? *a.tm_out32 : a2.tma16 ? *a.tm_out16 : tmA;      // unused by the other epilogues
    const int m_tiles = ceil_div(a.M, BM), n_tiles = a.N / BN;
    ProfScope _ps((EPI == EPI_F32 || EPI == EPI_F32_LN || EPI == EPI_F32_RED) ? KC_GEMM_F32 : EPI == EPI_F16_SILU ? KC_GEMM_SILU : EPI == EPI_QKV ? KC_GEMM_QKV : KC_GEMM_GEGLU, st);
    if constexpr (MC || TWO) {
        const int pairs = ((m_tiles + 1) / 2) * n_tiles;
        const int clusters = std::min(pairs, num_sms() / 2);
        RAP_CUDA(launch_pdl_cluster(gemm_kernel<BN, EPI, MC, TWO>, dim3(2 * clusters), dim3(kThreads), SMEM, st, 2u, tmA, tmB, tmC, a2));
        note_launch();
        return RAP_OK;
    }
    const int tiles = m_tiles * n_tiles;
    const int grid = (EPI == EPI_F32_LN || tiles < num_sms()) ? tiles : num_sms();
    if constexpr (EPI == EPI_F32_LN) {
        // one 128x256 tile per CTA, the two CTAs of a 128-row block (column halves) are a cluster: CTA rank = blockIdx.x & 1 = N-tile index
        RAP_REQUIRE(a.N == 2 * BN && a.ln_out && a.ln_mul && a.ln_add && a.ln_row_of_tok && !a.table && !a.out16 && !a.conv_pitch,
                    "gemm: the LayerNorm epilogue needs N = 512 and the ln_* arguments (no table / fp16 shadow / convolution)");
        RAP_CUDA(launch_pdl_cluster(gemm_kernel<BN, EPI, false>, dim3(grid), dim3(kThreads), SMEM, st, 2u, tmA, tmB, tmC, a2));
        note_launch();
        return RAP_OK;
    }
    RAP_CUDA(launch_pdl(gemm_kernel<BN, EPI, false>, dim3(grid), dim3(kThreads), SMEM, st, tmA, tmB, tmC, a2));
    note_launch();
    return RAP_OK;
}

// RAP_GEMM_MC=1: 2-CTA clusters share the B tile by multicast.  Built, parity-tested and measured (r02, same-box alternating A/B, 3 rounds):
// 99.59 / 99.66 / 99.50 ms per pair with it, 99.46 / 99.56 / 99.54 without; in isolation 2-10 % SLOWER (QKV 16.6 vs 15.9 us, out-proj 10.3 vs
// 9.4).  It removes 27 % of the L2 reads (ncu lts__throughput 37.9 % -> 27.6 % on the GEGLU GEMM) but the K loop is bound by the round trip of
// a stage (commit -> producer -> 360-400 ns of load latency -> MMAs; three stages = 375-384 ns per step), not by L2 or SM bandwidth (the TMA
// alone delivers a 48 KB stage per 214-227 ns on all SMs, tools/microbench/load_rate.cu), and it couples the two CTAs' pipelines.  Off by default.
// RAP_GEMM_2CTA=1: pairs of CTAs run 256x256 UMMAs with `cta_group::2` (each SM ingests half of the weight tile).
bool gemm_two_cta() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_GEMM_2CTA"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v == 1;
}
bool gemm_multicast() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_GEMM_MC"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v == 1;
}

}  // namespace

int launch_gemm(GemmEpi epi, int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
    RAP_REQUIRE(a.M > 0 && a.N % BN == 0 && a.K % BK == 0 && a.K >= BK, "gemm: bad shape M=%d N=%d K=%d BN=%d", a.M, a.N, a.K, BN);
    if (BN == 256) {
        const bool mc = (a.mc == 1 || (a.mc == 0 && gemm_multicast())) && !a.conv_pitch && a.M > BM;      // a single row of tiles has nothing to share
        const bool two = (a.mc == 3 || (a.mc == 0 && gemm_two_cta())) && !a.conv_pitch && a.M > BM;
        if (two) {
            switch (epi) {
                case EPI_F32: return launch_t<256, EPI_F32, false, true>(tmA, tmB, a, st);
                case EPI_QKV: return launch_t<256, EPI_QKV, false, true>(tmA, tmB, a, st);
                case EPI_GEGLU: return launch_t<256, EPI_GEGLU, false, true>(tmA, tmB, a, st);
                default: break;
            }
        }
        switch (epi) {
            case EPI_F32: return mc ? launch_t<256, EPI_F32, true>(tmA, tmB, a, st) : launch_t<256, EPI_F32>(tmA, tmB, a, st);
            case EPI_QKV: return mc ? launch_t<256, EPI_QKV, true>(tmA, tmB, a, st) : launch_t<256, EPI_QKV>(tmA, tmB, a, st);
            case EPI_GEGLU: return mc ? launch_t<256, EPI_GEGLU, true>(tmA, tmB, a, st) : launch_t<256, EPI_GEGLU>(tmA, tmB, a, st);
            case EPI_F32_LN: return launch_t<256, EPI_F32_LN>(tmA, tmB, a, st);
            case EPI_F32_RED: return launch_t<256, EPI_F32_RED>(tmA, tmB, a, st);
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
