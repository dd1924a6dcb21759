// Persistent warp-specialised tcgen05 GEMM for the DiT projections:
//     C[M,N] = A[M,K] (fp16, K contiguous) x W[N,K]^T (fp16, K contiguous), fp32 accumulate in TMEM,
// with the DiT-specific epilogues fused (qk-RMSnorm, GEGLU, bias+residual, bias+SiLU).
// One CTA per SM; warp 0 = TMA producer, warp 1 = UMMA issuer (+TMEM owner), warps 2..5 = epilogue.
// Accumulators are double-buffered in TMEM (2 x BN columns) so tile i's epilogue overlaps tile i+1's MMAs.
#pragma once
#include "common.h"

namespace rap {

enum GemmEpi : int {
    EPI_F32 = 0,     // out32 = acc (+bias[col]) (+addend[row,col]) (+table[sel[row]][col]); optional fp16 shadow
    EPI_F16_SILU,    // out16 = act(acc + bias[col]), act = SiLU (args.act = 0) or ReLU (args.act = 1)
    EPI_QKV,         // per-head RMS norm * gamma * sqrt(hd) on q,k; v passthrough; out16 [M,N]
    EPI_GEGLU,       // tile cols [0,BN/2)=h, [BN/2,BN)=g (weights pre-interleaved); out16[M,N/2] = h*gelu(g)
    EPI_F32_LN,      // EPI_F32 (bias + residual -> out32) AND the LayerNorm of the finished rows, modulated, as the next GEMM's fp16 operand:
                     // ln_out[row] = LN(out32[row]) * (ln_one_plus + ln_mul[sel(row)]) + ln_add[sel(row)].  N must be 512 = 2 tiles of 256: the
                     // two CTAs that own a 128-row block form a thread-block cluster and exchange per-row (sum, sum of squares) through
                     // distributed shared memory, so no LayerNorm kernel runs between the GEMMs (norm.py:76, layer.py:163)
    EPI_F32_RED,     // out32 += acc + bias[col]: the in-place residual form of EPI_F32 (addend == out32) with NO residual read in the SM at all --
                     // each epilogue warp writes (acc + bias) into a swizzled shared-memory box and hands it to the TMA as a bulk tensor
                     // reduction (cp.reduce.async.bulk.tensor ... .add, fp32 add in L2).  Needs args.tm_out32 (fp32 map, 32x32 boxes).
};

struct GemmArgs {
    int M, N, K;
    const float* bias;        // [N] (packed order) or null
    const float* addend;      // [M, ld32] or null (may alias out32)
    float* out32;             // [M, ld32]
    __half* out16;            // [M, ld16] (shadow for EPI_F32, main output otherwise) or null
    int ld32, ld16;
    const float* table;       // [2, N] row-select table (anchor embedding) or null
    const unsigned char* sel; // [M] row selector for `table`
    const float* gamma;       // EPI_QKV: [2(q,k), heads, 64]
    int act;                  // EPI_F16_SILU: 0 = SiLU, 1 = ReLU
    long long* dbg;           // optional clock64() timeline of CTA 0: [tile][8] (tools/gemm_timeline.py), else null
    const CUtensorMap* tm_out32;   // EPI_F32_RED: host pointer to the fp32 tensor map of out32 (make_tmap_f32_sw128, 32-row boxes); copied into the launch
    const CUtensorMap* tm_out16;   // EPI_QKV / EPI_GEGLU (plain 256-wide tiles): host pointer to the fp16 tensor map of out16 with 32-row boxes
                                   // (make_tmap_f16_sw128(..., 32)) -> the epilogue leaves through bulk tensor stores; null = per-thread stores
    int tma16;                     // set by launch_gemm
    int raster;                    // tile order of the plain kernels: 0 = row-block major (default), 1 = column-tile major
    int mc;                   // pair modes: 0 = as RAP_GEMM_MC / RAP_GEMM_2CTA say (default off), 1 = B-tile multicast, 2 = plain, 3 = 2-CTA UMMA
    // EPI_F32_LN: modulation rows ln_mul / ln_add are selected per token, sel(row) = ln_row_of_tok[row] * ln_row_stride (AdaLN table row of the
    // token's scene; stride 0 + an all-zero index vector for the plain affine LayerNorm in front of the feed-forward block)
    const float* ln_mul;
    const float* ln_add;
    const int* ln_row_of_tok;
    int ln_row_stride;
    float ln_one_plus;
    __half* ln_out;           // [M, ln_ld] fp16
    int ln_ld;
    // Implicit 3x3 convolution (spinnet.cu): A is an activation map stored per item as (conv_h + 1) rows x conv_pitch = conv_w + 2 columns
    // of positions (azimuth halo columns hold the circular neighbours, the extra row is an all-zero separator = elevation padding of
    // this and the next item), so tap (de, da) of an output tile is the SAME 2-D TMA box shifted by de * conv_pitch + da rows: no im2col
    // operand exists.  k-block kb = tap * conv_chunks + 64-channel chunk.  The fp16 epilogue writes only real positions (+ their halo
    // copies), so separators stay zero.  conv_pitch = 0: plain GEMM.
    int conv_pitch, conv_rows, conv_h, conv_w, conv_chunks;
};

int launch_gemm(GemmEpi epi, int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args, cudaStream_t st);

}  // namespace rap
