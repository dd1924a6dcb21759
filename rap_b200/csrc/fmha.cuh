// Varlen FMHA launcher (see fmha.cu).
#pragma once
#include "common.h"

namespace rap {

struct FmhaArgs {
    const int4* tiles;   // device: {q_row0, seg_start, seg_end, 0} per 256-row query block (two 128-row tiles)
    __half* out;         // [T, d_model] fp16, head h at columns [64h, 64h+64)
    int d_model;         // heads * 64; qkv rows are [q | k | v], each d_model wide
    float scale_log2;    // log2(e) / sqrt(head_dim)
    const float* score_bound;  // [heads] upper bound of |score * scale_log2| (from the qk-norm gammas), or null = unknown
    long long* dbg;      // optional clock64() timeline of CTA (0,0): [iter][16] slots (tools/fmha_timeline.py), else null
};

// tm_qkv: tensor map over the [T, 3*d_model] fp16 qkv buffer, box = 128 rows x 64 cols, SWIZZLE_128B.
// Query rows per CTA (256: two tiles ping-ponged by one UMMA warp; 128: one tile, two CTAs per SM) -- tile lists must match.
int fmha_rows_per_cta();
int launch_fmha(const CUtensorMap& tm_qkv, const FmhaArgs& args, int n_tiles, int heads, cudaStream_t st);

}  // namespace rap
