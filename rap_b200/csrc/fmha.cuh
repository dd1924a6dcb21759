// Varlen FMHA launcher (see fmha.cu).
#pragma once
#include <vector>

#include "common.h"

namespace rap {

// One CTA's work: a query block (rows_per_cta rows) of one head against the keys [kv_lo, kv_hi) of its segment.  Blocks whose key range
// was split to fill the last wave (slot >= 0) leave un-normalised partial rows in the workspace; the last piece to finish combines them.
struct FmhaItem {
    int q_row0, seg_start, seg_end, head;
    int kv_lo, kv_hi;      // key rows of this piece, kv_lo a multiple of 128 rows from seg_start, kv_hi = seg_end or block aligned
    int slot;              // partial-result slot of the (query block, head) unit, or -1: whole key range, write the output directly
    int piece;             // piece index | (number of pieces << 8)
};

struct FmhaSchedule {
    std::vector<FmhaItem> items;   // launch order: whole units first (longest first), then the big pieces, then the small ones
    std::vector<int> cta_ptr;      // persistent schedule only: CTA c walks items[cta_ptr[c] .. cta_ptr[c+1]); empty for the one-item-per-CTA kernel
    int n_slots = 0;               // units that were split
    int max_pieces = 0;            // workspace stride: pieces per slot
    size_t workspace_floats() const;
    int n_ctas() const { return cta_ptr.empty() ? (int)items.size() : (int)cta_ptr.size() - 1; }
};

// Work list for one attention launch over segments cu[0..n] (query blocks never straddle a segment).  With split enabled (default;
// RAP_FMHA_SPLIT=0 disables) the units of the last, partially filled wave are cut along the key axis so that every SM ends together.
// force_split (tests): cut the last wave even where the cost model says it does not pay.
FmhaSchedule build_fmha_schedule(const std::vector<int>& cu, int heads, bool force_split = false);

// Persistent ("stream-K") schedule (opt-in, RAP_FMHA_PERSIST=1|2; faster as an isolated launch, not inside the sampler): ONE CTA per SM, each walking a contiguous range of the linearised (unit, key block)
// space of equal estimated cost, so that every SM finishes together whatever the number of (query block, head) units is -- 256 units on
// 148 SMs (cfg2) otherwise waste a quarter of the second wave.  A unit that straddles a CTA boundary is cut along its keys; its pieces
// leave un-normalised partial rows in the workspace and the piece that finishes last combines them in piece order (deterministic).
// No CTA ever waits for another one.  n_ctas <= 0: one per SM.  force_split (tests): tiny per-CTA budget so that small inputs are cut too.
// Default is the one-item-per-CTA kernel above.
FmhaSchedule build_fmha_schedule_persistent(const std::vector<int>& cu, int heads, int n_ctas = 0, bool force_split = false);
bool fmha_persistent_enabled();
extern int g_persist_contiguous_override;   // -1: as RAP_FMHA_PERSIST says; 0 / 1: wave-major / contiguous persistent schedule (micro entry points)
// The schedule the product uses for this layout (persistent unless disabled).
FmhaSchedule build_fmha_schedule_auto(const std::vector<int>& cu, int heads, bool force_split = false);

struct FmhaArgs {
    const FmhaItem* items;   // device copy of FmhaSchedule::items
    const int* cta_ptr;      // device copy of FmhaSchedule::cta_ptr (persistent kernel) or null (one item per CTA)
    __half* out;         // [T, d_model] fp16, head h at columns [64h, 64h+64)
    int d_model;         // heads * 64; qkv rows are [q | k | v], each d_model wide
    float scale_log2;    // log2(e) / sqrt(head_dim)
    const float* score_bound;  // [heads] upper bound of |score * scale_log2| (from the qk-norm gammas), or null = unknown
    int mma2;            // one-item-per-CTA kernel: a second UMMA-issuing warp (warp 10) takes query tile 1 (set by launch_fmha from RAP_FMHA_MMA2)
    int early_release;   // hand S back before the exponentials when score_bound rules out the slow path (set by launch_fmha from RAP_FMHA_EARLY)
    float* ws;           // partial rows [slot][piece][rows_per_cta][68] (64 un-normalised outputs, reference exponent, row sum, pad), or null
    int* counters;       // [n_slots] arrival counters, zero between launches (the combining CTA resets its counter)
    int ws_pieces;       // FmhaSchedule::max_pieces
    int gen;             // persistent kernel: generation number of this launch (never 0, different from the previous launch on this workspace):
                         // the value of the per-row "partial row published" flags inside the workspace
    long long* dbg;      // optional clock64() timeline of CTA 0: [iter][16] slots (tools/fmha_timeline.py), else null
};

// tm_qkv: tensor map over the [T, 3*d_model] fp16 qkv buffer, box = 128 rows x 64 cols, SWIZZLE_128B.
// Query rows per CTA (256: two tiles ping-ponged by one UMMA warp; 128: one tile, two CTAs per SM; 384: three tiles on 64-key blocks) --
// schedules must match.
int fmha_rows_per_cta();
// n_ctas = FmhaSchedule::n_ctas(): the grid (number of items for the one-item-per-CTA kernel).
// tm_kv64: the same buffer with 64-row boxes -- needed by the three-tile kernel (RAP_FMHA_NT=3: 384 query rows per CTA, 64-key blocks).
int launch_fmha(const CUtensorMap& tm_qkv, const FmhaArgs& args, int n_ctas, cudaStream_t st, const CUtensorMap* tm_kv64 = nullptr);
extern int g_fmha_rows_override;            // micro entry points / tests: force the query rows per CTA (0 = RAP_FMHA_NT)

}  // namespace rap
