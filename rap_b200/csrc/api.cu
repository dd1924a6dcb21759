// C ABI (include/rap_b200.h): weight packing, batch plans, the DiT forward schedule, the fused Euler
// sampler and the Procrustes entry points.  Host code only enqueues kernels on the caller's stream.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rap_b200.h"
#include "fmha.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace rap {

__device__ unsigned int g_dev_error = 0;

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};
void note_launch(int n) { g_launches += n; }
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

struct ProfEvent { int kc; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfEvent> g_prof_events;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_open[KC_COUNT];
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
bool profiling_enabled() { return g_prof_on; }
void profile_mark(int kc, cudaStream_t st, bool end) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, st);
    if (!end) g_prof_open[kc] = e;
    else g_prof_events.push_back({kc, g_prof_open[kc], e});
}

// Tile width of the N=512 fp32-epilogue GEMMs (out-proj, FF2, embedding): 256 halves the A-operand re-reads and the shared-memory
// traffic per flop; RAP_F32_BN=128 restores the two-wave 128x128 tiling for A/B measurements.
int f32_bn() {
    static int bn = 0;
    if (!bn) { const char* e = getenv("RAP_F32_BN"); bn = (e && atoi(e) == 128) ? 128 : 256; }
    return bn;
}

// RAP_LN_FUSE=1: LayerNorm in the producing GEMM's epilogue (EPI_F32_LN, 2-CTA clusters exchange the row moments over DSMEM) instead of
// ln_modulate_kernel.  Implemented, parity-tested and MEASURED SLOWER (r02: 103.95 vs 100.57 ms/pair): the normalisation pass costs the
// 8 epilogue warps of 128 CTAs ~8 us per launch where the stand-alone kernel, spread over every SM, takes 3.6 us -- these GEMMs are
// epilogue-bound, so work moved into an epilogue lands on their critical path.  Hence off by default.
bool ln_fused() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_LN_FUSE"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v == 1;
}

// RAP_GEMM_RED=1: the in-place residual GEMMs (out-proj, FF2) leave through bulk tensor reductions (EPI_F32_RED) instead of reading the
// residual into the SM and storing the sum.
bool gemm_red() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_GEMM_RED"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v == 1;
}

// RAP_GEGLU_SCALAR=1: the GEGLU epilogue's first form (scalar Abramowitz-Stegun erf, two MUFUs per output) for A/B runs; default is the
// packed f32x2 form with erfc = 2^poly (gemm.cu geglu2).
int geglu_act() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_GEGLU_SCALAR"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v;
}

int qkv_act() {                                             // RAP_QKV_SCALAR=1: scalar RMS-norm epilogue of the QKV GEMM (A/B)
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_QKV_SCALAR"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v;
}

// The QKV and GEGLU GEMMs write their fp16 tiles with bulk tensor stores (UTMASTG) from swizzled boxes instead of transposing them through
// padded staging rows for per-thread stores: 14.3 vs 14.5-14.7 us (QKV), 28.5 vs 29.5 us (GEGLU), 97.71 / 97.95 vs 98.21 / 98.24 ms per pair
// (same box, alternating).  RAP_GEMM_TMA16=0 restores the per-thread stores.
bool gemm_tma16() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_GEMM_TMA16"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v == 1;
}

int gemm_raster() {                                          // RAP_GEMM_RASTER=1: column-tile-major tile order in the QKV / GEGLU GEMMs (A/B)
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_GEMM_RASTER"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RAP_NO_PDL"); v = (e && atoi(e) == 1) ? 0 : 1; }
    return v == 1;
}

int num_sms() {                                             // per device: a process may drive different devices
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& n = cache[dev & 63];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// cuTensorMapEncodeTiled is fetched through the runtime so the library links (and loads on a CPU-only
// box) without libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_f16_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(RAP_ENODEV, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    RAP_REQUIRE(box_rows >= 1 && box_rows <= 256 && cols % 64 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0,
                "tensor map: bad geometry rows=%llu cols=%llu box_rows=%u", (unsigned long long)rows, (unsigned long long)cols, box_rows);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RAP_ECUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return RAP_OK;
}

int make_tmap_f32_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(RAP_ENODEV, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    RAP_REQUIRE(box_rows >= 1 && box_rows <= 256 && cols % 32 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && row_stride_elems % 4 == 0,
                "tensor map: bad geometry rows=%llu cols=%llu box_rows=%u", (unsigned long long)rows, (unsigned long long)cols, box_rows);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_elems * 4};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RAP_ECUDA, "cuTensorMapEncodeTiled (fp32) failed with %d", (int)r);
    return RAP_OK;
}

}  // namespace rap

using namespace rap;

// ======================================================================================= model
namespace {

constexpr int D = 512, HEADS = 8, HD = 64, FF_INNER = 2048, ADA_CH = 256;

// |q| <= 8 max|gamma_q|, |k| <= 8 max|gamma_k| after the RMS norm (norm.py:33)  =>  |q.k| log2(e)/8 <= 64 gq gk log2(e)/8.
__global__ void score_bound_kernel(const float* gamma /* [2][8][64] */, float* bound /* [8] */) {
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gq = 0.f, gk = 0.f;
    for (int c = lane; c < 64; c += 32) { gq = fmaxf(gq, fabsf(gamma[h * 64 + c])); gk = fmaxf(gk, fabsf(gamma[512 + h * 64 + c])); }
    for (int o = 16; o; o >>= 1) { gq = fmaxf(gq, __shfl_xor_sync(0xffffffffu, gq, o)); gk = fmaxf(gk, __shfl_xor_sync(0xffffffffu, gk, o)); }
    if (lane == 0) bound[h] = 64.0f * gq * gk * (1.4426950408889634f / 8.0f) * 1.001f;
}

struct WeightEntry {
    void* dst = nullptr;
    bool half = false;
    PackSpec spec{};
    int64_t src_numel = 0;
    bool loaded = false;
};

struct Slab {
    char* base = nullptr;
    size_t size = 0, used = 0;
    template <typename T> T* take(size_t n) {
        used = (used + 255) & ~size_t(255);
        T* p = reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
};

}  // namespace

struct rap_dit {
    int L = 0, F = 0, M = 0;   // layers, local feature dim, AdaLN modules (2 per layer)
    int device = 0;
    Slab slab;
    // embedding
    __half *w_emb_x = nullptr, *w_emb_s = nullptr;
    float *b_emb = nullptr, *anchor_emb = nullptr;
    // per layer
    std::vector<__half*> wqkv[2], wout[2], w1, w2;
    std::vector<float*> bout[2], gamma[2] /* [2(q,k),8,64] */, b1, b2, ffn_w, ffn_b;
    float* score_bound = nullptr;   // [2L][8]: per (layer, block, head) bound of |q.k| * log2(e)/8 implied by the qk-norm gammas
    // AdaLN MLPs, module m = 2*layer + blk
    __half *ada_w1 = nullptr, *ada_w2 = nullptr, *ada_w3 = nullptr;
    float *ada_b1 = nullptr, *ada_b2 = nullptr, *ada_b3 = nullptr;
    // head
    __half *wh0 = nullptr, *wh2 = nullptr;
    float *bh0 = nullptr, *bh2 = nullptr, *wh4 = nullptr;
    // B-operand tensor maps
    CUtensorMap tm_emb_x, tm_emb_s, tm_h0, tm_h2;
    std::vector<CUtensorMap> tm_qkv[2], tm_out[2], tm_w1, tm_w2;
    std::map<std::string, WeightEntry> entries;
    std::vector<std::string> key_order;
    bool finalized = false;
};

namespace {

PackSpec plain(int rows, int cols) {
    PackSpec s{};
    s.rows = rows; s.cols = cols; s.src_ld = cols; s.row_mode = 0; s.n_ranges = 1; s.range_start[0] = 0; s.range_len[0] = cols;
    return s;
}

void add_entry(rap_dit* d, const std::string& key, void* dst, bool half, const PackSpec& spec, int64_t src_numel) {
    WeightEntry e;
    e.dst = dst; e.half = half; e.spec = spec; e.src_numel = src_numel;
    d->entries[key] = e;
    d->key_order.push_back(key);
}

}  // namespace

extern "C" int rap_version(void) { return 100; }
extern "C" const char* rap_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t rap_launch_count(void) { return g_launches.load(); }

extern "C" int rap_device_error(unsigned int* code_out) {
    unsigned int code = 0, zero = 0;
    RAP_CUDA(cudaDeviceSynchronize());
    RAP_CUDA(cudaMemcpyFromSymbol(&code, g_dev_error, sizeof code));
    if (code) RAP_CUDA(cudaMemcpyToSymbol(g_dev_error, &zero, sizeof zero));
    if (code_out) *code_out = code;
    return RAP_OK;
}

extern "C" int rap_dit_create(rap_dit_t** out, int embed_dim, int num_layers, int num_heads, int local_feat_dim) {
    RAP_REQUIRE(out, "rap_dit_create: out is NULL");
    RAP_REQUIRE(embed_dim == D && num_heads == HEADS, "rap_dit_create: the sm_100a kernels are specialised for embed_dim=512, 8 heads of 64 (got %d, %d)", embed_dim, num_heads);
    RAP_REQUIRE(num_layers >= 1 && num_layers <= 64 && local_feat_dim >= 0 && local_feat_dim <= 44, "rap_dit_create: bad num_layers/local_feat_dim");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RAP_ENODEV, "rap_dit_create: no CUDA device (this library has no CPU fallback)");
    rap_dit* d = new rap_dit();
    d->L = num_layers; d->F = local_feat_dim; d->M = 2 * num_layers;
    cudaGetDevice(&d->device);
    const int L = d->L, M = d->M;
    size_t halfs = (size_t)D * 64 + (size_t)D * 128 + (size_t)L * (2 * 3 * D * D + 2 * D * D + 2 * FF_INNER * D + D * FF_INNER) +
                   (size_t)M * (D * ADA_CH + D * D + 2 * D * D) + (size_t)D * D + (size_t)(D / 2) * D;
    size_t floats = (size_t)D + 2 * D + (size_t)L * (2 * D + 2 * 2 * HEADS * HD + 2 * FF_INNER + D + 2 * D) + (size_t)M * (D + D + 2 * D) + D + D / 2 + 3 * (D / 2);
    d->slab.size = halfs * 2 + floats * 4 + 256 * (64 + 20 * (size_t)L);
    if (cudaMalloc(&d->slab.base, d->slab.size) != cudaSuccess) { delete d; return fail(RAP_ENOMEM, "rap_dit_create: cudaMalloc(%zu) failed", d->slab.size); }
    cudaMemset(d->slab.base, 0, d->slab.size);
    Slab& s = d->slab;
    const int EIN = 126 + 21 + d->F;   // 63 + 63 + 21 + F

    d->w_emb_x = s.take<__half>((size_t)D * 64);
    d->w_emb_s = s.take<__half>((size_t)D * 128);
    d->b_emb = s.take<float>(D);
    d->anchor_emb = s.take<float>(2 * D);
    add_entry(d, "anchor_part_emb.weight", d->anchor_emb, false, plain(2, D), 2 * D);
    {   // emb_proj.weight (512, 179): columns [cond 0:63 | x_t 63:126 | scale 126:147 | feat 147:179] (embedding.py:161-177)
        PackSpec sx{}; sx.rows = D; sx.cols = 64; sx.src_ld = EIN; sx.n_ranges = 1; sx.range_start[0] = 63; sx.range_len[0] = 63;
        PackSpec ss{}; ss.rows = D; ss.cols = 128; ss.src_ld = EIN; ss.n_ranges = 3;
        ss.range_start[0] = 0; ss.range_len[0] = 63; ss.range_start[1] = 126; ss.range_len[1] = 21; ss.range_start[2] = 147; ss.range_len[2] = d->F;
        add_entry(d, "encoding_manager.emb_proj.weight", d->w_emb_x, true, sx, (int64_t)D * EIN);
        add_entry(d, "encoding_manager.emb_proj.weight#static", d->w_emb_s, true, ss, (int64_t)D * EIN);
    }
    add_entry(d, "encoding_manager.emb_proj.bias", d->b_emb, false, plain(1, D), D);

    d->ada_w1 = s.take<__half>((size_t)M * D * ADA_CH); d->ada_b1 = s.take<float>((size_t)M * D);
    d->ada_w2 = s.take<__half>((size_t)M * D * D);      d->ada_b2 = s.take<float>((size_t)M * D);
    d->ada_w3 = s.take<__half>((size_t)M * 2 * D * D);  d->ada_b3 = s.take<float>((size_t)M * 2 * D);

    const char* blk_name[2] = {"self", "global"};
    for (int l = 0; l < L; ++l) {
        const std::string p = "transformer_layers." + std::to_string(l) + ".";
        for (int b = 0; b < 2; ++b) {
            const int m = 2 * l + b;
            const std::string q = p + blk_name[b];
            add_entry(d, q + "_prenorm.timestep_embedder.linear_1.weight", d->ada_w1 + (size_t)m * D * ADA_CH, true, plain(D, ADA_CH), (int64_t)D * ADA_CH);
            add_entry(d, q + "_prenorm.timestep_embedder.linear_1.bias", d->ada_b1 + (size_t)m * D, false, plain(1, D), D);
            add_entry(d, q + "_prenorm.timestep_embedder.linear_2.weight", d->ada_w2 + (size_t)m * D * D, true, plain(D, D), (int64_t)D * D);
            add_entry(d, q + "_prenorm.timestep_embedder.linear_2.bias", d->ada_b2 + (size_t)m * D, false, plain(1, D), D);
            add_entry(d, q + "_prenorm.linear.weight", d->ada_w3 + (size_t)m * 2 * D * D, true, plain(2 * D, D), (int64_t)2 * D * D);
            add_entry(d, q + "_prenorm.linear.bias", d->ada_b3 + (size_t)m * 2 * D, false, plain(1, 2 * D), 2 * D);
            d->wqkv[b].push_back(s.take<__half>((size_t)3 * D * D));
            d->wout[b].push_back(s.take<__half>((size_t)D * D));
            d->bout[b].push_back(s.take<float>(D));
            d->gamma[b].push_back(s.take<float>(2 * HEADS * HD));
            add_entry(d, q + "_qkv_proj.weight", d->wqkv[b][l], true, plain(3 * D, D), (int64_t)3 * D * D);
            add_entry(d, q + "_out_proj.weight", d->wout[b][l], true, plain(D, D), (int64_t)D * D);
            add_entry(d, q + "_out_proj.bias", d->bout[b][l], false, plain(1, D), D);
            add_entry(d, q + "_q_norm.gamma", d->gamma[b][l], false, plain(HEADS, HD), HEADS * HD);
            add_entry(d, q + "_k_norm.gamma", d->gamma[b][l] + HEADS * HD, false, plain(HEADS, HD), HEADS * HD);
        }
        d->ffn_w.push_back(s.take<float>(D)); d->ffn_b.push_back(s.take<float>(D));
        d->w1.push_back(s.take<__half>((size_t)2 * FF_INNER * D)); d->b1.push_back(s.take<float>(2 * FF_INNER));
        d->w2.push_back(s.take<__half>((size_t)D * FF_INNER)); d->b2.push_back(s.take<float>(D));
        add_entry(d, p + "ff_norm.weight", d->ffn_w[l], false, plain(1, D), D);
        add_entry(d, p + "ff_norm.bias", d->ffn_b[l], false, plain(1, D), D);
        PackSpec g = plain(2 * FF_INNER, D); g.row_mode = 1;          // [h rows | g rows] -> 128/128 interleave per 256-wide tile
        PackSpec gb{}; gb.rows = 2 * FF_INNER; gb.cols = 1; gb.src_ld = 1; gb.row_mode = 1; gb.n_ranges = 1; gb.range_start[0] = 0; gb.range_len[0] = 1;
        add_entry(d, p + "ff.net.0.proj.weight", d->w1[l], true, g, (int64_t)2 * FF_INNER * D);
        add_entry(d, p + "ff.net.0.proj.bias", d->b1[l], false, gb, 2 * FF_INNER);
        add_entry(d, p + "ff.net.2.weight", d->w2[l], true, plain(D, FF_INNER), (int64_t)D * FF_INNER);
        add_entry(d, p + "ff.net.2.bias", d->b2[l], false, plain(1, D), D);
    }
    d->score_bound = s.take<float>((size_t)M * HEADS);
    d->wh0 = s.take<__half>((size_t)D * D); d->bh0 = s.take<float>(D);
    d->wh2 = s.take<__half>((size_t)(D / 2) * D); d->bh2 = s.take<float>(D / 2);
    d->wh4 = s.take<float>(3 * (D / 2));
    add_entry(d, "final_mlp.0.weight", d->wh0, true, plain(D, D), (int64_t)D * D);
    add_entry(d, "final_mlp.0.bias", d->bh0, false, plain(1, D), D);
    add_entry(d, "final_mlp.2.weight", d->wh2, true, plain(D / 2, D), (int64_t)(D / 2) * D);
    add_entry(d, "final_mlp.2.bias", d->bh2, false, plain(1, D / 2), D / 2);
    add_entry(d, "final_mlp.4.weight", d->wh4, false, plain(3, D / 2), 3 * (D / 2));
    if (s.used > s.size) { cudaFree(s.base); delete d; return fail(RAP_ESTATE, "rap_dit_create: slab overflow (%zu > %zu)", s.used, s.size); }
    *out = d;
    return RAP_OK;
}

extern "C" void rap_dit_destroy(rap_dit_t* d) {
    if (!d) return;
    cudaFree(d->slab.base);
    delete d;
}

extern "C" int rap_dit_num_keys(const rap_dit_t* d) {
    int n = 0;
    for (auto& k : d->key_order) n += (k.find('#') == std::string::npos);
    return n;
}
extern "C" const char* rap_dit_key(const rap_dit_t* d, int index, int64_t* numel_out) {
    int n = 0;
    for (auto& k : d->key_order) {
        if (k.find('#') != std::string::npos) continue;
        if (n++ == index) {
            if (numel_out) *numel_out = d->entries.at(k).src_numel;
            return k.c_str();
        }
    }
    return nullptr;
}

extern "C" int rap_dit_load_weight(rap_dit_t* d, const char* key, const float* src, int64_t numel, void* stream) {
    RAP_REQUIRE(d && key && src, "rap_dit_load_weight: NULL argument");
    auto it = d->entries.find(key);
    RAP_REQUIRE(it != d->entries.end() && std::string(key).find('#') == std::string::npos, "rap_dit_load_weight: unexpected key '%s'", key);
    RAP_REQUIRE(numel == it->second.src_numel, "rap_dit_load_weight: '%s' has %lld elements, expected %lld", key, (long long)numel, (long long)it->second.src_numel);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RAP_TRY(launch_pack(src, it->second.dst, it->second.half, it->second.spec, st));
    it->second.loaded = true;
    auto aux = d->entries.find(std::string(key) + "#static");
    if (aux != d->entries.end()) {
        RAP_TRY(launch_pack(src, aux->second.dst, aux->second.half, aux->second.spec, st));
        aux->second.loaded = true;
    }
    d->finalized = false;
    return RAP_OK;
}

extern "C" int rap_dit_finalize(rap_dit_t* d) {
    RAP_REQUIRE(d, "rap_dit_finalize: NULL");
    for (auto& kv : d->entries)
        if (!kv.second.loaded) return fail(RAP_ESTATE, "rap_dit_finalize: missing key '%s'", kv.first.c_str());
    const int L = d->L;
    RAP_TRY(make_tmap_f16_sw128(&d->tm_emb_x, d->w_emb_x, D, 64, 64, 128));       // B operands: 128-row boxes (a 256-wide tile = two boxes per stage,
    RAP_TRY(make_tmap_f16_sw128(&d->tm_emb_s, d->w_emb_s, D, 128, 128, 128));    // or one per CTA of a multicast pair)
    RAP_TRY(make_tmap_f16_sw128(&d->tm_h0, d->wh0, D, D, D, 128));
    RAP_TRY(make_tmap_f16_sw128(&d->tm_h2, d->wh2, D / 2, D, D, 128));
    for (int b = 0; b < 2; ++b) { d->tm_qkv[b].resize(L); d->tm_out[b].resize(L); }
    d->tm_w1.resize(L); d->tm_w2.resize(L);
    for (int l = 0; l < L; ++l) {
        for (int b = 0; b < 2; ++b) {
            score_bound_kernel<<<1, 256>>>(d->gamma[b][l], d->score_bound + (size_t)(2 * l + b) * HEADS);
            RAP_TRY(make_tmap_f16_sw128(&d->tm_qkv[b][l], d->wqkv[b][l], 3 * D, D, D, 128));
            RAP_TRY(make_tmap_f16_sw128(&d->tm_out[b][l], d->wout[b][l], D, D, D, 128));
        }
        RAP_TRY(make_tmap_f16_sw128(&d->tm_w1[l], d->w1[l], 2 * FF_INNER, D, D, 128));
        RAP_TRY(make_tmap_f16_sw128(&d->tm_w2[l], d->w2[l], D, FF_INNER, FF_INNER, 128));
    }
    d->finalized = true;
    return RAP_OK;
}

// ======================================================================================= plan
struct rap_plan {
    rap_dit* dit = nullptr;
    int T = 0, B = 0, P = 0, VP = 0, n_cap = 0;
    int n_items_part = 0, n_items_batch = 0, fmha_pieces_part = 0, fmha_pieces_batch = 0;   // n_items_*: grid of the launch (CTAs)
    int *cta_ptr_part = nullptr, *cta_ptr_batch = nullptr;
    int fmha_gen = 0;              // generation number of the last attention launch on this plan's workspace (FmhaArgs::gen)
    // Whole-sampler CUDA graph (small scenes are launch-bound: 137 kernels of a few microseconds per flow step).  The graph works on plan-owned
    // staging buffers, so the caller's pointers never enter it; it is captured on a private stream and launched into the caller's.
    struct SamplerGraph { int steps = 0, rigid = 0, outs = 0; cudaGraphExec_t exec = nullptr; long long launches = 0; };
    std::vector<SamplerGraph> graphs;
    int eager_calls = 0;           // the first sampler call runs eagerly (function attributes, lazy module loading), the second one captures
    bool graphs_failed = false;
    cudaStream_t gstream = nullptr;
    float *g_x1 = nullptr, *g_ep = nullptr, *g_tr = nullptr, *g_xf = nullptr, *g_feat = nullptr, *cond_copy = nullptr;
    int g_steps_cap = 0;   // persistent attention schedule: item range of every CTA (null: one item per CTA)
    char* ws = nullptr;
    // layout
    int *cu_part = nullptr, *slot = nullptr, *scene_of_tok = nullptr, *zero_rows = nullptr;
    FmhaItem *items_part = nullptr, *items_batch = nullptr;
    float* fmha_ws = nullptr;      // partial rows of split attention units (shared by the two attention launches of a layer)
    int* fmha_counters = nullptr;
    // activations
    float *h = nullptr, *h_static = nullptr, *x_t = nullptr, *tbuf = nullptr;
    __half *x16 = nullptr, *attn16 = nullptr, *qkv = nullptr, *ffmid = nullptr, *a_s = nullptr, *a_x = nullptr, *hh16 = nullptr, *y1 = nullptr, *y2 = nullptr;
    float *ada_e0 = nullptr, *ada_e1 = nullptr, *ada_e2 = nullptr, *ada_tab = nullptr;
    CUtensorMap tm_x16, tm_attn16, tm_ffmid, tm_as, tm_ax, tm_hh16, tm_y1, tm_qkv_attn, tm_qkv_attn64, tm_h32, tm_qkv_st, tm_ffmid_st;
    const float* cond = nullptr;
    const uint8_t* anchor = nullptr;
    bool condition_set = false;
    size_t resid_bytes = 0;        // bytes of [h | x16]: the fp32 residual stream and its LayerNorm'd fp16 copy
};

namespace {

__global__ void sampler_timesteps_kernel(float* t, int num_steps) {   // t_i = 1 - i * (1/num_steps), rounded like float(python double)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < num_steps) t[i] = (float)(1.0 - (double)i * (1.0 / (double)num_steps));
}

}  // namespace

extern "C" int rap_plan_create(rap_plan_t** out, rap_dit_t* dit, int n_scenes, const int32_t* cu_b, const int32_t* ppp, int max_parts, int max_timesteps) {
    RAP_REQUIRE(out && dit && cu_b && ppp, "rap_plan_create: NULL argument");
    RAP_REQUIRE(dit->finalized, "rap_plan_create: call rap_dit_finalize first");
    RAP_REQUIRE(n_scenes >= 1 && max_parts >= 1 && cu_b[0] == 0, "rap_plan_create: bad layout");
    const int T = cu_b[n_scenes];
    RAP_REQUIRE(T >= 1, "rap_plan_create: empty batch");
    std::vector<int> cu_batch(cu_b, cu_b + n_scenes + 1), cu_part{0}, slot, scene_of_tok(T);
    for (int b = 0; b < n_scenes; ++b) {
        RAP_REQUIRE(cu_b[b + 1] >= cu_b[b], "rap_plan_create: cu_seqlens_batch not monotone");
        int sum = 0;
        for (int p = 0; p < max_parts; ++p) {
            const int n = ppp[b * max_parts + p];
            RAP_REQUIRE(n >= 0, "rap_plan_create: negative points_per_part");
            if (n > 0) { cu_part.push_back(cu_part.back() + n); slot.push_back(b * max_parts + p); }
            sum += n;
        }
        // same invariant the reference asserts in utils/point_clouds.py:41-44
        RAP_REQUIRE(sum == cu_b[b + 1] - cu_b[b], "rap_plan_create: sum(points_per_part[%d])=%d != scene length %d", b, sum, cu_b[b + 1] - cu_b[b]);
        for (int i = cu_b[b]; i < cu_b[b + 1]; ++i) scene_of_tok[i] = b;
    }
    const FmhaSchedule sp = build_fmha_schedule_auto(cu_part, HEADS), sb = build_fmha_schedule_auto(cu_batch, HEADS);
    const std::vector<FmhaItem>&tp = sp.items, &tb = sb.items;
    const size_t fmha_ws_floats = std::max(sp.workspace_floats(), sb.workspace_floats());
    const size_t fmha_slots = (size_t)std::max(sp.n_slots, sb.n_slots);

    rap_plan* pl = new rap_plan();
    pl->dit = dit; pl->T = T; pl->B = n_scenes; pl->P = max_parts; pl->VP = (int)slot.size();
    pl->n_cap = std::max(std::max(max_timesteps, n_scenes), 1);
    pl->n_items_part = sp.n_ctas(); pl->n_items_batch = sb.n_ctas(); pl->fmha_pieces_part = sp.max_pieces; pl->fmha_pieces_batch = sb.max_pieces;
    const int M = dit->M;
    const size_t Tp = (size_t)((T + 127) / 128) * 128;     // padded rows so tail tiles stay in-bounds for plain loads/stores
    size_t bytes = 0;
    auto need = [&](size_t n) { bytes = ((bytes + 255) & ~size_t(255)) + n; };
    need(4 * cu_part.size()); need(4 * slot.size() + 4); need(4 * (size_t)T); need(4 * (size_t)T); need(sizeof(FmhaItem) * tp.size()); need(sizeof(FmhaItem) * tb.size()); need(4 * sp.cta_ptr.size() + 4); need(4 * sb.cta_ptr.size() + 4);
    need(4 * fmha_ws_floats); need(4 * fmha_slots + 4);
    need(4 * Tp * D); need(4 * Tp * D); need(4 * Tp * 3); need(4 * (size_t)pl->n_cap);
    need(2 * Tp * D); need(2 * Tp * D); need(2 * Tp * 3 * D); need(2 * Tp * FF_INNER); need(2 * Tp * 128); need(2 * Tp * 64);
    need(2 * Tp * D); need(2 * Tp * D); need(2 * Tp * (D / 2));
    need(4 * (size_t)pl->n_cap * ADA_CH); need(4 * (size_t)pl->n_cap * M * D); need(4 * (size_t)pl->n_cap * M * D); need(4 * (size_t)pl->n_cap * M * 2 * D);
    bytes += 4096;
    if (cudaMalloc(&pl->ws, bytes) != cudaSuccess) { delete pl; return fail(RAP_ENOMEM, "rap_plan_create: cudaMalloc(%zu) failed", bytes); }
    cudaMemset(pl->ws, 0, bytes);
    Slab s; s.base = pl->ws; s.size = bytes;
    pl->cu_part = s.take<int>(cu_part.size()); pl->slot = s.take<int>(slot.size() + 1); pl->scene_of_tok = s.take<int>(T); pl->zero_rows = s.take<int>(T);
    pl->items_part = s.take<FmhaItem>(tp.size()); pl->items_batch = s.take<FmhaItem>(tb.size());
    if (!sp.cta_ptr.empty()) pl->cta_ptr_part = s.take<int>(sp.cta_ptr.size());
    if (!sb.cta_ptr.empty()) pl->cta_ptr_batch = s.take<int>(sb.cta_ptr.size());
    pl->fmha_ws = s.take<float>(fmha_ws_floats); pl->fmha_counters = s.take<int>(fmha_slots + 1);
    pl->h = s.take<float>(Tp * D); pl->x16 = s.take<__half>(Tp * D);       // adjacent: one L2 persistence window covers the residual stream
    pl->resid_bytes = (size_t)(reinterpret_cast<char*>(pl->x16 + Tp * D) - reinterpret_cast<char*>(pl->h));
    pl->h_static = s.take<float>(Tp * D); pl->x_t = s.take<float>(Tp * 3); pl->tbuf = s.take<float>(pl->n_cap);
    pl->attn16 = s.take<__half>(Tp * D); pl->qkv = s.take<__half>(Tp * 3 * D); pl->ffmid = s.take<__half>(Tp * FF_INNER);
    pl->a_s = s.take<__half>(Tp * 128); pl->a_x = s.take<__half>(Tp * 64);
    pl->hh16 = s.take<__half>(Tp * D); pl->y1 = s.take<__half>(Tp * D); pl->y2 = s.take<__half>(Tp * (D / 2));
    pl->ada_e0 = s.take<float>((size_t)pl->n_cap * ADA_CH); pl->ada_e1 = s.take<float>((size_t)pl->n_cap * M * D);
    pl->ada_e2 = s.take<float>((size_t)pl->n_cap * M * D); pl->ada_tab = s.take<float>((size_t)pl->n_cap * M * 2 * D);
    if (s.used > s.size) { cudaFree(pl->ws); delete pl; return fail(RAP_ESTATE, "rap_plan_create: workspace overflow"); }
    auto up = [&](void* dst, const void* src, size_t n) { return n == 0 ? cudaSuccess : cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice); };
    cudaError_t e = up(pl->cu_part, cu_part.data(), 4 * cu_part.size());
    if (e == cudaSuccess) e = up(pl->slot, slot.data(), 4 * slot.size());
    if (e == cudaSuccess) e = up(pl->scene_of_tok, scene_of_tok.data(), 4 * (size_t)T);
    if (e == cudaSuccess) e = up(pl->items_part, tp.data(), sizeof(FmhaItem) * tp.size());
    if (e == cudaSuccess) e = up(pl->items_batch, tb.data(), sizeof(FmhaItem) * tb.size());
    if (e == cudaSuccess && pl->cta_ptr_part) e = up(pl->cta_ptr_part, sp.cta_ptr.data(), 4 * sp.cta_ptr.size());
    if (e == cudaSuccess && pl->cta_ptr_batch) e = up(pl->cta_ptr_batch, sb.cta_ptr.data(), 4 * sb.cta_ptr.size());
    int rc = e == cudaSuccess ? RAP_OK : fail(RAP_ECUDA, "rap_plan_create: upload failed: %s", cudaGetErrorString(e));
    // A-operand tensor maps (rows = T so out-of-range rows of the last tile are zero-filled by TMA)
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_x16, pl->x16, T, D, D, 128);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_attn16, pl->attn16, T, D, D, 128);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_ffmid, pl->ffmid, T, FF_INNER, FF_INNER, 128);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_as, pl->a_s, T, 128, 128, 128);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_ax, pl->a_x, T, 64, 64, 128);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_hh16, pl->hh16, T, D, D, 128);
    if (!rc) rc = make_tmap_f32_sw128(&pl->tm_h32, pl->h, T, D, D, 32);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_qkv_st, pl->qkv, T, 3 * D, 3 * D, 32);          // 32-row boxes: one per epilogue warp and store
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_ffmid_st, pl->ffmid, T, FF_INNER, FF_INNER, 32);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_y1, pl->y1, T, D, D, 128);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_qkv_attn, pl->qkv, T, 3 * D, 3 * D, 128);
    if (!rc) rc = make_tmap_f16_sw128(&pl->tm_qkv_attn64, pl->qkv, T, 3 * D, 3 * D, 64);
    if (rc) { cudaFree(pl->ws); delete pl; return rc; }
    *out = pl;
    return RAP_OK;
}

extern "C" void rap_plan_destroy(rap_plan_t* pl) {
    if (!pl) return;
    for (auto& g : pl->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (pl->gstream) cudaStreamDestroy(pl->gstream);
    for (float* p : {pl->g_x1, pl->g_ep, pl->g_tr, pl->g_xf, pl->g_feat, pl->cond_copy}) cudaFree(p);
    cudaFree(pl->ws);
    delete pl;
}
extern "C" int rap_plan_tokens(const rap_plan_t* pl) { return pl ? pl->T : 0; }
extern "C" int rap_plan_valid_parts(const rap_plan_t* pl) { return pl ? pl->VP : 0; }

extern "C" int rap_plan_set_condition(rap_plan_t* pl, const float* cond, const float* feats, const float* scales, const uint8_t* anchor, void* stream) {
    RAP_REQUIRE(pl && cond && scales && anchor && (feats || pl->dit->F == 0), "rap_plan_set_condition: NULL argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rap_dit* d = pl->dit;
    RAP_TRY(launch_posenc_static(cond, scales, pl->scene_of_tok, feats, d->F, pl->a_s, pl->T, st));
    GemmArgs g{};
    g.M = pl->T; g.N = D; g.K = 128; g.bias = d->b_emb; g.out32 = pl->h_static; g.ld32 = D; g.table = d->anchor_emb; g.sel = anchor;
    RAP_TRY(launch_gemm(EPI_F32, f32_bn(), pl->tm_as, d->tm_emb_s, g, st));
    if (!pl->cond_copy) RAP_CUDA(cudaMalloc(&pl->cond_copy, sizeof(float) * 3 * (size_t)pl->T));
    RAP_CUDA(cudaMemcpyAsync(pl->cond_copy, cond, sizeof(float) * 3 * (size_t)pl->T, cudaMemcpyDeviceToDevice, st));
    pl->cond = pl->cond_copy; pl->anchor = anchor; pl->condition_set = true;
    return RAP_OK;
}

namespace {

// AdaLN scale/shift for n timesteps already in pl->tbuf: table[n][module][scale 512 | shift 512]
// The residual stream h (fp32, re-read and re-written by three GEMM epilogues and three LayerNorms per layer) and its fp16 LN copy are
// kept in L2 by an access-policy window on the launching stream: the per-layer working set (qkv, attention output, GEGLU mid) is larger
// than L2, so without it every epilogue write of h turns into an HBM write-back and every re-read into an HBM read.
// RAP_L2_PERSIST=0 disables (A/B measurements).  Failures are ignored: it is a hint.
struct L2Policy { int enabled = -1; size_t carve = 0, max_window = 0; };
L2Policy& l2_policy() {                                   // per device: limits and the carve-out differ between devices of one process
    static L2Policy pol[64];
    int dev = 0;
    cudaGetDevice(&dev);
    L2Policy& p = pol[dev & 63];
    if (p.enabled < 0) {
        const char* e = getenv("RAP_L2_PERSIST");
        p.enabled = (e && atoi(e) == 0) ? 0 : 1;
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess) p.carve = (size_t)v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev) == cudaSuccess) p.max_window = (size_t)v;
        p.carve = std::min(p.carve, (size_t)48 << 20);
        if (!p.carve || !p.max_window || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, p.carve) != cudaSuccess) p.enabled = 0;
        cudaGetLastError();
    }
    return p;
}
void pin_residual_stream(rap_plan* pl, cudaStream_t st) {
    const L2Policy& pol = l2_policy();
    if (!pol.enabled || !pl->resid_bytes) return;
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = pl->h;
    attr.accessPolicyWindow.num_bytes = std::min(pl->resid_bytes, pol.max_window);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)pol.carve / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}
// The window is a property of the CALLER's stream: take it off again when the call has enqueued its kernels (launches already enqueued
// keep the policy they were launched with), so later work of the application on that stream is not steered into the carve-out.
void unpin_residual_stream(cudaStream_t st) {
    if (!l2_policy().enabled) return;
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

int ada_table(rap_plan* pl, int n, cudaStream_t st) {
    rap_dit* d = pl->dit;
    RAP_TRY(launch_timestep_sincos(pl->tbuf, pl->ada_e0, n, st));
    RAP_TRY(launch_ada_linear(d->ada_w1, d->ada_b1, pl->ada_e0, 0, pl->ada_e1, n, d->M, D, ADA_CH, 1, st));
    RAP_TRY(launch_ada_linear(d->ada_w2, d->ada_b2, pl->ada_e1, 1, pl->ada_e2, n, d->M, D, D, 1, st));
    RAP_TRY(launch_ada_linear(d->ada_w3, d->ada_b3, pl->ada_e2, 1, pl->ada_tab, n, d->M, 2 * D, D, 0, st));
    return RAP_OK;
}

// One velocity evaluation.  x: (T,3).  tab_row0: first row of the AdaLN table for this call; per_scene: rows indexed by scene.
// Leaves y2 (T,256) ready for the head tail; h holds the transformer features.
int dit_body(rap_plan* pl, const float* x, int tab_row0, bool per_scene, cudaStream_t st) {
    rap_dit* d = pl->dit;
    const int T = pl->T, M = d->M;
    const float scale_log2 = 1.4426950408889634f / 8.0f;   // log2(e) / sqrt(64)
    RAP_TRY(launch_posenc_x(x, pl->a_x, T, st));
    const int row_stride = M * 2 * D;
    const int* rows = per_scene ? pl->scene_of_tok : pl->zero_rows;
    const bool fuse = ln_fused();
    // LayerNorm of the finished residual rows as the NEXT GEMM's fp16 operand (x16), either from the producing GEMM's epilogue (default:
    // EPI_F32_LN, no LayerNorm kernel at all) or from ln_modulate_kernel (RAP_LN_FUSE=0, A/B measurements).  `m` = AdaLN module 2*layer+blk
    // whose [scale | shift] row modulates it, or -1 - layer for the plain affine LayerNorm in front of layer's feed-forward block.
    auto ln_args = [&](GemmArgs& g, int m) {
        if (m >= 0) {
            const float* ss = pl->ada_tab + ((size_t)tab_row0 * M + m) * 2 * D;
            g.ln_mul = ss; g.ln_add = ss + D; g.ln_row_stride = row_stride; g.ln_row_of_tok = rows; g.ln_one_plus = 1.0f;
        } else {
            const int l = -1 - m;
            g.ln_mul = d->ffn_w[l]; g.ln_add = d->ffn_b[l]; g.ln_row_stride = 0; g.ln_row_of_tok = pl->zero_rows; g.ln_one_plus = 0.0f;
        }
        g.ln_out = pl->x16; g.ln_ld = D;
    };
    auto residual_gemm = [&](GemmArgs& g, const CUtensorMap& ta, const CUtensorMap& tb, int next_ln, bool has_next) -> int {
        if (fuse && has_next && f32_bn() == 256) {
            ln_args(g, next_ln);
            return launch_gemm(EPI_F32_LN, 256, ta, tb, g, st);
        }
        if (gemm_red() && g.addend == g.out32 && !g.out16 && !g.table && f32_bn() == 256) {
            g.tm_out32 = &pl->tm_h32;
            RAP_TRY(launch_gemm(EPI_F32_RED, 256, ta, tb, g, st));
        } else {
            RAP_TRY(launch_gemm(EPI_F32, f32_bn(), ta, tb, g, st));
        }
        if (has_next) {
            GemmArgs t{};
            ln_args(t, next_ln);
            RAP_TRY(launch_ln_modulate(pl->h, pl->x16, T, t.ln_mul, t.ln_add, t.ln_row_stride, t.ln_row_of_tok, t.ln_one_plus, st));
        }
        return RAP_OK;
    };
    {
        GemmArgs g{};
        g.M = T; g.N = D; g.K = 64; g.addend = pl->h_static; g.out32 = pl->h; g.ld32 = D;
        RAP_TRY(residual_gemm(g, pl->tm_ax, d->tm_emb_x, 0, true));                 // + LN for layer 0's part-attention block
    }
    for (int l = 0; l < d->L; ++l) {
        for (int b = 0; b < 2; ++b) {
            GemmArgs q{};
            q.M = T; q.N = 3 * D; q.K = D; q.out16 = pl->qkv; q.ld16 = 3 * D; q.gamma = d->gamma[b][l]; q.act = qkv_act();
            if (gemm_tma16()) q.tm_out16 = &pl->tm_qkv_st;
            q.raster = gemm_raster();
            RAP_TRY(launch_gemm(EPI_QKV, 256, pl->tm_x16, d->tm_qkv[b][l], q, st));
            FmhaArgs f{};
            f.items = b == 0 ? pl->items_part : pl->items_batch; f.cta_ptr = b == 0 ? pl->cta_ptr_part : pl->cta_ptr_batch; f.ws = pl->fmha_ws; f.counters = pl->fmha_counters;
            f.ws_pieces = b == 0 ? pl->fmha_pieces_part : pl->fmha_pieces_batch; f.out = pl->attn16; f.d_model = D; f.scale_log2 = scale_log2;
            f.score_bound = d->score_bound + (size_t)(2 * l + b) * HEADS;
            if (++pl->fmha_gen <= 0) pl->fmha_gen = 1;
            f.gen = pl->fmha_gen;
            RAP_TRY(launch_fmha(pl->tm_qkv_attn, f, b == 0 ? pl->n_items_part : pl->n_items_batch, st, &pl->tm_qkv_attn64));
            GemmArgs o{};
            o.M = T; o.N = D; o.K = D; o.bias = d->bout[b][l]; o.addend = pl->h; o.out32 = pl->h; o.ld32 = D;
            RAP_TRY(residual_gemm(o, pl->tm_attn16, d->tm_out[b][l], b == 0 ? 2 * l + 1 : -1 - l, true));   // + LN for the global block / the feed-forward
        }
        GemmArgs f1{};
        f1.M = T; f1.N = 2 * FF_INNER; f1.K = D; f1.bias = d->b1[l]; f1.out16 = pl->ffmid; f1.ld16 = FF_INNER; f1.act = geglu_act();
        if (gemm_tma16()) f1.tm_out16 = &pl->tm_ffmid_st;
        f1.raster = gemm_raster();
        RAP_TRY(launch_gemm(EPI_GEGLU, 256, pl->tm_x16, d->tm_w1[l], f1, st));
        GemmArgs f2{};
        f2.M = T; f2.N = D; f2.K = FF_INNER; f2.bias = d->b2[l]; f2.addend = pl->h; f2.out32 = pl->h; f2.ld32 = D;
        const bool last = l == d->L - 1;
        if (last) { f2.out16 = pl->hh16; f2.ld16 = D; }                             // the head reads a plain fp16 copy, no LayerNorm
        RAP_TRY(residual_gemm(f2, pl->tm_ffmid, d->tm_w2[l], 2 * (l + 1), !last));  // + LN for the next layer's part-attention block
    }
    GemmArgs h0{};
    h0.M = T; h0.N = D; h0.K = D; h0.bias = d->bh0; h0.out16 = pl->y1; h0.ld16 = D;
    RAP_TRY(launch_gemm(EPI_F16_SILU, 128, pl->tm_hh16, d->tm_h0, h0, st));
    GemmArgs h2{};
    h2.M = T; h2.N = D / 2; h2.K = D; h2.bias = d->bh2; h2.out16 = pl->y2; h2.ld16 = D / 2;
    RAP_TRY(launch_gemm(EPI_F16_SILU, 128, pl->tm_y1, d->tm_h2, h2, st));
    return RAP_OK;
}

}  // namespace

extern "C" int rap_dit_forward(rap_plan_t* pl, const float* x, const float* timesteps, float* velocity, float* features, void* stream) {
    RAP_REQUIRE(pl && x && timesteps && velocity, "rap_dit_forward: NULL argument");
    RAP_REQUIRE(pl->condition_set, "rap_dit_forward: call rap_plan_set_condition first");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RAP_CUDA(cudaMemcpyAsync(pl->tbuf, timesteps, sizeof(float) * pl->B, cudaMemcpyDeviceToDevice, st));
    pin_residual_stream(pl, st);
    RAP_TRY(ada_table(pl, pl->B, st));
    RAP_TRY(dit_body(pl, x, 0, true, st));
    RAP_TRY(launch_head_final(pl->y2, pl->dit->wh4, pl->T, velocity, nullptr, 0.f, 0.f, nullptr, nullptr, st));
    if (features) RAP_CUDA(cudaMemcpyAsync(features, pl->h, sizeof(float) * (size_t)pl->T * D, cudaMemcpyDeviceToDevice, st));
    unpin_residual_stream(st);
    return RAP_OK;
}

namespace {

// Enqueues the whole S-step sampler on `st` (kernels + device-to-device copies only, no host synchronisation).
int enqueue_sampler(rap_plan* pl, const float* x_1, int num_steps, int rigidity_forcing, float* end_traj, float* traj, float* x_final, float* features,
                    cudaStream_t st) {
    const size_t row = (size_t)pl->T * 3;
    sampler_timesteps_kernel<<<ceil_div(num_steps, 64), 64, 0, st>>>(pl->tbuf, num_steps);
    RAP_CUDA(cudaGetLastError());
    note_launch(1);
    RAP_TRY(ada_table(pl, num_steps, st));
    RAP_CUDA(cudaMemcpyAsync(pl->x_t, x_1, sizeof(float) * row, cudaMemcpyDeviceToDevice, st));   // x_t = x_1.clone()
    const double dt = 1.0 / (double)num_steps;
    for (int step = 0; step < num_steps; ++step) {
        const double t = 1.0 - step * dt;
        RAP_TRY(dit_body(pl, pl->x_t, step, false, st));
        float* x0hat = end_traj ? end_traj + step * row : reinterpret_cast<float*>(pl->a_x);   // scratch when not recorded (a_x is rebuilt next step)
        float* xt_out = (traj && !rigidity_forcing) ? traj + step * row : nullptr;
        RAP_TRY(launch_head_final(pl->y2, pl->dit->wh4, pl->T, nullptr, pl->x_t, (float)t, (float)dt, x0hat, xt_out, st));
        if (rigidity_forcing) {
            // x_t = rigid(x0_hat) * (1 - t + dt) + x_1 * (t - dt)      (sampler.py:59-60)
            RAP_TRY(launch_procrustes(pl->cond, x0hat, pl->cu_part, nullptr, pl->VP, nullptr, nullptr, pl->x_t, traj ? traj + step * row : nullptr,
                                      x_1, (float)(1.0 - t + dt), (float)(t - dt), st, nullptr, nullptr, nullptr, pl->T));
        }
    }
    if (x_final) RAP_CUDA(cudaMemcpyAsync(x_final, pl->x_t, sizeof(float) * row, cudaMemcpyDeviceToDevice, st));
    if (features) RAP_CUDA(cudaMemcpyAsync(features, pl->h, sizeof(float) * (size_t)pl->T * D, cudaMemcpyDeviceToDevice, st));
    return RAP_OK;
}

// RAP_GRAPH: 1 = always replay the sampler as a CUDA graph, 0 = never, unset = whenever the trajectory staging buffers stay below
// kGraphStagingBytes.  Measured (tools/graph_ab.py, one B200): 6.32 -> 6.22 ms at 2 x 512 points / 5 steps, 29.83 -> 29.47 ms at 2 x 1024 / 20
// steps, 99.17 -> 98.39 ms at 2 x 4096 / 20 steps: 1-2 % everywhere -- the path is not launch-bound even at 1024 tokens (690 kernels in 6.3 ms:
// the kernels' own fixed latencies, ~9 us each, are the floor there), the graph only removes the host-side launch gaps.
// A plan captures on its THIRD sampler call: layouts that are used once or twice (dynamic batch samplers) never pay for an instantiation.
constexpr size_t kGraphStagingBytes = (size_t)256 << 20;
constexpr int kEagerCallsBeforeCapture = 2;
int graph_mode() {                                            // read per call (one getenv per sampler call): tests flip it inside one process
    const char* e = getenv("RAP_GRAPH");
    return e ? (atoi(e) != 0 ? 1 : 0) : -1;
}

int ensure_graph_buffers(rap_plan* pl, int num_steps) {
    const size_t row = (size_t)pl->T * 3;
    if (!pl->gstream) RAP_CUDA(cudaStreamCreateWithFlags(&pl->gstream, cudaStreamNonBlocking));
    if (!pl->g_x1) {
        RAP_CUDA(cudaMalloc(&pl->g_x1, sizeof(float) * row));
        RAP_CUDA(cudaMalloc(&pl->g_xf, sizeof(float) * row));
        RAP_CUDA(cudaMalloc(&pl->g_feat, sizeof(float) * (size_t)pl->T * D));
    }
    if (num_steps > pl->g_steps_cap) {           // trajectories grow with the step count: re-allocate and drop the graphs that point at the old ones
        if (!pl->graphs.empty()) RAP_CUDA(cudaDeviceSynchronize());            // (a replay may still be running on them)
        for (auto& g : pl->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        pl->graphs.clear();
        cudaFree(pl->g_ep); cudaFree(pl->g_tr);
        pl->g_ep = pl->g_tr = nullptr;
        RAP_CUDA(cudaMalloc(&pl->g_ep, sizeof(float) * row * num_steps));
        RAP_CUDA(cudaMalloc(&pl->g_tr, sizeof(float) * row * num_steps));
        pl->g_steps_cap = num_steps;
    }
    return RAP_OK;
}

}  // namespace

extern "C" int rap_sample_euler(rap_plan_t* pl, const float* x_1, int num_steps, int rigidity_forcing, float* end_traj, float* traj,
                                float* x_final, float* features, void* stream) {
    RAP_REQUIRE(pl && x_1, "rap_sample_euler: NULL argument");
    RAP_REQUIRE(pl->condition_set, "rap_sample_euler: call rap_plan_set_condition first");
    RAP_REQUIRE(num_steps >= 1 && num_steps <= pl->n_cap, "rap_sample_euler: num_steps=%d outside [1,%d] (max_timesteps of the plan)", num_steps, pl->n_cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t row = (size_t)pl->T * 3;
    const int mode = graph_mode();
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &capturing) != cudaSuccess) { cudaGetLastError(); capturing = cudaStreamCaptureStatusNone; }
    const bool want_graph = !pl->graphs_failed && !profiling_enabled() && capturing == cudaStreamCaptureStatusNone &&    // (a caller capturing us gets plain launches)
                            (mode == 1 || (mode == -1 && 2 * sizeof(float) * row * (size_t)num_steps <= kGraphStagingBytes));
    if (!want_graph || pl->eager_calls++ < kEagerCallsBeforeCapture) {
        pin_residual_stream(pl, st);
        const int rc = enqueue_sampler(pl, x_1, num_steps, rigidity_forcing, end_traj, traj, x_final, features, st);
        unpin_residual_stream(st);
        return rc;
    }
    // ---- graph route: inputs -> staging, replay (capturing first if this (steps, rigidity, outputs) combination is new), staging -> outputs
    RAP_TRY(ensure_graph_buffers(pl, num_steps));
    const int outs = (end_traj ? 1 : 0) | (traj ? 2 : 0) | (x_final ? 4 : 0) | (features ? 8 : 0);
    rap_plan::SamplerGraph* g = nullptr;
    for (auto& c : pl->graphs) if (c.steps == num_steps && c.rigid == (rigidity_forcing ? 1 : 0) && c.outs == outs) g = &c;
    if (!g) {
        cudaGraph_t graph = nullptr;
        const int64_t before = rap_launch_count();
        pin_residual_stream(pl, pl->gstream);                 // kernel nodes inherit the capturing stream's access-policy window
        cudaError_t e = cudaStreamBeginCapture(pl->gstream, cudaStreamCaptureModeRelaxed);
        int rc = e == cudaSuccess ? enqueue_sampler(pl, pl->g_x1, num_steps, rigidity_forcing, end_traj ? pl->g_ep : nullptr, traj ? pl->g_tr : nullptr,
                                                    x_final ? pl->g_xf : nullptr, features ? pl->g_feat : nullptr, pl->gstream)
                                  : RAP_ECUDA;
        if (e == cudaSuccess) e = cudaStreamEndCapture(pl->gstream, &graph);
        unpin_residual_stream(pl->gstream);
        cudaGraphExec_t exec = nullptr;
        if (e == cudaSuccess && rc == RAP_OK && graph) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess || rc != RAP_OK || !exec) {      // capture is an optimisation: fall back to stream launches for good
            cudaGetLastError();
            pl->graphs_failed = true;
            note_launch((int)(before - rap_launch_count()));
            pin_residual_stream(pl, st);
            rc = enqueue_sampler(pl, x_1, num_steps, rigidity_forcing, end_traj, traj, x_final, features, st);
            unpin_residual_stream(st);
            return rc;
        }
        rap_plan::SamplerGraph ng;
        ng.steps = num_steps; ng.rigid = rigidity_forcing ? 1 : 0; ng.outs = outs; ng.exec = exec; ng.launches = rap_launch_count() - before;
        note_launch((int)-ng.launches);                       // counted when the graph actually runs (below)
        pl->graphs.push_back(ng);
        g = &pl->graphs.back();
    }
    RAP_CUDA(cudaMemcpyAsync(pl->g_x1, x_1, sizeof(float) * row, cudaMemcpyDeviceToDevice, st));
    RAP_CUDA(cudaGraphLaunch(g->exec, st));
    note_launch((int)g->launches);
    if (end_traj) RAP_CUDA(cudaMemcpyAsync(end_traj, pl->g_ep, sizeof(float) * row * num_steps, cudaMemcpyDeviceToDevice, st));
    if (traj) RAP_CUDA(cudaMemcpyAsync(traj, pl->g_tr, sizeof(float) * row * num_steps, cudaMemcpyDeviceToDevice, st));
    if (x_final) RAP_CUDA(cudaMemcpyAsync(x_final, pl->g_xf, sizeof(float) * row, cudaMemcpyDeviceToDevice, st));
    if (features) RAP_CUDA(cudaMemcpyAsync(features, pl->g_feat, sizeof(float) * (size_t)pl->T * D, cudaMemcpyDeviceToDevice, st));
    return RAP_OK;
}

// ======================================================================================= procrustes
extern "C" int rap_procrustes_fit(const float* source, const float* target, const int32_t* cu_part, const int32_t* slot, int n_parts,
                                  int64_t n_points, float* rotations, float* translations, void* stream) {
    RAP_REQUIRE(source && target && cu_part && rotations && translations && n_parts >= 0, "rap_procrustes_fit: bad argument");
    return launch_procrustes(source, target, cu_part, slot, n_parts, rotations, translations, nullptr, nullptr, nullptr, 0.f, 0.f,
                             static_cast<cudaStream_t>(stream), nullptr, nullptr, nullptr, n_points);
}
extern "C" int rap_procrustes_rigidify(const float* prediction, const float* condition, const int32_t* cu_part, int n_parts, int64_t n_points,
                                       float* out, const float* mix, float w_rigid, float w_mix, void* stream) {
    RAP_REQUIRE(prediction && condition && cu_part && out && n_parts >= 0, "rap_procrustes_rigidify: bad argument");
    return launch_procrustes(condition, prediction, cu_part, nullptr, n_parts, nullptr, nullptr, out, nullptr, mix, w_rigid, w_mix,
                             static_cast<cudaStream_t>(stream), nullptr, nullptr, nullptr, n_points);
}
extern "C" int rap_procrustes_fit_sse(const float* source, const float* target, const int32_t* cu_part, const int32_t* slot, int n_parts,
                                      int64_t n_points, float* rotations, float* translations, float* sse, void* stream) {
    RAP_REQUIRE(source && target && cu_part && rotations && translations && sse && n_parts >= 0, "rap_procrustes_fit_sse: bad argument");
    return launch_procrustes(source, target, cu_part, slot, n_parts, rotations, translations, nullptr, nullptr, nullptr, 0.f, 0.f,
                             static_cast<cudaStream_t>(stream), sse, nullptr, nullptr, n_points);
}
extern "C" int rap_rigid_sse(const float* source, const float* target, const int32_t* cu_part, const int32_t* slot, int n_parts,
                             int64_t n_points, const float* rotations, const float* translations, float* sse, void* stream) {
    RAP_REQUIRE(source && target && cu_part && rotations && translations && sse && n_parts >= 0, "rap_rigid_sse: bad argument");
    return launch_procrustes(source, target, cu_part, slot, n_parts, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f,
                             static_cast<cudaStream_t>(stream), sse, rotations, translations, n_points);
}
extern "C" int rap_plan_fit(rap_plan_t* pl, const float* source, const float* target, float* rotations, float* translations, void* stream) {
    RAP_REQUIRE(pl, "rap_plan_fit: NULL plan");
    return rap_procrustes_fit(source, target, pl->cu_part, pl->slot, pl->VP, pl->T, rotations, translations, stream);
}
extern "C" int rap_plan_rigidify(rap_plan_t* pl, const float* prediction, const float* condition, float* out, void* stream) {
    RAP_REQUIRE(pl, "rap_plan_rigidify: NULL plan");
    return rap_procrustes_rigidify(prediction, condition, pl->cu_part, pl->VP, pl->T, out, nullptr, 1.0f, 0.0f, stream);
}

// ======================================================================================= micro entry points
extern "C" int rap_test_gemm(int epi, int bn, int M, int N, int K, const void* A, const void* W, const float* bias, const float* addend,
                             float* out32, void* out16, const float* gamma, void* dbg, void* stream) {
    CUtensorMap ta, tb;
    RAP_TRY(make_tmap_f16_sw128(&ta, A, M, K, K, 128));
    RAP_TRY(make_tmap_f16_sw128(&tb, W, N, K, K, 128));
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K; g.bias = bias; g.addend = addend; g.out32 = out32; g.out16 = static_cast<__half*>(out16); g.gamma = gamma;
    g.mc = (epi >> 8) & 3;                                       // test hook: epi | 0x100 forces the multicast variant, | 0x200 forbids it
    g.raster = (epi >> 12) & 1;                                  // | 0x1000: column-tile-major tile order
    const bool tma16 = (epi >> 11) & 1;                          // | 0x800: fp16 tiles leave through bulk tensor stores (QKV / GEGLU, BN = 256)
    g.act = (epi >> 10) & 1;                                     // | 0x400: ReLU instead of SiLU (EPI_F16_SILU) / the scalar GELU form (EPI_GEGLU)
    epi &= 0xff;
    g.ld32 = N; g.ld16 = (epi == EPI_GEGLU) ? N / 2 : N; g.dbg = static_cast<long long*>(dbg);
    CUtensorMap tc;
    if (epi == EPI_F32_RED) { RAP_TRY(make_tmap_f32_sw128(&tc, out32, M, N, N, 32)); g.tm_out32 = &tc; }
    if (tma16) { RAP_TRY(make_tmap_f16_sw128(&tc, out16, M, g.ld16, g.ld16, 32)); g.tm_out16 = &tc; }
    return launch_gemm(static_cast<GemmEpi>(epi), bn, ta, tb, g, static_cast<cudaStream_t>(stream));
}

// EPI_F32_LN in isolation: out32 = A W^T + bias + addend (N = 512) and ln_out = LN(out32) * (one_plus + ln_mul[sel]) + ln_add[sel],
// sel(row) = row_of_tok[row] * row_stride -- the residual GEMM with the next block's LayerNorm in its epilogue (2-CTA clusters).
extern "C" int rap_test_gemm_ln(int M, int K, const void* A, const void* W, const float* bias, const float* addend, float* out32, const float* ln_mul,
                                const float* ln_add, const int32_t* row_of_tok, int row_stride, float one_plus, void* ln_out, void* stream) {
    CUtensorMap ta, tb;
    RAP_TRY(make_tmap_f16_sw128(&ta, A, M, K, K, 128));
    RAP_TRY(make_tmap_f16_sw128(&tb, W, D, K, K, 128));
    GemmArgs g{};
    g.M = M; g.N = D; g.K = K; g.bias = bias; g.addend = addend; g.out32 = out32; g.ld32 = D;
    g.ln_mul = ln_mul; g.ln_add = ln_add; g.ln_row_of_tok = row_of_tok; g.ln_row_stride = row_stride; g.ln_one_plus = one_plus;
    g.ln_out = static_cast<__half*>(ln_out); g.ln_ld = D;
    return launch_gemm(EPI_F32_LN, 256, ta, tb, g, static_cast<cudaStream_t>(stream));
}

// Host-only: the attention work list (whole units + key-split pieces of the last wave) for segments cu[0..n]; 8 ints per item
// {q_row0, seg_start, seg_end, head, kv_lo, kv_hi, slot, piece | n_pieces << 8}.  Returns the item count (negative on error).
extern "C" int rap_fmha_schedule(const int32_t* cu_host, int n_segments, int heads, int force_split, int32_t* items_out, int capacity, int* n_slots_out,
                                 int* max_pieces_out) {
    RAP_REQUIRE(cu_host && n_segments >= 1 && heads >= 1, "rap_fmha_schedule: bad argument");
    const FmhaSchedule sch = build_fmha_schedule(std::vector<int>(cu_host, cu_host + n_segments + 1), heads, force_split != 0);
    if (n_slots_out) *n_slots_out = sch.n_slots;
    if (max_pieces_out) *max_pieces_out = sch.max_pieces;
    if (items_out) {
        RAP_REQUIRE((size_t)capacity >= sch.items.size(), "rap_fmha_schedule: capacity %d < %zu items", capacity, sch.items.size());
        memcpy(items_out, sch.items.data(), sizeof(FmhaItem) * sch.items.size());
    }
    return (int)sch.items.size();
}

// Timing aid: `iters` back-to-back launches of one attention call between two events (schedule built once, like a plan does).
namespace {
// Device copy of an attention schedule for the micro entry points (a plan holds the same data for the product path).
struct DevSchedule {
    FmhaItem* items = nullptr; int* cta_ptr = nullptr; float* ws = nullptr; int* counters = nullptr;
    int upload(const FmhaSchedule& sch) {
        RAP_CUDA(cudaMalloc(&items, sizeof(FmhaItem) * std::max<size_t>(sch.items.size(), 1)));
        RAP_CUDA(cudaMemcpy(items, sch.items.data(), sizeof(FmhaItem) * sch.items.size(), cudaMemcpyHostToDevice));
        if (!sch.cta_ptr.empty()) {
            RAP_CUDA(cudaMalloc(&cta_ptr, 4 * sch.cta_ptr.size()));
            RAP_CUDA(cudaMemcpy(cta_ptr, sch.cta_ptr.data(), 4 * sch.cta_ptr.size(), cudaMemcpyHostToDevice));
        }
        RAP_CUDA(cudaMalloc(&ws, 4 * sch.workspace_floats() + 16));
        RAP_CUDA(cudaMemset(ws, 0, 4 * sch.workspace_floats() + 16));             // the persistent kernel's "published" flags live in here
        RAP_CUDA(cudaMalloc(&counters, 4 * (size_t)sch.n_slots + 4));
        RAP_CUDA(cudaMemset(counters, 0, 4 * (size_t)sch.n_slots + 4));
        return RAP_OK;
    }
    ~DevSchedule() { cudaFree(items); cudaFree(cta_ptr); cudaFree(ws); cudaFree(counters); }
};
}  // namespace

// Host-only: the persistent ("stream-K") attention schedule for segments cu[0..n] on n_ctas CTAs (<= 0: one per SM of the current device,
// 148 without a device).  items as in rap_fmha_schedule; cta_ptr_out[n_ctas_used + 1] = item range of every CTA.  Returns the item count.
extern "C" int rap_fmha_schedule_persistent(const int32_t* cu_host, int n_segments, int heads, int n_ctas, int force_split, int32_t* items_out,
                                            int capacity, int32_t* cta_ptr_out, int cta_capacity, int* n_ctas_out, int* n_slots_out, int* max_pieces_out) {
    RAP_REQUIRE(cu_host && n_segments >= 1 && heads >= 1, "rap_fmha_schedule_persistent: bad argument");
    const FmhaSchedule sch = build_fmha_schedule_persistent(std::vector<int>(cu_host, cu_host + n_segments + 1), heads, n_ctas, force_split != 0);
    if (n_ctas_out) *n_ctas_out = sch.n_ctas();
    if (n_slots_out) *n_slots_out = sch.n_slots;
    if (max_pieces_out) *max_pieces_out = sch.max_pieces;
    if (items_out) {
        RAP_REQUIRE((size_t)capacity >= sch.items.size(), "rap_fmha_schedule_persistent: capacity %d < %zu items", capacity, sch.items.size());
        memcpy(items_out, sch.items.data(), sizeof(FmhaItem) * sch.items.size());
    }
    if (cta_ptr_out) {
        RAP_REQUIRE((size_t)cta_capacity >= sch.cta_ptr.size(), "rap_fmha_schedule_persistent: cta capacity %d < %zu", cta_capacity, sch.cta_ptr.size());
        memcpy(cta_ptr_out, sch.cta_ptr.data(), 4 * sch.cta_ptr.size());
    }
    return (int)sch.items.size();
}

// Timing aid: `iters` back-to-back launches of one attention call between two events (schedule built once, like a plan does).
// variant: -1 = what the product launches (RAP_FMHA_PERSIST), 0 = one item per CTA, 1 = persistent wave-major, 2 = persistent contiguous.
extern "C" int rap_bench_fmha_variant(const void* qkv, int T, int heads, const int32_t* cu_host, int n_segments, void* out, int iters, int variant,
                                      float* ms_per_launch) {
    RAP_REQUIRE(heads == HEADS && iters > 0 && ms_per_launch && variant >= -1 && variant <= 3, "rap_bench_fmha: bad argument");
    const std::vector<int> cu(cu_host, cu_host + n_segments + 1);
    FmhaSchedule sch;
    struct RowsGuard { ~RowsGuard() { g_fmha_rows_override = 0; } } rows_guard;
    if (variant >= 0) g_fmha_rows_override = variant == 3 ? 384 : 256;
    if (variant < 0) sch = build_fmha_schedule_auto(cu, heads);
    else if (variant == 0 || variant == 3) sch = build_fmha_schedule(cu, heads);
    else { g_persist_contiguous_override = variant == 2 ? 1 : 0; sch = build_fmha_schedule_persistent(cu, heads); g_persist_contiguous_override = -1; }
    DevSchedule ds;
    RAP_TRY(ds.upload(sch));
    CUtensorMap tm, tm64;
    RAP_TRY(make_tmap_f16_sw128(&tm, qkv, T, 3 * D, 3 * D, 128));
    RAP_TRY(make_tmap_f16_sw128(&tm64, qkv, T, 3 * D, 3 * D, 64));
    FmhaArgs f{};
    f.items = ds.items; f.cta_ptr = ds.cta_ptr; f.ws = ds.ws; f.counters = ds.counters; f.ws_pieces = sch.max_pieces;
    f.out = static_cast<__half*>(out); f.d_model = D; f.scale_log2 = 1.4426950408889634f / 8.0f;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = RAP_OK;
    for (int i = 0; i < 3 && !rc; ++i) { ++f.gen; rc = launch_fmha(tm, f, sch.n_ctas(), nullptr, &tm64); }
    cudaEventRecord(e0, nullptr);
    for (int i = 0; i < iters && !rc; ++i) { ++f.gen; rc = launch_fmha(tm, f, sch.n_ctas(), nullptr, &tm64); }
    cudaEventRecord(e1, nullptr);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(ms_per_launch, e0, e1);
    *ms_per_launch /= iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}
extern "C" int rap_bench_fmha(const void* qkv, int T, int heads, const int32_t* cu_host, int n_segments, void* out, int iters, float* ms_per_launch) {
    return rap_bench_fmha_variant(qkv, T, heads, cu_host, n_segments, out, iters, -1, ms_per_launch);
}

extern "C" int rap_test_fmha(const void* qkv, int T, int heads, const int32_t* cu_host, int n_segments, int force_split, void* out, void* dbg, void* stream) {
    RAP_REQUIRE(heads == HEADS, "rap_test_fmha: heads must be 8");
    std::vector<int> cu(cu_host, cu_host + n_segments + 1);
    // force_split: 0 / 1 = persistent kernel (wave-major schedule; 1 = tiny per-CTA share so small inputs are cut too), 4 / 5 = the same with one
    // contiguous range per CTA, 2 = one item per CTA with the last wave force-split, 3 = one item per CTA as the product launches it
    // 6 / 7 = as 2 / 3 with a second UMMA-issuing warp (one per query tile)
    // 8 / 9 = the three-tile kernel (384 query rows per CTA, 64-key blocks), last wave as the cost model says / force-split
    FmhaSchedule sch;
    struct RowsGuard { ~RowsGuard() { g_fmha_rows_override = 0; } } rows_guard;
    g_fmha_rows_override = (force_split == 8 || force_split == 9) ? 384 : 256;
    if (force_split == 8) force_split = 3;
    if (force_split == 9) force_split = 2;
    const bool mma2 = force_split == 6 || force_split == 7;
    if (mma2) force_split -= 4;
    if (force_split == 2 || force_split == 3) sch = build_fmha_schedule(cu, heads, force_split == 2);
    else { g_persist_contiguous_override = force_split >= 4 ? 1 : 0; sch = build_fmha_schedule_persistent(cu, heads, 0, (force_split & 1) != 0); g_persist_contiguous_override = -1; }
    DevSchedule ds;
    RAP_TRY(ds.upload(sch));
    CUtensorMap tm, tm64;
    int rc = make_tmap_f16_sw128(&tm, qkv, T, 3 * D, 3 * D, 128);
    if (!rc) rc = make_tmap_f16_sw128(&tm64, qkv, T, 3 * D, 3 * D, 64);
    if (!rc) {
        FmhaArgs f{};
        f.items = ds.items; f.cta_ptr = ds.cta_ptr; f.ws = ds.ws; f.counters = ds.counters; f.ws_pieces = sch.max_pieces;
        f.out = static_cast<__half*>(out); f.d_model = D; f.scale_log2 = 1.4426950408889634f / 8.0f; f.dbg = static_cast<long long*>(dbg);
        f.mma2 = mma2 ? 1 : 0;
        f.gen = 1;
        rc = launch_fmha(tm, f, sch.n_ctas(), static_cast<cudaStream_t>(stream), &tm64);
        f.gen = 2;                                                                                          // a second launch on the same workspace:
        if (!rc && sch.n_slots) rc = launch_fmha(tm, f, sch.n_ctas(), static_cast<cudaStream_t>(stream), &tm64);   // counters reset / stale row flags ignored
    }
    cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    return rc;
}

extern "C" int rap_test_ln_modulate(const float* h, void* out16, int T, const float* mul, const float* add, int row_stride, const int32_t* row_of_tok,
                                    float one_plus, void* stream) {
    return launch_ln_modulate(h, static_cast<__half*>(out16), T, mul, add, row_stride, row_of_tok, one_plus, static_cast<cudaStream_t>(stream));
}
// Copies plan-internal buffers out for debugging/tests: which = 0 scene_of_tok (T i32), 1 AdaLN table (n_cap*M*1024 f32), 2 h (T*512 f32),
// 3 h_static (T*512 f32), 4 x16 (T*512 f16), 5 qkv (T*1536 f16), 6 attn16 (T*512 f16).
extern "C" int rap_debug_read(rap_plan_t* pl, int which, void* dst, int64_t bytes) {
    RAP_REQUIRE(pl && dst, "rap_debug_read: NULL");
    const void* src = nullptr;
    switch (which) {
        case 0: src = pl->scene_of_tok; break;
        case 1: src = pl->ada_tab; break;
        case 2: src = pl->h; break;
        case 3: src = pl->h_static; break;
        case 4: src = pl->x16; break;
        case 5: src = pl->qkv; break;
        case 6: src = pl->attn16; break;
        default: return fail(RAP_EINVAL, "rap_debug_read: bad selector");
    }
    RAP_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice));
    return RAP_OK;
}

// ======================================================================================= profiling
extern "C" int rap_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = true;
    return RAP_OK;
}
// Host-synchronising.  ms_out / count_out: arrays of rap_profile_classes() entries (total ms, launches) per kernel class.
extern "C" int rap_profile_end(float* ms_out, int* count_out) {
    RAP_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = false;
    for (int i = 0; i < KC_COUNT; ++i) { if (ms_out) ms_out[i] = 0.f; if (count_out) count_out[i] = 0; }
    for (auto& ev : g_prof_events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev.a, ev.b);
        if (ms_out) ms_out[ev.kc] += ms;
        if (count_out) count_out[ev.kc] += 1;
        g_prof_pool.push_back(ev.a); g_prof_pool.push_back(ev.b);
    }
    g_prof_events.clear();
    return RAP_OK;
}
extern "C" int rap_profile_classes(void) { return KC_COUNT; }
extern "C" const char* rap_profile_class_name(int kc) {
    static const char* names[KC_COUNT] = {"gemm_f32_bias_residual", "gemm_f16_silu", "gemm_qkv_rmsnorm", "gemm_geglu", "fmha_varlen", "ln_modulate",
                                          "posenc", "head_final_euler", "adaln_table", "procrustes", "weight_pack"};
    return (kc >= 0 && kc < KC_COUNT) ? names[kc] : "";
}
