// Local feature extractor (MiniSpinNet) on the GPU: reference dataset_process/utils/spinnet/{patch_embedder,patchnet}.py and
// utils/common.py, called once per view from extract_sample_features.py:151-220.
//
//   patch_kernel        X1+X2  first-512-in-index-order ball query (pytorch3d.ops.ball_query semantics) over the shuffled cloud, empty
//                              slots = keypoint, centre = LAST slot, / des_r            (patch_embedder.py:93-139,141-166,185-188)
//   voxel_feature_kernel X3+C1 420 spherical voxels x first 10 patch points within 0.8/3, zero padding, slot-0 quirk, azimuth
//                              de-rotation, 1x1 conv 3->16 + BN + ReLU, max over the 10 samples   (patch_embedder.py:168-183,74-76;
//                              utils/common.py:396-469).  The (K,420,10,3) tensor the reference materialises (206 MB per view) never
//                              leaves the SM.
//   tcgen05 GEMM as IMPLICIT convolution (gemm.cu/gemm.cuh conv_* arguments, bias+ReLU epilogue)  the cylindrical CNN: Conv3d 16->64 (radial
//                              3->1) and 7 x Conv2d 3x3.  Maps are stored halo-padded per keypoint (8 x 22 positions: circular azimuth
//                              neighbours in the halo columns, one zero separator row = elevation padding), so each of the 9 taps is the
//                              same TMA box shifted by a constant number of rows -- no im2col operand (the first version wrote and re-read
//                              6 GB of it per view) (patchnet.py:49-84, utils/common.py:230-275); BatchNorm folded into weights / bias.
//   pool_kernel          C2    attention pooling (two 1x1 convs + BN + ReLU), weighted mean over 7x20, L2 norm  (patch_embedder.py:78-80)
#include "../../include/rap_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"

#include <cmath>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace rap {
namespace {

constexpr int PATCH = 512, NVOX = 420, RAD = 3, ELE = 7, AZI = 20, NSAMP = 10, POS = ELE * AZI;   // 140 output positions per keypoint
// Activation maps are stored halo-padded: per keypoint (ELE + 1) rows of PITCH = AZI + 2 positions (gemm.cuh, implicit convolution)
constexpr int PITCH = AZI + 2, KROWS = (ELE + 1) * PITCH;                                         // 22, 176

__constant__ float c_vox[NVOX * 3];       // voxel centres, order (rad, ele, azi)
__constant__ float c_rot[AZI * 2];        // cos / sin of -a * 18 deg
__constant__ float c_pnt[16 * 5];         // folded pnt_layer: w0 w1 w2 (x BN scale), bias (folded), per output channel (+pad)

// ---------------------------------------------------------------------------------------------------------------- X1 + X2
__global__ void gather_points_kernel(const float* __restrict__ pts, const int* __restrict__ perm, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = perm ? perm[i] : i;
    out[3 * i] = pts[3 * s]; out[3 * i + 1] = pts[3 * s + 1]; out[3 * i + 2] = pts[3 * s + 2];
}

// One warp per keypoint, 8 keypoints per block share every 512-point tile of the cloud; tiles are double-buffered with cp.async so the
// next tile's 6 KB arrive under the current tile's distance tests (the single-buffered version spent most of its 384 us per 2048 keypoints
// waiting on one L2 round trip per tile).
constexpr int kTilePts = 512;
__global__ void __launch_bounds__(256)
patch_kernel(const float* __restrict__ pts /* shuffled (N,3) */, const float* __restrict__ kpts, int N, int K, float des_r,
             float* __restrict__ patches /* (K,512,3) */, int* __restrict__ n_real /* (K) neighbours found, <= 512; may be null */) {
    __shared__ __align__(16) float tile[2][kTilePts * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + warp;
    const bool active = k < K;
    float kx = 0, ky = 0, kz = 0;
    if (active) { kx = kpts[3 * k]; ky = kpts[3 * k + 1]; kz = kpts[3 * k + 2]; }
    const float r2 = des_r * des_r;
    float* out = patches + (size_t)(active ? k : 0) * PATCH * 3;
    auto fetch = [&](int buf, int base) {                 // tile base is a multiple of 512 points = 6144 bytes: 16-byte aligned chunks
        const int floats = min(kTilePts, N - base) * 3;
        const float* src = pts + (size_t)base * 3;
        for (int i = threadIdx.x * 4; i < floats; i += 256 * 4) {
            if (i + 4 <= floats) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tile[buf][i]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + i) : "memory");
            } else {
                for (int j = i; j < floats; ++j) tile[buf][j] = src[j];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int count = 0;
    fetch(0, 0);
    for (int base = 0, buf = 0; base < N; base += kTilePts, buf ^= 1) {
        const int n_tile = min(kTilePts, N - base);
        const bool more = base + kTilePts < N;
        if (more) fetch(buf ^ 1, base + kTilePts);
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float* tl = tile[buf];
        if (active && count < PATCH) {
            for (int j = 0; j < n_tile; j += 32) {
                const int i = j + lane;
                float px = 0, py = 0, pz = 0;
                bool in = false;
                if (i < n_tile) {
                    px = tl[3 * i]; py = tl[3 * i + 1]; pz = tl[3 * i + 2];
                    const float dx = kx - px, dy = ky - py, dz = kz - pz;
                    in = dx * dx + dy * dy + dz * dz < r2;                     // strict <, like pytorch3d
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                const int pos = count + __popc(m & ((1u << lane) - 1));
                if (in && pos < PATCH) { out[3 * pos] = px; out[3 * pos + 1] = py; out[3 * pos + 2] = pz; }
                count += __popc(m);
                if (count >= PATCH) break;
            }
        }
        if (__syncthreads_and(!active || count >= PATCH)) break;      // also orders this tile's reads before its buffer is refilled
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!active) return;
    count = min(count, PATCH);
    if (n_real && lane == 0) n_real[k] = count;
    __syncwarp();
    // centre = last slot: the 512th neighbour if the patch is full, else the keypoint (patch_embedder.py:142 + :122-131)
    float cx = kx, cy = ky, cz = kz;
    if (count == PATCH) { cx = out[3 * (PATCH - 1)]; cy = out[3 * (PATCH - 1) + 1]; cz = out[3 * (PATCH - 1) + 2]; }
    __syncwarp();
    for (int i = lane; i < PATCH; i += 32) {
        float px = kx, py = ky, pz = kz;
        if (i < count) { px = out[3 * i]; py = out[3 * i + 1]; pz = out[3 * i + 2]; }
        out[3 * i] = (px - cx) / des_r; out[3 * i + 1] = (py - cy) / des_r; out[3 * i + 2] = (pz - cz) / des_r;
    }
}

// ---------------------------------------------------------------------------------------------------------------- X3 + first conv
// Block = one keypoint (patch in shared memory), thread = one spherical voxel.
__global__ void __launch_bounds__(448)
voxel_feature_kernel(const float* __restrict__ patches, const int* __restrict__ n_real, int K,
                     __half* __restrict__ feat /* (K*176, 64) halo-padded map */) {
    __shared__ float sp[PATCH * 3];
    const int k = blockIdx.x;
    // Slots behind the real neighbours of a non-full patch are copies of the keypoint = exactly (0,0,0) after centring: as samples they
    // are indistinguishable from the zero padding below, so only the real neighbours need to be scanned (typically 100-300 of 512).
    const int n_pts = n_real[k];
    for (int i = threadIdx.x; i < n_pts * 3; i += blockDim.x) sp[i] = patches[(size_t)k * PATCH * 3 + i];
    __syncthreads();
    const int v = threadIdx.x;
    if (v >= NVOX) return;
    const float cx = c_vox[3 * v], cy = c_vox[3 * v + 1], cz = c_vox[3 * v + 2];
    const float r2 = (0.8f / 3.0f) * (0.8f / 3.0f);
    float gx[NSAMP], gy[NSAMP], gz[NSAMP];
#pragma unroll
    for (int s = 0; s < NSAMP; ++s) gx[s] = gy[s] = gz[s] = 0.f;
    int cnt = 0, first = -1;
    for (int i = 0; i < n_pts && cnt < NSAMP; ++i) {
        const float px = sp[3 * i], py = sp[3 * i + 1], pz = sp[3 * i + 2];
        const float dx = cx - px, dy = cy - py, dz = cz - pz;
        if (dx * dx + dy * dy + dz * dz < r2) {
            if (cnt == 0) first = i;
#pragma unroll
            for (int s = 0; s < NSAMP; ++s)
                if (s == cnt) { gx[s] = px; gy[s] = py; gz[s] = pz; }
            ++cnt;
        }
    }
    if (first == 0) gx[0] = gy[0] = gz[0] = 0.f;                  // mask1 quirk (utils/common.py:422-424)
    const int a = v % AZI;
    const float ca = c_rot[2 * a], sa = c_rot[2 * a + 1];          // R = Rz(-a*18deg): x' = c x - s y, y' = s x + c y
    float best[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) best[o] = 0.f;                    // ReLU output >= 0, and every voxel has 10 slots (zeros included)
#pragma unroll
    for (int s = 0; s < NSAMP; ++s) {
        const float x = ca * gx[s] - sa * gy[s], y = sa * gx[s] + ca * gy[s], z = gz[s];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            const float t = fmaf(c_pnt[5 * o], x, fmaf(c_pnt[5 * o + 1], y, fmaf(c_pnt[5 * o + 2], z, c_pnt[5 * o + 3])));
            best[o] = fmaxf(best[o], t);                           // max(relu(.)) == max(0, .)
        }
    }
    // position (ele, azi) of the padded map, channels [rad*16, rad*16+16) of 64 (48 used: the radial axis is the Conv3d's depth, which
    // its 3-deep kernel collapses, so it folds into the channel dimension); azimuth 0 / 19 also fill the circular halo columns
    const int rad = v / POS, e = (v / AZI) % ELE;
    __half2 h[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) h[o] = __floats2half2_rn(best[2 * o], best[2 * o + 1]);
    __half* row = feat + ((size_t)k * KROWS + e * PITCH + a + 1) * 64 + rad * 16;
    auto put = [&](__half* d) { reinterpret_cast<uint4*>(d)[0] = *reinterpret_cast<uint4*>(&h[0]); reinterpret_cast<uint4*>(d)[1] = *reinterpret_cast<uint4*>(&h[4]); };
    put(row);
    if (a == AZI - 1) put(row - AZI * 64);
    if (a == 0) put(row + AZI * 64);
}

// ---------------------------------------------------------------------------------------------------------------- weights
// out Wp[128][9 * cpad] fp16 = W / sqrt(var+eps) re-ordered to [co][tap(ke,ka)][input channel, padded to cpad]; bp[128] fp32 = (b - mean) / sqrt(var+eps).
// Conv3d (taps = 27, source [co][ci][kd][ke][ka]): the depth index joins the channel, input channel = kd * Cin + ci.
__global__ void fold_conv_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ mean, const float* __restrict__ var,
                                 int Cout, int Cin, int taps, int cpad, __half* __restrict__ Wp, float* __restrict__ bp) {
    const int Kp = 9 * cpad;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * Kp) return;
    const int co = idx / Kp, col = idx % Kp;
    const int tap = col / cpad, ch = col % cpad, depth = taps / 9;
    float v = 0.f;
    const float inv = (co < Cout && var) ? rsqrtf(var[co] + 1e-5f) : 1.0f;
    if (co < Cout && ch < depth * Cin) {
        const int kd = ch / Cin, ci = ch % Cin;
        v = W[((size_t)co * Cin + ci) * taps + kd * 9 + tap] * inv;
    }
    Wp[idx] = __float2half_rn(v);
    if (col == 0) bp[co] = co < Cout ? (b[co] - (mean ? mean[co] : 0.f)) * inv : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------- pooling
struct PoolW { float w0[16 * 32], s0[16], t0[16], w1[16], s1, t1; };   // 1x1 convs with BN folded: y = relu(s * (W x) + t)
__constant__ PoolW c_pool;

// One warp per keypoint, lane = output position (140 positions -> 5 rounds): every lane holds its position's 32 channels in registers, the
// two 1x1 convs run as register dot products against weights broadcast from constant memory, and only the final 32 channel sums cross
// lanes (through a padded shared-memory transpose).  The first version put the channel on the lane and paid 16 five-step shuffle
// reductions per position (1.2 ms per 2048 keypoints, a third of the extractor).  x: halo-padded map (K*176, ld) fp32, first 32 columns valid.
__global__ void __launch_bounds__(128)
pool_kernel(const float* __restrict__ x, int ld, int K, float* __restrict__ desc) {
    __shared__ float part[4][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + warp;
    if (k >= K) return;                                   // whole warp
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int p = lane; p < POS; p += 32) {
        const float4* row = reinterpret_cast<const float4*>(x + ((size_t)k * KROWS + (p / AZI) * PITCH + p % AZI + 1) * ld);
        float xv[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const float4 v = row[q]; xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w; }
        float z = 0.f;
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) d = fmaf(c_pool.w0[o * 32 + c], xv[c], d);
            z = fmaf(c_pool.w1[o], fmaxf(fmaf(c_pool.s0[o], d, c_pool.t0[o]), 0.f), z);
        }
        const float w = fmaxf(fmaf(c_pool.s1, z, c_pool.t1), 0.f);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = fmaf(xv[c], w, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) part[warp][lane][c] = acc[c];
    __syncwarp();
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) sum += part[warp][l][lane];      // lane = channel from here on
    sum *= (1.0f / POS);
    float n2 = sum * sum;
#pragma unroll
    for (int s2 = 16; s2; s2 >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, s2);
    desc[(size_t)k * 32 + lane] = sum / fmaxf(sqrtf(n2), 1e-12f);
}

struct Layer { int Cin, Cout, taps, D, cpad, Kp; std::string w, b, bn; __half* Wp; float* bp; CUtensorMap tmB; };

}  // namespace
}  // namespace rap

using namespace rap;

struct rap_spinnet {
    std::map<std::string, float*> raw;          // device fp32 copies of every state_dict tensor
    std::map<std::string, int64_t> expect;
    std::set<std::string> loaded;
    std::vector<Layer> layers;
    char* slab = nullptr;
    char* ws = nullptr;
    size_t ws_bytes = 0;
    int ws_chunk = 0, ws_points = 0;
    bool finalized = false;
};

namespace {
const int kChunk = 2048;   // keypoints per pass (activation maps: 2048 * 176 rows)

void spec(rap_spinnet* h) {
    auto add = [&](const std::string& k, int64_t n) { h->expect[k] = n; };
    auto bn = [&](const std::string& p, int c, bool affine) {
        if (affine) { add(p + "weight", c); add(p + "bias", c); }
        add(p + "running_mean", c); add(p + "running_var", c);
    };
    add("pnt_layer.0.weight", 48); add("pnt_layer.0.bias", 16); bn("pnt_layer.1.", 16, true);
    add("pool_layer.0.weight", 512); add("pool_layer.0.bias", 16); bn("pool_layer.1.", 16, true);
    add("pool_layer.3.weight", 16); add("pool_layer.3.bias", 1); bn("pool_layer.4.", 1, true);
    const int cfg[8][4] = {{0, 16, 64, 27}, {3, 64, 64, 9}, {6, 64, 128, 9}, {9, 128, 128, 9}, {12, 128, 64, 9}, {15, 64, 64, 9}, {18, 64, 32, 9}, {21, 32, 32, 9}};
    for (auto& c : cfg) {
        const std::string p = "conv_net.ops." + std::to_string(c[0]) + ".";
        add(p + "weight", (int64_t)c[2] * c[1] * c[3]); add(p + "bias", c[2]);
        Layer L{};
        L.Cin = c[1]; L.Cout = c[2]; L.taps = c[3]; L.D = c[3] == 27 ? 3 : 1; L.cpad = ((c[1] * L.D + 63) / 64) * 64; L.Kp = 9 * L.cpad;
        L.w = p + "weight"; L.b = p + "bias";
        if (c[0] != 21) { L.bn = "conv_net.ops." + std::to_string(c[0] + 1) + "."; bn(L.bn, c[2], false); }
        h->layers.push_back(L);
    }
}
}  // namespace

extern "C" int rap_spinnet_create(rap_spinnet_t** out) {
    RAP_REQUIRE(out, "rap_spinnet_create: NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(RAP_ENODEV, "rap_spinnet_create: no CUDA device (this library has no CPU fallback)");
    rap_spinnet* h = new rap_spinnet();
    spec(h);
    size_t bytes = 0;
    for (auto& kv : h->expect) bytes += ((kv.second * 4 + 255) / 256) * 256;
    for (auto& L : h->layers) bytes += (size_t)128 * L.Kp * 2 + 512 + 512;
    if (cudaMalloc(&h->slab, bytes) != cudaSuccess) { delete h; return fail(RAP_ENOMEM, "rap_spinnet_create: cudaMalloc failed"); }
    char* p = h->slab;
    for (auto& kv : h->expect) { h->raw[kv.first] = reinterpret_cast<float*>(p); p += ((kv.second * 4 + 255) / 256) * 256; }
    for (auto& L : h->layers) { L.Wp = reinterpret_cast<__half*>(p); p += (size_t)128 * L.Kp * 2; L.bp = reinterpret_cast<float*>(p); p += 512; }
    *out = h;
    return RAP_OK;
}

extern "C" void rap_spinnet_destroy(rap_spinnet_t* h) {
    if (!h) return;
    cudaFree(h->slab);
    cudaFree(h->ws);
    delete h;
}

extern "C" int rap_spinnet_num_keys(const rap_spinnet_t* h) { return (int)h->expect.size(); }
extern "C" const char* rap_spinnet_key(const rap_spinnet_t* h, int index, int64_t* numel) {
    int i = 0;
    for (auto& kv : h->expect)
        if (i++ == index) { if (numel) *numel = kv.second; return kv.first.c_str(); }
    return nullptr;
}

extern "C" int rap_spinnet_load_weight(rap_spinnet_t* h, const char* key, const float* src, int64_t numel, void* stream) {
    RAP_REQUIRE(h && key && src, "rap_spinnet_load_weight: NULL argument");
    auto it = h->expect.find(key);
    RAP_REQUIRE(it != h->expect.end(), "rap_spinnet_load_weight: unexpected key '%s'", key);
    RAP_REQUIRE(it->second == numel, "rap_spinnet_load_weight: '%s' has %lld elements, expected %lld", key, (long long)numel, (long long)it->second);
    RAP_CUDA(cudaMemcpyAsync(h->raw[key], src, numel * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    h->loaded.insert(key);
    h->finalized = false;
    return RAP_OK;
}

extern "C" int rap_spinnet_finalize(rap_spinnet_t* h) {
    RAP_REQUIRE(h, "rap_spinnet_finalize: NULL");
    for (auto& kv : h->expect)
        if (!h->loaded.count(kv.first)) return fail(RAP_ESTATE, "rap_spinnet_finalize: missing key '%s'", kv.first.c_str());
    RAP_CUDA(cudaDeviceSynchronize());
    auto host = [&](const std::string& k) {
        std::vector<float> v(h->expect[k]);
        cudaMemcpy(v.data(), h->raw[k], v.size() * 4, cudaMemcpyDeviceToHost);
        return v;
    };
    // voxel grid, de-rotation table (double precision on the host, like numpy in the reference)
    std::vector<float> vox(NVOX * 3), rot(AZI * 2), pnt(16 * 5);
    for (int r = 0; r < RAD; ++r)
        for (int e = 0; e < ELE; ++e)
            for (int a = 0; a < AZI; ++a) {
                const double beta = M_PI * e / ELE + M_PI / ELE / 2, alpha = 2 * M_PI * a / AZI + M_PI / AZI, sc = (double)r / RAD + 1.0 / (2 * RAD);
                const int v = (r * ELE + e) * AZI + a;
                vox[3 * v] = (float)(sc * sin(beta) * cos(alpha)); vox[3 * v + 1] = (float)(sc * sin(beta) * sin(alpha)); vox[3 * v + 2] = (float)(sc * cos(beta));
            }
    for (int a = 0; a < AZI; ++a) { const double ang = -a * (2 * M_PI / AZI); rot[2 * a] = (float)cos(ang); rot[2 * a + 1] = (float)sin(ang); }
    {
        auto W = host("pnt_layer.0.weight"), b = host("pnt_layer.0.bias"), g = host("pnt_layer.1.weight"), be = host("pnt_layer.1.bias"),
             mu = host("pnt_layer.1.running_mean"), va = host("pnt_layer.1.running_var");
        for (int o = 0; o < 16; ++o) {
            const float s = g[o] / sqrtf(va[o] + 1e-5f);
            for (int c = 0; c < 3; ++c) pnt[5 * o + c] = W[3 * o + c] * s;
            pnt[5 * o + 3] = (b[o] - mu[o]) * s + be[o];
            pnt[5 * o + 4] = 0.f;
        }
    }
    PoolW pw{};
    {
        auto W0 = host("pool_layer.0.weight"), b0 = host("pool_layer.0.bias"), g0 = host("pool_layer.1.weight"), be0 = host("pool_layer.1.bias"),
             mu0 = host("pool_layer.1.running_mean"), va0 = host("pool_layer.1.running_var");
        auto W1 = host("pool_layer.3.weight"), b1 = host("pool_layer.3.bias"), g1 = host("pool_layer.4.weight"), be1 = host("pool_layer.4.bias"),
             mu1 = host("pool_layer.4.running_mean"), va1 = host("pool_layer.4.running_var");
        for (int i = 0; i < 512; ++i) pw.w0[i] = W0[i];
        for (int o = 0; o < 16; ++o) { const float s = g0[o] / sqrtf(va0[o] + 1e-5f); pw.s0[o] = s; pw.t0[o] = (b0[o] - mu0[o]) * s + be0[o]; pw.w1[o] = W1[o]; }
        const float s = g1[0] / sqrtf(va1[0] + 1e-5f);
        pw.s1 = s; pw.t1 = (b1[0] - mu1[0]) * s + be1[0];
    }
    RAP_CUDA(cudaMemcpyToSymbol(c_vox, vox.data(), vox.size() * 4));
    RAP_CUDA(cudaMemcpyToSymbol(c_rot, rot.data(), rot.size() * 4));
    RAP_CUDA(cudaMemcpyToSymbol(c_pnt, pnt.data(), pnt.size() * 4));
    RAP_CUDA(cudaMemcpyToSymbol(c_pool, &pw, sizeof pw));
    for (auto& L : h->layers) {
        const float* mean = L.bn.empty() ? nullptr : h->raw[L.bn + "running_mean"];
        const float* var = L.bn.empty() ? nullptr : h->raw[L.bn + "running_var"];
        fold_conv_kernel<<<ceil_div(128 * L.Kp, 256), 256>>>(h->raw[L.w], h->raw[L.b], mean, var, L.Cout, L.Cin, L.taps, L.cpad, L.Wp, L.bp);
        RAP_CUDA(cudaGetLastError());
        RAP_TRY(make_tmap_f16_sw128(&L.tmB, L.Wp, 128, L.Kp, L.Kp, 128));
    }
    RAP_CUDA(cudaDeviceSynchronize());
    h->finalized = true;
    return RAP_OK;
}

extern "C" int rap_spinnet_extract(rap_spinnet_t* h, const float* points, int n_points, const int32_t* perm, const float* keypoints, int n_keypoints,
                                   float des_r, float* desc_out, void* stream) {
    RAP_REQUIRE(h && points && keypoints && desc_out, "rap_spinnet_extract: NULL argument");
    RAP_REQUIRE(h->finalized, "rap_spinnet_extract: call rap_spinnet_finalize first");
    RAP_REQUIRE(n_points > 0 && n_keypoints > 0 && des_r > 0.f, "rap_spinnet_extract: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int chunk = n_keypoints < kChunk ? n_keypoints : kChunk;
    if (chunk > h->ws_chunk || n_points > h->ws_points) {       // (re)allocate the workspace; not on the steady-state path
        RAP_CUDA(cudaStreamSynchronize(st));
        cudaFree(h->ws);
        h->ws = nullptr;
        const size_t rows = (size_t)chunk * KROWS;
        h->ws_bytes = (size_t)n_points * 12 + (size_t)chunk * PATCH * 12 + (size_t)chunk * 4 + 256 + rows * 64 * 2 + 2 * rows * 128 * 2 + rows * 128 * 4 + 8192;
        if (cudaMalloc(&h->ws, h->ws_bytes) != cudaSuccess) { h->ws_chunk = h->ws_points = 0; return fail(RAP_ENOMEM, "rap_spinnet_extract: cudaMalloc(%zu) failed", h->ws_bytes); }
        RAP_CUDA(cudaMemsetAsync(h->ws, 0, h->ws_bytes, st));       // separator rows / padded channels of the maps are never written again
        h->ws_chunk = chunk; h->ws_points = n_points;
    }
    char* p = h->ws;
    auto take = [&](size_t n) { char* q = p; p += (n + 255) / 256 * 256; return q; };
    const size_t rows_cap = (size_t)h->ws_chunk * KROWS;
    float* pts_s = reinterpret_cast<float*>(take((size_t)h->ws_points * 12));
    float* patches = reinterpret_cast<float*>(take((size_t)h->ws_chunk * PATCH * 12));
    int* n_real = reinterpret_cast<int*>(take((size_t)h->ws_chunk * 4));
    __half* feat0 = reinterpret_cast<__half*>(take(rows_cap * 64 * 2));
    __half* fm[2] = {reinterpret_cast<__half*>(take(rows_cap * 128 * 2)), reinterpret_cast<__half*>(take(rows_cap * 128 * 2))};
    float* out32 = reinterpret_cast<float*>(take(rows_cap * 128 * 4));

    { ProfScope ps(KC_POSENC, st); gather_points_kernel<<<ceil_div(n_points, 256), 256, 0, st>>>(points, perm, pts_s, n_points); note_launch(); }
    RAP_CUDA(cudaGetLastError());
    for (int k0 = 0; k0 < n_keypoints; k0 += chunk) {
        const int kc = n_keypoints - k0 < chunk ? n_keypoints - k0 : chunk;
        const int M = kc * KROWS;
        patch_kernel<<<ceil_div(kc, 8), 256, 0, st>>>(pts_s, keypoints + (size_t)k0 * 3, n_points, kc, des_r, patches, n_real);
        RAP_CUDA(cudaGetLastError());
        voxel_feature_kernel<<<kc, 448, 0, st>>>(patches, n_real, kc, feat0);
        RAP_CUDA(cudaGetLastError());
        note_launch(2);
        const __half* in = feat0;
        int ld_in = 64;
        for (size_t li = 0; li < h->layers.size(); ++li) {
            Layer& L = h->layers[li];
            CUtensorMap tmA;                                     // the activation map itself: every tap is a shifted box of it
            RAP_TRY(make_tmap_f16_sw128(&tmA, in, M, ld_in, ld_in, 128));
            GemmArgs g{};
            g.M = M; g.N = 128; g.K = L.Kp; g.bias = L.bp;
            g.conv_pitch = PITCH; g.conv_rows = KROWS; g.conv_h = ELE; g.conv_w = AZI; g.conv_chunks = L.cpad / 64;
            const bool last = li + 1 == h->layers.size();
            if (last) { g.out32 = out32; g.ld32 = 128; RAP_TRY(launch_gemm(EPI_F32, 128, tmA, L.tmB, g, st)); }
            else { g.out16 = fm[li & 1]; g.ld16 = 128; g.act = 1; RAP_TRY(launch_gemm(EPI_F16_SILU, 128, tmA, L.tmB, g, st)); }
            in = fm[li & 1];
            ld_in = 128;
        }
        pool_kernel<<<ceil_div(kc, 4), 128, 0, st>>>(out32, 128, kc, desc_out + (size_t)k0 * 32);
        RAP_CUDA(cudaGetLastError());
        note_launch();
    }
    return RAP_OK;
}

// Test entry: the normalised 512-slot patches (index work -- compared bit-exactly with the reference/oracle).
extern "C" int rap_test_spinnet_patches(const float* points, int n_points, const int32_t* perm, const float* keypoints, int n_keypoints, float des_r,
                                        float* patches_out, void* stream) {
    RAP_REQUIRE(points && keypoints && patches_out && n_points > 0 && n_keypoints > 0, "rap_test_spinnet_patches: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* tmp = nullptr;
    RAP_CUDA(cudaMalloc(&tmp, (size_t)n_points * 12));
    gather_points_kernel<<<ceil_div(n_points, 256), 256, 0, st>>>(points, perm, tmp, n_points);
    patch_kernel<<<ceil_div(n_keypoints, 8), 256, 0, st>>>(tmp, keypoints, n_points, n_keypoints, des_r, patches_out, nullptr);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(RAP_ECUDA, "rap_test_spinnet_patches: %s", cudaGetErrorString(e));
    return RAP_OK;
}
