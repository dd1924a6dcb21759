// Keypoint selection in front of the feature extractor: voxel down-sampling, occupied-voxel count,
// statistical outlier removal and batched farthest point sampling.  HBM/latency-bound integer and fp32 work -- no tensor cores.
//
//   rap_voxel_downsample     <- dataset_process/utils/dataset_utils.py:279-322 (voxel_down_sample_torch)
//   rap_voxel_coverage       <- dataset_process/utils/point_sampling_utils.py:11-31 (calculate_voxel_coverage)
//   rap_fps_batched          <- pytorch3d sample_farthest_points as called at point_sampling_utils.py:263-309
//   rap_statistical_outliers <- open3d remove_statistical_outlier as called at dataset_process/extract_sample_features.py:360-413
//
// Index results are compared bit-exactly with the CPU oracle, so every fp32 expression below is written with explicit
// round-to-nearest intrinsics in the reference's operation order (no fused multiply-add contraction).
// The two key sorts and the run-head compaction use CUB (CUDA toolkit library code); everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <cfloat>
#include <cmath>
#include <vector>

#include "common.h"

using namespace rap;

namespace {

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }   // order-preserving
__host__ __device__ __forceinline__ float ord2f(int i) {
    const int b = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

struct Temp {                                       // stream-ordered scratch, released when the struct goes out of scope
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Temp(cudaStream_t s) : st(s) {}
    template <typename T> T* get(size_t n) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    ~Temp() { for (void* p : ptrs) cudaFreeAsync(p, st); }
};

// ------------------------------------------------------------------------------------------------ voxel down-sampling
// stats[0..2] = min xyz, stats[3..5] = max xyz, stats[6] = max distance to the voxel centre (all as order-preserving ints)
__device__ __forceinline__ float centre_dist(float x, float y, float z, float vs) {
    const float gx = floorf(__fdiv_rn(x, vs)), gy = floorf(__fdiv_rn(y, vs)), gz = floorf(__fdiv_rn(z, vs));
    const float dx = __fsub_rn(x, __fmul_rn(__fadd_rn(gx, 0.5f), vs));
    const float dy = __fsub_rn(y, __fmul_rn(__fadd_rn(gy, 0.5f), vs));
    const float dz = __fsub_rn(z, __fmul_rn(__fadd_rn(gz, 0.5f), vs));
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

__global__ void voxel_stats_kernel(const float* __restrict__ p, int n, float vs, int* __restrict__ stats) {
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN}, md = INT_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
        mn[0] = min(mn[0], f2ord(x)); mn[1] = min(mn[1], f2ord(y)); mn[2] = min(mn[2], f2ord(z));
        mx[0] = max(mx[0], f2ord(x)); mx[1] = max(mx[1], f2ord(y)); mx[2] = max(mx[2], f2ord(z));
        md = max(md, f2ord(centre_dist(x, y, z, vs)));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = __reduce_min_sync(0xffffffffu, mn[k]); mx[k] = __reduce_max_sync(0xffffffffu, mx[k]); }
    md = __reduce_max_sync(0xffffffffu, md);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(stats + k, mn[k]); atomicMax(stats + 3 + k, mx[k]); }
        atomicMax(stats + 6, md);
    }
}

// gid = flattened voxel id with the reference's stride v = grid.max(); key = index + quantised distance * dec
__global__ void voxel_keys_kernel(const float* __restrict__ p, int n, float vs, float dmax, long long ox, long long oy, long long oz,
                                  long long v, long long dec, long long* __restrict__ gid, long long* __restrict__ key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    const long long gx = (long long)floorf(__fdiv_rn(x, vs)) - ox, gy = (long long)floorf(__fdiv_rn(y, vs)) - oy,
                    gz = (long long)floorf(__fdiv_rn(z, vs)) - oz;
    const long long q = (long long)__fmul_rn(__fdiv_rn(centre_dist(x, y, z, vs), dmax), 999.0f);
    gid[i] = gx + gy * v + gz * v * v;
    key[i] = (long long)i + q * dec;
}

__global__ void mod_kernel(const long long* __restrict__ key, long long dec, int n, long long* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = key[i] % dec;
}

int bits_for(unsigned long long max_value) { int b = 1; while (b < 64 && (max_value >> b)) ++b; return b; }

// occupied-voxel count: 3 x 21-bit cell coordinates relative to the minimum corner
template <typename T>
__global__ void coverage_minmax_kernel(const T* __restrict__ p, int n, T vs, long long* __restrict__ mm) {
    long long mn[3] = {LLONG_MAX, LLONG_MAX, LLONG_MAX}, mx[3] = {LLONG_MIN, LLONG_MIN, LLONG_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const long long g = (long long)floor(p[3 * i + k] / vs);     // IEEE division in the array's own precision
            mn[k] = min(mn[k], g); mx[k] = max(mx[k], g);
        }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o; o >>= 1) { mn[k] = min(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o)); mx[k] = max(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(mm + k, mn[k]); atomicMax(mm + 3 + k, mx[k]); }
    }
}
template <typename T>
__global__ void coverage_keys_kernel(const T* __restrict__ p, int n, T vs, long long ox, long long oy, long long oz,
                                     unsigned long long* __restrict__ key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long gx = (long long)floor(p[3 * i] / vs) - ox, gy = (long long)floor(p[3 * i + 1] / vs) - oy, gz = (long long)floor(p[3 * i + 2] / vs) - oz;
    key[i] = gx | (gy << 21) | (gz << 42);
}
__global__ void count_heads_kernel(const unsigned long long* __restrict__ key, int n, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool head = i < n && (i == 0 || key[i] != key[i - 1]);
    const unsigned b = __ballot_sync(0xffffffffu, head);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// ------------------------------------------------------------------------------------------------ farthest point sampling
// One thread-block cluster per cloud.  Every point lives in the registers of exactly one thread (coordinates + running minimum
// distance); a step is: distance update + per-thread arg-max (registers), two redux instructions per warp, one __syncthreads, and
// for clusters one DSMEM slot exchange behind a cluster barrier.  Ties go to the lowest index, already-picked points carry -1.
constexpr int kFpsThreads = 512;
constexpr int kFpsMaxCluster = 16;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t local_addr, uint32_t cta, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(remote), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct FpsSlot { int dist_bits; int index; float x, y; float z; int pad[3]; };   // 32 B: two v4 stores

template <int PPT, bool COORD_REGS>
__global__ void __launch_bounds__(kFpsThreads)
fps_kernel(const float* __restrict__ pts, int P, const int* __restrict__ lengths, const int* __restrict__ K, const int* __restrict__ start,
           int max_k, int C, long long* __restrict__ idx_out, float* __restrict__ xyz_out) {
    extern __shared__ float s_pts[];                       // this CTA's points, SoA: x[PPT*512] y[...] z[...]
    __shared__ int s_wd[2][kFpsThreads / 32], s_wi[2][kFpsThreads / 32];
    __shared__ __align__(16) FpsSlot s_slot[2][kFpsMaxCluster];
    const int cloud = blockIdx.x / C;
    const int rank = (C > 1) ? (int)cluster_rank() : 0;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int L = lengths[cloud];
    const int k_n = min(K[cloud], L);
    const float* p = pts + (size_t)cloud * P * 3;
    long long* out = idx_out + (size_t)cloud * max_k;
    float* out_xyz = xyz_out ? xyz_out + (size_t)cloud * max_k * 3 : nullptr;
    if (rank == 0) for (int k = max(k_n, 0) + t; k < max_k; k += kFpsThreads) {      // -1 / zero padding like masked_gather
        out[k] = -1;
        if (out_xyz) { out_xyz[3 * k] = 0.f; out_xyz[3 * k + 1] = 0.f; out_xyz[3 * k + 2] = 0.f; }
    }

    constexpr int NREG = COORD_REGS ? PPT : 1;              // wide clouds keep only the running minimum in registers
    float px[NREG], py[NREG], pz[NREG], md[PPT];
    const int stride = C * kFpsThreads, base = rank * kFpsThreads + t;          // point index of slot j: base + j * stride
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = base + j * stride;
        const bool ok = i < L;
        const float x = ok ? p[3 * i] : 0.f, y = ok ? p[3 * i + 1] : 0.f, z = ok ? p[3 * i + 2] : 0.f;
        if (COORD_REGS) { px[j] = x; py[j] = y; pz[j] = z; }
        md[j] = ok ? FLT_MAX : -1.0f;
        s_pts[j * kFpsThreads + t] = x;
        s_pts[(PPT + j) * kFpsThreads + t] = y;
        s_pts[(2 * PPT + j) * kFpsThreads + t] = z;
    }
    if (k_n <= 0) { if (C > 1) cluster_sync_all(); return; }
    int last = start[cloud];
    float lx = p[3 * last], ly = p[3 * last + 1], lz = p[3 * last + 2];
    if (rank == 0 && t == 0) { out[0] = last; if (out_xyz) { out_xyz[0] = lx; out_xyz[1] = ly; out_xyz[2] = lz; } }
    __syncthreads();
    if (C > 1) cluster_sync_all();                          // every CTA of the cluster is resident before remote stores start

    for (int k = 1; k < k_n; ++k) {
        const int buf = k & 1;
        float best = -2.0f; int best_i = INT_MAX;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int i = base + j * stride;
            const float cx = COORD_REGS ? px[j] : s_pts[j * kFpsThreads + t], cy = COORD_REGS ? py[j] : s_pts[(PPT + j) * kFpsThreads + t],
                        cz = COORD_REGS ? pz[j] : s_pts[(2 * PPT + j) * kFpsThreads + t];
            const float dx = __fsub_rn(lx, cx), dy = __fsub_rn(ly, cy), dz = __fsub_rn(lz, cz);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            float m = md[j];
            m = (m < 0.f || i == last) ? -1.0f : fminf(m, d2);       // picked (or padding) points never compete again
            md[j] = m;
            if (m > best) { best = m; best_i = i; }                  // j ascending = index ascending: strict > keeps the first maximum
        }
        const int db = __float_as_int(best);                         // best >= -2: int order == float order on {-2,-1} U [0,inf)
        const int wmax = __reduce_max_sync(0xffffffffu, db);
        const int wi = __reduce_min_sync(0xffffffffu, db == wmax ? best_i : INT_MAX);
        if (lane == 0) { s_wd[buf][warp] = wmax; s_wi[buf][warp] = wi; }
        __syncthreads();
        int bd = lane < kFpsThreads / 32 ? s_wd[buf][lane] : INT_MIN, bi = lane < kFpsThreads / 32 ? s_wi[buf][lane] : INT_MAX;
        const int cmax = __reduce_max_sync(0xffffffffu, bd);
        bi = __reduce_min_sync(0xffffffffu, bd == cmax ? bi : INT_MAX);
        bd = cmax;
        if (C == 1) {
            last = bi;
            const int j = last / kFpsThreads, tt = last % kFpsThreads;     // C == 1: index = j * 512 + t
            lx = s_pts[j * kFpsThreads + tt]; ly = s_pts[(PPT + j) * kFpsThreads + tt]; lz = s_pts[(2 * PPT + j) * kFpsThreads + tt];
        } else {
            if (t < C) {                                                   // thread c posts this CTA's candidate to CTA c
                float x = 0.f, y = 0.f, z = 0.f;
                if (bi != INT_MAX) {
                    const int j = bi / stride, tt = bi % kFpsThreads;
                    x = s_pts[j * kFpsThreads + tt]; y = s_pts[(PPT + j) * kFpsThreads + tt]; z = s_pts[(2 * PPT + j) * kFpsThreads + tt];
                }
                const uint32_t a = (uint32_t)__cvta_generic_to_shared(&s_slot[buf][rank]);
                st_cluster_v4(a, t, (uint32_t)bd, (uint32_t)bi, __float_as_uint(x), __float_as_uint(y));
                st_cluster_v4(a + 16, t, __float_as_uint(z), 0u, 0u, 0u);
            }
            cluster_sync_all();
            int sd = lane < C ? s_slot[buf][lane].dist_bits : INT_MIN, si = lane < C ? s_slot[buf][lane].index : INT_MAX;
            const int gmax = __reduce_max_sync(0xffffffffu, sd);
            si = __reduce_min_sync(0xffffffffu, sd == gmax ? si : INT_MAX);
            last = si;
            const int owner = (last / kFpsThreads) % C;
            lx = s_slot[buf][owner].x; ly = s_slot[buf][owner].y; lz = s_slot[buf][owner].z;
        }
        if (rank == 0 && t == 0) { out[k] = last; if (out_xyz) { out_xyz[3 * k] = lx; out_xyz[3 * k + 1] = ly; out_xyz[3 * k + 2] = lz; } }
    }
    if (C > 1) cluster_sync_all();                          // no CTA exits while a peer may still write its slots
}

template <int PPT, bool COORD_REGS>
int launch_fps(const float* pts, int N, int P, const int* lengths, const int* K, const int* start, int max_k, int C, long long* idx_out,
               float* xyz_out, cudaStream_t st) {
    const size_t smem = (size_t)3 * PPT * kFpsThreads * sizeof(float);
    RAP_CUDA(cudaFuncSetAttribute(fps_kernel<PPT, COORD_REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) RAP_CUDA(cudaFuncSetAttribute(fps_kernel<PPT, COORD_REGS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(N * C); cfg.blockDim = dim3(kFpsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    RAP_CUDA(cudaLaunchKernelEx(&cfg, fps_kernel<PPT, COORD_REGS>, pts, P, lengths, K, start, max_k, C, idx_out, xyz_out));
    note_launch();
    return RAP_OK;
}

// ------------------------------------------------------------------------------------------------ statistical outlier removal
// Thread = query, all points streamed through shared memory in tiles, the KNN smallest squared distances (fp64, self included)
// kept sorted in registers; a candidate only enters the insertion network when it beats the current k-th distance.
constexpr int kKnnMax = 32, kKnnThreads = 128, kKnnTile = 1024;

template <int KMAX>
__global__ void __launch_bounds__(kKnnThreads) knn_mean_kernel(const float* __restrict__ p, int n, int k, double* __restrict__ avg) {
    __shared__ float tile[kKnnTile * 3];
    const int q = blockIdx.x * kKnnThreads + threadIdx.x;
    const bool active = q < n;
    const double qx = active ? p[3 * q] : 0.0, qy = active ? p[3 * q + 1] : 0.0, qz = active ? p[3 * q + 2] : 0.0;
    double best[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) best[j] = j < k ? DBL_MAX : -1.0;     // slots >= k are never beaten (d2 >= 0 > -1)
    double kth = DBL_MAX;
    for (int t0 = 0; t0 < n; t0 += kKnnTile) {
        const int cnt = min(kKnnTile, n - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 3; i += kKnnThreads) tile[i] = p[(size_t)t0 * 3 + i];
        __syncthreads();
        if (!active) continue;
        for (int i = 0; i < cnt; ++i) {
            const double dx = qx - (double)tile[3 * i], dy = qy - (double)tile[3 * i + 1], dz = qz - (double)tile[3 * i + 2];
            double d = dx * dx + dy * dy + dz * dz;
            if (d < kth) {
#pragma unroll
                for (int j = 0; j < KMAX; ++j)                          // bubble into the ascending list
                    if (d < best[j]) { const double tmp = best[j]; best[j] = d; d = tmp; }
                kth = -1.0;
#pragma unroll
                for (int j = 0; j < KMAX; ++j) kth = fmax(kth, best[j]);
            }
        }
    }
    if (active) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) if (j < k) s += sqrt(best[j]);      // ascending order, like std::accumulate over the KNN result
        avg[q] = s / (double)k;
    }
}

__global__ void moments_kernel(const double* __restrict__ avg, int n, double mean, double* __restrict__ acc /* [0]=sum or sq sum */, int pass) {
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double a = avg[i];
        s += pass == 0 ? a : (a - mean) * (a - mean);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}

__global__ void threshold_kernel(const double* __restrict__ avg, int n, double thr, uint8_t* __restrict__ keep, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool k = i < n && avg[i] > 0.0 && avg[i] < thr;
    if (i < n) keep[i] = k;
    const unsigned b = __ballot_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// ------------------------------------------------------------------------------------------------ evaluation: nearest neighbours
// out[i] = min over targets j of the SAME segment (cu_seqlens) with label_t[j] != label_q[i] (when labels are given) of |q_i - t_j|^2,
// +inf if there is none.  Thread = query; the targets are read straight from L1/L2 (every lane of a warp reads the same address).
// Serves the chamfer distance (eval/metrics.py:14-48, no labels, both directions) and the overlap ratio (:625-691, labels = part ids).
__global__ void __launch_bounds__(256)
segment_nn_kernel(const float* __restrict__ q, const float* __restrict__ t, const int* __restrict__ cu, int n_seg, const int* __restrict__ label_q,
                  const int* __restrict__ label_t, int n_total, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    int lo = 0, hi = n_seg;                                  // segment of query i: cu[lo] <= i < cu[lo + 1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cu[mid] <= i) lo = mid; else hi = mid; }
    const int s = cu[lo], e = cu[lo + 1];
    const float x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
    const int lab = label_q ? label_q[i] : -1;
    float best = INFINITY;
    for (int j = s; j < e; ++j) {
        const float dx = x - t[3 * j], dy = y - t[3 * j + 1], dz = z - t[3 * j + 2];
        const float d = dx * dx + dy * dy + dz * dz;
        if ((!label_t || label_t[j] != lab) && d < best) best = d;
    }
    out[i] = best;
}

// ------------------------------------------------------------------------------------------------ pose application
// out = R p + t for every point of a full-resolution cloud (demo.py:1338-1340, open3d PointCloud.transform): 24 B/point of traffic.
struct Rigid { double r[9], t[3]; };
template <typename T>
__global__ void apply_transform_kernel(const T* __restrict__ in, T* __restrict__ out, long long n, Rigid m) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
        out[3 * i] = (T)((T)m.r[0] * x + (T)m.r[1] * y + (T)m.r[2] * z + (T)m.t[0]);
        out[3 * i + 1] = (T)((T)m.r[3] * x + (T)m.r[4] * y + (T)m.r[5] * z + (T)m.t[1]);
        out[3 * i + 2] = (T)((T)m.r[6] * x + (T)m.r[7] * y + (T)m.r[8] * z + (T)m.t[2]);
    }
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int rap_segment_nn(const float* query, const float* target, const int32_t* cu_seqlens, int n_segments, const int32_t* label_query,
                              const int32_t* label_target, float* min_dist2_out, void* stream) {
    RAP_REQUIRE(query && target && cu_seqlens && min_dist2_out && n_segments >= 1, "rap_segment_nn: bad argument");
    RAP_REQUIRE((label_query == nullptr) == (label_target == nullptr), "rap_segment_nn: give both label arrays or neither");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int n_total = 0;
    RAP_CUDA(cudaMemcpyAsync(&n_total, cu_seqlens + n_segments, sizeof(int), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    if (n_total <= 0) return RAP_OK;
    segment_nn_kernel<<<ceil_div(n_total, 256), 256, 0, st>>>(query, target, cu_seqlens, n_segments, label_query, label_target, n_total, min_dist2_out);
    RAP_CUDA(cudaGetLastError());
    note_launch();
    return RAP_OK;
}

extern "C" int rap_apply_transform(const void* points, long long n, int is_f64, const double* transform_4x4_host, void* out, void* stream) {
    RAP_REQUIRE(points && out && transform_4x4_host && n >= 0, "rap_apply_transform: bad argument");
    const double* T = transform_4x4_host;
    RAP_REQUIRE(T[12] == 0.0 && T[13] == 0.0 && T[14] == 0.0 && T[15] == 1.0, "rap_apply_transform: not an affine 4x4 (last row must be 0 0 0 1)");
    if (n == 0) return RAP_OK;
    Rigid m;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m.r[r * 3 + c] = T[r * 4 + c]; m.t[r] = T[r * 4 + 3]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (int)std::min<long long>((n + 255) / 256, 8ll * num_sms());
    if (is_f64) apply_transform_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(points), static_cast<double*>(out), n, m);
    else apply_transform_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(points), static_cast<float*>(out), n, m);
    RAP_CUDA(cudaGetLastError());
    note_launch();
    return RAP_OK;
}

extern "C" int rap_voxel_downsample(const float* points, int n, float voxel_size, int64_t* indices_out, int* count_out, void* stream) {
    RAP_REQUIRE(points && indices_out && count_out && n > 0 && voxel_size > 0.f, "rap_voxel_downsample: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Temp tmp(st);
    int* stats = tmp.get<int>(8);
    long long *gid = tmp.get<long long>(n), *key = tmp.get<long long>(n), *gid2 = tmp.get<long long>(n), *key2 = tmp.get<long long>(n);
    long long* sel = tmp.get<long long>(n);
    int* d_count = tmp.get<int>(1);
    RAP_REQUIRE(stats && gid && key && gid2 && key2 && sel && d_count, "rap_voxel_downsample: out of device memory");
    const int init[8] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, INT_MIN, 0};
    RAP_CUDA(cudaMemcpyAsync(stats, init, sizeof(init), cudaMemcpyHostToDevice, st));
    voxel_stats_kernel<<<std::min(ceil_div(n, 256), 4 * num_sms()), 256, 0, st>>>(points, n, voxel_size, stats);
    int h[8];
    RAP_CUDA(cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    long long off[3], v = 0;
    for (int k = 0; k < 3; ++k) {
        off[k] = (long long)std::floor(ord2f(h[k]) / voxel_size);                           // IEEE fp32 division, like the device
        v = std::max(v, (long long)std::floor(ord2f(h[3 + k]) / voxel_size) - off[k]);
    }
    const float dmax = ord2f(h[6]);
    long long dec = 10;
    for (long long m = n - 1; m >= 10; m /= 10) dec *= 10;                                   // 10 ** len(str(n - 1))
    RAP_REQUIRE(v < 2000000, "rap_voxel_downsample: %lld voxels per axis overflow the 64-bit voxel id", v);
    voxel_keys_kernel<<<ceil_div(n, 256), 256, 0, st>>>(points, n, voxel_size, dmax, off[0], off[1], off[2], v, dec, gid, key);
    // stable LSD order: by (quantised distance, index) first, then by voxel id -> the head of every voxel run is the reference's amin
    const int key_bits = bits_for((unsigned long long)(1000 * dec)), gid_bits = bits_for((unsigned long long)((v + 1) * (v * v + v + 1)));
    size_t ws1 = 0, ws2 = 0, ws3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, ws1, key, key2, gid, gid2, n, 0, key_bits, st);
    cub::DeviceRadixSort::SortPairs(nullptr, ws2, gid2, gid, key2, key, n, 0, gid_bits, st);
    cub::DeviceSelect::UniqueByKey(nullptr, ws3, gid, key, gid2, sel, d_count, n, st);
    void* ws = tmp.get<char>(std::max(ws1, std::max(ws2, ws3)));
    RAP_REQUIRE(ws, "rap_voxel_downsample: out of device memory");
    RAP_CUDA(cub::DeviceRadixSort::SortPairs(ws, ws1, key, key2, gid, gid2, n, 0, key_bits, st));
    RAP_CUDA(cub::DeviceRadixSort::SortPairs(ws, ws2, gid2, gid, key2, key, n, 0, gid_bits, st));
    RAP_CUDA(cub::DeviceSelect::UniqueByKey(ws, ws3, gid, key, gid2, sel, d_count, n, st));
    RAP_CUDA(cudaMemcpyAsync(count_out, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    mod_kernel<<<ceil_div(*count_out, 256), 256, 0, st>>>(sel, dec, *count_out, reinterpret_cast<long long*>(indices_out));
    RAP_CUDA(cudaGetLastError());
    note_launch(9);
    RAP_CUDA(cudaStreamSynchronize(st));
    return RAP_OK;
}

template <typename T>
static int voxel_coverage_impl(const T* points, int n, T vs, int* count_out, cudaStream_t st) {
    Temp tmp(st);
    long long* mm = tmp.get<long long>(6);
    unsigned long long *key = tmp.get<unsigned long long>(n), *key2 = tmp.get<unsigned long long>(n);
    int* d_count = tmp.get<int>(1);
    RAP_REQUIRE(mm && key && key2 && d_count, "rap_voxel_coverage: out of device memory");
    const long long init[6] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};
    RAP_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    RAP_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
    coverage_minmax_kernel<T><<<std::min(ceil_div(n, 256), 4 * num_sms()), 256, 0, st>>>(points, n, vs, mm);
    long long h[6];
    RAP_CUDA(cudaMemcpyAsync(h, mm, sizeof(h), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) RAP_REQUIRE(h[3 + k] - h[k] < (1ll << 21), "rap_voxel_coverage: more than 2^21 voxels along an axis");
    coverage_keys_kernel<T><<<ceil_div(n, 256), 256, 0, st>>>(points, n, vs, h[0], h[1], h[2], key);
    size_t wsn = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, wsn, key, key2, n, 0, 63, st);
    void* ws = tmp.get<char>(wsn);
    RAP_REQUIRE(ws, "rap_voxel_coverage: out of device memory");
    RAP_CUDA(cub::DeviceRadixSort::SortKeys(ws, wsn, key, key2, n, 0, 63, st));
    count_heads_kernel<<<ceil_div(n, 256), 256, 0, st>>>(key2, n, d_count);
    RAP_CUDA(cudaMemcpyAsync(count_out, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    note_launch(4);
    return RAP_OK;
}

extern "C" int rap_voxel_coverage(const void* points, int n, int is_f64, double voxel_size, int* count_out, void* stream) {
    RAP_REQUIRE(count_out && n >= 0 && voxel_size > 0.0, "rap_voxel_coverage: bad argument");
    if (n == 0) { *count_out = 0; return RAP_OK; }
    RAP_REQUIRE(points, "rap_voxel_coverage: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return is_f64 ? voxel_coverage_impl<double>(static_cast<const double*>(points), n, voxel_size, count_out, st)
                  : voxel_coverage_impl<float>(static_cast<const float*>(points), n, (float)voxel_size, count_out, st);
}

extern "C" int rap_fps_capacity(void) { return kFpsMaxCluster * kFpsThreads * 36; }

extern "C" int rap_fps_batched(const float* points, int n_clouds, int max_points, const int32_t* lengths, const int32_t* k, const int32_t* start_idx,
                               int max_k, int64_t* indices_out, float* sampled_out, void* stream) {
    RAP_REQUIRE(points && lengths && k && start_idx && indices_out && n_clouds > 0 && max_points > 0 && max_k > 0, "rap_fps_batched: bad argument");
    RAP_REQUIRE(max_points <= rap_fps_capacity(), "rap_fps_batched: %d points per cloud exceed the register-resident capacity (%d); pre-sample first "
                "(the reference caps at 20 x max_points_per_part, extract_sample_features.py:417-437)", max_points, rap_fps_capacity());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long* out = reinterpret_cast<long long*>(indices_out);
    // smallest footprint that holds the longest cloud: coordinates in registers (8 or 20 points per thread) up to 8 CTAs, beyond
    // that 36 points per thread with the coordinates in shared memory
    if (max_points <= 8 * kFpsThreads * 8) {
        int C = 1;
        while (C * kFpsThreads * 8 < max_points) C *= 2;
        return launch_fps<8, true>(points, n_clouds, max_points, lengths, k, start_idx, max_k, C, out, sampled_out, st);
    }
    if (max_points <= 8 * kFpsThreads * 20) return launch_fps<20, true>(points, n_clouds, max_points, lengths, k, start_idx, max_k, 8, out, sampled_out, st);
    // > 81920 points: 16-CTA clusters (opt-in size) keep the coordinates in registers; if this device cannot co-schedule 16 CTAs of this
    // footprint the launch fails cleanly and the shared-memory variant (36 points per thread, 8 CTAs) takes over.
    static int wide16 = -1;                               // RAP_FPS_CLUSTER16=0 disables (A/B measurements)
    if (wide16 < 0) { const char* e = getenv("RAP_FPS_CLUSTER16"); wide16 = (e && atoi(e) == 0) ? 0 : 1; }
    if (wide16 && max_points <= 16 * kFpsThreads * 20) {
        if (launch_fps<20, true>(points, n_clouds, max_points, lengths, k, start_idx, max_k, 16, out, sampled_out, st) == RAP_OK) return RAP_OK;
        cudaGetLastError();
        wide16 = 0;
    }
    return launch_fps<36, false>(points, n_clouds, max_points, lengths, k, start_idx, max_k, max_points <= 8 * kFpsThreads * 36 ? 8 : 16, out,
                                 sampled_out, st);
}

extern "C" int rap_statistical_outliers(const float* points, int n, int nb_neighbors, double std_ratio, uint8_t* keep_out, int* n_keep_out, void* stream) {
    RAP_REQUIRE(points && keep_out && n_keep_out && n > 1 && nb_neighbors >= 1, "rap_statistical_outliers: bad argument");
    RAP_REQUIRE(nb_neighbors <= kKnnMax, "rap_statistical_outliers: nb_neighbors %d > %d", nb_neighbors, kKnnMax);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Temp tmp(st);
    double* avg = tmp.get<double>(n);
    double* acc = tmp.get<double>(2);
    int* d_count = tmp.get<int>(1);
    RAP_REQUIRE(avg && acc && d_count, "rap_statistical_outliers: out of device memory");
    RAP_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
    RAP_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
    const int k = std::min(nb_neighbors, n);
    knn_mean_kernel<kKnnMax><<<ceil_div(n, kKnnThreads), kKnnThreads, 0, st>>>(points, n, k, avg);
    const int rb = std::min(ceil_div(n, 256), 2 * num_sms());
    moments_kernel<<<rb, 256, 0, st>>>(avg, n, 0.0, acc, 0);
    double h[2];
    RAP_CUDA(cudaMemcpyAsync(h, acc, sizeof(double), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    const double mean = h[0] / n;
    moments_kernel<<<rb, 256, 0, st>>>(avg, n, mean, acc + 1, 1);
    RAP_CUDA(cudaMemcpyAsync(h + 1, acc + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    const double thr = mean + std_ratio * std::sqrt(h[1] / (n - 1));
    threshold_kernel<<<ceil_div(n, 256), 256, 0, st>>>(avg, n, thr, keep_out, d_count);
    RAP_CUDA(cudaMemcpyAsync(n_keep_out, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    RAP_CUDA(cudaStreamSynchronize(st));
    note_launch(4);
    return RAP_OK;
}
