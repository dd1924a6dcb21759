// Host-side plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>

namespace rap {

// errno-style negative return codes of the C ABI (include/rap_b200.h)
enum : int { RAP_OK = 0, RAP_EINVAL = -22, RAP_ENOMEM = -12, RAP_ECUDA = -5, RAP_ENODEV = -19, RAP_ESTATE = -1 };

void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define RAP_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::rap::fail(::rap::RAP_ECUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,      \
                               cudaGetErrorString(_e));                                             \
    } while (0)

#define RAP_REQUIRE(cond, ...)                                                                      \
    do {                                                                                            \
        if (!(cond)) return ::rap::fail(::rap::RAP_EINVAL, __VA_ARGS__);                            \
    } while (0)

#define RAP_TRY(expr)                                                                               \
    do {                                                                                            \
        int _rc = (expr);                                                                           \
        if (_rc != 0) return _rc;                                                                   \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// 2D fp16 row-major tensor map with a (box_rows x 64 cols) box and 128-byte swizzle.
int make_tmap_f16_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                        uint32_t box_rows);
// 2D fp32 row-major tensor map with a (box_rows x 32 cols = 128 B) box and 128-byte swizzle (bulk stores / reductions of GEMM epilogues).
int make_tmap_f32_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                        uint32_t box_rows);

int num_sms();

// Launch with programmatic stream serialization (PDL): the kernel may start while its predecessor in the stream is still draining; it
// must call pdl_wait() before reading or writing global memory.  RAP_NO_PDL=1 falls back to plain stream order (A/B measurements).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// Same, as a thread-block cluster of `cluster_x` consecutive CTAs (distributed shared memory between them).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
void note_launch(int n = 1);   // counts kernels launched by this library (rap_launch_count)

// Optional per-kernel-class timing (rap_profile_begin/_end): CUDA events recorded on the launching stream
// around every launch of the library's kernels, so bench.py can report live per-kernel durations.
enum KernelClass : int { KC_GEMM_F32 = 0, KC_GEMM_SILU, KC_GEMM_QKV, KC_GEMM_GEGLU, KC_FMHA, KC_LN, KC_POSENC, KC_HEAD, KC_ADA,
                         KC_PROCRUSTES, KC_PACK, KC_COUNT };
bool profiling_enabled();
void profile_mark(int kc, cudaStream_t st, bool end);
struct ProfScope {
    int kc; cudaStream_t st; bool on;
    ProfScope(int k, cudaStream_t s) : kc(k), st(s), on(profiling_enabled()) { if (on) profile_mark(kc, st, false); }
    ~ProfScope() { if (on) profile_mark(kc, st, true); }
};

}  // namespace rap
