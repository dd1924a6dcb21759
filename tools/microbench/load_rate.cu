// Per-SM operand INGEST through the TMA on one B200: every CTA streams 48 KB stages (three 128-row x 128-byte fp16 boxes, the GEMM's A tile +
// 256-row weight tile per 64-deep K step) from L2-resident tensors into shared memory with S stages in flight, nothing consumes them.
// Output per (CTAs, S): ns and SM cycles per 48 KB stage, GB/s per SM and aggregate.  The GEMM K loops take 375-384 ns per stage with S = 3.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o load_rate load_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int STEPS = 512;   // stages streamed per CTA
template <int S>
__global__ void __launch_bounds__(32, 1) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bars = base + S * 49152;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    long long g0, g1; const long long c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const int m0 = (blockIdx.x % 64) * 128, n0 = (blockIdx.x % 6) * 256;
    auto issue = [&](int i) {
        const int s = i % S, kc = (i % 8) * 64;
        const uint32_t dst = base + s * 49152, bar = bars + 8 * s;
        mbar_expect_tx(bar, 49152);
        tma_load_2d(dst, &tmA, bar, kc, m0);
        tma_load_2d(dst + 16384, &tmB, bar, kc, n0);
        tma_load_2d(dst + 32768, &tmB, bar, kc, n0 + 128);
    };
    for (int i = 0; i < S; ++i) issue(i);
    for (int i = 0; i < STEPS; ++i) {
        mbar_wait(bars + 8 * (i % S), (i / S) & 1);
        if (i + S < STEPS) issue(i + S);
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[2 * blockIdx.x] = g1 - g0; out[2 * blockIdx.x + 1] = clock64() - c0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make(EncodeFn fn, void* base, uint64_t rows, uint64_t cols) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return tm;
}
template <int S> void run(const CUtensorMap& a, const CUtensorMap& b, long long* d, int grid) {
    const int smem = S * 49152 + 64 + 1024;
    cudaFuncSetAttribute(k<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int i = 0; i < 200; ++i) k<S><<<grid, 32, smem>>>(a, b, d);     // clocks up
    cudaDeviceSynchronize();
    long long h[2 * 148];
    cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double ns = 0, cyc = 0;
    for (int i = 0; i < grid; ++i) { ns += h[2 * i]; cyc += h[2 * i + 1]; }
    ns /= grid; cyc /= grid;
    printf("CTAs %3d, %d stages of 48 KB in flight: %6.1f ns = %6.1f cycles per stage (%.2f GHz)  ->  %6.1f GB/s per SM, %5.2f TB/s aggregate  [%s]\n", grid, S, ns / STEPS,
           cyc / STEPS, cyc / ns, 49152.0 * STEPS / ns, 49152.0 * STEPS / ns * grid / 1e3, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    __half *A, *B; long long* d;
    cudaMalloc(&A, 8192 * 512 * 2); cudaMalloc(&B, 1536 * 512 * 2); cudaMalloc(&d, sizeof(long long) * 2 * 148);
    cudaMemset(A, 0, 8192 * 512 * 2); cudaMemset(B, 0, 1536 * 512 * 2);
    void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
    const CUtensorMap ta = make((EncodeFn)fp, A, 8192, 512), tb = make((EncodeFn)fp, B, 1536, 512);
    for (int grid : {148, 128, 16, 1}) {
        run<1>(ta, tb, d, grid); run<2>(ta, tb, d, grid); run<3>(ta, tb, d, grid); run<4>(ta, tb, d, grid);
    }
    return 0;
}
