// Per-SM global WRITE throughput on one B200, in the access patterns of the fp32 GEMM epilogue (a 128-row x 256-column fp32 tile of a
// [8192 x 512] residual stream per CTA, 128 KB): what can one SM push out per clock when every SM does it at once?
//   stg_lin     STG.128, every warp instruction 512 contiguous bytes
//   stg_tile    STG.128, every warp instruction 4 rows x 128 B (row pitch 2 KB) -- the register epilogue's store pattern
//   rmw_tile    the same with the matching LDG.128 in front (residual read + add + store)
//   tma_store   32-row x 128-byte boxes from shared memory through cp.async.bulk.tensor (UTMASTG), one issuing lane per warp
//   tma_reduce  the same boxes as cp.reduce.async.bulk.tensor .add (UTMAREDG)
// Output: bytes / clk / SM (clock64 of CTA 0 over the whole loop) and the aggregate TB/s from CUDA events.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_rate store_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int ROWS = 8192, COLS = 512, TM = 128, TN = 256, REPS = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stg4(float* p, float4 v) {   // volatile: repeated stores to one address are not merged by the compiler
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* h, const __grid_constant__ CUtensorMap tm, long long* cyc) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x % (ROWS / TM * (COLS / TN));
    const int m0 = (tile / (COLS / TN)) * TM, n0 = (tile % (COLS / TN)) * TN;
    const int q = warp & 3, half = warp >> 2;
    const int row_base = m0 + q * 32, col0 = n0 + half * 128;
    const int rr0 = lane >> 3, cc = lane & 7;
    if (MODE >= 3) {   // fill this warp's two boxes once
        for (int i = lane; i < 2 * 1024; i += 32) reinterpret_cast<float*>(raw + (base - smem_u32(raw)))[warp * 2048 + i] = 1.0f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (MODE == 0) {            // linear: this CTA's 128 KB as one contiguous range
                float* p = h + (size_t)blockIdx.x * (TM * TN) + (size_t)(warp * 4 + c) * 1024;
#pragma unroll
                for (int it = 0; it < 8; ++it) stg4(p + it * 128 + lane * 4, v);
            } else if (MODE == 1 || MODE == 2) {
                float* p = h + (size_t)(row_base + rr0) * COLS + col0 + c * 32 + cc * 4;
                float4 a[8];
                if (MODE == 2) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) a[it] = *reinterpret_cast<const float4*>(p + (size_t)it * 4 * COLS);
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    float4 o = v;
                    if (MODE == 2) { o.x += a[it].x; o.y += a[it].y; o.z += a[it].z; o.w += a[it].w; }
                    stg4(p + (size_t)it * 4 * COLS, o);
                }
            } else {
                const uint32_t box = base + warp * 8192 + (c & 1) * 4096;
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    if (MODE == 3)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(box), "r"(col0 + c * 32), "r"(row_base) : "memory");
                    else
                        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(box), "r"(col0 + c * 32), "r"(row_base) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                __syncwarp();
            }
        }
        v.x += 1.f;
    }
    if (MODE >= 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE> void run(const char* name, float* h, const CUtensorMap& tm, long long* cyc, int grid) {
    const int smem = 8 * 8192 + 1024 + 100 * 1024;   // + ballast: one CTA per SM
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 200; ++i) k<MODE><<<grid, 256, smem>>>(h, tm, cyc);   // long enough for the clocks to ramp up
    cudaEventRecord(e0);
    const int N = 20;
    for (int i = 0; i < N; ++i) k<MODE><<<grid, 256, smem>>>(h, tm, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes_cta = (double)REPS * TM * TN * 4 * (MODE == 2 ? 2 : 1);
    printf("%-10s grid %3d: %6.1f B/clk/SM (%lld cycles per %d x 128 KB%s), kernel %.1f us (%.2f GHz) -> %.2f TB/s aggregate  [%s]\n", name, grid, bytes_cta / c, c, REPS,
           MODE == 2 ? " in + out" : "", ms / N * 1e3, c / (ms / N * 1e6), bytes_cta * grid / (ms / N * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* h; long long* cyc;
    cudaMalloc(&h, (size_t)148 * TM * TN * 4 + ROWS * COLS * 4); cudaMalloc(&cyc, 8);
    cudaMemset(h, 0, (size_t)ROWS * COLS * 4);
    void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
    CUtensorMap tm;
    cuuint64_t dims[2] = {COLS, ROWS}, strides[1] = {COLS * 4};
    cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map: %d\n", (int)r); return 1; }
    for (int grid : {128, 148, 16}) {
        run<0>("stg_lin", h, tm, cyc, grid);
        run<1>("stg_tile", h, tm, cyc, grid);
        run<2>("rmw_tile", h, tm, cyc, grid);
        run<3>("tma_store", h, tm, cyc, grid);
        run<4>("tma_reduce", h, tm, cyc, grid);
    }
    return 0;
}
