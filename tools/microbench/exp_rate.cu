// Throughput of exp2 variants on one B200: fp32 MUFU, packed f16x2 MUFU, FMA-pipe polynomial.  elements/clk/SM.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__global__ void k_f32(float* out, int iters) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f16x2(float* out, int iters) {
    unsigned x[8];
    for (int i = 0; i < 8; ++i) { __half2 h = __floats2half2_rn(-0.001f * (threadIdx.x + i), -0.5f); x[i] = *(unsigned*)&h; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
    unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
__global__ void k_poly(float* out, int iters) {   // 2^x, x<=0: floor split + degree-3 minimax on [0,1) + exponent add
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = -0.37f * (threadIdx.x % 7 + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float xc = fmaxf(x[i], -126.f);
            float fl = floorf(xc);
            float f = xc - fl;
            float p = fmaf(f, 0.0555041f, 0.2402265f);
            p = fmaf(p, f, 0.6931472f);
            p = fmaf(p, f, 1.0f);
            x[i] = __int_as_float(__float_as_int(p) + ((int)fl << 23)) - 1.5f;
        }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K> float run(K k, float* d, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148 * 4, 256>>>(d, 10);
    cudaEventRecord(a); k<<<148 * 4, 256>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    int iters = 20000; int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double ops = 148.0 * 4 * 256 * 8.0 * iters;
    float t1 = run(k_f32, d, iters), t2 = run(k_f16x2, d, iters), t3 = run(k_poly, d, iters);
    double clk = clk_khz * 1e3;
    printf("clock %.0f MHz (attr)\n", clk / 1e6);
    printf("ex2.f32    : %.3f ms  %.1f elem/clk/SM\n", t1, ops / (t1 * 1e-3) / clk / 148);
    printf("ex2.f16x2  : %.3f ms  %.1f elem/clk/SM (2 per op)\n", t2, 2 * ops / (t2 * 1e-3) / clk / 148);
    printf("poly exp2  : %.3f ms  %.1f elem/clk/SM\n", t3, ops / (t3 * 1e-3) / clk / 148);
    return 0;
}
