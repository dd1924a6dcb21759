// How fast can ONE warp per scheduler (SMSP) stream exp2?  cycles per MUFU for: pure MUFU, softmax-like mix (FFMA+MUFU+FADD
// + pack every other), with 1 and 2 warps per SMSP.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, long long* cyc, int iters, float c, float m) {
    float x[32], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned pk = 0;
    for (int i = 0; i < 32; ++i) x[i] = -0.01f * (threadIdx.x % 13 + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float a, b;
            if (MODE == 0) {
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(x[i]));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(x[i + 1]));
                x[i] = a * 0.0f - 1.0f - 0.001f * i; x[i + 1] = b;   // keep dependency short but real
            } else {
                float u = fmaf(x[i], c, -m), v = fmaf(x[i + 1], c, -m);
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(u));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(v));
                acc[(i >> 1) & 7] += a + b;
                __half2 h = __floats2half2_rn(a, b);
                pk ^= *reinterpret_cast<unsigned*>(&h);
                x[i] -= 1e-7f; x[i + 1] -= 1e-7f;
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i]; for (int i = 0; i < 32; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(pk);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int threads) {
    float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
    int iters = 2000;
    k<MODE><<<148, threads>>>(d, c, 10, 0.18f, 1.0f);
    k<MODE><<<148, threads>>>(d, c, iters, 0.18f, 1.0f);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SMSP=%d : %.2f cycles per MUFU warp-instr (per warp)  -> %.2f cycles per MUFU per SMSP\n", name, threads / 128,
           (double)h / (iters * 32.0), (double)h / (iters * 32.0) / (threads / 128));
}
int main() {
    run<0>("pure MUFU", 128); run<0>("pure MUFU", 256); run<0>("pure MUFU", 512);
    run<1>("FFMA+MUFU+FADD+pack mix", 128); run<1>("FFMA+MUFU+FADD+pack mix", 256); run<1>("FFMA+MUFU+FADD+pack mix", 512);
    return 0;
}
