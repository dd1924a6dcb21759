// What does ONE softmax warp per scheduler lose against the MUFU bound (8 cycles per exponential and scheduler)?  The attention kernel's two
// tiles take turns on the MUFU, so the block period is the sum of two single-warp exponential phases.  Variants of the inner loop over a
// 128-element row (registers only), 1 / 2 warps per scheduler, cycles per exponential per scheduler:
//   A  ex2 only                          B  + packed scale (FFMA2)           C  + packed row sum (FADD2)
//   D  + fp16 pack (F2FP) = the kernel   E  D with scalar FFMA / FADD        F  D, the row sum taken from the packed fp16 pairs (HADD2.F32? no: HFMA2 into f16x2 partials)
//   G  D, exps issued in groups of 8 before their consumers (source-level software pipelining)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o softmax_issue softmax_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b)); return r; }

template <int V> __global__ void k(float* out, long long* cyc, int iters, float sl2, float m) {
    float s[128];
    for (int i = 0; i < 128; ++i) s[i] = -0.01f * (threadIdx.x % 13 + i);
    uint32_t acc_p = 0;
    float total = 0.f;
    const f32x2 sl2x2 = pack2(sl2, sl2), nm2 = pack2(-m, -m);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        f32x2 ls[4] = {0ull, 0ull, 0ull, 0ull};
        float lsc[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t hs = 0;
        if (V == 6) {
#pragma unroll
            for (int g = 0; g < 128; g += 16) {
                float e[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float ua, ub;
                    unpack2(ffma2(pack2(s[g + i], s[g + i + 1]), sl2x2, nm2), ua, ub);
                    e[i] = ex2f(ua); e[i + 1] = ex2f(ub);
                }
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    ls[(i >> 1) & 3] = fadd2(ls[(i >> 1) & 3], pack2(e[i], e[i + 1]));
                    acc_p ^= pack_half2(e[i], e[i + 1]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
                float ua, ub;
                if (V == 0) { ua = s[i]; ub = s[i + 1]; }
                else if (V == 4) { ua = fmaf(s[i], sl2, -m); ub = fmaf(s[i + 1], sl2, -m); }
                else unpack2(ffma2(pack2(s[i], s[i + 1]), sl2x2, nm2), ua, ub);
                const float a = ex2f(ua), b = ex2f(ub);
                if (V == 0 || V == 1) { total += 0.f; acc_p ^= __float_as_uint(a) ^ __float_as_uint(b); }
                else if (V == 2) { ls[(i >> 1) & 3] = fadd2(ls[(i >> 1) & 3], pack2(a, b)); }
                else if (V == 3) { ls[(i >> 1) & 3] = fadd2(ls[(i >> 1) & 3], pack2(a, b)); acc_p ^= pack_half2(a, b); }
                else if (V == 4) { lsc[(i >> 1) & 3] += a + b; acc_p ^= pack_half2(a, b); }
                else if (V == 5) {      // row sum from the fp16 pairs: one HADD2 per pair instead of an FADD2 (precision aside: issue experiment)
                    const uint32_t pk = pack_half2(a, b);
                    acc_p ^= pk;
                    asm("add.rn.f16x2 %0, %0, %1;" : "+r"(hs) : "r"(pk));
                }
            }
        }
        float s0, s1;
        unpack2(fadd2(fadd2(ls[0], ls[1]), fadd2(ls[2], ls[3])), s0, s1);
        total += s0 + s1 + lsc[0] + lsc[1] + lsc[2] + lsc[3] + __uint_as_float(hs);
        s[it & 127] -= 1e-7f;
    }
    const long long t1 = clock64();
    float r = total + __uint_as_float(acc_p);
    for (int i = 0; i < 128; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int V> void run(const char* name) {
    float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
    printf("%-62s", name);
    for (int threads : {128, 256, 384}) {
        const int iters = 4000;
        k<V><<<148, threads>>>(d, c, 10, 0.18f, 1.0f);
        k<V><<<148, threads>>>(d, c, iters, 0.18f, 1.0f);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("  %d warp/sched: %5.2f", threads / 128, (double)h / (iters * 128.0) / (threads / 128));
    }
    printf("   [%s]\n", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(c);
}
int main() {
    run<0>("A ex2 only");
    run<1>("B + packed scale (FFMA2)");
    run<2>("C + packed row sum (FADD2)");
    run<3>("D + fp16 pack (F2FP): the kernel's loop");
    run<4>("E as D with scalar FFMA / FADD");
    run<5>("F as D, row sum from the fp16 pairs (HADD2)");
    run<6>("G as D, 16 exps issued before their consumers");
    return 0;
}
