// Softmax exp stream with a fraction QN/QD of the exponentials evaluated on the FMA pipe (Cody-Waite split + degree-3 polynomial,
// exponent spliced with integer ops) instead of MUFU.EX2: cycles per element (warp instruction) per SMSP, 1/2 warps per SMSP.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float exp2_poly(float t) {
    t = fmaxf(t, -126.0f);
    const float r = __fadd_rd(t, 12582912.0f);
    const float x = t - (r - 12582912.0f);
    float p = fmaf(0.077119089663028717f, x, 0.227564394474029541f);
    p = fmaf(p, x, 0.695146143436431885f);
    p = fmaf(p, x, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
template <int QN, int QD> __global__ void k(float* out, long long* cyc, int iters, float c, float m) {
    float x[32], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned pk = 0;
    for (int i = 0; i < 32; ++i) x[i] = -0.01f * (threadIdx.x % 13 + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float e[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float u = fmaf(x[i + j], c, -m);
                if (((i + j) % QD) < QN) e[j] = exp2_poly(u);
                else asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(u));
            }
            acc[(i >> 1) & 7] += e[0] + e[1];
            __half2 h = __floats2half2_rn(e[0], e[1]);
            pk ^= *reinterpret_cast<unsigned*>(&h);
            x[i] -= 1e-7f; x[i + 1] -= 1e-7f;
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i]; for (int i = 0; i < 32; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(pk);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// batched variant: the same work expressed stage by stage over 16 elements, so that every stage is 16 (or QN) independent instructions
template <int QN, int QD> __global__ void kb(float* out, long long* cyc, int iters, float c, float m) {
    float x[32], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned pk = 0;
    for (int i = 0; i < 32; ++i) x[i] = -0.01f * (threadIdx.x % 13 + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < 32; b += 16) {
            float u[16], r[16], e[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) u[j] = fmaf(x[b + j], c, -m);
#pragma unroll
            for (int j = 0; j < 16; ++j) if ((j % QD) < QN) u[j] = fmaxf(u[j], -126.0f);
#pragma unroll
            for (int j = 0; j < 16; ++j) if ((j % QD) < QN) r[j] = __fadd_rd(u[j], 12582912.0f); else asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(u[j]));
#pragma unroll
            for (int j = 0; j < 16; ++j) if ((j % QD) < QN) u[j] = u[j] - (r[j] - 12582912.0f);
#pragma unroll
            for (int j = 0; j < 16; ++j) if ((j % QD) < QN) e[j] = fmaf(0.077119089663028717f, u[j], 0.227564394474029541f);
#pragma unroll
            for (int j = 0; j < 16; ++j) if ((j % QD) < QN) e[j] = fmaf(e[j], u[j], 0.695146143436431885f);
#pragma unroll
            for (int j = 0; j < 16; ++j) if ((j % QD) < QN) e[j] = fmaf(e[j], u[j], 1.0f);
#pragma unroll
            for (int j = 0; j < 16; ++j) if ((j % QD) < QN) e[j] = __int_as_float(__float_as_int(e[j]) + (__float_as_int(r[j]) << 23));
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                acc[(j >> 1) & 7] += e[j] + e[j + 1];
                __half2 h = __floats2half2_rn(e[j], e[j + 1]);
                pk ^= *reinterpret_cast<unsigned*>(&h);
                x[b + j] -= 1e-7f; x[b + j + 1] -= 1e-7f;
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i]; for (int i = 0; i < 32; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(pk);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// packed f32x2 variant: pairs (2k, 2k+1) share FFMA2 / FADD2 instructions; pair k takes the polynomial path when (k % QD) < QN
__device__ __forceinline__ unsigned long long pk2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long add2_rm(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int QN, int QD> __global__ void kp(float* out, long long* cyc, int iters, float c, float m) {
    float x[32];
    unsigned long long acc[4] = {0, 0, 0, 0};
    unsigned pk = 0;
    for (int i = 0; i < 32; ++i) x[i] = -0.01f * (threadIdx.x % 13 + i);
    const unsigned long long c2 = pk2(c, c), nm2 = pk2(-m, -m), magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f);
    const unsigned long long k3 = pk2(0.077119089663028717f, 0.077119089663028717f), k2 = pk2(0.227564394474029541f, 0.227564394474029541f);
    const unsigned long long k1 = pk2(0.695146143436431885f, 0.695146143436431885f), k0 = pk2(1.0f, 1.0f), eps = pk2(-1e-7f, -1e-7f), neg1 = pk2(-1.0f, -1.0f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            unsigned long long xv = pk2(x[i], x[i + 1]);
            unsigned long long u = fma2(xv, c2, nm2);
            float e0, e1;
            if (((i >> 1) % QD) < QN) {
                float u0, u1; up2(u, u0, u1);
                u = pk2(fmaxf(u0, -126.0f), fmaxf(u1, -126.0f));
                const unsigned long long r = add2_rm(u, magic);
                const unsigned long long f = add2(u, fma2(r, neg1, magic));
                unsigned long long p = fma2(k3, f, k2);
                p = fma2(p, f, k1);
                p = fma2(p, f, k0);
                float p0, p1, r0, r1; up2(p, p0, p1); up2(r, r0, r1);
                e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
                e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
            } else {
                float u0, u1; up2(u, u0, u1);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(u0));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(u1));
            }
            acc[(i >> 1) & 3] = add2(acc[(i >> 1) & 3], pk2(e0, e1));
            __half2 h = __floats2half2_rn(e0, e1);
            pk ^= *reinterpret_cast<unsigned*>(&h);
            up2(add2(xv, eps), x[i], x[i + 1]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 4; ++i) { float a, b; up2(acc[i], a, b); s += a + b; } for (int i = 0; i < 32; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(pk);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int QN, int QD> void runp() {
    float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
    for (int threads : {128, 256, 512}) {
        int iters = 2000;
        kp<QN, QD><<<148, threads>>>(d, c, 10, 0.18f, 1.0f);
        kp<QN, QD><<<148, threads>>>(d, c, iters, 0.18f, 1.0f);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("packed  poly %2d/%-2d warps/SMSP=%d : %.2f cycles per element per SMSP\n", QN, QD, threads / 128, (double)h / (iters * 32.0) / (threads / 128));
    }
}
// packed AND batched: the 16 pairs of a 32-element chunk go through every stage together (all the MUFUs of the chunk first, then the polynomial
// pairs stage by stage), so that no instruction waits on the one in front of it -- the form a softmax warp with one partner per scheduler needs
template <int QN, int QD> __global__ void kpb(float* out, long long* cyc, int iters, float c, float m) {
    float x[32];
    unsigned long long acc[4] = {0, 0, 0, 0};
    unsigned pk = 0;
    for (int i = 0; i < 32; ++i) x[i] = -0.01f * (threadIdx.x % 13 + i);
    const unsigned long long c2 = pk2(c, c), nm2 = pk2(-m, -m), magic = pk2(12582912.0f, 12582912.0f);
    const unsigned long long k3 = pk2(0.077119089663028717f, 0.077119089663028717f), k2 = pk2(0.227564394474029541f, 0.227564394474029541f);
    const unsigned long long k1 = pk2(0.695146143436431885f, 0.695146143436431885f), k0 = pk2(1.0f, 1.0f), eps = pk2(-1e-7f, -1e-7f), neg1 = pk2(-1.0f, -1.0f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        unsigned long long u[16], r[16], f[16], e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) u[j] = fma2(pk2(x[2 * j], x[2 * j + 1]), c2, nm2);
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) >= QN) {
            float u0, u1, e0, e1; up2(u[j], u0, u1);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(u0));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(u1));
            e[j] = pk2(e0, e1);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) { float u0, u1; up2(u[j], u0, u1); u[j] = pk2(fmaxf(u0, -126.0f), fmaxf(u1, -126.0f)); }
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) r[j] = add2_rm(u[j], magic);
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) f[j] = fma2(r[j], neg1, magic);
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) f[j] = add2(u[j], f[j]);
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) e[j] = fma2(k3, f[j], k2);
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) e[j] = fma2(e[j], f[j], k1);
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) e[j] = fma2(e[j], f[j], k0);
#pragma unroll
        for (int j = 0; j < 16; ++j) if ((j % QD) < QN) {
            float p0, p1, r0, r1; up2(e[j], p0, p1); up2(r[j], r0, r1);
            e[j] = pk2(__int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23)), __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23)));
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc[j & 3] = add2(acc[j & 3], e[j]);
            float e0, e1; up2(e[j], e0, e1);
            __half2 h = __floats2half2_rn(e0, e1);
            pk ^= *reinterpret_cast<unsigned*>(&h);
            up2(add2(pk2(x[2 * j], x[2 * j + 1]), eps), x[2 * j], x[2 * j + 1]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 4; ++i) { float a, b; up2(acc[i], a, b); s += a + b; } for (int i = 0; i < 32; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(pk);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int QN, int QD> void runpb() {
    float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
    for (int threads : {128, 256, 512}) {
        int iters = 2000;
        kpb<QN, QD><<<148, threads>>>(d, c, 10, 0.18f, 1.0f);
        kpb<QN, QD><<<148, threads>>>(d, c, iters, 0.18f, 1.0f);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("packed+batched poly %2d/%-2d warps/SMSP=%d : %.2f cycles per element per SMSP\n", QN, QD, threads / 128, (double)h / (iters * 32.0) / (threads / 128));
    }
}
template <int QN, int QD> void runb() {
    float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
    for (int threads : {128, 256, 512}) {
        int iters = 2000;
        kb<QN, QD><<<148, threads>>>(d, c, 10, 0.18f, 1.0f);
        kb<QN, QD><<<148, threads>>>(d, c, iters, 0.18f, 1.0f);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("batched poly %2d/%-2d warps/SMSP=%d : %.2f cycles per element per SMSP\n", QN, QD, threads / 128, (double)h / (iters * 32.0) / (threads / 128));
    }
}
template <int QN, int QD> void run() {
    float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
    for (int threads : {128, 256}) {
        int iters = 2000;
        k<QN, QD><<<148, threads>>>(d, c, 10, 0.18f, 1.0f);
        k<QN, QD><<<148, threads>>>(d, c, iters, 0.18f, 1.0f);
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("poly %2d/%-2d warps/SMSP=%d : %.2f cycles per element per SMSP\n", QN, QD, threads / 128, (double)h / (iters * 32.0) / (threads / 128));
    }
}
__global__ void acc_check(float* err) {
    float worst = 0;
    for (int i = threadIdx.x; i < 4000000; i += blockDim.x) {
        float t = -30.0f + i * 7.6e-6f, a; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(t));
        worst = fmaxf(worst, fabsf(exp2_poly(t) / exp2f(t) - 1.0f));
    }
    atomicMax(reinterpret_cast<int*>(err), __float_as_int(worst));
}
int main() {
    run<0, 4>(); run<1, 4>(); run<1, 1>();
    runp<0, 4>(); runp<1, 8>(); runp<1, 4>(); runp<5, 16>(); runp<3, 8>(); runp<1, 2>(); runp<1, 1>();
    runpb<0, 8>(); runpb<1, 8>(); runpb<1, 4>(); runpb<3, 8>(); runpb<1, 2>(); runpb<1, 1>();
    float* e; cudaMalloc(&e, 4); cudaMemset(e, 0, 4); acc_check<<<1, 1024>>>(e); float h; cudaMemcpy(&h, e, 4, cudaMemcpyDeviceToHost);
    printf("poly max rel err on [-30,0.4]: %.3e\n", h);
    return 0;
}
