// Fused per-part Procrustes / Kabsch (reference procrustes.py:6-118):
//   segmented means -> centred 3x3 covariance -> 3x3 SVD -> det fix -> (R,t) -> optional apply + re-interpolation
// in ONE launch with one CTA per part and no host synchronisation (the reference runs a Python loop over
// B x max_parts slots with a cuSOLVER SVD, a det() and >=1 device sync per part).
// Reductions and the SVD run in fp64 (a few hundred flops per part) so the result is at least as accurate
// as the reference's fp32 torch.linalg.svd; the point data path is fp32.
#include "../../include/rap_b200.h"
#include "kernels.cuh"
#include "sm100.cuh"

namespace rap {
namespace {

constexpr int kThreads = 512;

template <int N>
__device__ __forceinline__ void block_reduce(double (&v)[N], double* smem /* [16][N] */, double* result /* [N] smem */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int o = 16; o; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < N; ++i) smem[warp * N + i] = v[i];
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += smem[w * N + threadIdx.x];   // fixed order: deterministic
        result[threadIdx.x] = s;
    }
    __syncthreads();
}

// One-sided (Hestenes) Jacobi SVD of a 3x3: A V = U diag(s), singular values sorted descending,
// degenerate left vectors completed to a right-handed frame.  Then Kabsch: R = V diag(1,1,d) U^T with
// d = sign(det(V U^T)) -- identical to flipping the last row of Vt when det(R) < 0 (procrustes.py:28-33).
__host__ __device__ void kabsch_from_covariance(const double H[9], double R[9]) {
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { A[i][j] = H[i * 3 + j]; V[i][j] = (i == j); }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) { alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q]; }
                if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
                off = fmax(off, fabs(gamma) / sqrt(alpha * beta));
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
                for (int i = 0; i < 3; ++i) {
                    const double ap = A[i][p], aq = A[i][q];
                    A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    double sv[3];
    int ord[3] = {0, 1, 2};
    for (int j = 0; j < 3; ++j) sv[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (sv[ord[b]] > sv[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double U[3][3], Vs[3][3];
    const double tiny = 1e-12 * fmax(sv[ord[0]], 1e-300);
    for (int j = 0; j < 3; ++j) {
        const int c = ord[j];
        for (int i = 0; i < 3; ++i) { Vs[i][j] = V[i][c]; U[i][j] = sv[c] > tiny ? A[i][c] / sv[c] : 0.0; }
    }
    // complete rank-deficient U (sigma ~ 0) so that the frame is orthonormal
    if (sv[ord[0]] <= 1e-300) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) U[i][j] = Vs[i][j]; }
    else {
        if (sv[ord[1]] <= tiny) {   // rank 1: pick any unit vector orthogonal to U0
            const double ax = fabs(U[0][0]), ay = fabs(U[1][0]), az = fabs(U[2][0]);
            double e[3] = {0, 0, 0};
            e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
            const double d = e[0] * U[0][0] + e[1] * U[1][0] + e[2] * U[2][0];
            double n = 0;
            for (int i = 0; i < 3; ++i) { U[i][1] = e[i] - d * U[i][0]; n += U[i][1] * U[i][1]; }
            n = sqrt(n);
            for (int i = 0; i < 3; ++i) U[i][1] /= n;
        }
        if (sv[ord[2]] <= tiny) {   // third = first x second
            U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
        }
    }
    auto det3 = [](const double M[3][3]) {
        return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
               M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    };
    const double d = (det3(Vs) * det3(U) < 0) ? -1.0 : 1.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + d * Vs[i][2] * U[j][2];
}

__global__ void __launch_bounds__(kThreads)
procrustes_kernel(const float* __restrict__ src, const float* __restrict__ tgt, const int* __restrict__ cu_part,
                  const int* __restrict__ slot, float* __restrict__ R_out, float* __restrict__ t_out,
                  float* __restrict__ apply_out, float* __restrict__ apply_out2, const float* __restrict__ mix,
                  float w_rigid, float w_mix, float* __restrict__ sse_out, const float* __restrict__ R_in, const float* __restrict__ t_in,
                  long long n_points) {
    __shared__ double red[(kThreads / 32) * 9];
    __shared__ double res[9];
    __shared__ float Rt[12];
    const int part = blockIdx.x;
    const int s = cu_part[part], e = cu_part[part + 1], n = e - s;
    if (n <= 0) return;
    // The reference asserts sum(points_per_part) == N on the host (utils/point_clouds.py:41-44); here a part table that runs past the
    // clouds is refused on the device (no read or write out of bounds) and reported through the error word (rap_device_error).
    if (n_points > 0 && (s < 0 || e > n_points)) {
        if (threadIdx.x == 0) atomicCAS(&g_dev_error, 0u, 0xB1u | ((unsigned)part << 8));
        return;
    }
    const float* S = src + (size_t)s * 3;
    const float* G = tgt + (size_t)s * 3;
    const int sl_io = slot ? slot[part] : part;
    if (R_in) {                                   // residual of a GIVEN transform (eval/metrics.py:511-622): skip the fit
        if (threadIdx.x < 9) Rt[threadIdx.x] = R_in[(size_t)sl_io * 9 + threadIdx.x];
        else if (threadIdx.x < 12) Rt[threadIdx.x] = t_in[(size_t)sl_io * 3 + threadIdx.x - 9];
        __syncthreads();
    } else {
    double m[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += kThreads) {
        m[0] += S[3 * i]; m[1] += S[3 * i + 1]; m[2] += S[3 * i + 2];
        m[3] += G[3 * i]; m[4] += G[3 * i + 1]; m[5] += G[3 * i + 2];
    }
    block_reduce<6>(m, red, res);
    const double inv_n = 1.0 / n;
    const double ms[3] = {res[0] * inv_n, res[1] * inv_n, res[2] * inv_n};
    const double mt[3] = {res[3] * inv_n, res[4] * inv_n, res[5] * inv_n};
    __syncthreads();   // res is reused below

    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // H = src_c^T tgt_c
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const double a[3] = {S[3 * i] - ms[0], S[3 * i + 1] - ms[1], S[3 * i + 2] - ms[2]};
        const double b[3] = {G[3 * i] - mt[0], G[3 * i + 1] - mt[1], G[3 * i + 2] - mt[2]};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) h[r * 3 + c] += a[r] * b[c];
    }
    block_reduce<9>(h, red, res);

    if (threadIdx.x == 0) {
        double H[9], R[9];
        for (int i = 0; i < 9; ++i) H[i] = res[i];
        kabsch_from_covariance(H, R);
        for (int i = 0; i < 9; ++i) Rt[i] = (float)R[i];
        for (int i = 0; i < 3; ++i)     // t = mu_tgt - R mu_src
            Rt[9 + i] = (float)(mt[i] - (R[i * 3] * ms[0] + R[i * 3 + 1] * ms[1] + R[i * 3 + 2] * ms[2]));
        if (R_out) {
            const int sl = slot ? slot[part] : part;
            for (int i = 0; i < 9; ++i) R_out[(size_t)sl * 9 + i] = Rt[i];
            for (int i = 0; i < 3; ++i) t_out[(size_t)sl * 3 + i] = Rt[9 + i];
        }
    }
    __syncthreads();
    }
    if (sse_out) {                                // sum of squared residuals |R s + t - g|^2 of the part (fp64 accumulation)
        double acc[1] = {0.0};
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const float x = S[3 * i], y = S[3 * i + 1], z = S[3 * i + 2];
            const float d0 = Rt[0] * x + Rt[1] * y + Rt[2] * z + Rt[9] - G[3 * i];
            const float d1 = Rt[3] * x + Rt[4] * y + Rt[5] * z + Rt[10] - G[3 * i + 1];
            const float d2 = Rt[6] * x + Rt[7] * y + Rt[8] * z + Rt[11] - G[3 * i + 2];
            acc[0] += (double)(d0 * d0 + d1 * d1 + d2 * d2);
        }
        __syncthreads();
        block_reduce<1>(acc, red, res);
        if (threadIdx.x == 0) sse_out[sl_io] = (float)res[0];
    }
    if (apply_out) {
        const float r0 = Rt[0], r1 = Rt[1], r2 = Rt[2], r3 = Rt[3], r4 = Rt[4], r5 = Rt[5], r6 = Rt[6], r7 = Rt[7], r8 = Rt[8];
        const float t0 = Rt[9], t1 = Rt[10], t2 = Rt[11];
        for (int i = threadIdx.x; i < n; i += kThreads) {
            const float x = S[3 * i], y = S[3 * i + 1], z = S[3 * i + 2];
            float o0 = (r0 * x + r1 * y + r2 * z + t0) * w_rigid;
            float o1 = (r3 * x + r4 * y + r5 * z + t1) * w_rigid;
            float o2 = (r6 * x + r7 * y + r8 * z + t2) * w_rigid;
            const size_t g = (size_t)(s + i) * 3;
            if (mix) { o0 += mix[g] * w_mix; o1 += mix[g + 1] * w_mix; o2 += mix[g + 2] * w_mix; }
            apply_out[g] = o0; apply_out[g + 1] = o1; apply_out[g + 2] = o2;
            if (apply_out2) { apply_out2[g] = o0; apply_out2[g + 1] = o1; apply_out2[g + 2] = o2; }
        }
    }
}

}  // namespace

void kabsch_host(const double* H, double* R) { kabsch_from_covariance(H, R); }

int launch_procrustes(const float* src, const float* tgt, const int* cu_part, const int* slot, int n_parts,
                      float* R_out, float* t_out, float* apply_out, float* apply_out2, const float* mix,
                      float w_rigid, float w_mix, cudaStream_t st, float* sse_out, const float* R_in, const float* t_in, long long n_points) {
    if (n_parts <= 0) return RAP_OK;
    ProfScope _ps(KC_PROCRUSTES, st);
    procrustes_kernel<<<n_parts, kThreads, 0, st>>>(src, tgt, cu_part, slot, R_out, t_out, apply_out, apply_out2, mix, w_rigid, w_mix, sse_out, R_in, t_in, n_points);
    RAP_CUDA(cudaGetLastError());
    note_launch();
    return RAP_OK;
}

}  // namespace rap

extern "C" int rap_kabsch_host(const double* H9, double* R9) {
    if (!H9 || !R9) return rap::fail(rap::RAP_EINVAL, "rap_kabsch_host: NULL argument");
    rap::kabsch_host(H9, R9);
    return rap::RAP_OK;
}
