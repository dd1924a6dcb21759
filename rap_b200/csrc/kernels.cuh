// Launchers of the HBM-bound / small kernels around the tensor-core kernels (elementwise.cu, procrustes.cu).
#pragma once
#include "common.h"

namespace rap {

// ---- embedding front-end (flow_model/embedding.py:31-58,131-177) -------------------------------
// A_s[T,128] fp16 = [posenc(cond) 63 | posenc(scale) 21 | local feat 32 | 0 x12]   (step-invariant part)
int launch_posenc_static(const float* cond, const float* scales, const int* scene_of_tok, const float* feats, int feat_dim,
                         __half* out, int T, cudaStream_t st);
// A_x[T,64] fp16 = [posenc(x_t) 63 | 0]
int launch_posenc_x(const float* x, __half* out, int T, cudaStream_t st);

// ---- LayerNorm (+AdaLN modulation) -> fp16 GEMM operand (flow_model/norm.py:76, layer.py:163) ---
// out = LN(h) * (one_plus + mul[c]) + add[c]; mul/add rows selected per token via row_of_tok (never null).
int launch_ln_modulate(const float* h, __half* out, int T, const float* mul, const float* add, int row_stride,
                       const int* row_of_tok, float one_plus, cudaStream_t st);

// ---- AdaLN conditioning MLP for a list of timesteps (flow_model/norm.py:71-73) ------------------
int launch_timestep_sincos(const float* t, float* out, int n, cudaStream_t st);           // out[n,256]
// out[n,m,O] = (silu?)(W[m,O,K] . in[n,(m),K] + b[m,O]);  in_has_m = 0 -> in[n,K] shared by all modules
int launch_ada_linear(const __half* W, const float* b, const float* in, int in_has_m, float* out, int n, int M, int O, int K,
                      int silu_out, cudaStream_t st);

// ---- head tail + Euler step (point_cloud_dit.py:116 last Linear; sampler.py:88-90) --------------
// v = y2[T,256] . W4[3,256]^T.  If x_t != null: x0_hat = x_t - v t ; x_t <- x_t - dt v (written to x_t and xt_traj).
int launch_head_final(const __half* y2, const float* W4, int T, float* v_out, float* x_t, float t, float dt,
                      float* x0hat_out, float* xt_traj, cudaStream_t st);

// ---- weight packing (fp32 source -> fp16/fp32 packed destination) --------------------------------
struct PackSpec {
    int rows, cols;          // destination shape
    int src_ld;              // source leading dimension (elements)
    int row_mode;            // 0 identity, 1 GEGLU interleave (block 128 of h then 128 of g; g rows start at rows/2 in source)
    int n_ranges;            // column ranges copied back to back, remaining destination columns are zero
    int range_start[3], range_len[3];
};
int launch_pack(const float* src, void* dst, int dst_is_half, const PackSpec& spec, cudaStream_t st);

// ---- Procrustes (procrustes.py:6-118) ------------------------------------------------------------
// One CTA per part.  (R,t) = argmin |R src + t - tgt|.  Optional outputs:
//   R_out/t_out  : written at slot[part] (row-major 3x3 / 3)
//   apply_out    : (R src + t) * w_rigid + mix[i] * w_mix   (mix may be null -> pure rigid points)
//   apply_out2   : second copy of apply_out (trajectory row) or null
int launch_procrustes(const float* src, const float* tgt, const int* cu_part, const int* slot, int n_parts,
                      float* R_out, float* t_out, float* apply_out, float* apply_out2, const float* mix,
                      float w_rigid, float w_mix, cudaStream_t st, float* sse_out = nullptr, const float* R_in = nullptr,
                      const float* t_in = nullptr, long long n_points = 0);
//   sse_out      : per slot sum of squared residuals |R src + t - tgt|^2 (rigidity RMSE for generation selection)
//   R_in / t_in  : evaluate a GIVEN transform (slot-indexed) instead of fitting one
//   n_points     : rows of src/tgt (> 0: parts reaching past it are skipped and flagged in the device error word)

void kabsch_host(const double* H9, double* R9);

}  // namespace rap
