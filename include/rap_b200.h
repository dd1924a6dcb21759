/* rap_b200 -- C ABI of the B200-native flow-matching registration sampler (PRBonn/RAP hot path).
 *
 * The reference has no FFI/plugin registry; its drop-in boundary is three Python modules
 * (SURVEY.md 8(b)).  The Python mirrors in rap_b200/rectified_point_flow/ bind exactly these entry
 * points with ctypes; each one names the reference interface it replaces (paths relative to
 * /root/reference/rectified_point_flow).
 *
 * Conventions: all tensor pointers are DEVICE pointers borrowed from the caller (row-major, contiguous,
 * kept alive until the stream has run); fp32 for coordinates/poses, int32 for offsets; `stream` is a
 * cudaStream_t passed as void*.  No entry point synchronises the host except where stated.  Return 0 on
 * success or a negative errno-style code, with a message available from rap_last_error().
 */
#ifndef RAP_B200_H
#define RAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rap_dit rap_dit_t;   /* packed PointCloudDiT weights on one device                  */
typedef struct rap_plan rap_plan_t; /* batch layout (cu_seqlens), workspace, tensor maps, tile lists */

int rap_version(void);
const char* rap_last_error(void);
/* Reads (and clears) the device-side protocol error word; host-synchronising, for tests. 0 = clean. */
int rap_device_error(unsigned int* code_out);

/* ---- model: replaces flow_model.PointCloudDiT (flow_model/point_cloud_dit.py:10-117) -------------- */
int rap_dit_create(rap_dit_t** out, int embed_dim, int num_layers, int num_heads, int local_feat_dim);
void rap_dit_destroy(rap_dit_t* dit);
/* `key` is the reference state_dict key; `src` a device fp32 tensor of `numel`
 * elements in the reference layout.  The library converts / re-packs it (fp16 operands, GEGLU row
 * interleave, emb_proj column split). */
int rap_dit_load_weight(rap_dit_t* dit, const char* key, const float* src, int64_t numel, void* stream);
/* Fails unless every key of the reference state_dict has been loaded exactly as shaped. */
int rap_dit_finalize(rap_dit_t* dit);
int rap_dit_num_keys(const rap_dit_t* dit);
const char* rap_dit_key(const rap_dit_t* dit, int index, int64_t* numel_out);

/* ---- batch plan: replaces RectifiedPointFlow._prepare_data (modeling.py:203-231) ------------------
 * cu_seqlens_batch: HOST int32 [n_scenes+1]; points_per_part: HOST int32 [n_scenes*max_parts]
 * (zero = empty slot).  max_timesteps bounds num_steps of rap_sample_euler / n_scenes of rap_dit_forward. */
int rap_plan_create(rap_plan_t** out, rap_dit_t* dit, int n_scenes, const int32_t* cu_seqlens_batch,
                    const int32_t* points_per_part, int max_parts, int max_timesteps);
void rap_plan_destroy(rap_plan_t* plan);
int rap_plan_tokens(const rap_plan_t* plan);
int rap_plan_valid_parts(const rap_plan_t* plan);
/* Step-invariant inputs (cond (T,3) f32, local_features (T,F) f32, scales (n_scenes) f32, anchor (T) u8):
 * evaluates the step-invariant half of the embedding once (embedding.py:155-177 minus the x_t term). */
int rap_plan_set_condition(rap_plan_t* plan, const float* cond, const float* local_features, const float* scales,
                           const uint8_t* anchor_indices, void* stream);

/* PointCloudDiT.forward (point_cloud_dit.py:141-191): velocity (T,3) f32; transformer_features (T,512) f32
 * or NULL.  timesteps: device f32 [n_scenes]. */
int rap_dit_forward(rap_plan_t* plan, const float* x, const float* timesteps, float* velocity,
                    float* transformer_features, void* stream);

/* get_sampler('euler') / flow_sampler (sampler.py:11-92) fused with the model: all num_steps Euler steps
 * with x_t resident on the device.  end_point_trajectory / trajectory: (num_steps,T,3) f32 or NULL;
 * x_final: (T,3) f32 or NULL; transformer_features: (T,512) of the last step or NULL. */
int rap_sample_euler(rap_plan_t* plan, const float* x_1, int num_steps, int rigidity_forcing,
                     float* end_point_trajectory, float* trajectory, float* x_final,
                     float* transformer_features, void* stream);

/* ---- Procrustes: replaces procrustes.py:40-118 --------------------------------------------------
 * cu_seqlens_part: device int32 [n_parts+1] over the packed (T,3) clouds; slot: device int32 [n_parts]
 * giving b*max_parts+p of each non-empty part (NULL = identity).  rotations (slots,3,3) / translations
 * (slots,3) must be zero-initialised by the caller so empty slots read as zeros (procrustes.py:71-76).
 * n_points = rows of the packed clouds: the reference asserts sum(points_per_part) == N on the host (utils/point_clouds.py:41-44); here
 * a part reaching past n_points is skipped on the device (nothing read or written out of bounds) and flagged in the error word that
 * rap_device_error() returns (code 0xB1).  n_points <= 0 disables the check. */
int rap_procrustes_fit(const float* source, const float* target, const int32_t* cu_seqlens_part, const int32_t* slot,
                       int n_parts, int64_t n_points, float* rotations, float* translations, void* stream);
/* out = (R cond + t) * w_rigid + mix * w_mix per part, (R,t) fitted cond -> prediction; mix may be NULL.
 * rigidify_prediction_with_procrustes is w_rigid=1, mix=NULL; sampler.py:60 is w_rigid=1-t+dt, w_mix=t-dt. */
int rap_procrustes_rigidify(const float* prediction, const float* condition, const int32_t* cu_seqlens_part, int n_parts, int64_t n_points,
                            float* out, const float* mix, float w_rigid, float w_mix, void* stream);
/* Generation selection (modeling.py:456-547, eval/metrics.py:511-622): per slot sum of squared residuals |R src + t - tgt|^2, either
 * fused with the fit (one launch per trajectory step instead of fit_transformations + compute_rigidity_rmse host loops) or for a
 * given (R,t).  sse (slots) must be zero-initialised by the caller. */
int rap_procrustes_fit_sse(const float* source, const float* target, const int32_t* cu_seqlens_part, const int32_t* slot, int n_parts,
                           int64_t n_points, float* rotations, float* translations, float* sse, void* stream);
int rap_rigid_sse(const float* source, const float* target, const int32_t* cu_seqlens_part, const int32_t* slot, int n_parts,
                  int64_t n_points, const float* rotations, const float* translations, float* sse, void* stream);
/* Plan-bound convenience wrappers using the plan's own part table. */
int rap_plan_fit(rap_plan_t* plan, const float* source, const float* target, float* rotations, float* translations, void* stream);
int rap_plan_rigidify(rap_plan_t* plan, const float* prediction, const float* condition, float* out, void* stream);

/* ---- local feature extractor: replaces dataset_process.extract_sample_features.FeatureExtractor / MiniSpinNet.forward
 * (dataset_process/extract_sample_features.py:151-220, dataset_process/utils/spinnet/patch_embedder.py:49-91) ------------------
 * Weights: the MiniSpinNet state_dict (the `Desc.`-stripped keys, fp32 device tensors; num_batches_tracked is not needed).
 * extract: points (n_points,3) f32, perm int32 [n_points] = the shuffle the reference draws with np.random.choice (NULL = identity),
 * keypoints (n_keypoints,3) f32, des_r = descriptor radius -> desc (n_keypoints,32) f32, L2-normalised. */
typedef struct rap_spinnet rap_spinnet_t;
int rap_spinnet_create(rap_spinnet_t** out);
void rap_spinnet_destroy(rap_spinnet_t* h);
int rap_spinnet_num_keys(const rap_spinnet_t* h);
const char* rap_spinnet_key(const rap_spinnet_t* h, int index, int64_t* numel_out);
int rap_spinnet_load_weight(rap_spinnet_t* h, const char* key, const float* src, int64_t numel, void* stream);
int rap_spinnet_finalize(rap_spinnet_t* h);
int rap_spinnet_extract(rap_spinnet_t* h, const float* points, int n_points, const int32_t* perm, const float* keypoints, int n_keypoints,
                        float des_r, float* desc_out, void* stream);

int rap_test_spinnet_patches(const float* points, int n_points, const int32_t* perm, const float* keypoints, int n_keypoints, float des_r,
                             float* patches_out, void* stream);

/* ---- keypoint selection in front of the extractor ------------------------------------------------------------
 * All pointers are device pointers unless noted; the calls synchronise the stream (they return counts to the host).
 *
 * rap_voxel_downsample  replaces dataset_utils.voxel_down_sample_torch (dataset_process/utils/dataset_utils.py:279-322): one index per
 *                       occupied voxel (point nearest the voxel centre in 1000 quantised steps, ties to the lowest index), ordered by
 *                       the reference's flattened voxel id.  indices_out: capacity n (int64, device); count_out: host.
 * rap_voxel_coverage    replaces calculate_voxel_coverage (dataset_process/utils/point_sampling_utils.py:11-31): occupied voxel count of
 *                       floor(points / voxel_size) evaluated in the array's own precision (is_f64 selects double input).
 * rap_fps_batched       replaces pytorch3d sample_farthest_points(points (N,P,3), lengths, K, random_start_point) as called by
 *                       apply_batched_fps (point_sampling_utils.py:263-309); start_idx holds the per-cloud start indices (the host draws
 *                       them with torch.randint like the reference).  indices_out (N,max_k) int64, -1 padded; sampled_out (N,max_k,3)
 *                       optional.  One thread-block cluster per cloud, points register-resident: at most rap_fps_capacity() per cloud.
 * rap_statistical_outliers replaces open3d PointCloud.remove_statistical_outlier(nb_neighbors, std_ratio) as called at
 *                       dataset_process/extract_sample_features.py:360-413: keep_out (n) uint8 mask (device), n_keep_out host. */
int rap_voxel_downsample(const float* points, int n, float voxel_size, int64_t* indices_out, int* count_out, void* stream);
int rap_voxel_coverage(const void* points, int n, int is_f64, double voxel_size, int* count_out, void* stream);
int rap_fps_capacity(void);
int rap_fps_batched(const float* points, int n_clouds, int max_points, const int32_t* lengths, const int32_t* k, const int32_t* start_idx,
                    int max_k, int64_t* indices_out, float* sampled_out, void* stream);
int rap_statistical_outliers(const float* points, int n, int nb_neighbors, double std_ratio, uint8_t* keep_out, int* n_keep_out, void* stream);

/* Evaluation metrics with ground truth: squared distance of every query point to its nearest target point of the same
 * segment (cu_seqlens: device, n_segments + 1 entries shared by both packed clouds), optionally only over targets with a DIFFERENT label.
 * Building block of compute_cd (reference eval/metrics.py:14-48) and compute_overlap_ratio (:625-691).  All pointers device. */
int rap_segment_nn(const float* query, const float* target, const int32_t* cu_seqlens, int n_segments, const int32_t* label_query,
                   const int32_t* label_target, float* min_dist2_out, void* stream);

/* Pose application (demo.py:1330-1340, open3d PointCloud.transform): out = R p + t over a full-resolution cloud,
 * fp32 or fp64 points (device), the affine 4x4 row-major in HOST memory.  In-place (out == points) is allowed. */
int rap_apply_transform(const void* points, long long n, int is_f64, const double* transform_4x4_host, void* out, void* stream);

/* Timing aid (tools/fmha_split_sweep.py): mean duration of `iters` back-to-back attention launches on the default stream. */
int rap_bench_fmha(const void* qkv, int T, int heads, const int32_t* cu_seqlens_host, int n_segments, void* out, int iters, float* ms_per_launch);
/* Same for a named kernel variant: -1 what the product launches, 0 one item per CTA (default), 1 persistent stream-K (whole waves dealt round-robin),
 * 2 persistent with one contiguous range per CTA, 3 three query tiles per CTA on 64-key blocks (RAP_FMHA_NT=3). */
int rap_bench_fmha_variant(const void* qkv, int T, int heads, const int32_t* cu_seqlens_host, int n_segments, void* out, int iters, int variant,
                           float* ms_per_launch);

/* Host-only helper (no GPU): the attention work list of one launch -- (query block, head) units, the units of the last partially filled
 * wave cut along the key axis so every SM finishes together; 8 int32 per item {q_row0, seg_start, seg_end, head, kv_lo, kv_hi, slot,
 * piece | n_pieces << 8}.  Returns the number of items (items_out may be NULL to query it). */
int rap_fmha_schedule(const int32_t* cu_seqlens_host, int n_segments, int heads, int force_split, int32_t* items_out, int capacity,
                      int* n_slots_out, int* max_pieces_out);

/* Host-only helper (no GPU): the PERSISTENT attention schedule (opt-in kernel, RAP_FMHA_PERSIST=1) -- one CTA per SM, each walking a contiguous,
 * equal-cost range of the linearised (query block x head, key block) space; units that straddle a CTA boundary are cut along the keys and
 * recombined by the piece that finishes last.  items as above, ordered by CTA; cta_ptr_out[n_ctas_used + 1] = item range of every CTA.
 * n_ctas <= 0: the SM count of the current device (148 without one).  Returns the number of items. */
int rap_fmha_schedule_persistent(const int32_t* cu_seqlens_host, int n_segments, int heads, int n_ctas, int force_split, int32_t* items_out,
                                 int capacity, int32_t* cta_ptr_out, int cta_capacity, int* n_ctas_out, int* n_slots_out, int* max_pieces_out);

/* Host-only helper (no GPU): the 3x3 Kabsch solve used by the kernel, for CPU unit tests. */
int rap_kabsch_host(const double* H9, double* R9);

/* Kernel launch counter (incremented by every kernel this library launches; for bench.py's gpu_launches). */
int64_t rap_launch_count(void);

/* Per-kernel-class device timing (CUDA events around every launch on the launching stream) between
 * rap_profile_begin and rap_profile_end (host-synchronising); arrays have rap_profile_classes() entries. */
int rap_profile_begin(void);
int rap_profile_end(float* total_ms_per_class, int* launches_per_class);
int rap_profile_classes(void);
const char* rap_profile_class_name(int kernel_class);

/* Micro entry points used by tests/ and profiling to exercise one kernel at a time.  rap_test_gemm: epi = 0 f32 bias+residual, 1 f16 SiLU,
 * 2 QKV + RMS-norm, 3 GEGLU, 5 the in-place residual form out32 += A W^T + bias through bulk tensor reductions (addend must equal out32, no out16);
 * epi | 0x100 runs the 2-CTA-cluster variant that shares the weight tile by TMA multicast (BN = 256), | 0x300 the 2-CTA UMMA variant, | 0x200
 * forbids both (default: RAP_GEMM_MC / RAP_GEMM_2CTA, off); | 0x400 ReLU (epi 1) / the scalar epilogue forms (epi 2, 3); | 0x800 fp16 tiles written by
 * bulk tensor stores (epi 2, 3).  dbg: NULL or a device buffer of 128 + 4 * 148 int64 (clock64 timeline of CTA 0 in
 * [tile][8], then per CTA {globaltimer, clock64} at entry and exit). */
int rap_test_gemm(int epi, int bn, int M, int N, int K, const void* A_f16, const void* W_f16, const float* bias,
                  const float* addend, float* out32, void* out16, const float* gamma, void* dbg, void* stream);
/* The residual GEMM with the next block's LayerNorm in its epilogue (N = 512; the two CTAs of a 128-row block exchange row moments over
 * distributed shared memory): out32 = A W^T + bias + addend, ln_out (fp16) = LN(out32) * (one_plus + ln_mul[sel]) + ln_add[sel],
 * sel(row) = row_of_tok[row] * row_stride.  Replaces AdaptiveLayerNorm / ff_norm as separate kernels (norm.py:76, layer.py:163). */
int rap_test_gemm_ln(int M, int K, const void* A_f16, const void* W_f16, const float* bias, const float* addend, float* out32,
                     const float* ln_mul, const float* ln_add, const int32_t* row_of_tok, int row_stride, float one_plus, void* ln_out_f16,
                     void* stream);
/* dbg: optional device buffer of n_kv_blocks*16 int64 receiving a clock64() timeline of CTA (0,0), or NULL. */
int rap_test_fmha(const void* qkv_f16, int T, int heads, const int32_t* cu_seqlens_host, int n_segments, int force_split, void* out_f16,
                  void* dbg, void* stream);

int rap_test_ln_modulate(const float* h, void* out_f16, int T, const float* mul, const float* add, int row_stride,
                         const int32_t* row_of_tok, float one_plus, void* stream);
/* Debug/test: copy a plan-internal device buffer to `dst` (device). 0 scene_of_tok, 1 AdaLN table, 2 h, 3 h_static,
 * 4 x16, 5 qkv, 6 attn16. */
int rap_debug_read(rap_plan_t* plan, int which, void* dst, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* RAP_B200_H */
