/* spr_b200.h -- C ABI of libspr_b200.so, the sm_100a implementation of the Superpoints_Registration
 * hot path (KPConv preprocessing, KPConv layer, superpoint matching, pose solve).
 *
 * Conventions
 *   - every pointer prefixed d_ is a DEVICE pointer; the caller (PyTorch in the shipped host code)
 *     owns every buffer, including workspaces: the library never allocates or frees device memory
 *     and keeps no state between calls;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - every call returns 0 on success, a negative SPR_E* code on failure; spr_last_error() gives the
 *     message of the last failure on the calling thread.  The reference's extensions raise
 *     RuntimeError in the same situations (cpp_neighbors/wrapper.cpp:77,95,133,203;
 *     cpp_subsampling/wrapper.cpp:266-270) and the Python host layer does the same;
 *   - stacked-cloud layout as in the reference: points of all clouds concatenated [N,3] fp32 row-major,
 *     plus per-cloud lengths int32 [B] (kpconv.py:322-323);
 *   - neighbour matrices are row-major [Nq, row_stride] with the shadow value Ns_total marking padding
 *     (neighbors.cpp:321-327), element type int64 (the reference's collate dict dtype, kpconv.py:396-398)
 *     or int32, selected by idx_is_64.
 *
 * Paths are relative to /root/reference/src/.
 */
#ifndef SPR_B200_H_
#define SPR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPR_API __attribute__((visibility("default")))
#else
#define SPR_API
#endif

#define SPR_OK 0
#define SPR_EINVAL (-1)  /* bad argument (empty input, limit out of range, ...) */
#define SPR_ECUDA (-2)   /* a CUDA runtime call or launch failed */
#define SPR_ENOSPACE (-3) /* caller-provided workspace too small */
#define SPR_EUNSUPPORTED (-4)

#define SPR_MAX_NEIGHBOR_LIMIT 128

SPR_API int spr_version(void);
SPR_API const char* spr_last_error(void);
/* Number of kernels this library has launched from the calling process (all threads); bench.py reports
 * the delta over the timed region as gpu_launches. */
SPR_API unsigned long long spr_launch_count(void);
/* Sticky numeric flags of the current device, raised by kernels that write fp16 (hi, lo) operand images for the
 * tensor-core GEMMs (spr_gemm_prepare_input, spr_layernorm256_prepare, spr_instance_norm_lrelu_ex, spr_attention_varlen
 * with an image output, spr_gemm_tc with plane / image output).  SPR_FLAG_FP16_OVERFLOW: an activation times the
 * image scale exceeded the fp16 range (|x| > 65504 / scale, i.e. ~4094 at the host layer's scale of 16, or infinite);
 * the affected outputs are NaN (never silently wrong).  A NaN input is not flagged: it propagates as NaN.  Synchronises the device.  reset != 0 clears them. */
#define SPR_FLAG_FP16_OVERFLOW 1u
SPR_API unsigned int spr_numeric_flags(int reset);

/* ---------------------------------------------------------------------------------------------
 * Grid (voxel) subsampling.
 * Replaces grid_subsampling.subsample_batch(points, batches, sampleDl=..., max_p=0)
 *   models/backbone_kpconv/cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333
 *   -> batch_grid_subsampling  cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211
 * (the features/classes variants are not on the path: kpconv.py:370 passes points only).
 * Barycentres are bit-identical to the reference's (sequential fp32 sum in input order times
 * (float)(1.0/count)); they are emitted per cloud in FIRST-OCCURRENCE order (voxels ordered by their
 * lowest member index) instead of the reference's std::unordered_map iteration order.
 *
 *   d_points [n_points,3] f32, d_lengths [n_clouds] i32
 *   d_out_points: room for n_points*3 f32; d_out_lengths [n_clouds] i32; d_out_total [1] i32
 * The caller reads d_out_total / d_out_lengths back to learn M.
 * ------------------------------------------------------------------------------------------- */
SPR_API size_t spr_grid_subsample_workspace_bytes(int n_points, int n_clouds);
SPR_API int spr_grid_subsample_batch(const float* d_points, const int32_t* d_lengths, int n_points, int n_clouds,
                             float sample_dl, float* d_out_points, int32_t* d_out_lengths, int32_t* d_out_total,
                             void* d_workspace, size_t workspace_bytes, void* stream);
/* Same binning with a choice of voxel lattice and representative:
 *   SPR_SUBSAMPLE_REFERENCE  the CPU reference (spr_grid_subsample_batch): per-cloud origin, barycentre = sum * (1/count)
 *   SPR_SUBSAMPLE_MEAN       voxel = floor(p / dl) on the global lattice, mean = sum / count: MinkowskiEngine's
 *                            UNWEIGHTED_AVERAGE quantisation as the reference's PreprocessorGPU uses it (kpconv.py:221-243)
 *   SPR_SUBSAMPLE_FIRST      same lattice, the first point (input order) of every voxel unchanged: the KITTI loader's
 *                            down-sampler (data_loaders/kitti_pred.py:12-14,203-204)
 * Voxels come out per cloud in first-occurrence order in every mode. */
#define SPR_SUBSAMPLE_REFERENCE 0
#define SPR_SUBSAMPLE_MEAN 1
#define SPR_SUBSAMPLE_FIRST 2
SPR_API int spr_grid_subsample_batch_ex(const float* d_points, const int32_t* d_lengths, int n_points, int n_clouds,
                                float sample_dl, int mode, float* d_out_points, int32_t* d_out_lengths,
                                int32_t* d_out_total, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched fixed-radius neighbour search.
 * Replaces radius_neighbors.batch_query(queries, supports, q_batches, s_batches, radius=...)
 *   models/backbone_kpconv/cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238
 *   -> batch_nanoflann_neighbors  cpp_neighbors/neighbors/neighbors.cpp:211-332
 * fused with the truncation the Python caller applies (kpconv.py:259-260, [:, :max_neighbors]).
 * Row i holds the supports of query i's cloud with fp32 d2 < radius^2, nearest first (ties by lower
 * index), at most `limit` of them, then the shadow value n_supports.  *d_out_max_count receives the
 * un-truncated batch-wide maximum in-radius count, i.e. the width the reference matrix had before
 * truncation (neighbors.cpp:296-297); min(max_count, limit) leading columns are the reference's output.
 *
 * Two-step form so that one cell grid over the supports serves several query sets (the conv and the
 * pool searches of a level share supports and radius, kpconv.py:353,380):
 *   spr_cell_grid_build   bins the supports of every cloud into a uniform grid with cell >= radius
 *   spr_radius_query      answers queries against a built grid
 * ------------------------------------------------------------------------------------------- */
SPR_API size_t spr_cell_grid_workspace_bytes(int n_supports, int n_clouds);
SPR_API int spr_cell_grid_build(const float* d_supports, const int32_t* d_s_lengths, int n_supports, int n_clouds, float radius,
                        void* d_grid_workspace, size_t workspace_bytes, void* stream);
/* order[i] = index of the i-th support in (cloud, z, y, x) cell order of a built grid: a spatially coherent walk of
 * the stacked clouds, used as the query processing order of the KPConv kernels (d_order) */
SPR_API int spr_cell_grid_order(const void* d_grid_workspace, int n_supports, int n_clouds, int32_t* d_order, void* stream);
SPR_API int spr_radius_query(const float* d_queries, const int32_t* d_q_lengths, int n_queries, int n_clouds,
                     const void* d_grid_workspace, int n_supports, float radius, int limit, void* d_out_idx,
                     int idx_is_64, int row_stride, int32_t* d_out_max_count, void* stream);
/* Same search with a choice of which in-radius supports a row keeps when there are more than `limit`:
 *   SPR_ORDER_NEAREST  the `limit` nearest, nearest first (the reference's CPU Preprocessor; what spr_radius_query does)
 *   SPR_ORDER_INDEX    the first `limit` in support-index order (pytorch3d.ops.ball_query as the reference's
 *                      PreprocessorGPU calls it, models/backbone_kpconv/kpconv.py:265-292; the matrix then keeps all
 *                      `limit` columns) */
#define SPR_ORDER_NEAREST 0
#define SPR_ORDER_INDEX 1
SPR_API int spr_radius_query_ex(const float* d_queries, const int32_t* d_q_lengths, int n_queries, int n_clouds,
                        const void* d_grid_workspace, int n_supports, float radius, int limit, int order, void* d_out_idx,
                        int idx_is_64, int row_stride, int32_t* d_out_max_count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * KPConv layer forward (rigid kernel, linear influence, sum aggregation -- the only mode any shipped
 * config uses: conf/qk_regtr_full_*.yaml KP_influence: linear, aggregation_mode: sum).
 * Replaces KPConv.forward(q_pts, s_pts, neighb_inds, x)  models/backbone_kpconv/kpconv_blocks.py:269-414
 *   out[n,:] = ( sum_k ( sum_h max(0, 1 - |s[idx[n,h]] - q[n] - kp[k]| / extent) * x[idx[n,h],:] ) @ W[k] )
 *              / max(1, #{h : rowsum(x)[idx[n,h]] > 0})
 * One fused kernel: neighbour-feature gather -> kernel-point influence -> (K*Cin)xCout contraction.
 *   d_q [nq,3], d_s [ns,3], d_idx [nq,row_stride] (first H columns used; value ns = shadow),
 *   d_x [ns,cin], d_w [K,cin,cout], d_kp [K,3], d_out [nq,cout]; all fp32 row-major.
 *   mode: 0 = fp32 CUDA-core contraction (parity anchor); 1 = tcgen05 tensor-core contraction with
 *         split-precision operands (when built in; SPR_EUNSUPPORTED otherwise).
 *   d_order (optional, [nq] i32): a permutation of the queries = the order they are processed in (results are written
 *         at the original rows); used by the Cin = 1 kernel, whose 32 queries per warp then share their gathers.
 * ------------------------------------------------------------------------------------------- */
SPR_API size_t spr_kpconv_workspace_bytes(int nq, int ns, int cin, int cout, int n_kernel_points);
SPR_API int spr_kpconv_forward(const float* d_q, const float* d_s, const void* d_idx, int idx_is_64, int row_stride, int H,
                       const float* d_x, int cin, const float* d_w, int cout, const float* d_kp, int n_kernel_points,
                       float extent, float* d_out, int nq, int ns, int mode, void* d_workspace, size_t workspace_bytes,
                       const int32_t* d_order, void* stream);

/* KPConv (tensor-core path) with operands prepared by the producing kernels instead of the built-in pre-pass:
 *   spr_kpconv_prepare_weights: [W_hi | W_lo] stage images + max|W| (one word), once per weight
 *   spr_kpconv_forward_prepared: d_pts4 [ns] float4, d_x16 [ns, c] (hi | lo<<16), d_amax_x (one word) as written
 *   by spr_instance_norm_lrelu_ex.  Cin = Cout = c in {32, 64, 128, 256}.  d_order (optional, [nq] i32): a
 *   permutation of the queries giving the processing order (results are written at the original rows). */
SPR_API size_t spr_kpconv_weight_image_bytes(int c);
SPR_API size_t spr_kpconv_scratch_bytes(int H, int c);   /* d_scratch of spr_kpconv_forward_prepared (0 for c = 32) */
SPR_API int spr_kpconv_prepare_weights(const float* d_w, int c, void* d_img, void* d_amax_w, void* stream);
SPR_API int spr_kpconv_forward_prepared(const float* d_q, const void* d_idx, int idx_is_64, int row_stride, int H,
                                const void* d_pts4, const void* d_x16, const void* d_amax_x, int c, const void* d_wimg,
                                const void* d_amax_w, const float* d_kp, float extent, float* d_out, int nq, int ns,
                                void* d_scratch, const int32_t* d_order, void* stream);

/* Per-cloud instance normalisation + LeakyReLU (+ optional residual add before the activation).
 * Replaces BatchNormBlock.forward with nn.InstanceNorm1d (kpconv_blocks.py:474-530: per cloud, per
 * channel, biased variance, eps, no affine, no running stats) followed by nn.LeakyReLU
 * (kpconv_blocks.py:556-561,645,727,741).
 *   y = lrelu( (x - mean_b,c) / sqrt(var_b,c + eps) [+ residual] ),  slope = 1 disables the activation,
 *   d_residual may be NULL.  d_x may alias d_out.  Deterministic: fp64 partial sums over fixed row chunks. */
SPR_API size_t spr_instance_norm_workspace_bytes(int n_rows, int n_clouds, int c);
SPR_API int spr_instance_norm_lrelu(const float* d_x, const int32_t* d_lengths, int n, int n_clouds, int c, float eps,
                            float slope, const float* d_residual, float* d_out, void* d_workspace,
                            size_t workspace_bytes, void* stream);

/* Same operator with format-aware outputs, so that the consumer of the normalised rows needs no conversion pass.
 * Any of: d_out_f32 (plain rows), d_out_img (operand image of the next spr_gemm_tc with K = c, see below),
 * d_out_x16 + d_out_pts4 + d_amax (pre-split feature rows, packed support points (x, y, z, +-2^-e) built from
 * d_points [n,3], and max|y|: the inputs of spr_kpconv_forward_prepared).  c must be a multiple of 32.
 *
 * d_stats16 (optional): [ceil(n/16), c] pairs (sum, sum of squares) of every 16-row block of d_x, as written by the
 * producer of d_x (spr_gemm_tc / spr_kpconv_forward_prepared `d_stats16`).  With it the statistics pass over d_x is
 * skipped: blocks inside a cloud come from d_stats16, the ragged rows at the two ends of each cloud from d_x.
 * d_residual_stats16 (optional, needs d_residual): the residual is itself a raw producer output with these block sums;
 * it is normalised with its own per-cloud statistics (no activation) inside the apply kernel before it is added --
 * the shortcut branch of a bottleneck block (kpconv_blocks.py:706-741) without a pass of its own.
 * x16_planar selects the layout of d_out_x16: 0 = one 32-bit word (hi | lo << 16) per channel (spr_kpconv_forward_prepared),
 * 1 = per group of 32 channels 32 hi halves then 32 lo halves (spr_kpconv_forward_gather, spr_kpconv_forward_staged). */
SPR_API int spr_instance_norm_lrelu_ex(const float* d_x, const int32_t* d_lengths, int n, int n_clouds, int c, float eps,
                               float slope, const float* d_residual, float* d_out_f32, void* d_out_img, float a_scale,
                               void* d_out_x16, void* d_out_pts4, const float* d_points, void* d_amax,
                               const float* d_stats16, const float* d_residual_stats16, int x16_planar,
                               void* d_workspace, size_t workspace_bytes,
                               void* stream);

/* Second-generation tensor-core KPConv (csrc/kpconv_g.cu): the neighbour features are gathered by the TMA engine
 * (cp.async.bulk.tensor tile::gather4 on the pre-split rows) and BOTH products of the layer -- influences x features per
 * query, then (K*Cin) x Cout per tile -- run on tcgen05.  Same operands and result as spr_kpconv_forward_prepared; its
 * own weight image (channel-major K order) and scratch.  c in {32, 64, 128}, H <= 64 (spr_kpconv_gather_supported). */
SPR_API int spr_kpconv_gather_supported(int c, int H);
SPR_API size_t spr_kpconv_gather_weight_image_bytes(int c);
SPR_API size_t spr_kpconv_gather_scratch_bytes(int c);
SPR_API int spr_kpconv_gather_prepare_weights(const float* d_w, int c, void* d_img, void* d_amax_w, void* stream);
SPR_API int spr_kpconv_forward_gather(const float* d_q, const void* d_idx, int idx_is_64, int row_stride, int H,
                              const void* d_pts4, const void* d_x16, const void* d_amax_x, int c, const void* d_wimg,
                              const void* d_amax_w, const float* d_kp, float extent, float* d_out, int nq, int ns,
                              void* d_scratch, const int32_t* d_order, void* stream);

/* Third-generation tensor-core KPConv (csrc/kpconv_s.cu): the structure of spr_kpconv_forward_prepared with the
 * neighbours' rows staged through per-warp shared-memory rings by asynchronous copies (two 8-neighbour blocks ahead) and
 * the warp-level MMA fragments read with ldmatrix.  Consumes the PLANAR pre-split rows (x16_planar = 1 of
 * spr_instance_norm_lrelu_ex) and its own weight image (channels of a pass permuted); no scratch.  c in {32, 64, 128,
 * 256}, H <= 96 (spr_kpconv_staged_supported).  Replaces the same reference operator, kpconv_blocks.py:269-414. */
SPR_API int spr_kpconv_staged_supported(int c, int H);
SPR_API size_t spr_kpconv_staged_weight_image_bytes(int c);
SPR_API int spr_kpconv_staged_prepare_weights(const float* d_w, int c, void* d_img, void* d_amax_w, void* stream);
SPR_API int spr_kpconv_forward_staged(const float* d_q, const void* d_idx, int idx_is_64, int row_stride, int H,
                                      const void* d_pts4, const void* d_x16, const void* d_amax_x, int c,
                                      const void* d_wimg, const void* d_amax_w, const float* d_kp, float extent,
                                      float* d_out, int nq, int ns, const int32_t* d_order, void* stream);

/* max_pool(x, inds)  kpconv_blocks.py:127-143: out[n,c] = max_h xpad[idx[n,h],c] where xpad has a zero
 * row appended for the shadow index.  d_order (optional, [nq] i32): a permutation of the pooled points giving the
 * processing order (results are written at the original rows; cell order keeps the gathers in cache). */
SPR_API int spr_max_pool(const float* d_x, const void* d_idx, int idx_is_64, int row_stride, int H, int nq, int ns, int c,
                 float* d_out, const int32_t* d_order, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Superpoint matching.  Replaces the per-pair body of RegTR.softmax_correlation
 *   models/qk_regtr_full.py:445-576 : corr = S T^T / sqrt(D); attn = softmax(corr, dim=-2) * softmax(corr, dim=-1);
 *   N > M : val,ind = max(attn, dim=1) (one source per target), else val,ind = max(attn, dim=2).
 * Batched over pairs with packed features: d_src [sum N_p, D], d_tgt [sum M_p, D], offsets [P+1] i32.
 *   d_corr: caller-provided scratch/result, packed per pair, pair p at d_corr_offsets[p] (i64), N_p*M_p f32.
 *   d_val / d_ind: packed per pair; pair p writes M_p entries if N_p > M_p, else N_p entries, starting at
 *   d_out_offsets[p] (i32 [P+1]).  max_n / max_m: largest N_p / M_p of the batch (grid sizing).
 *   d_attn (optional, may be NULL): same layout as d_corr, receives the dual-softmax matrix
 *   (the reference returns it as outputs['attn'], qk_regtr_full.py:295).
 * ------------------------------------------------------------------------------------------- */
SPR_API size_t spr_match_workspace_bytes(int total_src, int total_tgt, int n_pairs);
SPR_API int spr_dual_softmax_match(const float* d_src, const float* d_tgt, const int32_t* d_src_offsets,
                           const int32_t* d_tgt_offsets, const int64_t* d_corr_offsets, const int32_t* d_out_offsets,
                           int n_pairs, int total_src, int total_tgt, int D, int max_n, int max_m, float* d_corr,
                           float* d_attn, float* d_val, int64_t* d_ind, void* d_workspace, size_t workspace_bytes,
                           void* stream);

/* Sinkhorn soft assignment + weighted targets for the 3DMatch configuration.
 * Replaces qk_regtr_full.py:532-536 (affinity from the correlation) + utils/se3_torch.py:166-202 (sinkhorn with
 * slack) + :204-231 (perm = exp(log_alpha), weighted_t = perm @ tgt / (rowsum + 1e-6), weights = rowsum):
 *   score = clamp(corr, 0); aff = -(score - softplus(alpha)) / (exp(beta) + 0.02)
 * d_corr as produced by spr_dual_softmax_match.  Outputs packed per source point:
 *   d_weighted_tgt [total_src,3], d_weights [total_src]. */
SPR_API size_t spr_sinkhorn_workspace_bytes(int total_src, int total_tgt, int n_pairs);
SPR_API int spr_sinkhorn_weighted_targets(const float* d_corr, const int64_t* d_corr_offsets, const int32_t* d_src_offsets,
                                  const int32_t* d_tgt_offsets, int n_pairs, int total_src, int total_tgt, int max_n,
                                  int max_m, const float* d_tgt_xyz, float softplus_alpha, float exp_beta, int n_iters, int slack,
                                  float* d_weighted_tgt, float* d_weights, void* d_workspace, size_t workspace_bytes,
                                  void* stream);

/* The same Sinkhorn normalisation on a caller-supplied affinity (log_alpha) matrix per pair, packed like d_corr:
 * utils/se3_torch.py:166-202 `sinkhorn(log_alpha, n_iters, slack)` -> d_log_perm (packed N_p x M_p, optional) and
 * :204-239 `compute_rigid_transform_with_sinkhorn` up to the pose solve -> d_weighted_tgt / d_weights (optional, need
 * d_tgt_xyz).  At least one of the two outputs must be requested.  Workspace: spr_sinkhorn_workspace_bytes. */
SPR_API int spr_sinkhorn_affinity(const float* d_affinity, const int64_t* d_mat_offsets, const int32_t* d_src_offsets,
                          const int32_t* d_tgt_offsets, int n_pairs, int total_src, int total_tgt, int max_n, int max_m,
                          int n_iters, int slack, float* d_log_perm, const float* d_tgt_xyz, float* d_weighted_tgt,
                          float* d_weights, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched weighted Procrustes (Kabsch).  Replaces compute_rigid_transform(a, b, weights)
 *   utils/se3_torch.py:109-163 : w~ = w / max(sum w, 1e-6); weighted centroids; cov = (a-ca)^T ((b-cb) * w~);
 *   U,S,V = svd(cov); R = V U^T, third column of V negated when det(R) <= 0; t = -R ca + cb; out = [R | t].
 * Packed correspondences: d_a, d_b [total,3], d_w [total] (NULL = uniform weights, the `weights is None`
 * branch :145-150), pair p owns rows offsets[p]..offsets[p+1].  d_out [n_pairs,12] row-major 3x4.
 * Moments are accumulated in fp64 and the 3x3 SVD is a one-sided Jacobi in fp64, one warp per pair.
 * ------------------------------------------------------------------------------------------- */
SPR_API int spr_weighted_procrustes(const float* d_a, const float* d_b, const float* d_w, const int32_t* d_offsets, int n_pairs,
                            float* d_out, void* stream);

/* Gather rows: out[i,:] = src[base_offset(pair of i) + ind[i], :]  (the torch.gather at qk_regtr_full.py:478,589
 * that turns argmax indices into corresponding points).  d_src has n_src rows; a row number outside [0, n_src) is
 * clamped into range (never dereferenced out of bounds). */
SPR_API int spr_gather_rows3(const float* d_src, int n_src, const int64_t* d_ind, const int32_t* d_row_pair_base, int n_rows,
                     float* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optional correspondence refinements of RegTR.softmax_correlation (use_ratio_test / use_lgr / use_ransac; all off in
 * the shipped configs).  Pairs are packed back to back like everywhere else.
 *
 * spr_top2_ratio       RegTR.ratio_test, qk_regtr_full.py:370-384: per output element (one per target when
 *                      N_p > M_p, one per source otherwise) the two largest attention values v1 >= v2 over the other
 *                      axis; val = v1 if v2/v1 < lowe_thres else 0; ind = position of v1.  d_attn is the packed
 *                      attention written by spr_dual_softmax_match.  The reduced axis needs >= 2 entries (torch.topk
 *                      raises otherwise; the host mirror checks).
 * spr_inlier_reweight  RegTR.recompute_weights, :386-391: w_out = w * (|b - (R a + t)| < acceptance_radius), the pose
 *                      of each row's pair taken from d_poses [n_pairs, 12].  local_global_registration (:393-398)
 *                      alternates it with spr_weighted_procrustes num_refinement_steps times.
 * spr_select_hypothesis RegTR.ransac, :400-421, after the hypotheses are solved (spr_weighted_procrustes over the
 *                      sampled rows): d_loss[p, h] = mean_r |b_r - T_{p,h} a_r| over all rows of pair p, d_out[p] =
 *                      the first hypothesis with a strictly smaller loss than all before it, d_best[p] (optional)
 *                      its number.  d_poses is [n_pairs, n_hypotheses, 12].
 * ------------------------------------------------------------------------------------------- */
SPR_API int spr_top2_ratio(const float* d_attn, const int64_t* d_corr_offsets, const int32_t* d_src_offsets,
                           const int32_t* d_tgt_offsets, const int32_t* d_out_offsets, int n_pairs, int total_out,
                           float lowe_thres, float* d_val, int64_t* d_ind, void* stream);
SPR_API int spr_inlier_reweight(const float* d_a, const float* d_b, const float* d_w, const float* d_poses,
                                const int32_t* d_offsets, int n_pairs, int total, float acceptance_radius,
                                float* d_w_out, void* stream);
SPR_API int spr_select_hypothesis(const float* d_a, const float* d_b, const int32_t* d_offsets, int n_pairs,
                                  const float* d_poses, int n_hypotheses, float* d_loss, float* d_out, int32_t* d_best,
                                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cross-encoder building blocks (reference: models/transformer/transformers.py:18-259, nn.MultiheadAttention
 * with d_model 256 / 8 heads).  Tokens of all clouds are PACKED ([total_tokens, d]); the reference pads every
 * cloud to the longest one and masks (utils/seq_manipulation.py:6-48).
 *
 * spr_split_f16: fp32 matrix -> fp16 (hi, lo) planes, x = hi + lo to ~22 bits; the first n_scaled columns are
 *   multiplied by `scale` first (the query block of a packed QKV projection: log2(e)/sqrt(head_dim)).
 * spr_attention_varlen: O[q, h] = softmax_k(Q[q,h] . K[k,h]) V[k,h] over the key segment of each query tile.
 *   d_hi / d_lo: planes with `ld` halves per row holding Q (pre-scaled), K and V at column offsets q_col / k_col /
 *   v_col (+ head * head_dim).  d_tiles: n_tiles x int32[4] = {first query row, query rows in the tile (<= 64),
 *   first key row, key rows}; an entry with 0 query rows is skipped (padding of a fixed-size list).  d_out: fp32 [rows, out_ld].  Products are hi*hi + lo*hi + hi*lo on the tensor
 *   cores with fp32 accumulation (fp32-level accuracy); the soft-max is online, nothing N x M is materialised.
 * ------------------------------------------------------------------------------------------- */
SPR_API int spr_split_f16(const float* d_x, int rows, int cols, int ld_in, void* d_hi, void* d_lo, int ld_out,
                          int n_scaled, float scale, void* stream);
SPR_API int spr_attention_varlen(const void* d_hi, const void* d_lo, int ld, int q_col, int k_col, int v_col,
                                 int n_heads, int head_dim, const int32_t* d_tiles, int n_tiles, float* d_out,
                                 int out_ld, void* d_out_img, float img_scale, void* stream);
/* Same operator on the tcgen05 tensor cores (csrc/attention_tc.cu): a tile holds up to 128 query rows, the scores
 * and the per-tile P V product live in tensor memory, one soft-max thread per query row.  Same operands, same
 * split-precision products, same outputs; out_ld must be a multiple of 4. */
SPR_API int spr_attention_varlen_tc(const void* d_hi, const void* d_lo, int ld, int q_col, int k_col, int v_col,
                                    int n_heads, int head_dim, const int32_t* d_tiles, int n_tiles, float* d_out,
                                    int out_ld, void* d_out_img, float img_scale, void* stream);

/* Dense layers on the tcgen05 tensor cores with fp32-level accuracy (nn.Linear of the cross-encoder:
 * in/out projections of nn.MultiheadAttention and the FFN, transformers.py:184-245):
 *     Y[T, N] = act( X[T, K] W[N, K]^T + b ) (+ residual)
 * Operands are fp16 (hi, lo) pairs held in global memory as ready-made shared-memory images (K-major
 * SWIZZLE_128B tiles) that the kernel fetches with bulk async copies:
 *   A image (spr_gemm_a_image_bytes): written by spr_layernorm256_prepare (LayerNorm + positional embedding),
 *     spr_gemm_prepare_input (plain fp32 rows), spr_attention_varlen (d_out_img) or a previous spr_gemm_tc
 *     (out_mode 2); activations are multiplied by a power-of-two a_scale first.
 *   W image (spr_gemm_w_image_bytes): spr_gemm_prepare_weight, once per weight, w_scale a power of two.
 * spr_gemm_tc out_mode: 0 = fp32 rows [T, ld_out] (+ residual, ReLU), 1 = fp16 hi / lo planes [T, ld_out] with the
 * first n_scaled columns multiplied by col_scale, 2 = A image of the next GEMM (K_next = N) scaled by next_scale.
 * out_scale must be 1 / (a_scale * w_scale).  d_out may alias d_residual.  d_stats16 (optional, out_mode 0):
 * [ceil(T/16), N] pairs (sum, sum of squares) over each 16-row block of the rows written, for the InstanceNorm that
 * follows a UnaryBlock's Linear (spr_instance_norm_lrelu_ex d_stats16). */
SPR_API size_t spr_gemm_a_image_bytes(int T, int K);
SPR_API size_t spr_gemm_w_image_bytes(int N, int K);
SPR_API int spr_gemm_prepare_weight(const float* d_w, int N, int K, float w_scale, void* d_img, void* stream);
SPR_API int spr_gemm_prepare_input(const float* d_x, int T, int K, int ld, float a_scale, void* d_img, void* stream);
SPR_API int spr_layernorm256_prepare(const float* d_x, const float* d_gamma, const float* d_beta, const float* d_pos,
                                     int T, float eps, float a_scale, void* d_img, float* d_out_f32, void* stream);
SPR_API int spr_gemm_tc(const void* d_a_img, const void* d_w_img, const float* d_bias, const float* d_residual,
                        int ld_res, int T, int N, int K, float out_scale, int relu, int out_mode, void* d_out,
                        void* d_out_lo, int ld_out, int n_scaled, float col_scale, float next_scale, float* d_stats16,
                        void* stream);

/* The whole packed cross-encoder (transformers.py:18-259: n_layers pre-norm TransformerCrossEncoderLayer with
 * positional values + the final LayerNorm) issued from one call: per layer LN(+pos) -> QKV projection -> attention ->
 * output projection (+x), the same with the partner cloud's keys, LN -> FFN1 (ReLU) -> FFN2 (+x); 11 launches per layer
 * through the entry points above, nothing else -- it exists to take the host interpreter out of a launch-bound forward.
 * d_x [T, 256] is updated in place; d_out (optional) receives the final LayerNorm.  Host arrays, layer-major:
 *   ptrs[18 l + ..]: 0-5 norm1/2/3 weight, bias; 6-9 self-attention in_proj image, bias, out_proj image, bias; 10-13 the
 *                    same for the cross-attention; 14-17 linear1 image, bias, linear2 image, bias   (device pointers)
 *   scal[9 l + ..]:  0-2 eps of norm1/2/3; 3-6 weight-image scales (self in, self out, cross in, cross out); 7, 8
 *                    linear1, linear2 scales (the `scale` given to spr_gemm_prepare_weight)
 * Scratch, caller-owned: d_img (A image, K = 256), d_img_ffn (A image, K = d_ff), d_hi / d_lo (fp16 [T, 768] planes).
 * attention_generation: 2 = spr_attention_varlen_tc (tiles of <= 128 rows), 1 = spr_attention_varlen (<= 64). */
SPR_API int spr_cross_encoder_forward(float* d_x, const float* d_pos, int T, int d_model, int n_heads, int d_ff,
                                      int n_layers, const void* const* ptrs, const float* scal,
                                      const int32_t* d_sa_tiles, int n_sa_tiles, const int32_t* d_ca_tiles,
                                      int n_ca_tiles, void* d_img, void* d_img_ffn, void* d_hi, void* d_lo,
                                      float a_scale, int attention_generation, const float* d_final_gamma,
                                      const float* d_final_beta, float final_eps, float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPR_B200_H_ */
