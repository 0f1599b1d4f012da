/*
 * fairguide -- C ABI of the B200 (sm_100a) fairness-guidance kernels.
 *
 * This is the drop-in boundary.  The reference (sail-sg/finetune-fair-diffusion) is pure Python
 * with no FFI of its own: the interface it exposes for this path is the set of Python closures in
 * exp-*-/1-main-debias.py.  Each entry point below names the reference
 * function it replaces; the Python mirror of those closures lives in
 * finetune-fair-diffusion_b200/api.py and reaches these symbols through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every tensor is contiguous, row-major, device memory unless
 *     the parameter says "host";
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises, allocates or keeps
 *     global mutable state; calls on different streams with disjoint buffers are safe;
 *   - return value: 0 = launched; < 0 = argument error (FG_ERR_*); > 0 = cudaError_t of a launch;
 *   - `dtype` selects the element type of the `void*` image / probability tensors (fg_dtype_t);
 *   - absence is signalled the way the reference does it: -1 sentinels (boxes, preds, probs,
 *     targets, losses), never by an error.
 *
 * Reference citations: E1 = exp-1-debias-gender/1-main-debias.py, E3 = exp-3-debias-gender-race/
 * 1-main-debias.py, E4 = exp-4-debias-gender-race-age/1-main-debias.py.
 */
#ifndef FAIRGUIDE_H
#define FAIRGUIDE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_ABI_VERSION 1

#define FG_OK 0
#define FG_ERR_INVALID_ARG (-1)
#define FG_ERR_DTYPE (-2)
#define FG_ERR_LIMIT (-3)
#define FG_ERR_WORKSPACE (-4)

typedef enum { FG_F32 = 0, FG_BF16 = 1, FG_F16 = 2 } fg_dtype_t;

int fg_abi_version(void);
/* Static string for any value the functions below return (FG_ERR_* or a cudaError_t). */
const char* fg_error_string(int code);

/* ------------------------------------------------------------------ boxes ------------------
 * get_largest_face_app (E1:1292-1304) + expand_bbox (E1:238-265), batched.
 * boxes [n, max_faces, 4] float32 (x0,y0,x1,y1); counts [n] int32 (0 = no detection), NULL means
 * every image has exactly max_faces candidates.  Largest area clipped to [0, dim_max]^2, strict
 * '>' so the first maximum wins.  Output boxes_out [n,4] int64, rounded half-to-even in float32
 * exactly like the reference; rows without a face get `fill` (E1:1328-1332) and indicator 0. */
int fg_select_expand_boxes(const float* boxes, const int32_t* counts, int n, int max_faces, int dim_max,
                           double expand_coef, double target_ratio, int64_t fill,
                           int64_t* boxes_out, uint8_t* indicators_out, void* stream);

/* ------------------------------------------------------------------ crop / resize ---------
 * crop_face (E1:267-290) for every image of a batch, fused with transforms.Resize (E1:1905).
 * images [n,C,H,W]; boxes [n,4] int64; indicators [n] uint8 or NULL (= all faces present).
 * chips [n,C,chip_h,chip_w]: bilinear (align_corners=False, no antialias) resample of the box, the
 * part of the box outside the image reading `fill_value`; images without a face, or with an empty /
 * fully outside box, become all-`fill_value` chips.  small [n,C,small_h,small_w]: the same sampler
 * over the whole image.  Either output may be NULL.  One launch reads each image once from HBM. */
int fg_crop_resize_fwd(const void* images, int n, int C, int H, int W,
                       const int64_t* boxes, const uint8_t* indicators,
                       void* chips, int chip_h, int chip_w,
                       void* small, int small_h, int small_w,
                       float fill_value, int dtype, void* stream);

/* apply_grad_hook_face (E1:1584-1617, E3:1751-1784, E4:1823-1867) and gen_dynamic_weights
 * (E1:1619-1633, E3:1787-1803, E4:1870-1895) reduced to their per-image parameters.
 * bbox / bbox_ori [n,4] int64; targets_a / preds_ori_a [n] int64 for attribute a < n_attr (unused
 * slots NULL); hook_factors / weight_factors: HOST arrays of n_attr floats.  e1_rule != 0 selects
 * E1's branch in which a -1 target always takes the factor.
 * Outputs (each may be NULL): region [n,4] int32 = x0,y0,x1,y1 (half open, empty if x1<=x0) of
 * bbox ^ bbox_ori ^ image with the reference's python-slice semantics for a -1 original box;
 * scale [n] float32 = factor applied to the gradient inside the region (1 if bbox is all -1);
 * weights [n] float32 = dynamic loss weight. */
int fg_guidance_factors(const uint8_t* face_indicators, const int64_t* bbox, const int64_t* bbox_ori,
                        const int64_t* targets0, const int64_t* targets1, const int64_t* targets2,
                        const int64_t* preds_ori0, const int64_t* preds_ori1, const int64_t* preds_ori2,
                        const float* hook_factors, const float* weight_factors, int n_attr, int e1_rule,
                        int n, int H, int W, int32_t* region, float* scale, float* weights, void* stream);

/* Backward of the fused forward, into the image: for every pixel
 *   g_images = scale_in_region * resize_bwd(g_small) + crop_bwd(g_chips)
 * (the fairness branch is taken before the hook, E3:2103 vs E3:2106, so it is not scaled).
 * Gather formulation: deterministic, no atomics, g_images is written exactly once.
 * g_chips / g_small / region+scale may be NULL (term dropped / scale 1). */
int fg_image_grad(const void* g_chips, const void* g_small,
                  const int64_t* boxes, const uint8_t* indicators,
                  const int32_t* region, const float* scale,
                  void* g_images, int n, int C, int H, int W,
                  int chip_h, int chip_w, int small_h, int small_w, int dtype, void* stream);

/* Backward of apply_grad_hook_face on its own: g_out = g_in * (scale inside region, 1 outside). */
int fg_region_scale(const void* g_in, const int32_t* region, const float* scale, void* g_out,
                    int n, int C, int H, int W, int dtype, void* stream);

/* ------------------------------------------------------------------ attribute head --------
 * torchvision MobileNetV3 `classifier` (mobilenetv3.py:190-216 as used at E1:929-935):
 *   logits = W2 * hardswish(W1 * pooled + b1) + b2      (Dropout is identity in eval mode)
 * pooled [m,d_in], w1 [d_hid,d_in], b1 [d_hid], w2 [k_head,d_hid], b2 [k_head] in `dtype`;
 * hidden_pre [m,d_hid] (dtype) receives the pre-activation for the backward; logits [m,k_head] f32. */
size_t fg_head_workspace_bytes(int m, int d_in, int d_hid, int k_head, int dtype);
int fg_head_fwd(const void* pooled, const void* w1, const void* b1, const void* w2, const void* b2,
                int m, int d_in, int d_hid, int k_head, void* hidden_pre, float* logits,
                void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* g_pooled [m,d_in] (dtype) from g_logits [m,k_head] f32; weights are frozen (no weight grads). */
int fg_head_bwd(const float* g_logits, const void* hidden_pre, const void* w1, const void* w2,
                int m, int d_in, int d_hid, int k_head, void* g_pooled,
                void* workspace, size_t workspace_bytes, int dtype, void* stream);

/* get_face_gender (E1:1355-1401), get_face_gender_race (E3:1387-1457), get_face_gender_race_age
 * (E4:1378-1475) after the classifier: per-attribute logit slice, softmax, argmax, scatter into
 * `fill`-initialised per-image rows.
 * logits [m,k_head] f32; src_row [n] int32 = row of `logits` for image i, or NULL for i;
 * selector [n] uint8 or NULL (= all selected); col_start / width: HOST arrays [n_attr] (E1: {40},{2};
 * E3: {0,2},{2,4}; E4: {0,2,6},{2,4,2}).  Outputs are attribute-major and contiguous per attribute:
 * preds [n_attr][n] int64; probs / logits_out: attribute a starts at element n*sum(width[<a]) and is
 * [n,width[a]] in `dtype`. */
int fg_head_attributes(const float* logits, int m, int k_head, const int32_t* src_row, const uint8_t* selector,
                       int n, int n_attr, const int32_t* col_start, const int32_t* width, float fill,
                       int64_t* preds, void* probs, void* logits_out, int dtype, void* stream);

/* Backward of fg_head_attributes (the autograd edge probs / sliced logits -> classifier logits that the reference's
 * micro-batch loop differentiates through, E3:2104 -> E3:2119-2122): g_logits_full [m,k_head] f32 receives, in the
 * columns of attribute a, g_logits_attr[a] + softmax-backward(g_probs[a]); every other element is zero.
 * probs: the forward's `probs` output; g_probs / g_logits_attr: HOST arrays [n_attr] of device pointers to [n,width[a]]
 * tensors in `dtype` (NULL entries = no gradient through that output). */
int fg_head_attributes_bwd(const void* probs, const void* const* g_probs, const void* const* g_logits_attr,
                           const int32_t* src_row, const uint8_t* selector, int n, int m, int k_head, int n_attr,
                           const int32_t* col_start, const int32_t* width, float* g_logits_full, int dtype, void* stream);

/* ------------------------------------------------------------------ fairness loss ---------
 * CE_loss(logits[idx], targets[idx]) on idx = face & target != -1, `fill` elsewhere
 * (E1:1912-1915, E3:2114-2122, E4:2238-2251).  logits [n,k] dtype; loss [n] dtype. */
int fg_fair_ce_fwd(const void* logits, const int64_t* targets, const uint8_t* face_indicators,
                   int n, int k, float fill, void* loss, int dtype, void* stream);
/* g_logits [n,k] = g_loss[i] * (softmax(logits[i]) - onehot(target_i)) on active rows, else 0. */
int fg_fair_ce_bwd(const void* logits, const int64_t* targets, const uint8_t* face_indicators,
                   const void* g_loss, int n, int k, void* g_logits, int dtype, void* stream);

/* The three calls above for every attribute of a head, the per-image loss assembly and d(mean loss)/d(head logits)
 * in ONE launch (E3:2114-2147, E4:2238-2283):
 *   loss_fair[a][i] = CE_a or `fill`;   loss[i] = sum_a loss_fair[a][i] + weight_img * dyn_weights[i] *
 *   (loss_clip[i] + loss_dino[i]) + weight_face * loss_face[i]   (fp32, left to right; the semantic / face terms
 *   are skipped when their pointers are NULL);   g_logits [n,k_head] f32 = g_coef * (softmax - onehot) in the
 *   columns [col_start[a], col_start[a]+width[a]) of active rows, 0 elsewhere (rounded through `dtype` like
 *   fg_fair_ce_bwd).  logits_attr / targets: HOST arrays of n_attr DEVICE pointers ([n,width[a]] dtype and [n]
 *   int64); width / col_start: HOST arrays; loss_fair [n_attr][n] dtype; loss [n] f32 or NULL; g_logits or NULL. */
int fg_fair_loss_fused(const void* const* logits_attr, const int64_t* const* targets, const int32_t* width,
                       const int32_t* col_start, int n_attr, const uint8_t* face_indicators, int n, int k_head,
                       float fill, float g_coef, const float* dyn_weights, const void* loss_clip,
                       const void* loss_dino, const void* loss_face, float weight_img, float weight_face,
                       void* loss_fair, float* loss, float* g_logits, int dtype, void* stream);

/* ------------------------------------------------------------------ assignment ------------
 * generate_dynamic_targets (E1:1403-1447): rank split at N*target_ratio and binomial-CDF
 * uncertainty, N = rows of probs [n_all,2] without a -1.  Ties in probs[:,1] resolve by row index.
 * targets [n_all] int64 (-1 for skipped rows); uncertainty [n_all] dtype or NULL.  If
 * threshold >= 0, targets whose uncertainty exceeds it are set to -1 as at E1:1835. */
size_t fg_rank_binom_workspace_bytes(int n_all);
int fg_assign_rank_binom(const void* probs, int n_all, double target_ratio, float threshold,
                         int64_t* targets, void* uncertainty, void* workspace, size_t workspace_bytes,
                         int dtype, void* stream);

/* generate_dynamic_targets_gender_race (E3:1459-1569; probs_age NULL, K = 8) and
 * generate_dynamic_targets_gender_race_age (E4:1477-1615; K = 16), split at the all-reduce:
 *
 * fg_ot_plan_counts: this rank's S Monte-Carlo draws.  rand_* [S,n_valid] in `dtype` are the uniform
 * draws (the caller draws them with torch.rand in the reference's order so seeded runs match);
 * n_valid = rows with a face, which the caller knows (it fixed the shape of rand_*).  For every draw
 * the class histogram b is formed and the exact transport problem ot.emd(ones, b, M) is solved on the
 * device; counts [n_valid,K] int32 receives the number of draws that sent row i to class j
 * (= target_probs before E3:1534).  Integer, so the cross-rank sum is order independent.
 *
 * fg_ot_targets: epilogue E3:1536-1565 on the rank-summed counts, in `dtype` arithmetic with the
 * reference's operation order; optional thresholding (E3:2022-2023) when threshold >= 0.
 * Outputs [n_all]: targets_* int64, unc_* dtype (NULL to skip); the age pair is used when K = 16. */
size_t fg_ot_workspace_bytes(int n_all, int K, int S);
int fg_ot_plan_counts(const void* probs_gender, const void* probs_race, const void* probs_age, int n_all,
                      const void* rand_gender, const void* rand_race, const void* rand_age, int S, int n_valid,
                      int32_t* counts, void* workspace, size_t workspace_bytes, int dtype, void* stream);
int fg_ot_targets(const int32_t* counts, const void* probs_gender, const void* probs_race, int n_all, int n_valid,
                  int K, float threshold,
                  int64_t* targets_gender, void* unc_gender, int64_t* targets_race, void* unc_race,
                  int64_t* targets_age, void* unc_age, void* workspace, size_t workspace_bytes,
                  int dtype, void* stream);

/* Test hook: exact assignment for ONE demand vector b (host array [K] int64) on a given cost matrix
 * M [n,K] float64 (device); assign [n] int32.  Exercises the same solver kernel as fg_ot_plan_counts. */
int fg_ot_solve_single(const double* M, int n, int K, const int64_t* b_host, int32_t* assign,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Test hook: the cost matrix the solver uses, M [n_valid,K] float64, rows in compacted order. */
int fg_ot_cost_matrix(const void* probs_gender, const void* probs_race, const void* probs_age, int n_all,
                      int n_valid, double* M, void* workspace, size_t workspace_bytes, int dtype, void* stream);

/* ------------------------------------------------------------------ next rows -
 * f2: detector staging (E1:1317 + 1326, same in E3 / E4).  out_bgr_hwc [n,H,W,3] uint8 =
 *   ((images*0.5 + 0.5)*255) with every operation rounded through `dtype` like the eager expression, truncated to
 *   uint8 like numpy's astype (out-of-range values wrap modulo 256, non-finite -> 0), permuted to HWC and swapped
 *   to BGR -- the array face_app.get() receives.  images [n,3,H,W] dtype. */
int fg_stage_detector_input(const void* images, int n, int C, int H, int W, uint8_t* out_bgr_hwc, int dtype, void* stream);

/* f3: get_evaluate_metrics (E3:1716-1749; E4:1780-1821 when probs_age != NULL).  probs_* [n,2] / [n,4] / [n,2] dtype,
 * rows of -1 are skipped.  out (DEVICE, fp64): [0] gender_gap [1] gender_pred_below_08 [2] race_gap
 * [3] race_pred_below_08 [4] gender_race_gap, and with age [5] age0_freq [6] age1_freq [7] age_pred_below_08 [8] age_gap.
 * One launch, no host synchronisation (the reference makes one blocking .item() per number).
 * probs_gender == NULL selects exp-6's race-only get_evaluate_metrics (exp-6-debias-race/1-main-debias.py:1624-1638):
 * out [0..3] race0..3_freq [4] race_gap [5] race_pred_below_08. */
int fg_bias_metrics(const void* probs_gender, const void* probs_race, const void* probs_age, int n, double* out,
                    int dtype, void* stream);

/* f3: generate_dynamic_targets_race (exp-6-debias-race/1-main-debias.py:1413-1482), the enumerated-composition
 * assignment of the race-only experiment.  The caller enumerates the compositions (n1..n4) of N = n_valid, weights them
 * with the normalised multinomial coefficients, sorts by weight and keeps the 95 % head (E6:1438-1459; host code in
 * the reference as well): demands [S,16] int32 (classes 4..15 zero), weights [S] float64, both on the DEVICE.
 * Per composition the exact transport problem ot.emd(ones, b, M) with M = ot.dist(probs, eye(4), "euclidean")
 * (E6:1461-1464) is solved by the solver of fg_ot_plan_counts; the plans are accumulated with their weights in
 * the given order in fp64, rows L1-normalised, argmax / 1 - max taken (E6:1465-1472) and scattered to
 * targets [n_all] int64 / uncertainty [n_all] dtype (NULL to skip), -1 for rows without a face.  threshold >= 0
 * applies targets[uncertainty > threshold] = -1. */
size_t fg_race_workspace_bytes(int n_all, int S);
int fg_assign_race_enumerated(const void* probs_race, int n_all, int n_valid, const int32_t* demands,
                              const double* weights, int S, float threshold, int64_t* targets, void* uncertainty,
                              void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* Test hook: the E6 cost matrix, M4 [n_valid,4] float64, rows in compacted order. */
int fg_race_cost_matrix(const void* probs_race, int n_all, int n_valid, double* M4, void* workspace,
                        size_t workspace_bytes, int dtype, void* stream);

/* f1: aligned 112x112 face chip, image_pipeline (E1:292-312) for a batch.
 *
 * fg_align_matrices: landmarks [n,5,2] float32 (the detector's `kps`, image pixels); indicators [n] uint8 or NULL.
 *   Least-squares similarity of the landmarks onto the 112x112 five-point template (skimage SimilarityTransform.estimate,
 *   E1:304-305; closed form in 2-D), rounded to float32 like `.to(img.dtype)` (E1:307).  M_out [n,2,3] float64 or NULL
 *   receives that matrix (-1 rows without a face); params [n,16] float32 (fg_align_params_bytes) receives what the two
 *   warp kernels need: the composed `output pixel -> source pixel` affine map of kornia.warp_affine(align_corners=False)
 *   (pixel<->[-1,1] with 2/(size-1), grid_sample un-normalisation with align_corners=False), its inverse and the bounding
 *   box of the taps.
 * fg_aligned_warp_fwd: images [n,C,Hs,Ws] -> out [n,C,Hd,Wd]; bilinear, taps outside the image read -1 (zero padding in
 *   the 0..255 domain of E1:293); rows without a face are filled with `fill` (E1:1332).
 * fg_aligned_warp_bwd: g_out [n,C,Hd,Wd] -> g_images [n,C,Hs,Ws], C <= 4; gather form (no atomics).  accumulate = 0
 *   overwrites the whole gradient (zeros away from the face), accumulate = 1 adds to what fg_image_grad wrote. */
size_t fg_align_params_bytes(int n);
int fg_align_matrices(const float* landmarks, const uint8_t* indicators, int n, int Hs, int Ws, int Hd, int Wd,
                      double* M_out, float* params, void* stream);
int fg_aligned_warp_fwd(const void* images, int n, int C, int Hs, int Ws, const float* params, const uint8_t* indicators,
                        void* out, int Hd, int Wd, float fill, int dtype, void* stream);
int fg_aligned_warp_bwd(const void* g_out, int n, int C, int Hd, int Wd, const float* params, const uint8_t* indicators,
                        void* g_images, int Hs, int Ws, int accumulate, int dtype, void* stream);

/* f1: face-realism loss.  Feature rows are [n,d] (d = 512 for the reference's SFNet), d % 4 == 0.
 *
 * fg_feats_normalize_fwd / _bwd: the tail of get_face_feats (E1:1186-1189): raw [n,d] dtype (net(x) + net(flip x)) ->
 *   f [n,d] float32 = raw / max(|raw|, 1e-12), inv_norm [n] float32 (or NULL); the backward maps g_f to g_raw (dtype).
 * fg_face_search_top1: FaceFeatsModel.semantic_search (E1:96-117), top-1 dot product.  queries [m,d] float32, selector [m]
 *   uint8 or NULL, db [D,d] float32 (L2-normalised rows, E1:88) -> best_row [m] int64 (-1 where not selected; the lowest
 *   row wins ties), similarity [m] float32 or NULL (-1 where not selected).
 * fg_face_loss_fwd: loss_face of E1:1917-1929 / E3:2124-2143 / E4:2253-2272 for the whole micro-batch in one call:
 *   normalise raw_feats; a row with a face whose targets all equal the original predictions with confidence
 *   max(probs_ori) >= confidence_level takes the original image's features feats_ori [n,d] float32 as target; the other
 *   rows with a face (search_needs_target = 1, E1: and with a target) take the nearest database row; loss = 1 - <f, target>
 *   in `dtype`, `fill` elsewhere.  targets / preds_ori: HOST arrays of n_attr device pointers ([n] int64 each);
 *   probs_ori: host array of device pointers [n,widths[a]] dtype.
 * fg_face_loss_bwd: g_loss [n] dtype -> g_raw_feats [n,d] dtype, reading the forward's state from the same workspace. */
int fg_feats_normalize_fwd(const void* raw, int n, int d, float* f, float* inv_norm, int dtype, void* stream);
int fg_feats_normalize_bwd(const float* g_f, const float* f, const float* inv_norm, int n, int d, void* g_raw, int dtype,
                           void* stream);
size_t fg_face_search_workspace_bytes(int m);
int fg_face_search_top1(const float* queries, const uint8_t* selector, int m, const float* db, int D, int d,
                        int64_t* best_row, float* similarity, void* workspace, size_t workspace_bytes, void* stream);
/* The same search for query BATCHES (m >= ~16) on the tensor cores: one TF32 tcgen05 pass keeps, per query, the rows whose
 * approximate score is within the TF32 error bound of the query's best one; those (a handful per query) are re-scored with the
 * exact fp32 expression of fg_face_search_top1, so best_row / similarity are identical to it.  db_norm_bound >= the largest
 * L2 norm of a database row (1 for the reference's normalised features; the caller computes it once per database);
 * d a multiple of 32, queries / db 16-byte aligned. */
size_t fg_face_search_tc_workspace_bytes(int m, int D);
int fg_face_search_top1_tc(const float* queries, const uint8_t* selector, int m, const float* db, int D, int d,
                           float db_norm_bound, int64_t* best_row, float* similarity, void* workspace,
                           size_t workspace_bytes, void* stream);
size_t fg_face_loss_workspace_bytes(int n, int d);
int fg_face_loss_fwd(const void* raw_feats, const float* feats_ori, const float* db, int n, int d, int D,
                     const uint8_t* face_indicators, const int64_t* const* targets, const int64_t* const* preds_ori,
                     const void* const* probs_ori, const int32_t* widths, int n_attr, float confidence_level,
                     int search_needs_target, float fill, void* loss, void* workspace, size_t workspace_bytes,
                     int dtype, void* stream);
int fg_face_loss_bwd(const void* g_loss, const float* feats_ori, const float* db, int n, int d, void* g_raw_feats,
                     void* workspace, size_t workspace_bytes, int dtype, void* stream);
/* Test hook: which target every row of the last fg_face_loss_fwd used (>= 0 database row, -2 feats_ori, -1 none). */
int fg_face_loss_target_rows(const void* workspace, int n, int d, int64_t* target_row, void* stream);

/* f4: bucketed gradient synchronisation, the manual all-reduce of E1:1996-2011 (same loop in E3:2217-2229).
 * grad_ptrs [n_tensors] (DEVICE array of device addresses of the .grad tensors, all of `dtype`), offsets [n_tensors+1]
 * int64 (DEVICE; prefix sums of the element counts, offsets[n_tensors] == total).
 * fg_grad_bucket_pack: bucket [total+1] float32 receives the gradients back to back and, in the last slot, the number of
 *   non-finite values found (also written to nonfinite_count, int32 on the device) -- the reference's
 *   `torch.isfinite(p.grad).all()` for every tensor, without a host round trip per tensor.  The caller all-reduces the
 *   bucket (SUM), which also sums the counts over the ranks.
 * fg_grad_bucket_unpack: p.grad = bucket / divisor_a / divisor_b (num_processes, N_backward), rounded like the two eager
 *   divisions. */
int fg_grad_bucket_pack(const uint64_t* grad_ptrs, const int64_t* offsets, int n_tensors, int64_t total, float* bucket,
                        int32_t* nonfinite_count, int dtype, void* stream);
int fg_grad_bucket_unpack(const uint64_t* grad_ptrs, const int64_t* offsets, int n_tensors, int64_t total,
                          const float* bucket, double divisor_a, double divisor_b, int dtype, void* stream);

/* ------------------------------------------------------------------ multi-GPU exchanges -
 * The two exchanges of the path as one-shot NVLink transfers between peer-mapped ("symmetric") buffers, replacing
 *   customized_all_gather (E1:222-235; call sites E3:1978-1986)   ->  fg_peer_push + fg_peer_wait_copy
 *   dist.all_reduce(plan counts, SUM) (E3:1535, E4:1569)            ->  fg_peer_push(self_only) + fg_peer_wait_sum
 * Every rank owns one allocation of the same layout; peer_base_dev is a DEVICE array of the `world` base addresses (rank
 * order).  Offsets are bytes from a base: flags at flags_off ([2][64] uint32, flag_index 0 = rows, 1 = counts), a region at
 * region_off + parity * parity_stride where parity = *epoch_dev & 1.  epoch_dev (uint32, device) is bumped once per step by
 * fg_peer_epoch_advance; flags carry the epoch.  All sizes / offsets are multiples of 16 bytes.  Plain stream launches: no
 * host synchronisation, capturable in a CUDA graph.  A wait that sees no flag within ~1 s ORs 0x100 into status[0]
 * (int32, device; may be NULL) and returns control to the stream instead of hanging.
 *   fg_peer_push       copies src[bytes] to slot_off inside the region of every peer (self_only = 0) or of this rank only
 *                      (self_only = 1), then raises this rank's flag `flag_index` on every peer.  done_counter: one zeroed
 *                      uint32 of device scratch, shared by consecutive pushes on a stream.
 *   fg_peer_wait_copy  waits for the `world` flags, then copies bytes from this rank's region to dst.
 *   fg_peer_wait_sum   waits for the `world` flags, then out[i] = sum over ranks (rank order) of the int32 slabs at the
 *                      region of every peer (n_elems a multiple of 4). */
int fg_peer_epoch_advance(uint32_t* epoch_dev, void* stream);
int fg_peer_push(const void* src, size_t bytes, const void* peer_base_dev, size_t region_off, size_t parity_stride,
                 size_t slot_off, int self_only, size_t flags_off, int flag_index, int rank, int world,
                 const uint32_t* epoch_dev, uint32_t* done_counter, void* stream);
int fg_peer_wait_copy(const void* peer_base_dev, int rank, int world, size_t flags_off, int flag_index, size_t region_off,
                      size_t parity_stride, const uint32_t* epoch_dev, void* dst, size_t bytes, int32_t* status, void* stream);
int fg_peer_wait_sum(const void* peer_base_dev, int rank, int world, size_t flags_off, int flag_index, size_t region_off,
                     size_t parity_stride, const uint32_t* epoch_dev, int32_t* out, size_t n_elems, int32_t* status, void* stream);
/* Exchange 1 fused with its packing / unpacking (what dist.PeerExchange uses): the row block {indicator, probs_0 .. probs_2}
 * [n, 1 + sum(widths)] of `dtype` is assembled in shared memory from the head's outputs (indicators uint8 [n], probs[a]
 * [n, widths[a]]; HOST arrays of device pointers / widths) and stored into slot_off of every peer's row region, flag 0 raised;
 * fg_peer_wait_unpack waits for the `world` flags and takes the gathered [n_all, 1 + sum(widths)] block of this rank's region
 * apart into indicators_all uint8 [n_all] and probs_all[a] [n_all, widths[a]].  n a multiple of 8. */
int fg_peer_push_rows(const uint8_t* indicators, const void* const* probs, const int32_t* widths, int n_attr, int n,
                      const void* peer_base_dev, size_t region_off, size_t parity_stride, size_t slot_off, size_t flags_off,
                      int rank, int world, const uint32_t* epoch_dev, uint32_t* done_counter, int dtype, void* stream);
int fg_peer_wait_unpack(const void* peer_base_dev, int rank, int world, size_t flags_off, size_t region_off, size_t parity_stride,
                        const uint32_t* epoch_dev, uint8_t* indicators_all, void* const* probs_all, const int32_t* widths,
                        int n_attr, int n_all, int32_t* status, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FAIRGUIDE_H */
