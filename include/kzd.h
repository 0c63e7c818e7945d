/*
 * kzd.h - C ABI of libkzd: the sm_100a CUDA implementation of the dense-detection hot path of
 * Kotomiya07/kuzushiji-vision (its vendored Ultralytics 8.3.63 / YOLOv12).
 *
 * The reference has no FFI for this path - the boundary is three Python call signatures
 * (SURVEY.md section 8b). Each entry point below is what a ctypes binding for one of those
 * signatures binds; the citation names the reference code it replaces. Paths are relative to
 * /root/reference/yolov12/ultralytics/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (host);
 *   - the caller owns all memory, including the scratch `workspace` (size it with the matching
 *     *_workspace_bytes call); the library allocates nothing and keeps no state between calls;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work and never synchronise;
 *   - every call returns 0 on success, a negative KZD_E* code for a bad argument, or a positive
 *     cudaError_t; kzd_last_error() returns a thread-local message for the last failure;
 *   - tensors are dense row-major fp32 unless stated otherwise.
 */
#ifndef KZD_H
#define KZD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KZD_VERSION 100 /* 0.1.0 */
#define KZD_MAX_LEVELS 8
#define KZD_REG_MAX 16 /* DFL bins per box side: nn/modules/head.py:39 */
#define KZD_MAX_TOPK 32

#define KZD_EINVAL (-1)   /* bad argument */
#define KZD_ENOSPACE (-2) /* workspace too small */

/* One pyramid level of the anchor grid (make_anchors, utils/tal.py:333-345): h*w cells, row-major
 * (y outer), cell centre (x+0.5, y+0.5)*stride, first anchor of the level at `offset`. */
typedef struct kzd_level {
    int32_t h, w;
    int32_t offset;
    float stride;
} kzd_level;

int kzd_version(void);
const char* kzd_last_error(void);
/* Number of SMs / compute capability of the current device (diagnostics only). */
int kzd_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * decode
 * ------------------------------------------------------------------------------------------ */

/* Detect._inference (nn/modules/head.py:100-131) + DFL.forward (nn/modules/block.py:71-74) +
 * dist2bbox(xywh=True, dim=1) (utils/tal.py:348-357) + cls.sigmoid(), fused.
 *   h_level_ptrs[l] : device pointer to level l's map (b, 64+nc, h_l, w_l), contiguous NCHW
 *   in_dtype        : element type of the maps: 0 fp32, 1 fp16, 2 bf16 (AMP / half inference,
 *                     engine/validator.py:115-117); the arithmetic and y are always fp32
 *   y               : (b, 4+nc, A)  [cx, cy, w, h] in pixels, then sigmoid(class logits)
 * The per-level `cat` copy (head.py:104) is skipped; anchors and strides come from `levels`. */
int kzd_decode_infer(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels,
                     int n_levels, int b, int nc, float* y, void* stream);

/* Both decodes of the same head maps in one pass: y as kzd_decode_infer, plus the training-side boxes
 * out_xyxy (b,A,4) in grid units (v8DetectionLoss.bbox_decode, utils/loss.py:197-204) and, when
 * out_xyxy_px != NULL, their pixel-space copy (utils/loss.py:236). For validation during training
 * (engine/validator.py:180-185: eval-mode Detect, then model.loss on the same maps): the maps are read
 * once instead of twice and the (b,A,64) transposed copy is not needed for the decode. */
int kzd_decode_dual(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels,
                    int n_levels, int b, int nc, float* y, float* out_xyxy, float* out_xyxy_px,
                    void* stream);

/* kzd_decode_infer / kzd_decode_dual (out_xyxy may be NULL) + the first stage of
 * ops.non_max_suppression (utils/ops.py:234-286, best-class path) in the same kernel: for the
 * Detect -> non_max_suppression flow (models/yolo/detect/predict.py:25, detect/val.py:94) with a known
 * conf_thres. y's box rows are written as xyxy - the state ops.py:243-244 (in_place=True) leaves the
 * caller's tensor in - and every anchor whose best class score exceeds conf_thres is appended to the
 * candidate pool exactly as kzd_nms_filter would after re-reading y (same keys, same rows; slot order is
 * unspecified in both). cand_counts is zeroed on the stream first. Follow with kzd_nms_suppress. */
int kzd_decode_filter(const void* const* h_level_ptrs, int in_dtype, const kzd_level* h_levels,
                      int n_levels, int b, int nc, float* y, float* out_xyxy, float* out_xyxy_px,
                      float conf_thres, uint64_t* cand_keys, float* cand_rows, int32_t* cand_counts,
                      int cand_cap, void* stream);

/* v8DetectionLoss.bbox_decode (utils/loss.py:197-204) + dist2bbox(xywh=False).
 *   pred_dist     : (b, A, 64)  side-major, 16 bins per side; in_dtype 0 fp32, 1 fp16, 2 bf16
 *                   (under AMP the head emits fp16, engine/trainer.py:379); outputs are fp32
 *   anchor_points : (A, 2)      grid units
 *   out_xyxy      : (b, A, 4)   grid units
 * Optional fused caller glue (utils/loss.py:236, `pred_bboxes * stride_tensor`, the assigner's
 * pd_bboxes): when out_xyxy_px != NULL, stride_tensor (A) must be given and out_xyxy_px (b,A,4)
 * receives out_xyxy * stride (one fp32 multiply, as the reference rounds it). */
int kzd_decode_train(const void* pred_dist, int in_dtype, const float* anchor_points,
                     const float* stride_tensor, int b, int A, float* out_xyxy, float* out_xyxy_px,
                     void* stream);

/* Backward of kzd_decode_train wrt pred_dist (the CIoU/DFL losses back-propagate through the
 * decoded boxes, utils/loss.py:252-254): grad_pred_dist (b,A,64) from grad_xyxy (b,A,4) fp32. pred_dist and
 * grad_pred_dist are elements of in_dtype (0 fp32, 1 fp16, 2 bf16: the AMP flow's tensor is read as it is and
 * the gradient is rounded once into the same dtype). */
int kzd_decode_train_bwd(const void* pred_dist, int in_dtype, const float* grad_xyxy, int b, int A,
                         void* grad_pred_dist, void* stream);

/* ---------------------------------------------------------------------------------------------
 * assign
 * ------------------------------------------------------------------------------------------ */

size_t kzd_assign_workspace_bytes(int b, int A, int G);
/* Byte offset inside the workspace of an int32 the last kzd_tal_assign call on it left behind: non-zero
 * when that call's device check found anc_points different from h_levels (it then used the slow
 * all-anchors scan). A host shim may copy it out asynchronously and rebuild its level table. */
size_t kzd_assign_report_offset(int b, int A);

/* TaskAlignedAssigner.forward (utils/tal.py:39-118, helpers :120-295; CIoU utils/metrics.py:96-130).
 *   pd_scores (b,A,nc) sigmoided, pd_bboxes (b,A,4) xyxy px, anc_points (A,2) px,
 *   gt_labels (b,G) class ids as float, gt_bboxes (b,G,4) xyxy px, mask_gt (b,G) 0/1 as float.
 *   h_levels/n_levels: the grid the anchors come from. It is only an accelerator: anc_points is
 *     checked against it on the device and, on any mismatch (or n_levels == 0), candidate
 *     enumeration falls back to scanning all A anchors per GT - same results, slower.
 *   zero_fill: how torch.topk's padding with zero-metric entries is reproduced when a GT has fewer
 *     than topk positive candidates: 0 = padding never lands inside the GT (true whenever anchors
 *     0..topk-1 are outside every GT); 1 = emulate "lowest anchor indices first".
 *   score_dtype: dtype pd_scores had at the reference call site (0 fp32, 1 fp16, 2 bf16). The kernel
 *     reads fp32 values (or the half-precision tensor itself: KZD_ASSIGN_NATIVE_SCORES); under AMP the reference rounds pd_scores.pow(alpha) to that dtype
 *     (utils/tal.py:137,150) and the kernel reproduces that one rounding.
 * Outputs (all written for every anchor):
 *   target_labels int64 (b,A), target_bboxes (b,A,4), target_scores (b,A,nc), fg_mask uint8 (b,A),
 *   target_gt_idx int64 (b,A); counters double[2] = { sum(target_scores), count(fg_mask) }
 *   (may be NULL). G must be >= 1 (the G == 0 early-out of tal.py:64-71 is host logic).
 *   flags: bit mask.
 *     KZD_ASSIGN_PERSISTENT      the caller keeps ONE workspace per (b, A) - sized for the largest G it
 *       will pass - zeroed ONCE after allocation, and passes it unchanged on every call (one call at a time
 *       per workspace). Every call leaves the per-anchor state all-zero again (the patch kernel swaps the words
 *       it reads for zero, the last CTA of each image clears the counters), so no per-call memset is needed (CUDA-graph friendly). Without this
 *       bit the call zeroes the state itself first (one extra 8 B/anchor memset).
 *     KZD_ASSIGN_TRUSTED_ANCHORS the caller vouches that anc_points equals the grid described by h_levels
 *       bit for bit (e.g. it built both, or checked them once on the host); the per-call device check is
 *       skipped.
 *     KZD_ASSIGN_NATIVE_SCORES   pd_scores points to elements of score_dtype (fp16 / bf16, the tensor the AMP
 *       flow passes: utils/loss.py:233 under autocast) instead of fp32: the kernels read them as they are, so
 *       the caller needs no up-cast pass over (b,A,nc). No effect with score_dtype 0.
 *     KZD_ASSIGN_WARP_PER_GT     scheduling hint, same results: a whole warp per GT instead of two GTs per warp -
 *       faster when the call has the GPU to itself (an eager training step), slower when other kernels share the
 *       SMs with it (several graph pipelines in flight). */
#define KZD_ASSIGN_PERSISTENT 1
#define KZD_ASSIGN_TRUSTED_ANCHORS 2
#define KZD_ASSIGN_NATIVE_SCORES 4
#define KZD_ASSIGN_WARP_PER_GT 8
int kzd_tal_assign(const void* pd_scores, const float* pd_bboxes, const float* anc_points,
                   const float* gt_labels, const float* gt_bboxes, const float* mask_gt,
                   int b, int A, int G, int nc, int topk, float alpha, float beta, float eps,
                   const kzd_level* h_levels, int n_levels, int zero_fill, int score_dtype,
                   int64_t* target_labels, float* target_bboxes, float* target_scores,
                   uint8_t* fg_mask, int64_t* target_gt_idx, double* counters,
                   void* workspace, size_t workspace_bytes, int flags, void* stream);

/* ---------------------------------------------------------------------------------------------
 * NMS
 * ------------------------------------------------------------------------------------------ */

/* Stage 1 of ops.non_max_suppression (utils/ops.py:234-286): xywh->xyxy (written back into
 * `prediction` when in_place, :243-244), confidence filter, best class (:274-275) or multi-label
 * (:271-272), optional class filter (:278-279). Candidates are appended (unordered) to a
 * caller-owned pool of `cand_cap` records per image.
 *   prediction   : (b, 4+nc, A) in/out
 *   in_place     : 0 = boxes are xywh, leave them; 1 = boxes are xywh, write xyxy back (the
 *                  reference's default side effect); 2 = boxes are already xyxy (re-run after a
 *                  pool overflow), leave them
 *   classes      : device int32[n_classes] or NULL
 *   cand_keys    : uint64 (b, cand_cap)   score bits << 32 | ~order   (order = anchor, or
 *                                          anchor*nc+class when multi_label)
 *   cand_rows    : float  (b, cand_cap, 6) [x1,y1,x2,y2,conf,cls]
 *   cand_counts  : int32 (b)  number of candidates found (may exceed cand_cap: the excess was
 *                  dropped and the caller must retry with a larger pool) */
int kzd_nms_filter(float* prediction, int b, int A, int nc, float conf_thres, int multi_label,
                   const int32_t* classes, int n_classes, int in_place,
                   uint64_t* cand_keys, float* cand_rows, int32_t* cand_counts, int cand_cap,
                   void* stream);

/* Apriori labels of the autolabelling path (utils/ops.py:256-261, `labels=` of non_max_suppression; the validator
 * passes them with save_hybrid, models/yolo/detect/val.py:55-64,92-102): labels (b, L, 5) rows [cls, cx, cy, w, h]
 * in pixels, zero padded, label_counts int32 (b). Every label joins its image's candidate pool as a detection
 * of its class with confidence 1.0, ordered after the image's predictions (as the reference's torch.cat leaves
 * them). Call between kzd_nms_filter (same A, nc, conf_thres, multi_label, classes) and kzd_nms_suppress. */
int kzd_nms_append_labels(const float* labels, const int32_t* label_counts, int b, int L, int A, int nc,
                          float conf_thres, int multi_label, const int32_t* classes, int n_classes,
                          uint64_t* cand_keys, float* cand_rows, int32_t* cand_counts, int cand_cap, void* stream);

size_t kzd_nms_workspace_bytes(int b, int cand_cap, int n_cap);

/* Stage 2 (utils/ops.py:285-311 + torchvision.ops.nms): per image, order the candidates by
 * (score desc, order asc) - torchvision's stable descending sort -, keep the first max_nms,
 * offset boxes by cls*max_wh unless agnostic, build the IoU > iou_thres bitmask, run the greedy
 * scan and write the first max_det kept rows in score order.
 *   n_cap     : capacity of the suppression stage per image (>= min(count, max_nms) for every
 *               image, else that image's out_counts entry is set to -1 and nothing is written)
 *   out_rows  : float (b, max_det, 6);  out_counts : int32 (b) */
int kzd_nms_suppress(const uint64_t* cand_keys, const float* cand_rows, const int32_t* cand_counts,
                     int b, int cand_cap, int n_cap, float iou_thres, int agnostic, int max_det,
                     int max_nms, float max_wh, float* out_rows, int32_t* out_counts,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * next rows of the path
 * ------------------------------------------------------------------------------------------ */

/* v8DetectionLoss.preprocess (utils/loss.py:180-195), step 1: per-image target counts.
 *   targets (n, 6) = [image index, class, cx, cy, w, h]; counts (b) int32; first (b) int32 = first row
 *   of each image; flags[0] bit 0: image indices not non-decreasing (caller must group them with a
 *   stable sort and call again), bit 1: an index outside [0, b). */
int kzd_gt_count(const float* targets, int n, int b, int32_t* counts, int32_t* first, int32_t* flags,
                 void* stream);

/* Step 2: out (b, G, 5) = [class, x1, y1, x2, y2] px, zero padded; scale4 = device float[4]
 * (w, h, w, h in pixels: utils/loss.py:224). Rows keep their order of appearance. */
int kzd_gt_pack(const float* targets, int n, int b, int G, const float* scale4, const int32_t* counts,
                const int32_t* first, float* out, void* stream);

/* Post-NMS tail: ops.scale_boxes (utils/ops.py:92-127, padding=True, xywh=False) + clip_boxes
 * (:319-338) for all detections of a batch, in place (models/yolo/detect/predict.py:37-41 does it per
 * image). rows (b, max_det, row_stride) with the xyxy box in columns 0..3; counts (b) int32 or NULL (all
 * rows); params (b,5) device floats = [pad_x, pad_y, 1/gain, w0, h0] of each image, 1/gain evaluated in double
 * precision and rounded to fp32 (that is how aten's `boxes /= gain` with a host scalar multiplies). */
int kzd_scale_boxes(float* rows, const int32_t* counts, int b, int max_det, int row_stride,
                    const float* params, void* stream);

/* Validator matching for all images of a batch: DetectionValidator._process_batch (models/yolo/detect/val.py:
 * 209-228: box_iou of labels x detections, utils/metrics.py:52-71) + BaseValidator.match_predictions
 * (engine/validator.py:222-260, use_scipy=False). A detection is correct at threshold iouv[i] when its best
 * equal-class label (highest IoU) has IoU >= iouv[i] and no earlier detection chose the same label at that
 * threshold (detections come in descending confidence order, as kzd_nms_suppress writes them).
 *   det_rows (b, max_det, row_stride >= 6) [x1,y1,x2,y2,conf,cls], det_counts int32 (b) or NULL (all rows);
 *   gt (b, G, 5) [cls, x1,y1,x2,y2] in the same coordinate space, gt_counts int32 (b) labels per image;
 *   iouv device float[n_iou], n_iou <= 32, every threshold > 0 (the reference uses linspace(0.5, 0.95, 10));
 *   correct uint8 (b, max_det, n_iou): 0/1, rows beyond det_counts are 0.
 * Exactly equal IoUs of one detection with several labels are ordered by the reference's unstable argsort;
 * here the label with the larger index wins. */
size_t kzd_match_workspace_bytes(int b, int max_det, int G, int n_iou);
int kzd_match_predictions(const float* det_rows, const int32_t* det_counts, int b, int max_det, int row_stride,
                          const float* gt, const int32_t* gt_counts, int G, const float* iouv, int n_iou,
                          uint8_t* correct, void* workspace, size_t workspace_bytes, void* stream);

/* Loss tail on the assigned targets (utils/loss.py:243-258; BboxLoss :91-113, DFLoss :65-88):
 *   loss3[0] = sum_fg (1 - CIoU(pred_box, target_box/stride)) * w / tss   (box, before hyp.box)
 *   loss3[1] = sum BCEWithLogits(pred_scores, target_scores) / tss         (cls)
 *   loss3[2] = sum_fg DFL(pred_dist, target ltrb) * w / tss                (dfl)
 * with w = target_scores.sum(-1), tss = max(*target_scores_sum, 1). The predicted box is re-decoded
 * from pred_dist (b,A,64) in the kernel. pred_scores are logits (b,A,nc); target_bboxes are in pixels
 * (the assigner's output, divided by stride_tensor (A) here); fg_mask uint8 (b,A).
 * sums_ws: double[3] scratch. The backward writes dense grad_pred_scores (b,A,nc) and grad_pred_dist
 * (b,A,64) for grad_loss3 = d(final)/d(loss3) (device float[3]). */
int kzd_loss_tail_fwd(const float* pred_scores, const float* pred_dist, const float* anchor_points,
                      const float* stride_tensor, const float* target_bboxes, const float* target_scores,
                      const uint8_t* fg_mask, const float* target_scores_sum, int b, int A, int nc,
                      double* sums_ws, float* loss3, void* stream);
int kzd_loss_tail_bwd(const float* pred_scores, const float* pred_dist, const float* anchor_points,
                      const float* stride_tensor, const float* target_bboxes, const float* target_scores,
                      const uint8_t* fg_mask, const float* target_scores_sum, int b, int A, int nc,
                      const float* grad_loss3, float* grad_pred_scores, float* grad_pred_dist, void* stream);

/* Sparse-target form of the same loss tail (utils/loss.py:243-247): the BCE never streams the
 * dense (b,A,nc) target_scores. BCE(x, t) = softplus(x) - t*x: the first term is one pass over pred_scores,
 * the second exists only at the foreground (anchor, target_labels[anchor]) elements, because a foreground row
 * of the assigner's target_scores is one-hot * norm at the anchor's target label and background rows are zero
 * (utils/tal.py:228-236). target_scores is read at those elements only (also the box / DFL weight, which the
 * reference takes as target_scores.sum(-1), utils/loss.py:93); target_labels int64 (b,A) is kzd_tal_assign's
 * first output. score_dtype (0 fp32, 1 fp16, 2 bf16): dtype the reference rounds the BCE target to under AMP
 * (target_scores.to(dtype), utils/loss.py:247) - the weight keeps the fp32 value. The backward writes
 * sigmoid(x)*scale everywhere (reads pred_scores only) and subtracts t*scale at the foreground elements.
 * logits_dtype / dist_dtype (0 fp32, 1 fp16, 2 bf16): element type of pred_scores + grad_pred_scores and of
 * pred_dist + grad_pred_dist. Under AMP the head outputs are half precision (engine/trainer.py:379): the kernels
 * read them as they are, compute in fp32 and round the gradients once into the same dtype (what autograd does
 * behind the reference's fp32-autocast BCE / softmax) - no up-cast copies of the (b,A,nc) and (b,A,64) tensors. */
int kzd_loss_tail_sparse_fwd(const void* pred_scores, const void* pred_dist, const float* anchor_points,
                             const float* stride_tensor, const float* target_bboxes, const float* target_scores,
                             const int64_t* target_labels, const uint8_t* fg_mask, const float* target_scores_sum,
                             int b, int A, int nc, int score_dtype, int logits_dtype, int dist_dtype,
                             double* sums_ws, float* loss3, void* stream);
int kzd_loss_tail_sparse_bwd(const void* pred_scores, const void* pred_dist, const float* anchor_points,
                             const float* stride_tensor, const float* target_bboxes, const float* target_scores,
                             const int64_t* target_labels, const uint8_t* fg_mask, const float* target_scores_sum,
                             int b, int A, int nc, int score_dtype, int logits_dtype, int dist_dtype,
                             const float* grad_loss3, void* grad_pred_scores, void* grad_pred_dist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KZD_H */
