/*
 * tlod_b200.h -- C ABI of libtlod_b200.so: the Faster R-CNN RoI / proposal hot
 * path of live-group/Transfer-Learning-Library-for-Object-Detection as
 * hand-written CUDA for sm_100a (B200).
 *
 * This is the drop-in boundary.  Every entry point replaces one symbol the
 * reference binds through its cffi bridge (`torch.utils.ffi`), cited per
 * function as /root/reference path:line.  Conventions (same as the reference's
 * tensor-free launcher layer, lib/model/roi_align/src/roi_align_kernel.h:13-27):
 *
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it
 *     is named `h_*`; all tensors are contiguous, fp32 NCHW, int32 indices;
 *   - the CALLER allocates everything, including outputs and workspaces (query
 *     the `*_workspace_bytes` functions); the library never frees or retains;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default
 *     stream).  All work is enqueued on it; no entry point synchronises the
 *     host, reads device memory from the host, or allocates device memory;
 *   - return value: 0 = success; < 0 = TLOD_ERR_* argument error (nothing was
 *     launched); > 0 = the `cudaError_t` of the failed launch.  The library
 *     never calls exit() (the reference does: roi_align_kernel.cu:84-88);
 *   - re-entrant and device-correct: kernels run on the device that is current
 *     for the calling thread, like the reference under nn.DataParallel.
 *
 * There is no CPU implementation behind any of these symbols.
 */
#ifndef TLOD_B200_H
#define TLOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLOD_B200_VERSION 100 /* 0.1.0 */

enum {
  TLOD_OK = 0,
  TLOD_ERR_NULL_POINTER = -1,
  TLOD_ERR_BAD_SHAPE = -2,      /* non-positive or inconsistent sizes            */
  TLOD_ERR_UNSUPPORTED = -3,    /* valid request outside the implemented range   */
  TLOD_ERR_WORKSPACE = -4,      /* workspace missing or too small                */
  TLOD_ERR_INT32_OVERFLOW = -5  /* tensor has >= 2^31 elements (reference: UB)   */
};

int tlod_version(void);
/* Static string for a TLOD_ERR_* code or a cudaError_t. */
const char* tlod_error_string(int code);
/* Number of kernel launches enqueued by this library in this process (bench.py's `gpu_launches`). */
unsigned long long tlod_launch_count(void);

/* Optional per-kernel timing used by bench.py: while enabled, every kernel this
 * library launches is bracketed by two CUDA events on its stream.
 * tlod_profile_collect() waits for them (the one call here that synchronises the
 * host) and returns the number of distinct kernel names; tlod_profile_get() reads
 * entry `index`: name (static storage until the next reset), summed device time in
 * milliseconds and number of launches. */
void tlod_profile_enable(int on);
void tlod_profile_reset(void);
int tlod_profile_collect(void);
int tlod_profile_get(int index, const char** name, double* total_ms, long long* launches);

/* ------------------------------------------------------------------------ */
/* RoIAlign                                                                   */
/* replaces roi_align_forward_cuda / roi_align_backward_cuda                  */
/*   lib/model/roi_align/src/roi_align_cuda.c:7-40, :42-76                    */
/*   (kernels lib/model/roi_align/src/roi_align_kernel.cu:15-70, :94-143)     */
/* ------------------------------------------------------------------------ */
/* RoIAlign plan: everything that depends only on (rois, map geometry, aligned size):
 * per-RoI sampling tables, the RoI indices stably sorted by image, per-image offsets, the
 * backward column chains and, per (image, map row), the list of gradient rows that feed it.  Build it once per `rois` tensor with tlod_roi_align_plan
 * and pass it to the forward and the backward call (same batch, height, width, num_rois,
 * aligned_h, aligned_w, spatial_scale).  `plan` must be 256-byte aligned and at least
 * tlod_roi_align_plan_bytes(batch, num_rois) bytes.  Limits of the planned (shared-memory
 * resident, atomic-free) kernels: batch <= 1024, aligned_h/w <= 16 (backward also:
 * aligned_w == 8, channels % 32 == 0, batch * height <= 8192); outside them, or with
 * plan == NULL, the forward/backward entry points use the generic kernels (one CTA per
 * RoI and channel block; fp32 atomics in the backward). */
size_t tlod_roi_align_plan_bytes(int batch, int num_rois);
int tlod_roi_align_plan(const float* rois, int batch, int height, int width, int num_rois,
                        int aligned_h, int aligned_w, float spatial_scale, void* plan,
                        size_t plan_bytes, void* stream);

/* features (batch, channels, height, width); rois (num_rois, 5) =
 * [batch_idx, x1, y1, x2, y2] in image pixels; output (num_rois, channels,
 * aligned_h, aligned_w).  One bilinear sample per output cell at
 * start + p * extent/(aligned-1); see SURVEY.md appendix B items 1-7.
 * Every output element is written (RoIs whose batch index is outside
 * [0, batch) produce zeros; the reference reads out of bounds there).
 * Requires height >= 2, width >= 2, aligned_h >= 2, aligned_w >= 2. */
int tlod_roi_align_forward(const float* features, const float* rois, float* output, int batch,
                           int channels, int height, int width, int num_rois, int aligned_h,
                           int aligned_w, float spatial_scale, const void* plan, size_t plan_bytes,
                           void* stream);

/* top_grad (num_rois, channels, aligned_h, aligned_w) -> bottom_grad (batch,
 * channels, height, width).  bottom_grad is fully OVERWRITTEN with the gradient
 * (the reference accumulates into a buffer its caller has just zeroed,
 * functions/roi_align.py:42, so the observable result is the same). */
int tlod_roi_align_backward(const float* top_grad, const float* rois, float* bottom_grad,
                            int batch, int channels, int height, int width, int num_rois,
                            int aligned_h, int aligned_w, float spatial_scale, const void* plan,
                            size_t plan_bytes, void* stream);

/* 2x2 / stride-1 average pooling of every (height, width) tile -> (height-1, width-1):
 * the `avg_pool2d(x, kernel_size=2, stride=1)` of RoIAlignAvg
 * (lib/model/roi_align/modules/roi_align.py:26-29).  in (tiles, height, width),
 * out (tiles, height-1, width-1); tiles = num_rois * channels.  The backward entry is the
 * adjoint: grad_out (tiles, height-1, width-1) -> grad_in (tiles, height, width), fully
 * overwritten. */
int tlod_avgpool2x2_forward(const float* in, float* out, long long tiles, int height, int width,
                            void* stream);
int tlod_avgpool2x2_backward(const float* grad_out, float* grad_in, long long tiles, int height,
                             int width, void* stream);

/* RoIAlignAvg in one call (lib/model/roi_align/modules/roi_align.py:20-29): RoIAlign at
 * (pooled_h + 1, pooled_w + 1) samples followed by avg_pool2d(kernel_size=2, stride=1).
 * output / top_grad are (num_rois, channels, pooled_h, pooled_w); `plan` is the plan built for
 * aligned size (pooled_h + 1, pooled_w + 1).
 * Forward: for pooled 7 x 7 (the only size the reference uses), channels % 16 == 0 and a plan,
 * ONE kernel produces the pooled tensor; the (R, C, 8, 8) sample tensor is never written and
 * `scratch` may be NULL.  Other shapes run RoIAlign into `scratch` and then the 2x2 average
 * (TLOD_ERR_WORKSPACE if scratch is missing or smaller than tlod_roi_align_avg_scratch_bytes).
 * Backward: the adjoint of the average is written into `scratch` (always required), then
 * tlod_roi_align_backward runs on it; bottom_grad is fully overwritten. */
size_t tlod_roi_align_avg_scratch_bytes(int channels, int num_rois, int pooled_h, int pooled_w);
int tlod_roi_align_avg_forward(const float* features, const float* rois, float* output, int batch,
                               int channels, int height, int width, int num_rois, int pooled_h,
                               int pooled_w, float spatial_scale, const void* plan, size_t plan_bytes,
                               void* scratch, size_t scratch_bytes, void* stream);
int tlod_roi_align_avg_backward(const float* top_grad, const float* rois, float* bottom_grad,
                                int batch, int channels, int height, int width, int num_rois,
                                int pooled_h, int pooled_w, float spatial_scale, const void* plan,
                                size_t plan_bytes, void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* RoIPool                                                                    */
/* replaces roi_pooling_forward_cuda / roi_pooling_backward_cuda              */
/*   lib/model/roi_pooling/src/roi_pooling_cuda.c                             */
/*   (kernels lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93, :128-203)*/
/* ------------------------------------------------------------------------ */
/* argmax (num_rois, channels, pooled_h, pooled_w) int32: flat index into the
 * whole features buffer, -1 for an empty bin; may be NULL. */
int tlod_roi_pool_forward(const float* features, const float* rois, float* output, int* argmax,
                          int batch, int channels, int height, int width, int num_rois,
                          int pooled_h, int pooled_w, float spatial_scale, void* stream);

/* bottom_grad fully overwritten.  Reproduces the reference gather exactly up
 * to fp32 summation order: a top_grad element is counted only if its argmax
 * cell lies inside the rounded RoI and inside the bin's feasible set
 * (roi_pooling_kernel.cu:157-196). */
int tlod_roi_pool_backward(const float* top_grad, const int* argmax, const float* rois,
                           float* bottom_grad, int batch, int channels, int height, int width,
                           int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                           void* stream);

/* ------------------------------------------------------------------------ */
/* RoICrop (bilinear sampler on a per-RoI grid; BASELINE cfg3 "crop 14")      */
/* replaces BilinearSamplerBHWD_updateOutput_cuda / _updateGradInput_cuda     */
/*   lib/model/roi_crop/src/roi_crop_cuda.c, roi_crop_cuda.h:5-8              */
/*   (kernels lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:45-108, :111-190)*/
/* ------------------------------------------------------------------------ */
/* features (in_batch, channels, height, width); grid_yx (out_batch, grid_h, grid_w, 2) =
 * (y, x) in [-1, 1] (align_corners semantics: -1 / +1 are the centres of the first / last
 * cell); output (out_batch, channels, grid_h, grid_w).  Output batch b samples image
 * b / (out_batch / in_batch); corners outside the map contribute zero.
 * The backward entry overwrites grad_features with the gradient w.r.t. the features; like
 * the reference kernel it produces no gradient for the grid. */
int tlod_roi_crop_forward(const float* features, const float* grid_yx, float* output, int in_batch,
                          int channels, int height, int width, int out_batch, int grid_h, int grid_w,
                          void* stream);
int tlod_roi_crop_backward(const float* grad_output, const float* grid_yx, float* grad_features,
                           int in_batch, int channels, int height, int width, int out_batch,
                           int grid_h, int grid_w, void* stream);

/* RoICrop on an axis-aligned grid fused with the max_pool2d(2, 2) that follows it in the
 * detector (cfg.POOLING_MODE == 'crop': lib/model/faster_rcnn/faster_rcnn.py:73-80,
 * _affine_grid_gen lib/model/utils/net_utils.py:142-164): the (out_batch, C, 14, 14) sample
 * tensor is never written.  The rotation-free affine grid is the outer product of per-RoI
 * coordinate vectors: grid_y (out_batch, grid_h) and grid_x (out_batch, grid_w), normalised
 * to [-1, 1] exactly like the (y, x) grid of tlod_roi_crop_forward (grid[n, i, j] =
 * (grid_y[n, i], grid_x[n, j])); RoI n samples image n / (out_batch / in_batch).
 * output (out_batch, C, grid_h / 2, grid_w / 2); argmax (same shape, uint8, may be NULL in
 * the forward): which sample of the 2 x 2 window won, 2 * dy + dx, first maximum in row-major
 * order (max_pool2d's rule) -- the backward routes the gradient through it.
 * Implemented for grid_h == grid_w == 14 (the reference's POOLING_SIZE * 2), channels % 16 == 0
 * and maps whose 16 planes fit shared memory; anything else returns TLOD_ERR_UNSUPPORTED with
 * nothing launched (compose tlod_roi_crop_forward and a pooling pass instead).
 * workspace: tlod_roi_crop_pool_workspace_bytes(out_batch), 16-byte aligned; it holds the
 * per-RoI sampling tables and may be discarded after each call.
 * Backward: grad_features (in_batch, C, H, W) is fully overwritten. */
size_t tlod_roi_crop_pool_workspace_bytes(int out_batch);
int tlod_roi_crop_pool_forward(const float* features, const float* grid_y, const float* grid_x,
                               float* output, unsigned char* argmax, int in_batch, int channels,
                               int height, int width, int out_batch, int grid_h, int grid_w,
                               void* workspace, size_t workspace_bytes, void* stream);
int tlod_roi_crop_pool_backward(const float* grad_output, const unsigned char* argmax,
                                const float* grid_y, const float* grid_x, float* grad_features,
                                int in_batch, int channels, int height, int width, int out_batch,
                                int grid_h, int grid_w, void* workspace, size_t workspace_bytes,
                                void* stream);

/* ------------------------------------------------------------------------ */
/* NMS                                                                        */
/* replaces nms_cuda                lib/model/nms/src/nms_cuda.c:8-19         */
/*   (nms_kernel + host greedy scan lib/model/nms/src/nms_cuda_kernel.cu:41-161)*/
/* ------------------------------------------------------------------------ */
/* boxes: n rows of `box_stride` floats (>= 4; the reference passes 5 =
 * x1 y1 x2 y2 score), already sorted by score descending.  IoU uses the
 * reference's fp32 formula with every operation individually rounded
 * (no FMA contraction); a box is suppressed iff IoU > thresh (strict).
 * keep_out[0 .. *num_out) = ascending kept indices, both on the DEVICE; no host
 * synchronisation happens inside (the reference does 4).  max_keep > 0 stops
 * after that many survivors (the caller's `[:post_nms_topN]`).
 * The IoU mask is only computed for the leading boxes the scan can reach before max_keep
 * survivors exist (extended x4 on demand), so a small max_keep makes the call much cheaper.
 * workspace: tlod_nms_workspace_bytes(n) bytes, 32-byte aligned. */
size_t tlod_nms_workspace_bytes(int n);
int tlod_nms(const float* boxes, int n, int box_stride, float thresh, int max_keep, int* keep_out,
             int* num_out, void* workspace, size_t workspace_bytes, void* stream);

/* Test-time per-class NMS of one image, all classes in one call:
 * replaces the loop methods/DAF/DAF_test.py:302-320 (identical in every *_test.py):
 *   for j in 1..K-1: inds = scores[:, j] > thresh; sort desc; nms(cls_dets, cfg.TEST.NMS).
 * scores (num_rois, num_classes); boxes (num_rois, box_cols) with box_cols = 4 * num_classes
 * (class j at columns 4j..4j+3) or 4 (class agnostic).  Classes first_class .. num_classes-1 are
 * processed; c below = class - first_class; Rp = tlod_class_nms_padded_rows(num_rois).
 * Outputs (device): dets_out (nc, Rp, 5) candidates sorted by score (ties: lower row first)
 * = [x1, y1, x2, y2, score]; order_out (nc, Rp) their source rows (-1 = filler);
 * count_out (nc) candidates above the threshold; keep_out (nc, Rp) kept positions, ascending;
 * valid_out (nc): the class's detections are dets_out[c][keep_out[c][0 .. valid_out[c])].
 * num_out (nc) is scratch.  num_rois <= 2048.  No host synchronisation inside. */
int tlod_class_nms_padded_rows(int num_rois);
size_t tlod_class_nms_workspace_bytes(int num_rois, int num_classes);
int tlod_class_nms(const float* scores, const float* boxes, int num_rois, int num_classes,
                   int first_class, int box_cols, float score_thresh, float nms_thresh, float* dets_out,
                   int* order_out, int* keep_out, int* num_out, int* count_out, int* valid_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* Proposal layer (fused, batched)                                            */
/* replaces _ProposalLayer.forward     lib/model/rpn/proposal_layer.py:49-163 */
/*   = bbox_transform_inv + clip_boxes lib/model/rpn/bbox_transform.py:77-133 */
/*   + torch.sort + per-image nms() + padding                                 */
/* ------------------------------------------------------------------------ */
/* scores (batch, 2A, H, W): fg probabilities are channels [A, 2A);
 * deltas (batch, 4A, H, W); im_info (batch, 3) = [h, w, scale];
 * anchors (A, 4) base anchors (generate_anchors output as fp32);
 * rois_out (batch, post_nms_topN, 5): zero padded, column 0 = image index.
 * pre_nms_topN is applied only if 0 < pre_nms_topN < batch*H*W*A (the
 * reference's whole-batch numel test, proposal_layer.py:138).
 * Optional debug outputs (may be NULL): order_out (batch, n_sorted) int32 flat
 * anchor index per rank; sorted_boxes_out (batch, n_sorted, 4); num_out (batch).
 * Ties in score keep the lower anchor index first (stable descending sort). */
int tlod_proposals_n_sorted(int batch, int num_anchors, int height, int width, int pre_nms_topN);
size_t tlod_proposals_workspace_bytes(int batch, int num_anchors, int height, int width,
                                      int pre_nms_topN, int post_nms_topN);
int tlod_proposals(const float* scores, const float* deltas, const float* im_info,
                   const float* anchors, float* rois_out, int batch, int num_anchors, int height,
                   int width, int feat_stride, int pre_nms_topN, int post_nms_topN,
                   float nms_thresh, int* order_out, float* sorted_boxes_out, int* num_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* Box arithmetic used by the RPN / test scripts                              */
/*   lib/model/rpn/bbox_transform.py:36-75, :77-103, :125-133, :168-257       */
/* ------------------------------------------------------------------------ */
/* boxes (batch, n, 4) [or (n,4) shared by all images when boxes_batched == 0],
 * deltas (batch, n, 4) -> out (batch, n, 4); then clip to im_info if non-NULL. */
int tlod_bbox_transform_inv_clip(const float* boxes, int boxes_batched, const float* deltas,
                                 const float* im_info, float* out, int batch, int n, void* stream);
/* in-place clip_boxes: boxes (batch, n, 4*k) clamped to [0, w-1] x [0, h-1] */
int tlod_clip_boxes(float* boxes, const float* im_info, int batch, int n, int k, void* stream);
/* overlaps (batch, n, k): anchors (n,4) if anchors_batched == 0 else
 * (batch, n, anchor_stride) with the 4 coordinates at column anchor_offset;
 * gt (batch, k, gt_stride).  gt with w==h==1 -> 0, anchor with w==h==1 -> -1. */
int tlod_bbox_overlaps_batch(const float* anchors, int anchors_batched, int anchor_stride,
                             int anchor_offset, const float* gt, int gt_stride, float* overlaps,
                             int batch, int n, int k, void* stream);
/* targets (batch, n, 4) = (dx, dy, log dw, log dh); ex (n,4) or (batch,n,4). */
int tlod_bbox_transform_batch(const float* ex_rois, int ex_batched, const float* gt_rois,
                              float* targets, int batch, int n, void* stream);

/* ------------------------------------------------------------------------ */
/* Anchor-target assignment                                                   */
/* replaces the device part of _AnchorTargetLayer.forward                     */
/*   lib/model/rpn/anchor_target_layer.py:98-116 (labels), :147-191 (targets) */
/* The random subsampling (:123-145) stays on the host, as in the reference.  */
/* ------------------------------------------------------------------------ */
/* anchors (n, 4) inside-image anchors; gt (batch, k, gt_stride>=4).
 * labels (batch, n) fp32 in {-1, 0, 1}; argmax (batch, n) int32 (ties -> lowest
 * gt index); max_overlaps (batch, n) or NULL.
 * workspace: tlod_anchor_labels_workspace_bytes(batch, k). */
size_t tlod_anchor_labels_workspace_bytes(int batch, int k);
int tlod_anchor_labels(const float* anchors, const float* gt, int gt_stride, float* labels,
                       int* argmax, float* max_overlaps, int batch, int n, int k,
                       float negative_overlap, float positive_overlap, int clobber_positives,
                       void* workspace, size_t workspace_bytes, void* stream);
/* HOST function (no device work): the random subsampling between the two kernels,
 * lib/model/rpn/anchor_target_layer.py:118-145, on the host copy of labels (batch, n) in
 * {-1, 0, 1}, edited in place: per image, surplus foreground labels beyond num_fg and surplus
 * background labels beyond rpn_batchsize - (foreground kept) are set to -1, chosen with
 * np.random.permutation exactly as the reference does.  mt_key (624 words) / mt_pos are numpy's
 * global MT19937 state (np.random.get_state()[1:3]); they are advanced in place, to be put back
 * with np.random.set_state(), so the stream continues as after the reference's own calls.
 * num_examples_last = number of labels >= 0 of the LAST image (:156, stale loop variable). */
int tlod_anchor_subsample_host(float* labels, int batch, int n, int num_fg, int rpn_batchsize,
                               unsigned int* mt_key, int* mt_pos, int* num_examples_last);
/* HOST: MT19937 key blocks generated ahead of time, off the critical path (while the device
 * computes the labels the sampling waits for).  ahead ((blocks + 1) * 624 words): block 0 = a copy
 * of mt_key, block j = the key numpy will hold after j more refills.  numpy's state is not touched. */
int tlod_mt_pregen(const unsigned int* mt_key, unsigned int* ahead, int blocks);
/* tlod_anchor_subsample_host drawing from `ahead` (tlod_mt_pregen) instead of refilling.  The
 * blocks are used only if block 0 still equals mt_key (nothing was drawn since tlod_mt_pregen);
 * when they run out the stream continues in mt_key.  ahead = NULL: the plain call.  Results and
 * the final (mt_key, mt_pos) are identical either way. */
int tlod_anchor_subsample_host_ahead(float* labels, int batch, int n, int num_fg, int rpn_batchsize,
                                     unsigned int* mt_key, int* mt_pos, unsigned int* ahead, int ahead_blocks,
                                     int* num_examples_last);
/* HOST function: out[0..n) = np.random.permutation(n) drawn from the given MT19937 state
 * (numpy's legacy RandomState: Fisher-Yates from the top, masked rejection sampling). */
int tlod_numpy_permutation(unsigned int* mt_key, int* mt_pos, long long n, long long* out);
/* Final maps.  labels/argmax (batch, n) over inside anchors (after the host
 * subsampling); inv_index (total) int32: position of each of the total = H*W*A
 * anchors in the inside list or -1.
 * Outputs in the reference's layouts: labels_out (batch, 1, A*H, W),
 * targets_out / inside_w_out / outside_w_out (batch, 4A, H, W). */
int tlod_anchor_targets_finalize(const float* labels, const int* argmax, const float* anchors,
                                 const float* gt, int gt_stride, const int* inv_index,
                                 float* labels_out, float* targets_out, float* inside_w_out,
                                 float* outside_w_out, int batch, int n, int k, int num_anchors,
                                 int height, int width, float inside_weight,
                                 float positive_weight, float negative_weight, void* stream);
/* The same with {inside, positive, negative} weights read from DEVICE memory (3 floats): the
 * positive / negative weight is 1 / (number of sampled anchors), known only after the host
 * subsampling, so a launch captured in a CUDA graph cannot carry it as an argument. */
int tlod_anchor_targets_finalize_dev(const float* labels, const int* argmax, const float* anchors,
                                     const float* gt, int gt_stride, const int* inv_index,
                                     float* labels_out, float* targets_out, float* inside_w_out,
                                     float* outside_w_out, int batch, int n, int k, int num_anchors,
                                     int height, int width, const float* weights_dev, void* stream);

/* ------------------------------------------------------------------------ */
/* Proposal-target assignment                              */
/* replaces the device part of _ProposalTargetLayer.forward                   */
/*   lib/model/rpn/proposal_target_layer_cascade.py:118-130 (overlaps, max,   */
/*   gt class), :183-212 (gather, label clamp, targets, weights).             */
/* The fg / bg sampling (:140-181) stays on the host with numpy's RNG.        */
/* ------------------------------------------------------------------------ */
/* rois (batch, n, roi_stride) with x1,y1,x2,y2 at column roi_offset (the layer passes the
 * proposals with the gt boxes appended: stride 5, offset 1); gt (batch, k, gt_stride >= 5)
 * = [x1,y1,x2,y2,class].  max_overlaps / labels (batch, n) fp32, assignment (batch, n) int32
 * (ties -> lowest gt index); labels = class of the assigned gt box. */
int tlod_roi_gt_assign(const float* rois, int roi_stride, int roi_offset, const float* gt,
                       int gt_stride, float* max_overlaps, int* assignment, float* labels, int batch,
                       int n, int k, void* stream);
/* HOST function (no GPU work): the fg / bg sampling of _ProposalTargetLayer
 * (proposal_target_layer_cascade.py:140-181) on numpy's global MT19937 stream, draw for draw
 * (np.random.permutation(fg_num), np.random.rand(k)): h_max_overlaps (batch, n) host floats ->
 * h_keep (batch, rois_per_image) sampled candidate indices, foreground first; h_fg_count (batch).
 * mt_key / mt_pos: numpy's legacy state (624 words + position), updated in place.
 * TLOD_ERR_BAD_SHAPE if an image has neither fg nor bg candidates (the reference raises). */
int tlod_proposal_sample_host(const float* h_max_overlaps, int batch, int n, int rois_per_image,
                              int fg_rois_per_image, float fg_thresh, float bg_thresh_hi,
                              float bg_thresh_lo, unsigned int* mt_key, int* mt_pos, int* h_keep,
                              int* h_fg_count);
/* keep (batch, rois_per_image) int32: sampled candidate indices, foreground first;
 * fg_count (batch) int32: rows >= fg_count[b] get label 0.  Outputs: rois_out
 * (batch, P, 5) with column 0 = image index, labels_out (batch, P), targets_out / inside_out /
 * outside_out (batch, P, 4); targets = bbox_transform_batch(roi, assigned gt), optionally
 * (t - h_means) / h_stds, zero where label == 0.  h_* are HOST pointers to 4 floats. */
int tlod_proposal_targets(const float* rois, int roi_stride, int roi_offset, const float* gt,
                          int gt_stride, const int* assignment, const float* labels, const int* keep,
                          const int* fg_count, float* rois_out, float* labels_out, float* targets_out,
                          float* inside_out, float* outside_out, int batch, int n, int k,
                          int rois_per_image, const float* h_means, const float* h_stds,
                          const float* h_inside_w, int normalize, void* stream);

/* ------------------------------------------------------------------------ */
/* Gradient reversal + domain-classifier loss reduction                       */
/*   lib/DAF/DA.py:19-33 (GRLayer), lib/MAF/DA.py:34-53 (weighted GRL),       */
/*   lib/DAF/faster_rcnn.py:181-220 (image / instance / consistency losses)   */
/* ------------------------------------------------------------------------ */
/* out[i] = -alpha * grad[i]  (one kernel instead of neg() and mul()) */
int tlod_grl_backward(const float* grad, float* out, float alpha, long long n, void* stream);
/* out[r, :] = -alpha * weight[r] * grad[r, :]   grad (rows, cols) */
int tlod_grl_backward_weighted(const float* grad, const float* row_weight, float* out, float alpha,
                               int rows, int cols, void* stream);
/* img_score (batch, 2, H, W) logits, ins_prob (num_ins) sigmoid outputs,
 * ins_label (num_ins) or NULL (= domain_label everywhere), domain_label 0/1.
 * losses_out[4] = { image NLL (mean), instance BCE (mean), consistency MSE
 * (sum, target = mean softmax prob of channel domain_label, detached),
 * consistency target }.  workspace (tlod_da_loss_workspace_bytes(), 8-byte aligned) may be NULL:
 * it is only used when the map has more than 2^16 cells, to reduce the image head over many
 * CTAs first; without it such a map is reduced by one CTA (same result, slower). */
size_t tlod_da_loss_workspace_bytes(void);
int tlod_da_loss_forward(const float* img_score, const float* ins_prob, const float* ins_label,
                         int domain_label, float* losses_out, int batch, int height, int width,
                         int num_ins, void* workspace, size_t workspace_bytes, void* stream);
/* Gradients of w_img*img + w_ins*ins + w_cst*cst given losses_out from the
 * forward call: grad_img_score (batch,2,H,W), grad_ins_prob (num_ins).
 * upstream (3 floats on the DEVICE, may be NULL): autograd's incoming gradients of
 * the three losses, multiplied into the weights on the device so that the
 * caller never has to read them back. */
int tlod_da_loss_backward(const float* img_score, const float* ins_prob, const float* ins_label,
                          int domain_label, const float* losses_out, const float* upstream,
                          float w_img, float w_ins, float w_cst, float* grad_img_score,
                          float* grad_ins_prob, int batch, int height, int width, int num_ins,
                          void* stream);

/* Image-level domain-classifier losses of up to TLOD_DA_MAX_LEVELS feature levels, one
 * launch each way, any map size (multi-CTA, fp64 partial sums):
 *   lib/MAF/faster_rcnn.py:188-205   conv3 / conv4 / conv5 heads,
 *                                    F.nll_loss(F.log_softmax(score, 1), label) per level
 *   lib/ATF/faster_rcnn.py:303-321   the same with ignore_index = -1
 * The h_* arguments are HOST arrays of `levels` entries; h_scores[l] is the DEVICE pointer of
 * level l's logits (h_batch[l], 2, h_height[l], h_width[l]); h_labels (may be NULL) holds per
 * level the DEVICE pointer of an int64 label map (batch, height, width) -- what the reference's
 * ImageLabelResizeLayer returns -- or NULL = every cell carries `domain_label`.  Cells whose
 * label equals ignore_index are not counted.
 * out (levels, 4) on the device: { mean NLL over the counted cells (NaN if none, like torch),
 * mean softmax probability of class `domain_label` over them, number of counted cells, 0 }.
 * The MAF instance-level CrossEntropyLoss over (R, 2) logits (faster_rcnn.py:207-213) is the
 * same reduction with height = width = 1.
 * Backward: grad_scores[l] = h_weights[l] * upstream[l] * d(out[l][0]) / d(score[l]), fully
 * written (zeros at ignored cells); upstream: `levels` floats on the DEVICE or NULL (= 1),
 * h_weights: host floats or NULL (= 1). */
#define TLOD_DA_MAX_LEVELS 4
size_t tlod_da_image_loss_workspace_bytes(void);
int tlod_da_image_loss_forward(int levels, const float* const* h_scores,
                               const long long* const* h_labels, const int* h_batch,
                               const int* h_height, const int* h_width, int domain_label,
                               int ignore_index, float* out, void* workspace, size_t workspace_bytes,
                               void* stream);
int tlod_da_image_loss_backward(int levels, const float* const* h_scores,
                                const long long* const* h_labels, const int* h_batch,
                                const int* h_height, const int* h_width, int domain_label,
                                int ignore_index, const float* out, const float* upstream,
                                const float* h_weights, float* const* h_grad_scores, void* stream);

/* ------------------------------------------------------------------------ */
/* RPN head losses                                         */
/*   lib/model/rpn/rpn.py:90-108 (keep = label != -1, index_select,           */
/*   cross_entropy) and _smooth_l1_loss, lib/model/utils/net_utils.py:72-86   */
/* ------------------------------------------------------------------------ */
/* cls_score (batch, 2A, H, W) logits (bg = channels [0, A), fg = [A, 2A)); labels
 * (batch, 1, A*H, W) fp32 in {-1, 0, 1} (the anchor-target layer's output); bbox_pred /
 * bbox_targets / inside_w / outside_w (batch, 4A, H, W).
 * losses_out[4] = { cross-entropy (mean over labels != -1), smooth-L1 (sum / batch),
 * number of kept anchors, number of foreground anchors }.  Deterministic (fixed order). */
size_t tlod_rpn_loss_workspace_bytes(void);
int tlod_rpn_loss_forward(const float* cls_score, const float* labels, const float* bbox_pred,
                          const float* bbox_targets, const float* inside_w, const float* outside_w,
                          float* losses_out, int batch, int num_anchors, int height, int width,
                          float sigma, void* workspace, size_t workspace_bytes, void* stream);
/* Gradients of up[0]*loss_cls + up[1]*loss_box (upstream: 2 floats on the DEVICE or NULL = 1, 1)
 * w.r.t. cls_score and bbox_pred, given losses_out of the forward call. */
int tlod_rpn_loss_backward(const float* cls_score, const float* labels, const float* bbox_pred,
                           const float* bbox_targets, const float* inside_w, const float* outside_w,
                           const float* losses_out, const float* upstream, float* grad_cls_score,
                           float* grad_bbox_pred, int batch, int num_anchors, int height, int width,
                           float sigma, void* stream);

/* ------------------------------------------------------------------------ */
/* MAF / PT-MAF scale-reduce rearrangement and label layers */
/*   lib/MAF/drm.py:21-42 (chunk / reshape / cat loops = space-to-depth),     */
/*   lib/DAF/LabelResizeLayer.py:42-58 (instance labels in blocks of 256)     */
/* ------------------------------------------------------------------------ */
/* in (batch, C, H, W) -> out (batch, C*s*s, H/s, W/s):
 * out[b, c*s*s + dy*s + dx, i, j] = in[b, c, i*s + dy, j*s + dx]; the border beyond
 * s*floor(H/s), s*floor(W/s) is dropped (backward writes zeros there). */
int tlod_space_to_depth_forward(const float* in, float* out, int batch, int channels, int height,
                                int width, int scale, void* stream);
int tlod_space_to_depth_backward(const float* grad_out, float* grad_in, int batch, int channels,
                                 int height, int width, int scale, void* stream);
/* out[r] = domain_labels[r / minibatch] for r < images * minibatch, else `fill`
 * (DAF: np.ones -> 1, MAF / ATF: np.zeros -> 0); domain_labels (images) on the device. */
int tlod_instance_labels(const float* domain_labels, float* out, int rows, int images, int minibatch,
                         float fill, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TLOD_B200_H */
