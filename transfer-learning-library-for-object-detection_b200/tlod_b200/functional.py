"""Torch-facing wrappers over the C ABI (PyTorch is plumbing: device memory and streams).

Every function takes CUDA fp32 tensors, allocates outputs / workspaces with torch
on the tensors' device, and enqueues the kernels on torch's current stream.  None of
them synchronises the host.  A CPU tensor is an error -- there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib


def _require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("tlod_b200 has no CPU implementation: got a %s tensor" % t.device)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _workspace(device, nbytes: int, tag: str) -> torch.Tensor:
    """Scratch buffer for one call, from torch's caching allocator on the current stream.

    Never shared between calls: a cached buffer per (device, stream) would be baked into every
    CUDA graph captured on that stream, and graphs replayed concurrently (bench.py runs the source
    and the target domain on two streams) would then race on it.  A per-call allocation is
    stream-ordered when eager, and owned by the graph's private pool when captured."""
    del tag
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------
# RoIAlign / RoIPool
# ---------------------------------------------------------------------------
def roi_align_plan(rois, feature_size, aligned_h: int, aligned_w: int, spatial_scale: float):
    """Per-`rois` plan (sampling tables, image-sorted RoI list, backward column chains) shared by
    roi_align_forward and roi_align_backward.  Returns None where the planned kernels do not
    apply (the generic kernels are used then)."""
    _require_cuda(rois)
    rois = _f32(rois)
    B, _, H, W = [int(v) for v in feature_size]
    R = rois.size(0)
    if R == 0 or B > 1024 or aligned_h > 16 or aligned_w > 16:
        return None
    nbytes = lib.tlod_roi_align_plan_bytes(B, R)
    plan = torch.empty((nbytes,), dtype=torch.uint8, device=rois.device)
    with torch.cuda.device(rois.device):
        rc = lib.tlod_roi_align_plan(rois.data_ptr(), B, H, W, R, int(aligned_h), int(aligned_w),
                                     float(spatial_scale), plan.data_ptr(), plan.numel(), _stream(rois.device))
    if rc == _lib.ERR_UNSUPPORTED:  # map too large for the planned kernels: the generic ones serve it
        return None
    check(rc, "tlod_roi_align_plan")
    return plan


def roi_align_forward(features, rois, aligned_h: int, aligned_w: int, spatial_scale: float, plan=None,
                      use_plan: bool = True):
    _require_cuda(features, rois, plan)
    features, rois = _f32(features), _f32(rois)
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be (R, 5) [batch_idx, x1, y1, x2, y2]")
    B, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty((R, C, aligned_h, aligned_w), dtype=torch.float32, device=features.device)
    if plan is None and use_plan:
        plan = roi_align_plan(rois, features.shape, aligned_h, aligned_w, spatial_scale)
    with torch.cuda.device(features.device):
        check(lib.tlod_roi_align_forward(features.data_ptr(), rois.data_ptr(), out.data_ptr(), B, C, H, W, R,
                                         int(aligned_h), int(aligned_w), float(spatial_scale),
                                         _ptr(plan), 0 if plan is None else plan.numel(),
                                         _stream(features.device)), "tlod_roi_align_forward")
    return out


def roi_align_backward(top_grad, rois, feature_size, spatial_scale: float, plan=None, use_plan: bool = True):
    _require_cuda(top_grad, rois, plan)
    top_grad, rois = _f32(top_grad), _f32(rois)
    B, C, H, W = [int(v) for v in feature_size]
    R, _, AH, AW = top_grad.shape
    grad = torch.empty((B, C, H, W), dtype=torch.float32, device=top_grad.device)
    if plan is None and use_plan:
        plan = roi_align_plan(rois, feature_size, AH, AW, spatial_scale)
    with torch.cuda.device(top_grad.device):
        check(lib.tlod_roi_align_backward(top_grad.data_ptr(), rois.data_ptr(), grad.data_ptr(), B, C, H, W, R,
                                          AH, AW, float(spatial_scale), _ptr(plan),
                                          0 if plan is None else plan.numel(), _stream(top_grad.device)),
              "tlod_roi_align_backward")
    return grad


def roi_align_avg_forward(features, rois, pooled_h: int, pooled_w: int, spatial_scale: float, plan=None):
    """RoIAlignAvg (modules/roi_align.py:20-29): samples at (pooled_h + 1, pooled_w + 1), 2x2 / stride-1
    average -> (R, C, pooled_h, pooled_w).  7 x 7 with a plan is one kernel; the sample tensor never exists."""
    _require_cuda(features, rois, plan)
    features, rois = _f32(features), _f32(rois)
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be (R, 5) [batch_idx, x1, y1, x2, y2]")
    B, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty((R, C, pooled_h, pooled_w), dtype=torch.float32, device=features.device)
    args = (features.data_ptr(), rois.data_ptr(), out.data_ptr(), B, C, H, W, R, int(pooled_h), int(pooled_w),
            float(spatial_scale), _ptr(plan), 0 if plan is None else plan.numel())
    with torch.cuda.device(features.device):
        rc = lib.tlod_roi_align_avg_forward(*args, None, 0, _stream(features.device))
        if rc == _lib.ERR_WORKSPACE:  # not the fused shape: the composed path needs the sample tensor
            scratch = torch.empty((R, C, pooled_h + 1, pooled_w + 1), dtype=torch.float32, device=features.device)
            rc = lib.tlod_roi_align_avg_forward(*args, scratch.data_ptr(), scratch.numel() * 4,
                                                _stream(features.device))
        check(rc, "tlod_roi_align_avg_forward")
    return out


def roi_align_avg_backward(top_grad, rois, feature_size, spatial_scale: float, plan=None):
    """Adjoint of roi_align_avg_forward: (R, C, ph, pw) -> (B, C, H, W)."""
    _require_cuda(top_grad, rois, plan)
    top_grad, rois = _f32(top_grad), _f32(rois)
    B, C, H, W = [int(v) for v in feature_size]
    R, _, PH, PW = top_grad.shape
    grad = torch.empty((B, C, H, W), dtype=torch.float32, device=top_grad.device)
    scratch = torch.empty((R, C, PH + 1, PW + 1), dtype=torch.float32, device=top_grad.device)
    with torch.cuda.device(top_grad.device):
        check(lib.tlod_roi_align_avg_backward(top_grad.data_ptr(), rois.data_ptr(), grad.data_ptr(), B, C, H, W, R,
                                              PH, PW, float(spatial_scale), _ptr(plan),
                                              0 if plan is None else plan.numel(), scratch.data_ptr(),
                                              max(scratch.numel() * 4, 4), _stream(top_grad.device)),
              "tlod_roi_align_avg_backward")
    return grad


def avgpool2x2_forward(x):
    """(..., h, w) -> (..., h-1, w-1): avg_pool2d(kernel_size=2, stride=1)."""
    _require_cuda(x)
    x = _f32(x)
    h, w = int(x.shape[-2]), int(x.shape[-1])
    out = torch.empty(tuple(x.shape[:-2]) + (h - 1, w - 1), dtype=torch.float32, device=x.device)
    tiles = x.numel() // (h * w)
    with torch.cuda.device(x.device):
        check(lib.tlod_avgpool2x2_forward(x.data_ptr(), out.data_ptr(), tiles, h, w, _stream(x.device)),
              "tlod_avgpool2x2_forward")
    return out


def avgpool2x2_backward(grad_out):
    """Adjoint of avgpool2x2_forward: (..., h-1, w-1) -> (..., h, w)."""
    _require_cuda(grad_out)
    g = _f32(grad_out)
    h, w = int(g.shape[-2]) + 1, int(g.shape[-1]) + 1
    out = torch.empty(tuple(g.shape[:-2]) + (h, w), dtype=torch.float32, device=g.device)
    tiles = out.numel() // (h * w)
    with torch.cuda.device(g.device):
        check(lib.tlod_avgpool2x2_backward(g.data_ptr(), out.data_ptr(), tiles, h, w, _stream(g.device)),
              "tlod_avgpool2x2_backward")
    return out


def roi_pool_forward(features, rois, pooled_h: int, pooled_w: int, spatial_scale: float):
    _require_cuda(features, rois)
    features, rois = _f32(features), _f32(rois)
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be (R, 5) [batch_idx, x1, y1, x2, y2]")
    B, C, H, W = features.shape
    R = rois.size(0)
    out = torch.empty((R, C, pooled_h, pooled_w), dtype=torch.float32, device=features.device)
    argmax = torch.empty((R, C, pooled_h, pooled_w), dtype=torch.int32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib.tlod_roi_pool_forward(features.data_ptr(), rois.data_ptr(), out.data_ptr(), argmax.data_ptr(),
                                        B, C, H, W, R, int(pooled_h), int(pooled_w), float(spatial_scale),
                                        _stream(features.device)), "tlod_roi_pool_forward")
    return out, argmax


def roi_pool_backward(top_grad, argmax, rois, feature_size, spatial_scale: float):
    _require_cuda(top_grad, argmax, rois)
    top_grad, rois = _f32(top_grad), _f32(rois)
    argmax = argmax.contiguous()
    B, C, H, W = [int(v) for v in feature_size]
    R, _, PH, PW = top_grad.shape
    grad = torch.empty((B, C, H, W), dtype=torch.float32, device=top_grad.device)
    with torch.cuda.device(top_grad.device):
        check(lib.tlod_roi_pool_backward(top_grad.data_ptr(), argmax.data_ptr(), rois.data_ptr(), grad.data_ptr(),
                                         B, C, H, W, R, PH, PW, float(spatial_scale), _stream(top_grad.device)),
              "tlod_roi_pool_backward")
    return grad


def roi_crop_forward(features, grid_yx):
    """features (ib, C, H, W), grid_yx (ob, GH, GW, 2) -> (ob, C, GH, GW)."""
    _require_cuda(features, grid_yx)
    features, grid_yx = _f32(features), _f32(grid_yx)
    ib, C, H, W = features.shape
    ob, GH, GW, two = grid_yx.shape
    if two != 2:
        raise ValueError("grid must be (N, GH, GW, 2) = (y, x)")
    out = torch.empty((ob, C, GH, GW), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib.tlod_roi_crop_forward(features.data_ptr(), grid_yx.data_ptr(), out.data_ptr(), ib, C, H, W, ob,
                                        GH, GW, _stream(features.device)), "tlod_roi_crop_forward")
    return out


def roi_crop_backward(grad_output, grid_yx, feature_size):
    _require_cuda(grad_output, grid_yx)
    grad_output, grid_yx = _f32(grad_output), _f32(grid_yx)
    ib, C, H, W = [int(v) for v in feature_size]
    ob, GH, GW, _ = grid_yx.shape
    grad = torch.empty((ib, C, H, W), dtype=torch.float32, device=grad_output.device)
    with torch.cuda.device(grad_output.device):
        check(lib.tlod_roi_crop_backward(grad_output.data_ptr(), grid_yx.data_ptr(), grad.data_ptr(), ib, C, H, W,
                                         ob, GH, GW, _stream(grad_output.device)), "tlod_roi_crop_backward")
    return grad


def roi_crop_pool_supported(features, grid_h: int, grid_w: int) -> bool:
    """Shapes served by the fused RoICrop + max_pool2d(2, 2) kernels (the C side re-checks)."""
    return grid_h == 14 and grid_w == 14 and features.size(1) % 16 == 0 and \
        features.size(2) * features.size(3) * 64 <= 190 * 1024


def roi_crop_pool_forward(features, grid_y, grid_x):
    """RoICrop on the axis-aligned grid (grid_y (R, 14), grid_x (R, 14)) + max_pool2d(2, 2) in one kernel:
    -> (pooled (R, C, 7, 7), argmax (R, C, 7, 7) uint8)."""
    _require_cuda(features, grid_y, grid_x)
    features, grid_y, grid_x = _f32(features), _f32(grid_y), _f32(grid_x)
    ib, C, H, W = features.shape
    ob, GH = grid_y.shape
    GW = grid_x.size(1)
    dev = features.device
    out = torch.empty((ob, C, GH // 2, GW // 2), dtype=torch.float32, device=dev)
    arg = torch.empty((ob, C, GH // 2, GW // 2), dtype=torch.uint8, device=dev)
    ws = _workspace(dev, lib.tlod_roi_crop_pool_workspace_bytes(ob), "crop_pool")
    with torch.cuda.device(dev):
        check(lib.tlod_roi_crop_pool_forward(features.data_ptr(), grid_y.data_ptr(), grid_x.data_ptr(),
                                             out.data_ptr(), arg.data_ptr(), ib, C, H, W, ob, GH, GW, ws.data_ptr(),
                                             ws.numel(), _stream(dev)), "tlod_roi_crop_pool_forward")
    return out, arg


def roi_crop_pool_backward(grad_output, argmax, grid_y, grid_x, feature_size):
    _require_cuda(grad_output, argmax, grid_y, grid_x)
    grad_output, grid_y, grid_x = _f32(grad_output), _f32(grid_y), _f32(grid_x)
    ib, C, H, W = [int(v) for v in feature_size]
    ob, GH = grid_y.shape
    GW = grid_x.size(1)
    dev = grad_output.device
    grad = torch.empty((ib, C, H, W), dtype=torch.float32, device=dev)
    ws = _workspace(dev, lib.tlod_roi_crop_pool_workspace_bytes(ob), "crop_pool")
    with torch.cuda.device(dev):
        check(lib.tlod_roi_crop_pool_backward(grad_output.data_ptr(), argmax.contiguous().data_ptr(),
                                              grid_y.data_ptr(), grid_x.data_ptr(), grad.data_ptr(), ib, C, H, W, ob,
                                              GH, GW, ws.data_ptr(), ws.numel(), _stream(dev)),
              "tlod_roi_crop_pool_backward")
    return grad


# ---------------------------------------------------------------------------
# NMS / proposals
# ---------------------------------------------------------------------------
def nms_device(dets, thresh: float, max_keep: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """dets (n, >=4) sorted by score -> (keep int32 (n,), num int32 (1,)), both on the
    device, no host synchronisation (keep[num:] is unspecified)."""
    _require_cuda(dets)
    dets = _f32(dets)
    n, stride = dets.shape
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=dets.device)
    num = torch.zeros((1,), dtype=torch.int32, device=dets.device)
    if n == 0:
        return keep[:0], num
    nbytes = lib.tlod_nms_workspace_bytes(n)
    ws = _workspace(dets.device, nbytes, "nms")
    with torch.cuda.device(dets.device):
        check(lib.tlod_nms(dets.data_ptr(), n, stride, float(thresh), int(max_keep), keep.data_ptr(),
                           num.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dets.device)), "tlod_nms")
    return keep, num


def class_nms(scores, pred_boxes, score_thresh: float, nms_thresh: float, first_class: int = 1):
    """Per-class test-time NMS of one image (methods/*/*_test.py loop), all classes in one call.
    scores (R, K), pred_boxes (R, 4K) or (R, 4) -> list over classes first_class..K-1 of
    (n_j, 5) tensors [x1, y1, x2, y2, score] on the device.  One host synchronisation (the
    per-class counts) instead of the reference's one per class."""
    _require_cuda(scores, pred_boxes)
    scores, boxes = _f32(scores), _f32(pred_boxes)
    R, K = scores.shape
    dev = scores.device
    nc = K - first_class
    if R == 0 or nc <= 0:
        return [scores.new_zeros((0, 5)) for _ in range(max(nc, 0))]
    Rp = lib.tlod_class_nms_padded_rows(R)
    dets = torch.empty((nc, Rp, 5), dtype=torch.float32, device=dev)
    order = torch.empty((nc, Rp), dtype=torch.int32, device=dev)
    keep = torch.empty((nc, Rp), dtype=torch.int32, device=dev)
    meta = torch.empty((3, nc), dtype=torch.int32, device=dev)  # num, count, valid
    ws = _workspace(dev, lib.tlod_class_nms_workspace_bytes(R, K), "class_nms")
    with torch.cuda.device(dev):
        check(lib.tlod_class_nms(scores.data_ptr(), boxes.data_ptr(), R, K, int(first_class), boxes.size(1),
                                 float(score_thresh), float(nms_thresh), dets.data_ptr(), order.data_ptr(),
                                 keep.data_ptr(), meta[0].data_ptr(), meta[1].data_ptr(), meta[2].data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream(dev)), "tlod_class_nms")
    valid = meta[2].tolist()  # the one synchronisation
    return [dets[c].index_select(0, keep[c, :valid[c]].long()) for c in range(nc)]


def proposals(scores, deltas, im_info, anchors, feat_stride: int, pre_nms_topN: int, post_nms_topN: int,
              nms_thresh: float, return_debug: bool = False):
    """Fused batched proposal layer -> rois (B, post_nms_topN, 5)."""
    _require_cuda(scores, deltas, im_info, anchors)
    scores, deltas, im_info, anchors = _f32(scores), _f32(deltas), _f32(im_info), _f32(anchors)
    B, A2, H, W = scores.shape
    A = A2 // 2
    if anchors.shape != (A, 4) or deltas.shape != (B, 4 * A, H, W) or im_info.shape[0] != B:
        raise ValueError("inconsistent proposal-layer shapes")
    dev = scores.device
    rois = torch.empty((B, post_nms_topN, 5), dtype=torch.float32, device=dev)
    n_sorted = lib.tlod_proposals_n_sorted(B, A, H, W, int(pre_nms_topN))
    order = boxes = num = None
    if return_debug:
        order = torch.empty((B, n_sorted), dtype=torch.int32, device=dev)
        boxes = torch.empty((B, n_sorted, 4), dtype=torch.float32, device=dev)
        num = torch.empty((B,), dtype=torch.int32, device=dev)
    nbytes = lib.tlod_proposals_workspace_bytes(B, A, H, W, int(pre_nms_topN), int(post_nms_topN))
    ws = _workspace(dev, nbytes, "proposals")
    with torch.cuda.device(dev):
        check(lib.tlod_proposals(scores.data_ptr(), deltas.data_ptr(), im_info.data_ptr(), anchors.data_ptr(),
                                 rois.data_ptr(), B, A, H, W, int(feat_stride), int(pre_nms_topN),
                                 int(post_nms_topN), float(nms_thresh), _ptr(order), _ptr(boxes), _ptr(num),
                                 ws.data_ptr(), ws.numel(), _stream(dev)), "tlod_proposals")
    if return_debug:
        return rois, order, boxes, num
    return rois


# ---------------------------------------------------------------------------
# box arithmetic
# ---------------------------------------------------------------------------
def bbox_transform_inv(boxes, deltas, im_info=None):
    """boxes (B,N,4) or (N,4); deltas (B,N,4) -> (B,N,4), clipped if im_info given."""
    _require_cuda(boxes, deltas, im_info)
    boxes, deltas = _f32(boxes), _f32(deltas)
    B, N, _ = deltas.shape
    batched = 1 if boxes.dim() == 3 else 0
    out = torch.empty_like(deltas)
    info = None if im_info is None else _f32(im_info)
    with torch.cuda.device(deltas.device):
        check(lib.tlod_bbox_transform_inv_clip(boxes.data_ptr(), batched, deltas.data_ptr(), _ptr(info),
                                               out.data_ptr(), B, N, _stream(deltas.device)),
              "tlod_bbox_transform_inv_clip")
    return out


def clip_boxes_(boxes, im_info):
    """In place; boxes (B, N, 4k) contiguous fp32."""
    _require_cuda(boxes, im_info)
    if boxes.dtype != torch.float32 or not boxes.is_contiguous():
        raise ValueError("clip_boxes_ needs a contiguous fp32 tensor")
    B, N, K4 = boxes.shape
    info = _f32(im_info)
    with torch.cuda.device(boxes.device):
        check(lib.tlod_clip_boxes(boxes.data_ptr(), info.data_ptr(), B, N, K4 // 4, _stream(boxes.device)),
              "tlod_clip_boxes")
    return boxes


def bbox_overlaps_batch(anchors, gt_boxes):
    """anchors (N,4) | (B,N,4) | (B,N,5: batch idx first); gt (B,K,>=4) -> (B,N,K)."""
    _require_cuda(anchors, gt_boxes)
    anchors, gt = _f32(anchors), _f32(gt_boxes)
    B, K, gs = gt.shape
    if anchors.dim() == 2:
        batched, N, stride, off = 0, anchors.size(0), anchors.size(1), 0
    elif anchors.dim() == 3:
        batched, N, stride = 1, anchors.size(1), anchors.size(2)
        off = 0 if stride == 4 else 1
    else:
        raise ValueError("anchors input dimension is not correct.")
    out = torch.empty((B, N, K), dtype=torch.float32, device=gt.device)
    with torch.cuda.device(gt.device):
        check(lib.tlod_bbox_overlaps_batch(anchors.data_ptr(), batched, stride, off, gt.data_ptr(), gs,
                                           out.data_ptr(), B, N, K, _stream(gt.device)),
              "tlod_bbox_overlaps_batch")
    return out


def bbox_transform_batch(ex_rois, gt_rois):
    _require_cuda(ex_rois, gt_rois)
    ex, gt = _f32(ex_rois), _f32(gt_rois)
    B, N, _ = gt.shape
    batched = 1 if ex.dim() == 3 else 0
    out = torch.empty((B, N, 4), dtype=torch.float32, device=gt.device)
    with torch.cuda.device(gt.device):
        check(lib.tlod_bbox_transform_batch(ex.data_ptr(), batched, gt.data_ptr(), out.data_ptr(), B, N,
                                            _stream(gt.device)), "tlod_bbox_transform_batch")
    return out


# ---------------------------------------------------------------------------
# anchor targets
# ---------------------------------------------------------------------------
def anchor_labels(anchors, gt_boxes, negative_overlap: float, positive_overlap: float,
                  clobber_positives: bool = False, want_max: bool = False):
    _require_cuda(anchors, gt_boxes)
    anchors, gt = _f32(anchors), _f32(gt_boxes)
    B, K, gs = gt.shape
    N = anchors.size(0)
    dev = gt.device
    labels = torch.empty((B, N), dtype=torch.float32, device=dev)
    argmax = torch.empty((B, N), dtype=torch.int32, device=dev)
    mx = torch.empty((B, N), dtype=torch.float32, device=dev) if want_max else None
    ws = _workspace(dev, lib.tlod_anchor_labels_workspace_bytes(B, K), "anchor_labels")
    with torch.cuda.device(dev):
        check(lib.tlod_anchor_labels(anchors.data_ptr(), gt.data_ptr(), gs, labels.data_ptr(), argmax.data_ptr(),
                                     _ptr(mx), B, N, K, float(negative_overlap), float(positive_overlap),
                                     int(bool(clobber_positives)), ws.data_ptr(), ws.numel(), _stream(dev)),
              "tlod_anchor_labels")
    return (labels, argmax, mx) if want_max else (labels, argmax)


def anchor_targets_finalize(labels, argmax, anchors, gt_boxes, inv_index, num_anchors: int, height: int,
                            width: int, inside_weight: float, positive_weight: float, negative_weight: float,
                            weights_dev=None):
    """weights_dev: optional (3,) fp32 device tensor {inside, positive, negative}; when given the three
    float arguments are ignored (the launch can then be captured in a CUDA graph and replayed with
    other weights)."""
    _require_cuda(labels, argmax, anchors, gt_boxes, inv_index, weights_dev)
    labels, anchors, gt = _f32(labels), _f32(anchors), _f32(gt_boxes)
    argmax = argmax.contiguous()
    inv_index = inv_index.contiguous()
    B, N = labels.shape
    K, gs = gt.size(1), gt.size(2)
    A, H, W = int(num_anchors), int(height), int(width)
    dev = labels.device
    labels_out = torch.empty((B, 1, A * H, W), dtype=torch.float32, device=dev)
    targets = torch.empty((B, 4 * A, H, W), dtype=torch.float32, device=dev)
    inside = torch.empty_like(targets)
    outside = torch.empty_like(targets)
    if weights_dev is not None:
        if weights_dev.dtype != torch.float32 or weights_dev.numel() != 3 or not weights_dev.is_contiguous():
            raise ValueError("weights_dev must be a contiguous (3,) fp32 tensor")
        with torch.cuda.device(dev):
            check(lib.tlod_anchor_targets_finalize_dev(labels.data_ptr(), argmax.data_ptr(), anchors.data_ptr(),
                                                       gt.data_ptr(), gs, inv_index.data_ptr(),
                                                       labels_out.data_ptr(), targets.data_ptr(), inside.data_ptr(),
                                                       outside.data_ptr(), B, N, K, A, H, W, weights_dev.data_ptr(),
                                                       _stream(dev)), "tlod_anchor_targets_finalize_dev")
        return labels_out, targets, inside, outside
    with torch.cuda.device(dev):
        check(lib.tlod_anchor_targets_finalize(labels.data_ptr(), argmax.data_ptr(), anchors.data_ptr(),
                                               gt.data_ptr(), gs, inv_index.data_ptr(), labels_out.data_ptr(),
                                               targets.data_ptr(), inside.data_ptr(), outside.data_ptr(), B, N, K,
                                               A, H, W, float(inside_weight), float(positive_weight),
                                               float(negative_weight), _stream(dev)),
              "tlod_anchor_targets_finalize")
    return labels_out, targets, inside, outside


# ---------------------------------------------------------------------------
# proposal targets
# ---------------------------------------------------------------------------
def roi_gt_assign(all_rois, gt_boxes):
    """all_rois (B, n, 5) [image, x1, y1, x2, y2], gt (B, K, 5) -> (max_overlaps (B,n) f32,
    assignment (B,n) int32, labels (B,n) f32 = class of the assigned gt)."""
    _require_cuda(all_rois, gt_boxes)
    rois, gt = _f32(all_rois), _f32(gt_boxes)
    B, n, rs = rois.shape
    K, gs = gt.size(1), gt.size(2)
    dev = rois.device
    mx = torch.empty((B, n), dtype=torch.float32, device=dev)
    asg = torch.empty((B, n), dtype=torch.int32, device=dev)
    lab = torch.empty((B, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.tlod_roi_gt_assign(rois.data_ptr(), rs, 1 if rs == 5 else 0, gt.data_ptr(), gs, mx.data_ptr(),
                                     asg.data_ptr(), lab.data_ptr(), B, n, K, _stream(dev)), "tlod_roi_gt_assign")
    return mx, asg, lab


def proposal_targets(all_rois, gt_boxes, assignment, labels, keep, fg_count, means, stds, inside_weights,
                     normalize: bool):
    _require_cuda(all_rois, gt_boxes, assignment, labels, keep, fg_count)
    rois, gt, labels = _f32(all_rois), _f32(gt_boxes), _f32(labels)
    assignment, keep, fg_count = assignment.contiguous(), keep.contiguous(), fg_count.contiguous()
    B, n, rs = rois.shape
    K, gs = gt.size(1), gt.size(2)
    P = keep.size(1)
    dev = rois.device
    rois_out = torch.empty((B, P, 5), dtype=torch.float32, device=dev)
    labels_out = torch.empty((B, P), dtype=torch.float32, device=dev)
    targets = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    inside = torch.empty_like(targets)
    outside = torch.empty_like(targets)
    f4 = ctypes.c_float * 4
    with torch.cuda.device(dev):
        check(lib.tlod_proposal_targets(rois.data_ptr(), rs, 1 if rs == 5 else 0, gt.data_ptr(), gs,
                                        assignment.data_ptr(), labels.data_ptr(), keep.data_ptr(),
                                        fg_count.data_ptr(), rois_out.data_ptr(), labels_out.data_ptr(),
                                        targets.data_ptr(), inside.data_ptr(), outside.data_ptr(), B, n, K, P,
                                        f4(*[float(v) for v in means]), f4(*[float(v) for v in stds]),
                                        f4(*[float(v) for v in inside_weights]), int(bool(normalize)),
                                        _stream(dev)), "tlod_proposal_targets")
    return rois_out, labels_out, targets, inside, outside


# ---------------------------------------------------------------------------
# RPN head losses
# ---------------------------------------------------------------------------
def rpn_loss_forward(cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w, sigma: float = 3.0):
    """-> (4,) device tensor [loss_cls, loss_box, kept anchors, foreground anchors]."""
    _require_cuda(cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w)
    ts = [_f32(t) for t in (cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w)]
    B, A2, H, W = ts[0].shape
    A = A2 // 2
    if ts[1].numel() != B * A * H * W or any(t.numel() != B * 4 * A * H * W for t in ts[2:]):
        raise ValueError("inconsistent RPN loss shapes")
    dev = ts[0].device
    out = torch.empty((4,), dtype=torch.float32, device=dev)
    ws = _workspace(dev, lib.tlod_rpn_loss_workspace_bytes(), "rpn_loss")
    with torch.cuda.device(dev):
        check(lib.tlod_rpn_loss_forward(*[t.data_ptr() for t in ts], out.data_ptr(), B, A, H, W, float(sigma),
                                        ws.data_ptr(), ws.numel(), _stream(dev)), "tlod_rpn_loss_forward")
    return out


def rpn_loss_backward(cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w, losses, upstream=None,
                      sigma: float = 3.0):
    _require_cuda(cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w, losses, upstream)
    ts = [_f32(t) for t in (cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w)]
    B, A2, H, W = ts[0].shape
    A = A2 // 2
    g_score = torch.empty_like(ts[0])
    g_pred = torch.empty_like(ts[2])
    up = None if upstream is None else _f32(upstream)
    dev = ts[0].device
    with torch.cuda.device(dev):
        check(lib.tlod_rpn_loss_backward(*[t.data_ptr() for t in ts], losses.data_ptr(), _ptr(up),
                                         g_score.data_ptr(), g_pred.data_ptr(), B, A, H, W, float(sigma),
                                         _stream(dev)), "tlod_rpn_loss_backward")
    return g_score, g_pred


# ---------------------------------------------------------------------------
# MAF scale-reduce rearrangement, label layers
# ---------------------------------------------------------------------------
def space_to_depth_forward(x, scale: int):
    _require_cuda(x)
    x = _f32(x)
    B, C, H, W = x.shape
    s = int(scale)
    out = torch.empty((B, C * s * s, H // s, W // s), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.tlod_space_to_depth_forward(x.data_ptr(), out.data_ptr(), B, C, H, W, s, _stream(x.device)),
              "tlod_space_to_depth_forward")
    return out


def space_to_depth_backward(grad_out, input_size, scale: int):
    _require_cuda(grad_out)
    g = _f32(grad_out)
    B, C, H, W = [int(v) for v in input_size]
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        check(lib.tlod_space_to_depth_backward(g.data_ptr(), out.data_ptr(), B, C, H, W, int(scale),
                                               _stream(g.device)), "tlod_space_to_depth_backward")
    return out


def instance_labels(domain_labels, rows: int, minibatch: int = 256, fill: float = 1.0):
    """(images,) domain labels on the device -> (rows, 1): LabelResizeLayer.py:42-58 without the
    two .cpu() copies."""
    _require_cuda(domain_labels)
    d = _f32(domain_labels).view(-1)
    out = torch.empty((int(rows), 1), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        check(lib.tlod_instance_labels(d.data_ptr(), out.data_ptr(), int(rows), d.numel(), int(minibatch),
                                       float(fill), _stream(d.device)), "tlod_instance_labels")
    return out


# ---------------------------------------------------------------------------
# GRL + DA losses
# ---------------------------------------------------------------------------
def grl_backward(grad, alpha: float, row_weight=None):
    _require_cuda(grad, row_weight)
    grad = _f32(grad)
    out = torch.empty_like(grad)
    with torch.cuda.device(grad.device):
        if row_weight is None:
            check(lib.tlod_grl_backward(grad.data_ptr(), out.data_ptr(), float(alpha), grad.numel(),
                                        _stream(grad.device)), "tlod_grl_backward")
        else:
            w = _f32(row_weight).view(-1)
            rows = grad.size(0)
            cols = grad.numel() // max(rows, 1)
            if w.numel() != rows:
                raise ValueError("row_weight must have one entry per row")
            check(lib.tlod_grl_backward_weighted(grad.data_ptr(), w.data_ptr(), out.data_ptr(), float(alpha), rows,
                                                 cols, _stream(grad.device)), "tlod_grl_backward_weighted")
    return out


def da_loss_forward(img_score, ins_prob, domain_label: int, ins_label=None):
    """-> losses (4,) device tensor [img_nll_mean, ins_bce_mean, cst_mse_sum, consistency_target]."""
    _require_cuda(img_score, ins_prob, ins_label)
    img_score, ins_prob = _f32(img_score), _f32(ins_prob).view(-1)
    B, two, H, W = img_score.shape
    if two != 2:
        raise ValueError("img_score must be (B, 2, H, W)")
    lab = None if ins_label is None else _f32(ins_label).view(-1)
    dev = img_score.device
    out = torch.empty((4,), dtype=torch.float32, device=dev)
    ws = None
    if B * H * W > (1 << 16):  # large maps: the image head is reduced over many CTAs first
        ws = _workspace(dev, int(lib.tlod_da_loss_workspace_bytes()), "da_loss")
    with torch.cuda.device(dev):
        check(lib.tlod_da_loss_forward(img_score.data_ptr(), ins_prob.data_ptr(), _ptr(lab), int(domain_label),
                                       out.data_ptr(), B, H, W, ins_prob.numel(), _ptr(ws),
                                       0 if ws is None else ws.numel(), _stream(dev)),
              "tlod_da_loss_forward")
    return out


def da_loss_backward(img_score, ins_prob, domain_label: int, losses, w_img: float, w_ins: float, w_cst: float,
                     ins_label=None, upstream=None):
    """upstream: optional (3,) device tensor multiplied into (w_img, w_ins, w_cst) on the device."""
    _require_cuda(img_score, ins_prob, losses, ins_label, upstream)
    up = None if upstream is None else _f32(upstream)
    img_score, ins_prob = _f32(img_score), _f32(ins_prob)
    B, _, H, W = img_score.shape
    lab = None if ins_label is None else _f32(ins_label).view(-1)
    g_img = torch.empty_like(img_score)
    g_ins = torch.empty_like(ins_prob)
    dev = img_score.device
    with torch.cuda.device(dev):
        check(lib.tlod_da_loss_backward(img_score.data_ptr(), ins_prob.data_ptr(), _ptr(lab), int(domain_label),
                                        losses.data_ptr(), _ptr(up), float(w_img), float(w_ins), float(w_cst),
                                        g_img.data_ptr(), g_ins.data_ptr(), B, H, W, ins_prob.numel(),
                                        _stream(dev)), "tlod_da_loss_backward")
    return g_img, g_ins


def _da_levels(scores, labels):
    n = len(scores)
    if not 1 <= n <= 4:
        raise ValueError("1..4 feature levels")
    scores = [_f32(s) for s in scores]
    for s in scores:
        if s.dim() != 4 or s.size(1) != 2:
            raise ValueError("each score map must be (B, 2, H, W)")
    labs = [None] * n if labels is None else [None if t is None else t.long().contiguous() for t in labels]
    arr_p = (ctypes.c_void_p * n)
    arr_i = (ctypes.c_int * n)
    return (scores, labs, arr_p(*[s.data_ptr() for s in scores]), arr_p(*[_ptr(t) for t in labs]),
            arr_i(*[s.size(0) for s in scores]), arr_i(*[s.size(2) for s in scores]),
            arr_i(*[s.size(3) for s in scores]))


def da_image_loss_forward(scores, domain_label: int, labels=None, ignore_index: int = -100):
    """Image-level DA losses of 1..4 feature levels in one launch (MAF conv3/conv4/conv5 heads, ATF's
    ignore_index = -1).  scores: list of (B, 2, H, W) logits; labels: None (= domain_label everywhere) or a
    list of int64 (B, H, W) maps / None per level.  -> (levels, 4) device tensor
    [nll mean, mean softmax prob of class domain_label, counted cells, 0] per level."""
    _require_cuda(*scores)
    scores, labs, sp, lp, bb, hh, ww = _da_levels(scores, labels)
    dev = scores[0].device
    out = torch.empty((len(scores), 4), dtype=torch.float32, device=dev)
    ws = _workspace(dev, lib.tlod_da_image_loss_workspace_bytes(), "da_image")
    with torch.cuda.device(dev):
        check(lib.tlod_da_image_loss_forward(len(scores), sp, lp, bb, hh, ww, int(domain_label), int(ignore_index),
                                             out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)),
              "tlod_da_image_loss_forward")
    return out


def da_image_loss_backward(scores, domain_label: int, out, labels=None, ignore_index: int = -100, upstream=None,
                           weights=None):
    """-> list of gradients w.r.t. each level's score map (of weights[l] * upstream[l] * loss[l])."""
    _require_cuda(*scores)
    scores, labs, sp, lp, bb, hh, ww = _da_levels(scores, labels)
    n = len(scores)
    dev = scores[0].device
    grads = [torch.empty_like(s) for s in scores]
    gp = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads])
    up = None if upstream is None else _f32(upstream)
    wt = None if weights is None else (ctypes.c_float * n)(*[float(w) for w in weights])
    with torch.cuda.device(dev):
        check(lib.tlod_da_image_loss_backward(n, sp, lp, bb, hh, ww, int(domain_label), int(ignore_index),
                                              out.data_ptr(), _ptr(up), wt, gp, _stream(dev)),
              "tlod_da_image_loss_backward")
    return grads


launch_count = _lib.launch_count
