"""CPU oracle for the RoI / proposal hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``transfer-learning-library-for-object-detection_b200/``) never does: it fails
loudly when its CUDA extension is missing instead of falling back to this code.

Two layers:

* ``liboracle.so`` (``oracle/tlod_oracle.c``): byte/float-exact C restatement of
  the reference kernels -- RoIAlign, RoIPool, NMS, the proposal layer, batched
  IoU, anchor labels -- each citing the reference file:line it follows.
* numpy restatements in this file of the small host-side pieces
  (``generate_anchors``, the anchor-target layer's subsampling / targets /
  weights, ``bbox_transform_batch``, the DA losses).

Parity pinning: ``oracle/validate_against_reference.py`` (run in the build
container, where /root/reference exists) checks these functions bit-for-bit
against the reference's own C (``oracle/_ref/libref_cpu.so``) and Python layers,
and writes the golden vectors under ``tests/golden``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    """Compile oracle/tlod_oracle.c (and oracle/_ref when the reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "tlod_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.orc_nms.restype = ctypes.c_int
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(c_float_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_int_p)


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().orc_num_threads())


# ---------------------------------------------------------------------------
# RoIAlign / RoIPool   (lib/model/roi_align/src/roi_align_kernel.cu,
#                       lib/model/roi_pooling/src/roi_pooling_kernel.cu)
# ---------------------------------------------------------------------------
def roi_align_forward(features, rois, aligned_h, aligned_w, spatial_scale):
    feat, fp = _f(features)
    r, rp = _f(rois)
    B, C, H, W = feat.shape
    out = np.empty((r.shape[0], C, aligned_h, aligned_w), np.float32)
    lib().orc_roi_align_fwd(fp, ctypes.c_float(spatial_scale), r.shape[0], H, W, C, aligned_h,
                            aligned_w, rp, out.ctypes.data_as(c_float_p))
    return out


def roi_align_backward(top_grad, rois, feature_shape, spatial_scale, accumulate_double=False):
    g, gp = _f(top_grad)
    r, rp = _f(rois)
    B, C, H, W = feature_shape
    R, _, AH, AW = g.shape
    out = np.empty((B, C, H, W), np.float32)
    lib().orc_roi_align_bwd(gp, ctypes.c_float(spatial_scale), B, R, H, W, C, AH, AW, rp,
                            out.ctypes.data_as(c_float_p), int(bool(accumulate_double)))
    return out


def roi_pool_forward(features, rois, pooled_h, pooled_w, spatial_scale):
    feat, fp = _f(features)
    r, rp = _f(rois)
    B, C, H, W = feat.shape
    out = np.empty((r.shape[0], C, pooled_h, pooled_w), np.float32)
    arg = np.empty((r.shape[0], C, pooled_h, pooled_w), np.int32)
    lib().orc_roi_pool_fwd(fp, ctypes.c_float(spatial_scale), r.shape[0], H, W, C, pooled_h,
                           pooled_w, rp, out.ctypes.data_as(c_float_p), arg.ctypes.data_as(c_int_p))
    return out, arg


def roi_pool_backward(top_grad, argmax, rois, feature_shape, spatial_scale):
    g, gp = _f(top_grad)
    a, ap = _i(argmax)
    r, rp = _f(rois)
    B, C, H, W = feature_shape
    R, _, PH, PW = g.shape
    out = np.empty((B, C, H, W), np.float32)
    lib().orc_roi_pool_bwd(gp, ap, ctypes.c_float(spatial_scale), B, R, H, W, C, PH, PW, rp,
                           out.ctypes.data_as(c_float_p))
    return out


# ---------------------------------------------------------------------------
# test-time per-class NMS   (methods/DAF/DAF_test.py:302-320, same in every *_test.py)
# ---------------------------------------------------------------------------
def per_class_nms(scores, pred_boxes, score_thresh, nms_thresh, first_class=1):
    """scores (R, K), pred_boxes (R, 4K) or (R, 4) -> list of (n_j, 5) arrays for classes
    first_class..K-1.  Sort = score descending, lower row first on ties (the reference's
    torch.sort leaves tie order unspecified)."""
    sc = np.asarray(scores, np.float32)
    bx = np.asarray(pred_boxes, np.float32)
    out = []
    for j in range(first_class, sc.shape[1]):
        inds = np.nonzero(sc[:, j] > np.float32(score_thresh))[0]
        if inds.size == 0:
            out.append(np.zeros((0, 5), np.float32))
            continue
        cls_scores = sc[inds, j]
        order = np.argsort(-cls_scores.astype(np.float64), kind="stable")
        cls_boxes = bx[inds] if bx.shape[1] == 4 else bx[inds][:, j * 4:(j + 1) * 4]
        dets = np.concatenate([cls_boxes, cls_scores[:, None]], 1)[order]
        keep = nms(dets, nms_thresh)
        out.append(np.ascontiguousarray(dets[keep]))
    return out


# ---------------------------------------------------------------------------
# _ProposalTargetLayer   (lib/model/rpn/proposal_target_layer_cascade.py:33-212)
# ---------------------------------------------------------------------------
def proposal_target_layer(all_rois, gt_boxes, rng=np.random, batch_size=128, fg_fraction=0.25, fg_thresh=0.5,
                          bg_thresh_hi=0.5, bg_thresh_lo=0.1, means=(0.0, 0.0, 0.0, 0.0),
                          stds=(0.1, 0.1, 0.2, 0.2), inside_weights=(1.0, 1.0, 1.0, 1.0), normalize=True):
    """all_rois (B, N, 5), gt_boxes (B, K, 5) -> (rois, labels, bbox_targets, inside_w, outside_w).
    ``rng`` (permutation, rand) is consumed in exactly the reference's order (:140-181)."""
    rois = np.asarray(all_rois, np.float32)
    gt = np.asarray(gt_boxes, np.float32)
    B = gt.shape[0]
    app = np.zeros_like(gt)
    app[:, :, 1:5] = gt[:, :, :4]
    rois = np.concatenate([rois, app], 1)
    P = int(batch_size)
    fg_per = int(np.round(fg_fraction * P))
    fg_per = 1 if fg_per == 0 else fg_per
    ov = bbox_overlaps_batch(rois, gt)                      # (B, n, K); batched, column 0 dropped
    max_ov = ov.max(2)
    assign = ov.argmax(2)                                    # ties -> lowest index
    labels = np.take_along_axis(gt[:, :, 4], assign, axis=1)
    labels_b = np.zeros((B, P), np.float32)
    rois_b = np.zeros((B, P, 5), np.float32)
    gt_b = np.zeros((B, P, 5), np.float32)
    for i in range(B):
        fg = np.nonzero(max_ov[i] >= np.float32(fg_thresh))[0]
        bg = np.nonzero((max_ov[i] < np.float32(bg_thresh_hi)) & (max_ov[i] >= np.float32(bg_thresh_lo)))[0]
        if fg.size > 0 and bg.size > 0:
            fg_this = min(fg_per, fg.size)
            fg = fg[rng.permutation(fg.size)[:fg_this]]
            bg = bg[np.floor(rng.rand(P - fg_this) * bg.size).astype(np.int64)]
        elif fg.size > 0:
            fg = fg[np.floor(rng.rand(P) * fg.size).astype(np.int64)]
            fg_this = P
            bg = bg[:0]
        elif bg.size > 0:
            bg = bg[np.floor(rng.rand(P) * bg.size).astype(np.int64)]
            fg_this = 0
            fg = fg[:0]
        else:
            raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
        keep = np.concatenate([fg, bg])
        labels_b[i] = labels[i][keep]
        if fg_this < P:
            labels_b[i][fg_this:] = 0
        rois_b[i] = rois[i][keep]
        rois_b[i, :, 0] = i
        gt_b[i] = gt[i][assign[i][keep]]
    t = bbox_transform_batch(rois_b[:, :, 1:5], gt_b[:, :, :4])
    if normalize:
        t = (t - np.asarray(means, np.float32)) / np.asarray(stds, np.float32)
    targets = np.zeros((B, P, 4), np.float32)
    inside = np.zeros((B, P, 4), np.float32)
    pos = labels_b > 0
    targets[pos] = t[pos]
    inside[pos] = np.asarray(inside_weights, np.float32)
    outside = (inside > 0).astype(np.float32)
    return rois_b, labels_b, targets, inside, outside


# ---------------------------------------------------------------------------
# RoICrop   (lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:12-23, 45-108, 111-190)
# ---------------------------------------------------------------------------
def roi_crop_forward(features, grid_yx):
    """features (ib, C, H, W), grid_yx (ob, GH, GW, 2) in [-1, 1] -> (ob, C, GH, GW)."""
    feat, fp = _f(features)
    g, gp = _f(grid_yx)
    ib, C, H, W = feat.shape
    ob, GH, GW, _ = g.shape
    out = np.empty((ob, C, GH, GW), np.float32)
    lib().orc_roi_crop_fwd(fp, gp, out.ctypes.data_as(c_float_p), ib, C, H, W, ob, GH, GW)
    return out


def roi_crop_backward(grad_out, grid_yx, feature_shape):
    """-> gradient w.r.t. the features (the reference kernel produces none for the grid)."""
    go, gop = _f(grad_out)
    g, gp = _f(grid_yx)
    ib, C, H, W = feature_shape
    ob, GH, GW, _ = g.shape
    out = np.empty((ib, C, H, W), np.float32)
    lib().orc_roi_crop_bwd(gop, gp, out.ctypes.data_as(c_float_p), ib, C, H, W, ob, GH, GW)
    return out


def affine_grid_gen(rois, input_size, grid_size):
    """_affine_grid_gen (lib/model/utils/net_utils.py:142-164) with torch-0.4's affine_grid
    (= align_corners=True): returns grid_xy (R, G, G, 2)."""
    rois = np.asarray(rois, np.float32)
    x1, y1, x2, y2 = [rois[:, k] / np.float32(16.0) for k in (1, 2, 3, 4)]
    height, width = np.float32(input_size[0]), np.float32(input_size[1])
    t00 = (x2 - x1) / (width - 1)
    t02 = (x1 + x2 - width + 1) / (width - 1)
    t11 = (y2 - y1) / (height - 1)
    t12 = (y1 + y2 - height + 1) / (height - 1)
    lin = np.linspace(-1.0, 1.0, grid_size, dtype=np.float32) if grid_size > 1 else np.zeros(1, np.float32)
    gx = t00[:, None, None] * lin[None, None, :] + t02[:, None, None] + 0 * lin[None, :, None]
    gy = t11[:, None, None] * lin[None, :, None] + t12[:, None, None] + 0 * lin[None, None, :]
    return np.stack([gx, gy], axis=3).astype(np.float32)


# ---------------------------------------------------------------------------
# NMS   (lib/model/nms/src/nms_cuda_kernel.cu:31-39, 41-85, 132-144)
# ---------------------------------------------------------------------------
def nms(dets, thresh, max_keep=0):
    """dets (n, >=4) sorted by score descending -> ascending keep indices (int32)."""
    d, dp = _f(dets)
    n = d.shape[0]
    if n == 0:
        return np.zeros((0,), np.int32)
    keep = np.empty((n,), np.int32)
    k = lib().orc_nms(dp, n, d.shape[1], ctypes.c_float(thresh), int(max_keep),
                      keep.ctypes.data_as(c_int_p))
    return keep[:k].copy()


# ---------------------------------------------------------------------------
# Anchors   (lib/model/rpn/generate_anchors.py:45-104)
# ---------------------------------------------------------------------------
def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    w = h = float(base_size)
    cx = cy = 0.5 * (base_size - 1)
    out = []
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    for rw, rh in zip(ws, hs):
        for s in scales:
            sw, sh = rw * s, rh * s
            out.append([cx - 0.5 * (sw - 1), cy - 0.5 * (sh - 1), cx + 0.5 * (sw - 1),
                        cy + 0.5 * (sh - 1)])
    return np.asarray(out, dtype=np.float64)


def shifted_anchors(anchors, feat_h, feat_w, feat_stride):
    """(K*A, 4) float32, flat index (y*W + x)*A + a   (proposal_layer.py:80-93)."""
    a = np.asarray(anchors, np.float32)
    sx = (np.arange(feat_w) * feat_stride).astype(np.float32)
    sy = (np.arange(feat_h) * feat_stride).astype(np.float32)
    gx, gy = np.meshgrid(sx, sy)
    shifts = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], 1)
    return (a[None, :, :] + shifts[:, None, :]).reshape(-1, 4).astype(np.float32)


# ---------------------------------------------------------------------------
# Proposal layer   (lib/model/rpn/proposal_layer.py:49-163)
# ---------------------------------------------------------------------------
def proposal_layer(scores, deltas, im_info, anchors, feat_stride, pre_nms_topN, post_nms_topN,
                   nms_thresh, exp_deltas=None, return_debug=False):
    """scores (B,2A,H,W), deltas (B,4A,H,W), im_info (B,3) -> rois (B,post,5).

    exp_deltas: element-wise exp(deltas) computed by the library whose results the
    caller wants to be exact against (torch CPU for the golden vectors, torch CUDA
    on the GPU box); None -> C expf."""
    sc, scp = _f(scores)
    dl, dlp = _f(deltas)
    info, infop = _f(im_info)
    an, anp = _f(anchors)
    B, A2, H, W = sc.shape
    A = A2 // 2
    N = A * H * W
    n_sorted = pre_nms_topN if 0 < pre_nms_topN < B * N else N
    n_sorted = min(n_sorted, N)
    out = np.empty((B, post_nms_topN, 5), np.float32)
    order = np.empty((B, n_sorted), np.int32)
    boxes = np.empty((B, n_sorted, 4), np.float32)
    num = np.empty((B,), np.int32)
    if exp_deltas is not None:
        ex, exp_p = _f(exp_deltas)
    else:
        exp_p = None
    lib().orc_proposals(scp, dlp, exp_p, infop, anp, B, A, H, W, int(feat_stride),
                        int(pre_nms_topN), int(post_nms_topN), ctypes.c_float(nms_thresh),
                        out.ctypes.data_as(c_float_p), order.ctypes.data_as(c_int_p),
                        boxes.ctypes.data_as(c_float_p), num.ctypes.data_as(c_int_p))
    if return_debug:
        return out, order, boxes, num
    return out


# ---------------------------------------------------------------------------
# IoU / anchor targets   (lib/model/rpn/bbox_transform.py:168-257,
#                         lib/model/rpn/anchor_target_layer.py:48-191)
# ---------------------------------------------------------------------------
def bbox_overlaps_batch(anchors, gt_boxes):
    """anchors (N,4) | (B,N,4) | (B,N,5 with batch idx first); gt (B,K,>=4) -> (B,N,K)."""
    gt, gtp = _f(gt_boxes)
    B, K, gs = gt.shape
    an = np.asarray(anchors, np.float32)
    batched = an.ndim == 3
    if batched and an.shape[2] == 5:
        an = an[:, :, 1:5]
    an, anp = _f(an)
    N = an.shape[-2]
    ov = np.empty((B, N, K), np.float32)
    lib().orc_bbox_overlaps_batch(anp, int(batched), gtp, gs, B, N, K, ov.ctypes.data_as(c_float_p))
    return ov


def anchor_labels(overlaps, neg_thresh=0.3, pos_thresh=0.7, clobber_positives=False):
    ov, ovp = _f(overlaps)
    B, N, K = ov.shape
    labels = np.empty((B, N), np.float32)
    argmax = np.empty((B, N), np.int32)
    mx = np.empty((B, N), np.float32)
    lib().orc_anchor_labels(ovp, B, N, K, ctypes.c_float(neg_thresh), ctypes.c_float(pos_thresh),
                            int(bool(clobber_positives)), labels.ctypes.data_as(c_float_p),
                            argmax.ctypes.data_as(c_int_p), mx.ctypes.data_as(c_float_p))
    return labels, argmax, mx


def bbox_transform_batch(ex_rois, gt_rois):
    """lib/model/rpn/bbox_transform.py:36-75; fp32, op by op.  ex (N,4)|(B,N,4), gt (B,N,4)."""
    ex = np.asarray(ex_rois, np.float32)
    gt = np.asarray(gt_rois, np.float32)
    one, half = np.float32(1.0), np.float32(0.5)
    ew = ex[..., 2] - ex[..., 0] + one
    eh = ex[..., 3] - ex[..., 1] + one
    ecx = ex[..., 0] + half * ew
    ecy = ex[..., 1] + half * eh
    gw = gt[..., 2] - gt[..., 0] + one
    gh = gt[..., 3] - gt[..., 1] + one
    gcx = gt[..., 0] + half * gw
    gcy = gt[..., 1] + half * gh
    with np.errstate(divide="ignore", invalid="ignore"):
        dx = (gcx - ecx) / ew
        dy = (gcy - ecy) / eh
        dw = np.log(gw / ew)
        dh = np.log(gh / eh)
    return np.stack([dx, dy, dw, dh], -1).astype(np.float32)


def anchor_target_layer(feat_h, feat_w, gt_boxes, im_info, anchors, feat_stride, rng=np.random,
                        neg_overlap=0.3, pos_overlap=0.7, clobber_positives=False,
                        fg_fraction=0.5, batchsize=256, inside_weight=1.0, positive_weight=-1.0):
    """lib/model/rpn/anchor_target_layer.py:48-191.  Returns the list of four
    arrays the reference returns.  ``rng`` must expose ``permutation`` and is
    consumed in exactly the reference's order (:123-145)."""
    gt = np.asarray(gt_boxes, np.float32)
    info = np.asarray(im_info, np.float32)
    B = gt.shape[0]
    A = np.asarray(anchors).shape[0]
    all_anchors = shifted_anchors(anchors, feat_h, feat_w, feat_stride)
    total = all_anchors.shape[0]
    lim_w, lim_h = int(info[0][1]), int(info[0][0])  # :86-87: first image, truncated
    keep = ((all_anchors[:, 0] >= 0) & (all_anchors[:, 1] >= 0) &
            (all_anchors[:, 2] < lim_w) & (all_anchors[:, 3] < lim_h))
    inds_inside = np.nonzero(keep)[0]
    anc = all_anchors[inds_inside]
    ov = bbox_overlaps_batch(anc, gt)
    labels, argmax, _ = anchor_labels(ov, neg_overlap, pos_overlap, clobber_positives)
    num_fg = int(fg_fraction * batchsize)
    sum_fg = (labels == 1).sum(1)
    sum_bg = (labels == 0).sum(1)
    i = 0
    for i in range(B):
        if sum_fg[i] > num_fg:
            fg_inds = np.nonzero(labels[i] == 1)[0]
            perm = rng.permutation(fg_inds.shape[0])
            labels[i][fg_inds[perm[:fg_inds.shape[0] - num_fg]]] = -1
        num_bg = batchsize - int((labels[i] == 1).sum())
        if sum_bg[i] > num_bg:
            bg_inds = np.nonzero(labels[i] == 0)[0]
            perm = rng.permutation(bg_inds.shape[0])
            labels[i][bg_inds[perm[:bg_inds.shape[0] - num_bg]]] = -1
    gt_sel = np.take_along_axis(gt[:, :, :4], argmax[:, :, None].astype(np.int64), axis=1)
    targets = bbox_transform_batch(anc, gt_sel)
    inside = np.zeros_like(labels)
    inside[labels == 1] = np.float32(inside_weight)
    assert positive_weight < 0, "only the reference's default RPN_POSITIVE_WEIGHT=-1 path"
    num_examples = int((labels[i] >= 0).sum())  # :156 uses the LAST image only
    w = np.float32(1.0 / num_examples) if num_examples > 0 else np.float32(np.inf)
    outside = np.zeros_like(labels)
    outside[labels == 1] = w
    outside[labels == 0] = w

    def unmap(data, fill):
        shape = (B, total) + data.shape[2:]
        ret = np.full(shape, fill, np.float32)
        ret[:, inds_inside] = data
        return ret

    labels = unmap(labels, -1).reshape(B, feat_h, feat_w, A).transpose(0, 3, 1, 2)
    labels = np.ascontiguousarray(labels).reshape(B, 1, A * feat_h, feat_w)
    targets = unmap(targets, 0).reshape(B, feat_h, feat_w, A * 4).transpose(0, 3, 1, 2)
    inside = np.repeat(unmap(inside, 0)[:, :, None], 4, 2).reshape(B, feat_h, feat_w, 4 * A)
    outside = np.repeat(unmap(outside, 0)[:, :, None], 4, 2).reshape(B, feat_h, feat_w, 4 * A)
    return [labels, np.ascontiguousarray(targets),
            np.ascontiguousarray(inside.transpose(0, 3, 1, 2)),
            np.ascontiguousarray(outside.transpose(0, 3, 1, 2))]


# ---------------------------------------------------------------------------
# GRL + domain-classifier losses   (lib/DAF/DA.py:19-33, lib/DAF/faster_rcnn.py:181-220)
# ---------------------------------------------------------------------------
def da_losses(img_score, ins_sigmoid, domain_label, ins_label=None):
    """float64 restatement.  img_score (B,2,H,W) logits; ins_sigmoid (R,1) or (R,);
    domain_label scalar 0/1 (image label map is that scalar broadcast,
    LabelResizeLayer.py:25-39); ins_label (R,) defaults to domain_label everywhere.

    Returns dict(img_loss, ins_loss, cst_loss, consistency_prob) following
    F.nll_loss(F.log_softmax(.,1), label) (mean), nn.BCELoss() (mean, log clamped
    at -100 as torch does) and MSELoss(size_average=False) (sum) against the mean
    softmax probability of channel 1 (source) / 0 (target)."""
    s = np.asarray(img_score, np.float64)
    p = np.asarray(ins_sigmoid, np.float64).reshape(-1)
    d = int(domain_label)
    m = s.max(1, keepdims=True)
    lse = m + np.log(np.exp(s - m).sum(1, keepdims=True))
    logp = s - lse
    img_loss = -logp[:, d].mean()
    y = np.full_like(p, float(d)) if ins_label is None else np.asarray(ins_label, np.float64).reshape(-1)
    logp_i = np.maximum(np.log(p), -100.0)
    log1mp_i = np.maximum(np.log1p(-p), -100.0)
    ins_loss = -(y * logp_i + (1 - y) * log1mp_i).mean()
    cons = np.exp(logp[:, 1 if d == 1 else 0]).mean()
    cst_loss = ((p - cons) ** 2).sum()
    return dict(img_loss=img_loss, ins_loss=ins_loss, cst_loss=cst_loss, consistency_prob=cons)


def da_losses_grad(img_score, ins_sigmoid, domain_label, ins_label=None, w_img=1.0, w_ins=1.0,
                   w_cst=1.0):
    """Analytic float64 gradients of w_img*img + w_ins*ins + w_cst*cst w.r.t. the
    image logits and instance probabilities (consistency target is detached)."""
    s = np.asarray(img_score, np.float64)
    p = np.asarray(ins_sigmoid, np.float64).reshape(-1)
    d = int(domain_label)
    B, _, H, W = s.shape
    m = s.max(1, keepdims=True)
    e = np.exp(s - m)
    sm = e / e.sum(1, keepdims=True)
    onehot = np.zeros_like(s)
    onehot[:, d] = 1.0
    g_img = w_img * (sm - onehot) / (B * H * W)
    y = np.full_like(p, float(d)) if ins_label is None else np.asarray(ins_label, np.float64).reshape(-1)
    R = p.shape[0]
    # torch's BCE backward: (p - y) / max((1-p)*p, eps=1e-12) / R
    g_ins = w_ins * (p - y) / np.maximum((1 - p) * p, 1e-12) / R
    cons = sm[:, 1 if d == 1 else 0].mean()
    g_ins = g_ins + w_cst * 2.0 * (p - cons)
    return g_img, g_ins
