"""Autograd bindings (modern static Functions; the reference's are torch-0.4 instance
Functions: lib/model/roi_align/functions/roi_align.py:8-51,
lib/model/roi_pooling/functions/roi_pool.py:6-38, lib/DAF/DA.py:19-33)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import functional as F


class RoIAlignFunction(Function):
    """features (B,C,H,W), rois (R,5) -> (R,C,aligned_h,aligned_w); grad only w.r.t. features."""

    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale):
        ctx.feature_size = tuple(features.shape)
        ctx.scale = float(spatial_scale)
        # the plan (sampling tables, image-sorted RoI list) is built once and reused by backward
        plan = F.roi_align_plan(rois, ctx.feature_size, int(aligned_height), int(aligned_width), ctx.scale)
        ctx.has_plan = plan is not None
        ctx.save_for_backward(rois, *([plan] if ctx.has_plan else []))
        return F.roi_align_forward(features, rois, int(aligned_height), int(aligned_width), ctx.scale,
                                   plan=plan, use_plan=ctx.has_plan)

    @staticmethod
    def backward(ctx, grad_output):
        rois = ctx.saved_tensors[0]
        plan = ctx.saved_tensors[1] if ctx.has_plan else None
        assert grad_output.is_cuda  # functions/roi_align.py:38
        grad_input = F.roi_align_backward(grad_output, rois, ctx.feature_size, ctx.scale, plan=plan,
                                          use_plan=ctx.has_plan)
        return grad_input, None, None, None, None


class RoIAlignAvgFunction(Function):
    """RoIAlign at (h+1, w+1) followed by the 2x2 stride-1 average (modules/roi_align.py:26-29) as
    one autograd node.  Forward: one kernel for 7 x 7 (the samples are averaged in registers, the
    (R, C, 8, 8) tensor is never written).  Backward: the average's adjoint, then the row-resident
    RoIAlign backward."""

    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale):
        ctx.feature_size = tuple(features.shape)
        ctx.scale = float(spatial_scale)
        ph, pw = int(aligned_height), int(aligned_width)
        plan = F.roi_align_plan(rois, ctx.feature_size, ph + 1, pw + 1, ctx.scale)
        ctx.has_plan = plan is not None
        ctx.save_for_backward(rois, *([plan] if ctx.has_plan else []))
        return F.roi_align_avg_forward(features, rois, ph, pw, ctx.scale, plan=plan)

    @staticmethod
    def backward(ctx, grad_output):
        rois = ctx.saved_tensors[0]
        plan = ctx.saved_tensors[1] if ctx.has_plan else None
        assert grad_output.is_cuda
        grad_input = F.roi_align_avg_backward(grad_output, rois, ctx.feature_size, ctx.scale, plan=plan)
        return grad_input, None, None, None, None


class RoIPoolFunction(Function):
    @staticmethod
    def forward(ctx, features, rois, pooled_height, pooled_width, spatial_scale):
        out, argmax = F.roi_pool_forward(features, rois, int(pooled_height), int(pooled_width),
                                         float(spatial_scale))
        ctx.save_for_backward(rois, argmax)
        ctx.feature_size = tuple(features.shape)
        ctx.scale = float(spatial_scale)
        ctx.mark_non_differentiable(argmax)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        rois, argmax = ctx.saved_tensors
        assert grad_output.is_cuda
        grad_input = F.roi_pool_backward(grad_output, argmax, rois, ctx.feature_size, ctx.scale)
        return grad_input, None, None, None, None


class RoICropFunction(Function):
    """(features (ib,C,H,W), grid_yx (ob,GH,GW,2)) -> (ob,C,GH,GW); lib/model/roi_crop/functions/
    roi_crop.py:7-21.  Like the reference kernel, the gradient of the grid is all zeros."""

    @staticmethod
    def forward(ctx, input1, input2):
        ctx.save_for_backward(input2)
        ctx.feature_size = tuple(input1.shape)
        return F.roi_crop_forward(input1, input2)

    @staticmethod
    def backward(ctx, grad_output):
        (grid,) = ctx.saved_tensors
        return F.roi_crop_backward(grad_output, grid, ctx.feature_size), torch.zeros_like(grid)


class RoICropPoolFunction(Function):
    """RoICrop on the rotation-free affine grid of _affine_grid_gen + F.max_pool2d(., 2, 2) as one node
    (lib/model/faster_rcnn/faster_rcnn.py:73-80).  grid_y / grid_x (R, 14): row / column coordinates."""

    @staticmethod
    def forward(ctx, features, grid_y, grid_x):
        out, arg = F.roi_crop_pool_forward(features, grid_y, grid_x)
        ctx.save_for_backward(arg, grid_y, grid_x)
        ctx.feature_size = tuple(features.shape)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        arg, grid_y, grid_x = ctx.saved_tensors
        return F.roi_crop_pool_backward(grad_output, arg, grid_y, grid_x, ctx.feature_size), None, None


class GradReverse(Function):
    """Identity forward; backward -alpha * g (optionally * per-row weight, lib/MAF/DA.py:34-53)."""

    @staticmethod
    def forward(ctx, x, alpha=0.1, row_weight=None):
        ctx.alpha = float(alpha)
        ctx.has_w = row_weight is not None
        if ctx.has_w:
            ctx.save_for_backward(row_weight)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        w = ctx.saved_tensors[0] if ctx.has_w else None
        return F.grl_backward(grad_output, ctx.alpha, w), None, None


def grad_reverse(x, alpha=0.1, row_weight=None):
    return GradReverse.apply(x, alpha, row_weight)


class RPNLossFunction(Function):
    """(rpn_cls_score (B,2A,H,W), rpn_bbox_pred (B,4A,H,W), labels, targets, inside_w, outside_w) ->
    (rpn_loss_cls, rpn_loss_box): lib/model/rpn/rpn.py:90-108 as one launch each way."""

    @staticmethod
    def forward(ctx, cls_score, bbox_pred, labels, bbox_targets, inside_w, outside_w, sigma=3.0):
        losses = F.rpn_loss_forward(cls_score, labels, bbox_pred, bbox_targets, inside_w, outside_w, sigma)
        ctx.save_for_backward(cls_score, bbox_pred, labels, bbox_targets, inside_w, outside_w, losses)
        ctx.sigma = float(sigma)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_cls, g_box):
        cls_score, bbox_pred, labels, targets, inside_w, outside_w, losses = ctx.saved_tensors
        up = torch.stack([g_cls.reshape(()), g_box.reshape(())]).float()
        gs, gp = F.rpn_loss_backward(cls_score, labels, bbox_pred, targets, inside_w, outside_w, losses, up,
                                     ctx.sigma)
        return gs.view_as(cls_score), gp.view_as(bbox_pred), None, None, None, None, None


def rpn_losses(cls_score, bbox_pred, labels, bbox_targets, inside_w, outside_w, sigma=3.0):
    return RPNLossFunction.apply(cls_score, bbox_pred, labels, bbox_targets, inside_w, outside_w, sigma)


class SpaceToDepthFunction(Function):
    """The rearrangement of DRM.forward (lib/MAF/drm.py:21-42) as one launch each way."""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.input_size = tuple(x.shape)
        ctx.scale = int(scale)
        return F.space_to_depth_forward(x, ctx.scale)

    @staticmethod
    def backward(ctx, grad_output):
        return F.space_to_depth_backward(grad_output, ctx.input_size, ctx.scale), None


def space_to_depth(x, scale):
    return SpaceToDepthFunction.apply(x, scale)


class DALossFunction(Function):
    """(img_score (B,2,H,W), ins_prob (R,1)) -> (img_loss, ins_loss, cst_loss) for one domain,
    lib/DAF/faster_rcnn.py:181-220.  One launch forward, one backward, no label tensors."""

    @staticmethod
    def forward(ctx, img_score, ins_prob, domain_label, ins_label=None):
        losses = F.da_loss_forward(img_score, ins_prob, int(domain_label), ins_label)
        ctx.save_for_backward(img_score, ins_prob, losses, *([] if ins_label is None else [ins_label]))
        ctx.domain = int(domain_label)
        ctx.has_label = ins_label is not None
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_img, g_ins, g_cst):
        saved = ctx.saved_tensors
        img_score, ins_prob, losses = saved[:3]
        lab = saved[3] if ctx.has_label else None
        # the three upstream gradients stay on the device (no .item() sync)
        up = torch.stack([g_img.reshape(()), g_ins.reshape(()), g_cst.reshape(())]).float()
        gi, gp = F.da_loss_backward(img_score, ins_prob, ctx.domain, losses, 1.0, 1.0, 1.0, lab, upstream=up)
        return gi, gp.view_as(ins_prob), None, None


def da_losses(img_score, ins_prob, domain_label, ins_label=None):
    return DALossFunction.apply(img_score, ins_prob, domain_label, ins_label)


class ImageDALossFunction(Function):
    """Image-level DA losses of 1..4 feature levels (MAF's conv3 / conv4 / conv5 heads,
    lib/MAF/faster_rcnn.py:188-205; ATF's ignore_index = -1, lib/ATF/faster_rcnn.py:303-321): one launch
    forward, one backward.  Returns the (levels,) vector of mean NLLs; the reference adds them up."""

    @staticmethod
    def forward(ctx, domain_label, ignore_index, n_levels, *maps):
        scores, labels = list(maps[:n_levels]), list(maps[n_levels:])
        labels = labels if labels else None
        out = F.da_image_loss_forward(scores, int(domain_label), labels, int(ignore_index))
        ctx.save_for_backward(out, *scores, *([] if labels is None else [t for t in labels if t is not None]))
        ctx.n = n_levels
        ctx.has = None if labels is None else [t is not None for t in labels]
        ctx.domain, ctx.ignore = int(domain_label), int(ignore_index)
        return out[:, 0].clone()

    @staticmethod
    def backward(ctx, g):
        out, rest = ctx.saved_tensors[0], list(ctx.saved_tensors[1:])
        scores, labs = rest[:ctx.n], rest[ctx.n:]
        labels = None
        if ctx.has is not None:
            it = iter(labs)
            labels = [next(it) if h else None for h in ctx.has]
        grads = F.da_image_loss_backward(scores, ctx.domain, out, labels, ctx.ignore, upstream=g.contiguous().float())
        return (None, None, None, *grads, *([None] * (0 if ctx.has is None else len(ctx.has))))


def image_da_losses(scores, domain_label, labels=None, ignore_index=-100):
    """scores: list of (B, 2, H, W) logits -> (levels,) mean NLL per level."""
    scores = list(scores)
    extra = [] if labels is None else list(labels)
    return ImageDALossFunction.apply(int(domain_label), int(ignore_index), len(scores), *scores, *extra)
