"""The pieces of lib/model/utils/net_utils.py on the RoI path.

``_affine_grid_gen`` (net_utils.py:142-164): theta from the RoIs (image px / 16) and
``F.affine_grid``.  The reference ran on torch 0.4, whose affine_grid is what torch >= 1.3
calls ``align_corners=True``; that is passed explicitly here (the default changed)."""
import torch
import torch.nn.functional as F


def _affine_grid_gen(rois, input_size, grid_size):
    rois = rois.detach()
    x1 = rois[:, 1::4] / 16.0
    y1 = rois[:, 2::4] / 16.0
    x2 = rois[:, 3::4] / 16.0
    y2 = rois[:, 4::4] / 16.0
    height = input_size[0]
    width = input_size[1]
    zero = rois.new_zeros((rois.size(0), 1))
    theta = torch.cat([
        (x2 - x1) / (width - 1),
        zero,
        (x1 + x2 - width + 1) / (width - 1),
        zero,
        (y2 - y1) / (height - 1),
        (y1 + y2 - height + 1) / (height - 1)], 1).view(-1, 2, 3)
    return F.affine_grid(theta, torch.Size((rois.size(0), 1, grid_size, grid_size)), align_corners=True)


def roi_crop_pool(roi_crop, base_feat, rois, grid_size):
    """The cfg.POOLING_MODE == 'crop' branch of lib/model/faster_rcnn/faster_rcnn.py:77-83:
    affine grid -> (y, x) order -> RoICrop -> (optional) 2x2 max pool is left to the caller."""
    grid_xy = _affine_grid_gen(rois.view(-1, 5), base_feat.size()[2:], grid_size)
    grid_yx = torch.stack([grid_xy[:, :, :, 1], grid_xy[:, :, :, 0]], 3).contiguous()
    return roi_crop(base_feat, grid_yx.detach())


def roi_crop_max_pool(base_feat, rois, grid_size):
    """The whole 'crop' branch (faster_rcnn.py:73-80) -- affine grid, RoICrop, F.max_pool2d(., 2, 2) --
    as ONE kernel each way when cfg.CROP_RESIZE_WITH_MAX_POOL doubles the grid to 14: the
    (R, C, 14, 14) sample tensor is never written.  Other sizes compose RoICrop and torch's pooling."""
    from tlod_b200 import functional as TF
    from tlod_b200.autograd import RoICropFunction, RoICropPoolFunction
    grid_xy = _affine_grid_gen(rois.view(-1, 5), base_feat.size()[2:], grid_size)
    if TF.roi_crop_pool_supported(base_feat, grid_size, grid_size):
        # rotation-free theta: the grid is the outer product of its first column (y) and first row (x)
        grid_y = grid_xy[:, :, 0, 1].contiguous()
        grid_x = grid_xy[:, 0, :, 0].contiguous()
        return RoICropPoolFunction.apply(base_feat, grid_y.detach(), grid_x.detach())
    grid_yx = torch.stack([grid_xy[:, :, :, 1], grid_xy[:, :, :, 0]], 3).contiguous()
    return F.max_pool2d(RoICropFunction.apply(base_feat, grid_yx.detach()), 2, 2)


def _smooth_l1_loss(bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma=1.0, dim=[1]):
    """lib/model/utils/net_utils.py:72-86 (imported by the reference's rpn.py:10 and faster_rcnn.py):
    the eager torch expression.  The fused kernel pair behind tlod_b200.rpn_losses computes the same
    loss together with the RPN cross-entropy in one launch each way."""
    s2 = sigma ** 2
    diff = bbox_inside_weights * (bbox_pred - bbox_targets)
    a = diff.abs()
    quad = (a < 1.0 / s2).detach().float()
    loss = bbox_outside_weights * (diff.pow(2) * (s2 / 2.0) * quad + (a - 0.5 / s2) * (1.0 - quad))
    for i in sorted(dim, reverse=True):
        loss = loss.sum(i)
    return loss.mean()
