"""Box arithmetic of the RPN (lib/model/rpn/bbox_transform.py), same names and argument
meaning; each call is one kernel of libtlod_b200.so instead of ~25 eager launches."""
import torch

from tlod_b200 import functional as F


def bbox_transform_batch(ex_rois, gt_rois):
    """:36-75.  ex (N,4) | (B,N,4); gt (B,N,4) -> (B,N,4) targets."""
    if ex_rois.dim() not in (2, 3):
        raise ValueError('ex_roi input dimension is not correct.')
    return F.bbox_transform_batch(ex_rois, gt_rois[:, :, :4])


def bbox_transform_inv(boxes, deltas, batch_size):
    """:77-103.  boxes (B,N,4), deltas (B,N,4) -> pred boxes (B,N,4) (not clipped)."""
    if deltas.size(2) != 4:
        # class-specific deltas (B, N, 4*k): decode each group of four against the same box
        B, N, K4 = deltas.shape
        k = K4 // 4
        bx = boxes.unsqueeze(2).expand(B, N, k, 4).reshape(B, N * k, 4)
        return F.bbox_transform_inv(bx, deltas.reshape(B, N * k, 4)).view(B, N, K4)
    return F.bbox_transform_inv(boxes, deltas)


def clip_boxes(boxes, im_shape, batch_size):
    """:125-133.  In place on a contiguous fp32 (B, N, 4k) tensor; returns it."""
    if boxes.is_contiguous() and boxes.dtype == torch.float32:
        return F.clip_boxes_(boxes, im_shape)
    clipped = F.clip_boxes_(boxes.float().contiguous(), im_shape)
    boxes.copy_(clipped)
    return boxes


def bbox_overlaps_batch(anchors, gt_boxes):
    """:168-257.  anchors (N,4) | (B,N,4) | (B,N,5); gt (B,K,5) -> (B,N,K)."""
    return F.bbox_overlaps_batch(anchors, gt_boxes)


def bbox_overlaps(anchors, gt_boxes):
    """:136-166.  (N,4) x (K,4) -> (N,K); same formula without the zero-area masks, so it is
    only routed through the batched kernel when no box is degenerate; kept for API parity."""
    ov = F.bbox_overlaps_batch(anchors, gt_boxes.unsqueeze(0))
    return ov[0]
