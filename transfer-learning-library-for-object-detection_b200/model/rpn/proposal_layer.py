"""_ProposalLayer (lib/model/rpn/proposal_layer.py:26-176).

Same constructor and forward((scores, bbox_deltas, im_info, cfg_key)) -> (B, post, 5).
The reference runs ~30 eager launches, a batch-wide torch.sort and a Python loop with one
host-synchronising NMS per image; here it is three launches for the whole batch
(tlod_proposals), nothing returns to the host."""
import numpy as np
import torch
import torch.nn as nn

from model.utils.config import cfg
from tlod_b200 import functional as F

from .generate_anchors import generate_anchors


class _ProposalLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super(_ProposalLayer, self).__init__()
        self._feat_stride = feat_stride
        self._anchors = torch.from_numpy(
            generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float()
        self._num_anchors = self._anchors.size(0)

    def forward(self, input):
        scores_all, bbox_deltas, im_info, cfg_key = input[0], input[1], input[2], input[3]
        # cfg is read at call time (ATF mutates TEST.RPN_POST_NMS_TOP_N between calls)
        pre_nms_topN = cfg[cfg_key].RPN_PRE_NMS_TOP_N
        post_nms_topN = cfg[cfg_key].RPN_POST_NMS_TOP_N
        nms_thresh = cfg[cfg_key].RPN_NMS_THRESH
        if self._anchors.device != scores_all.device:
            self._anchors = self._anchors.to(scores_all.device)
        # fg scores are channels [A, 2A) of scores_all; the kernel indexes them in place
        return F.proposals(scores_all, bbox_deltas, im_info, self._anchors, self._feat_stride,
                           pre_nms_topN, post_nms_topN, nms_thresh)

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""
        pass

    def reshape(self, bottom, top):
        pass
