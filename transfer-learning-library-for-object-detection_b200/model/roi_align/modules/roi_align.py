"""RoIAlign / RoIAlignAvg / RoIAlignMax (lib/model/roi_align/modules/roi_align.py:6-42).

Same constructors: (aligned_height, aligned_width, spatial_scale); forward(features, rois)
with features (B,C,H,W) fp32 CUDA and rois (R,5) = [batch_idx, x1, y1, x2, y2]."""
from torch.nn.functional import avg_pool2d, max_pool2d
from torch.nn.modules.module import Module

from tlod_b200.autograd import RoIAlignAvgFunction

from ..functions.roi_align import RoIAlignFunction


class RoIAlign(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super(RoIAlign, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return RoIAlignFunction(self.aligned_height, self.aligned_width, self.spatial_scale)(features, rois)


class RoIAlignAvg(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super(RoIAlignAvg, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        # RoIAlign at (h+1, w+1), then avg_pool2d(kernel_size=2, stride=1): both halves in this
        # library's kernels, one autograd node (the (h+1, w+1) intermediate is not kept)
        return RoIAlignAvgFunction.apply(features, rois, self.aligned_height, self.aligned_width,
                                         self.spatial_scale)


class RoIAlignMax(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super(RoIAlignMax, self).__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1, self.spatial_scale)(features, rois)
        return max_pool2d(x, kernel_size=2, stride=1)
