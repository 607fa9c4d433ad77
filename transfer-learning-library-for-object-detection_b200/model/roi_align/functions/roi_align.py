"""``RoIAlignFunction(aligned_h, aligned_w, scale)(features, rois)`` -- the reference's
instance-style call (lib/model/roi_align/functions/roi_align.py:8-51) on top of the
static autograd Function torch 2.x requires."""
from tlod_b200.autograd import RoIAlignFunction as _Fn


class RoIAlignFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _Fn.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)

    forward = __call__
