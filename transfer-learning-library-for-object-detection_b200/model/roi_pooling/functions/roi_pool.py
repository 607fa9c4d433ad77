"""``RoIPoolFunction(ph, pw, scale)(features, rois)``
(lib/model/roi_pooling/functions/roi_pool.py:6-38)."""
from tlod_b200.autograd import RoIPoolFunction as _Fn


class RoIPoolFunction(object):
    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _Fn.apply(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale)

    forward = __call__
