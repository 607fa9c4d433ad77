"""``ROIAlign(output_size, spatial_scale, sampling_ratio)`` / ``ROIPool(output_size,
spatial_scale)``: the constructor names BASELINE.json's north_star quotes (the
``model.roi_layers`` API of faster-rcnn.pytorch's pytorch-1.0 branch, only hinted at in
this reference by methods/IDF/IDF_test.py:22).  This reference has exactly one bilinear
sample per output cell and no ``sampling_ratio``; the alias reproduces ``RoIAlignAvg``
(kernel at (P+1)x(P+1), then 2x2 stride-1 average) and rejects any other sampling mode
instead of silently computing something the reference cannot pin."""
from torch.nn.modules.module import Module

from ..roi_align.modules.roi_align import RoIAlignAvg
from ..roi_pooling.modules.roi_pool import _RoIPooling


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


class ROIAlign(Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio=0):
        super(ROIAlign, self).__init__()
        if sampling_ratio not in (0, None):
            raise NotImplementedError(
                "sampling_ratio=%r: this reference's RoIAlign takes one sample per cell "
                "(roi_align_kernel.cu:40-49); only the default mode exists" % (sampling_ratio,))
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = 0
        self._impl = RoIAlignAvg(self.output_size[0], self.output_size[1], self.spatial_scale)

    def forward(self, input, rois):
        return self._impl(input, rois)


class ROIPool(Module):
    def __init__(self, output_size, spatial_scale):
        super(ROIPool, self).__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self._impl = _RoIPooling(self.output_size[0], self.output_size[1], self.spatial_scale)

    def forward(self, input, rois):
        return self._impl(input, rois)
