"""ImageLabelResizeLayer / InstanceLabelResizeLayer (lib/DAF/LabelResizeLayer.py:17-58).

Same forward(x, need_backprop) signatures and outputs, built on the device: the reference copies
the score map and the domain labels to the host, calls cv2.resize on a 1-element array and copies
the label tensor back (two .cpu() synchronisations per call).  NEAREST-resizing a single value is a
broadcast, so the image label map is ``need_backprop[i]`` everywhere."""
import torch
import torch.nn as nn

from tlod_b200 import functional as F


class ImageLabelResizeLayer(nn.Module):
    def forward(self, x, need_backprop):
        lbs = need_backprop.detach().to(x.device).float().view(-1)
        return lbs.view(-1, 1, 1).expand(lbs.numel(), x.size(2), x.size(3)).long().contiguous()


class InstanceLabelResizeLayer(nn.Module):
    def __init__(self, fill=1.0):
        super(InstanceLabelResizeLayer, self).__init__()
        self.minibatch = 256
        self.fill = float(fill)  # DAF: np.ones (:53); MAF / ATF build the array with np.zeros

    def forward(self, x, need_backprop):
        lbs = need_backprop.detach().to(x.device).float().view(-1)
        return F.instance_labels(lbs, x.size(0), self.minibatch, self.fill)
