"""DRM -- the scale-reduce module of MAF / PT-MAF (lib/MAF/drm.py:10-42).

Same constructor (in_dim, inner_channel, scale) and forward(x): 1x1 conv -> ReLU -> the tile
rearrangement.  The reference performs the rearrangement with (H/scale)*(W/scale) Python-level
chunk / reshape / cat calls (drm.py:30-40); it is a space-to-depth, done here in one launch
(tlod_space_to_depth_forward / _backward)."""
import torch.nn as nn

from tlod_b200.autograd import space_to_depth


class DRM(nn.Module):
    def __init__(self, in_dim, inner_channel, scale):
        super(DRM, self).__init__()
        self.in_dim = in_dim
        self.inner_channel = inner_channel
        self.scale = scale
        self.conv_low_dim = nn.Conv2d(self.in_dim, self.inner_channel, kernel_size=1, stride=1, bias=False)
        self.relu = nn.ReLU(inplace=False)

    def forward(self, x):
        low_dim = self.relu(self.conv_low_dim(x))
        return space_to_depth(low_dim, self.scale)
