"""RoICrop 14x14 forward / backward timing at BASELINE cfg3 (8 x 1024 x 38 x 75, 2048 RoIs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import torch
from tools.synth import synth_rois
from tlod_b200 import functional as F
from model.utils.net_utils import _affine_grid_gen
dev = torch.device("cuda:0")
B, C, H, W, R, G = 8, 1024, 38, 75, 2048, 14
g = torch.Generator().manual_seed(5)
feat = torch.relu(torch.randn(B, C, H, W, generator=g)).to(dev)
rois = synth_rois(R, B, 41)
rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
grid_xy = _affine_grid_gen(rois, (H, W), G)
grid_yx = torch.stack([grid_xy[..., 1], grid_xy[..., 0]], 3).contiguous()
top = torch.randn(R, C, G, G, device=dev)
alg = feat.numel() * 4 + grid_yx.numel() * 4 + R * C * G * G * 4
res = []
for fn in (lambda: F.roi_crop_forward(feat, grid_yx), lambda: F.roi_crop_backward(top, grid_yx, feat.shape)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 10 * 1e3)
print("cfg3 crop path: fwd %.0f us (%.2f of HBM)  bwd %.0f us (%.2f)" % (
    res[0], alg / res[0] / 1e3 / 6546.2, res[1], alg / res[1] / 1e3 / 6546.2))
