"""RoICrop timing at BASELINE cfg3 (ii): 8x1024x38x75, 2048 RoIs, 14x14 grid -> max-pool 7x7."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import torch
from tools.synth import synth_rois
from tlod_b200 import functional as F
from model.utils.net_utils import _affine_grid_gen
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, C, H, W, R, G = 8, 1024, 38, 75, 2048, 14
g = torch.Generator().manual_seed(5)
feat = torch.relu(torch.randn(B, C, H, W, generator=g)).to(dev)
rois = synth_rois(R, B, 41)
rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
gxy = _affine_grid_gen(rois, (H, W), G)
gyx = torch.stack([gxy[..., 1], gxy[..., 0]], 3).contiguous()
gy, gx = gxy[:, :, 0, 1].contiguous(), gxy[:, 0, :, 0].contiguous()
top14 = torch.randn(R, C, G, G, device=dev)
top7 = torch.randn(R, C, 7, 7, device=dev)
out7, arg = F.roi_crop_pool_forward(feat, gy, gx)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


alg14 = feat.numel() * 4 + gyx.numel() * 4 + R * C * G * G * 4
alg7 = feat.numel() * 4 + R * 28 * 4 + R * C * 49 * 5
for name, fn, alg in (
        ("RoICrop 14x14 fwd (unfused)", lambda: F.roi_crop_forward(feat, gyx), alg14),
        ("RoICrop 14x14 bwd (unfused)", lambda: F.roi_crop_backward(top14, gyx, feat.shape), alg14),
        ("crop+maxpool fused fwd", lambda: F.roi_crop_pool_forward(feat, gy, gx), alg7),
        ("crop+maxpool fused bwd", lambda: F.roi_crop_pool_backward(top7, arg, gy, gx, feat.shape), alg7),
        ("torch max_pool2d(2,2) fwd on 14x14", lambda: torch.nn.functional.max_pool2d(top14, 2, 2), 0)):
    us = timed(fn)
    print("cfg3 %-36s %9.1f us  %s" % (name, us, "%.3f of HBM (%.0f MB)" % (alg / us / 1e3 / 6546.2, alg / 1e6) if alg else ""))
