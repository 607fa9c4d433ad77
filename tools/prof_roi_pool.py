"""RoIPool forward / backward timing (VGG16 conv5, 2 images, 512 RoIs) and a larger case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import torch
from tools.synth import synth_rois
from tlod_b200 import functional as F
dev = torch.device("cuda:0")
for (B, C, H, W, R) in ((2, 512, 37, 75, 512), (8, 1024, 38, 75, 2048)):
    g = torch.Generator().manual_seed(5)
    feat = torch.relu(torch.randn(B, C, H, W, generator=g)).to(dev)
    rois = synth_rois(R, B, 41).to(dev)
    top = torch.randn(R, C, 7, 7, device=dev)
    out, arg = F.roi_pool_forward(feat, rois, 7, 7, 1 / 16)
    alg = feat.numel() * 4 + R * 20 + R * C * 49 * 8
    res = []
    for fn in (lambda: F.roi_pool_forward(feat, rois, 7, 7, 1 / 16),
               lambda: F.roi_pool_backward(top, arg, rois, feat.shape, 1 / 16)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 20 * 1e3)
    print("B=%d C=%d R=%d: fwd %.1f us (%.2f of HBM)  bwd %.1f us (%.2f)" % (
        B, C, R, res[0], alg / res[0] / 1e3 / 6546.2, res[1], alg / res[1] / 1e3 / 6546.2))
