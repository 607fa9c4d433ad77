"""Small driver for ncu: RoIAlign 8x8 forward + backward at BASELINE cfg3
(8x1024x38x75 features, 2048 RoIs) and the cfg1-scale proposal layer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle.synth import synth_rois, synth_rpn  # noqa: E402
from tlod_b200 import functional as F  # noqa: E402
from model.rpn.generate_anchors import generate_anchors  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "roi"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if what == "roi":
    B, C, H, W, R = 8, 1024, 38, 75, 2048
    x = torch.relu(torch.randn(B, C, H, W, device=dev))
    rois = synth_rois(R, B, 41)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
    top = torch.randn(R, C, 8, 8, device=dev)
    plan = F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16)
    alg = B * C * H * W * 4 + R * 20 + R * C * 64 * 4

    def timed(fn, n):
        for _ in range(3):
            r = fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n, r
    if iters <= 3:  # ncu mode: few launches
        for _ in range(iters):
            plan = F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16)
            y = F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan)
            g = F.roi_align_backward(top, rois, x.shape, 1 / 16, plan=plan)
        torch.cuda.synchronize()
        print("ok", float(y.sum()), float(g.sum()))
    else:
        tp, _ = timed(lambda: F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16), iters)
        tf, y = timed(lambda: F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan), iters)
        tb, g = timed(lambda: F.roi_align_backward(top, rois, x.shape, 1 / 16, plan=plan), iters)
        print("ok", float(y.sum()), float(g.sum()))
        print("cfg3 plan %.1f us  fwd %.1f us (%.0f GB/s, %.3f of 6546)  bwd %.1f us (%.0f GB/s, %.3f)" % (
            tp * 1e3, tf * 1e3, alg / tf / 1e6, alg / tf / 1e6 / 6546.2, tb * 1e3, alg / tb / 1e6, alg / tb / 1e6 / 6546.2))
else:
    B, A, H, W = 2, 12, 37, 75
    prob, deltas = synth_rpn(B, A, H, W, 3)
    im_info = torch.tensor([[600.0, 1200.0, 0.5859375]] * B)
    anchors = torch.from_numpy(generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
    args = [t.to(dev) for t in (prob, deltas, im_info, anchors)]
    for _ in range(iters):
        rois = F.proposals(*args, 16, 12000, 2000, 0.7)
    torch.cuda.synchronize()
    print("ok", float(rois.sum()))
