"""Driver for timing / ncu: RoIAlign(8, 8) and RoIAlignAvg(7, 7) forward + backward at BASELINE cfg3
(8x1024x38x75 features, 2048 RoIs) or cfg2 (2x512x37x75, 512 RoIs), and the proposal layer.

    python tools/prof_roi_align.py roi  <iters> [cfg3|cfg2]   # iters <= 3: ncu mode (few launches)
    python tools/prof_roi_align.py prop <iters>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools.synth import synth_rois, synth_rpn  # noqa: E402
from tlod_b200 import functional as F  # noqa: E402
from model.rpn.generate_anchors import generate_anchors  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "roi"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
PEAK = 6546.2


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


if what == "roi":
    B, C, H, W, R = (8, 1024, 38, 75, 2048) if cfg == "cfg3" else (2, 512, 37, 75, 512)
    x = torch.relu(torch.randn(B, C, H, W, device=dev))
    rois = synth_rois(R, B, 41)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
    top = torch.randn(R, C, 8, 8, device=dev)
    top7 = torch.randn(R, C, 7, 7, device=dev)
    plan = F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16)
    feat_b = B * C * H * W * 4 + R * 20
    alg8 = feat_b + R * C * 64 * 4
    alg7 = feat_b + R * C * 49 * 4
    if iters <= 3:  # ncu mode
        for _ in range(iters):
            plan = F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16)
            y = F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan)
            y7 = F.roi_align_avg_forward(x, rois, 7, 7, 1 / 16, plan=plan)
            g = F.roi_align_backward(top, rois, x.shape, 1 / 16, plan=plan)
            g7 = F.roi_align_avg_backward(top7, rois, x.shape, 1 / 16, plan=plan)
        torch.cuda.synchronize()
        print("ok", float(y.sum()), float(y7.sum()), float(g.sum()), float(g7.sum()))
    else:
        def line(name, ms, alg):
            print("%s %-34s %8.1f us  %6.0f GB/s  %.3f of %.0f" % (cfg, name, ms * 1e3, alg / ms / 1e6,
                                                                   alg / ms / 1e6 / PEAK, PEAK))
        tp, _ = timed(lambda: F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16), iters)
        print("%s plan %.1f us" % (cfg, tp * 1e3))
        tf, y = timed(lambda: F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan), iters)
        line("RoIAlign(8,8) fwd [fwd8]", tf, alg8)
        os.environ["TLOD_DISABLE_FWD8"] = "1"
        tf0, y0 = timed(lambda: F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan), iters)
        line("RoIAlign(8,8) fwd [round-1 planes]", tf0, alg8)
        ta0, a0 = timed(lambda: F.avgpool2x2_forward(F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan)), iters)
        line("RoIAlignAvg fwd [round-1 composed]", ta0, alg7)
        del os.environ["TLOD_DISABLE_FWD8"]
        ta, a1 = timed(lambda: F.roi_align_avg_forward(x, rois, 7, 7, 1 / 16, plan=plan), iters)
        line("RoIAlignAvg(7,7) fwd [fused]", ta, alg7)
        print("max |fwd8 - planes| = %.3g   max |fused - composed| = %.3g" % (
            float((y - y0).abs().max()), float((a1 - a0).abs().max())))
        tb, g = timed(lambda: F.roi_align_backward(top, rois, x.shape, 1 / 16, plan=plan), iters)
        line("RoIAlign(8,8) bwd [rows]", tb, alg8)
        tb7, g7 = timed(lambda: F.roi_align_avg_backward(top7, rois, x.shape, 1 / 16, plan=plan), iters)
        line("RoIAlignAvg(7,7) bwd [expand+rows]", tb7, alg7)
        te, _ = timed(lambda: F.avgpool2x2_backward(top7), iters)
        print("%s avgpool2x2 bwd (expand) alone %.1f us" % (cfg, te * 1e3))
        print("%s fwd+bwd+plan: RoIAlign(8,8) %.3f of HBM; RoIAlignAvg(7,7) %.3f of HBM (fused bytes)" % (
            cfg, 2 * alg8 / ((tp + tf + tb) * 1e6) / PEAK, 2 * alg7 / ((tp + ta + tb7) * 1e6) / PEAK))
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
