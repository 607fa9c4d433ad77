"""RoIPool forward / backward timing (VGG16 conv5, 2 images, 512 RoIs) and a larger case, with the
plane-resident forward / tabulated backward and with the generic kernels (the A/B knobs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import torch
from tools.synth import synth_rois
from tlod_b200 import functional as F
dev = torch.device("cuda:0")
KNOBS = ("TLOD_DISABLE_POOL_PLANES", "TLOD_DISABLE_POOL_TAB")


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for (B, C, H, W, R) in ((2, 512, 37, 75, 512), (8, 1024, 38, 75, 2048)):
    g = torch.Generator().manual_seed(5)
    feat = torch.relu(torch.randn(B, C, H, W, generator=g)).to(dev)
    rois = synth_rois(R, B, 41).to(dev)
    top = torch.randn(R, C, 7, 7, device=dev)
    alg = feat.numel() * 4 + R * 20 + R * C * 49 * 8
    outs = []
    for generic in (False, True):
        for k in KNOBS:
            os.environ.pop(k, None)
            if generic:
                os.environ[k] = "1"
        out, arg = F.roi_pool_forward(feat, rois, 7, 7, 1 / 16)
        grad = F.roi_pool_backward(top, arg, rois, feat.shape, 1 / 16)
        outs.append((out, arg, grad))
        fwd = timed(lambda: F.roi_pool_forward(feat, rois, 7, 7, 1 / 16))
        bwd = timed(lambda: F.roi_pool_backward(top, arg, rois, feat.shape, 1 / 16))
        print("%s B=%d C=%d R=%d: fwd %.1f us (%.2f of HBM)  bwd %.1f us (%.2f)" % (
            "generic kernels" if generic else "planes / table ", B, C, R, fwd, alg / fwd / 1e3 / 6546.2, bwd,
            alg / bwd / 1e3 / 6546.2))
    same = torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    gerr = ((outs[0][2] - outs[1][2]).abs().max() / outs[1][2].abs().max()).item()
    print("   outputs and argmax identical: %s; backward max rel diff %.2e" % (same, gerr))
for k in KNOBS:
    os.environ.pop(k, None)
