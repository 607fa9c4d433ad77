"""RoIAlign forward/backward timing at cfg2 scale (2 images, 512 ch, 37x75, 512/600 RoIs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import torch  # noqa: E402

from tools.synth import synth_rois  # noqa: E402
from tlod_b200 import functional as F  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev)
for (B, C, H, W, R) in ((2, 512, 37, 75, 512), (2, 512, 37, 75, 600), (1, 512, 37, 75, 128)):
    x = torch.relu(torch.randn(B, C, H, W, device=dev))
    rois = synth_rois(R, B, 41)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
    top = torch.randn(R, C, 8, 8, device=dev)
    plan = F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16)
    alg = B * C * H * W * 4 + R * 20 + R * C * 64 * 4
    res = []
    for fn in (lambda: F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan),
               lambda: F.roi_align_backward(top, rois, x.shape, 1 / 16, plan=plan)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        res.append(ts[len(ts) // 2])
    print("B=%d R=%d: fwd %.1f us (%.2f of roofline)  bwd %.1f us (%.2f)  [L2 flushed, incl. launch]" % (
        B, R, res[0] * 1e3, alg / res[0] / 1e6 / 6546.2, res[1] * 1e3, alg / res[1] / 1e6 / 6546.2))
