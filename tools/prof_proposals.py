"""Driver for ncu / timing of the proposal layer: TRAIN (12000 -> 2000) then TEST (6000 -> 300), B images."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools.synth import synth_rpn  # noqa: E402
from tlod_b200 import functional as F  # noqa: E402
from model.rpn.generate_anchors import generate_anchors  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
A, H, W = 12, 37, 75
prob, deltas = synth_rpn(B, A, H, W, 3)
im_info = torch.tensor([[600.0, 1200.0, 0.5859375]] * B)
anchors = torch.from_numpy(generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
args = [t.to(dev) for t in (prob, deltas, im_info, anchors)]
for name, pre, post in (("TRAIN", 12000, 2000), ("TEST", 6000, 300)):
    for _ in range(3):
        rois = F.proposals(*args, 16, pre, post, 0.7)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        rois = F.proposals(*args, 16, pre, post, 0.7)
    b.record()
    torch.cuda.synchronize()
    print("%s B=%d: %.1f us per call, checksum %.1f" % (name, B, a.elapsed_time(b) / iters * 1e3, float(rois.sum())))
