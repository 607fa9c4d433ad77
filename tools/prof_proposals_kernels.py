"""Per-kernel device time of the proposal layer (the library's own event pairs around each
launch): TRAIN 12000 -> 2000 and TEST 6000 -> 300, 2 images."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import numpy as np, torch
from tools.synth import synth_rpn
from tlod_b200 import functional as F, _lib
from model.rpn.generate_anchors import generate_anchors
dev = torch.device("cuda:0")
prob, deltas = synth_rpn(2, 12, 37, 75, 3)
im_info = torch.tensor([[600.0, 1200.0, 0.5859375]] * 2)
anchors = torch.from_numpy(generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
args = [t.to(dev) for t in (prob, deltas, im_info, anchors)]
for name, pre, post in (("TRAIN", 12000, 2000), ("TEST", 6000, 300)):
    for _ in range(3):
        F.proposals(*args, 16, pre, post, 0.7)
    torch.cuda.synchronize()
    _lib.profile_reset(); _lib.profile(True)
    for _ in range(10):
        F.proposals(*args, 16, pre, post, 0.7)
    torch.cuda.synchronize()
    prof = _lib.profile_read(); _lib.profile(False)
    print(name, {k: "%.1f us x%d" % (ms / cnt * 1e3, cnt // 10) for k, (ms, cnt) in prof.items()},
          "sum %.1f us" % (sum(ms for ms, _ in prof.values()) / 10 * 1e3))
