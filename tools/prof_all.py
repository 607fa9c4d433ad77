"""One pass over every kernel family at its BASELINE shape, for `ncu --set full` (each op runs
`reps` times; capture the last repetition with -k / -s / -c) and for the launch list."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools.synth import synth_rois, synth_rpn  # noqa: E402
from tlod_b200 import functional as F  # noqa: E402
from model.rpn.generate_anchors import generate_anchors  # noqa: E402
from model.utils.net_utils import _affine_grid_gen  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator().manual_seed(7)
# ---- RoIAlign cfg3 ----
B, C, H, W, R = 8, 1024, 38, 75, 2048
x = torch.relu(torch.randn(B, C, H, W, generator=g)).to(dev)
rois = synth_rois(R, B, 41)
rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
top8, top7 = torch.randn(R, C, 8, 8, device=dev), torch.randn(R, C, 7, 7, device=dev)
for _ in range(reps):
    plan = F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16)
    F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan)
    F.roi_align_avg_forward(x, rois, 7, 7, 1 / 16, plan=plan)
    F.roi_align_backward(top8, rois, x.shape, 1 / 16, plan=plan)
    F.avgpool2x2_backward(top7)
# ---- RoICrop + max-pool cfg3 ----
gxy = _affine_grid_gen(rois, (H, W), 14)
gy, gx = gxy[:, :, 0, 1].contiguous(), gxy[:, 0, :, 0].contiguous()
for _ in range(reps):
    o7, a7 = F.roi_crop_pool_forward(x, gy, gx)
    F.roi_crop_pool_backward(top7, a7, gy, gx, x.shape)
del top8, top7, o7, a7, x
# ---- RoIPool cfg2 ----
xp = torch.relu(torch.randn(2, 512, 37, 75, generator=g)).to(dev)
rp = synth_rois(512, 2, 41)
rp = rp[torch.argsort(rp[:, 0], stable=True)].contiguous().to(dev)
tp = torch.randn(512, 512, 7, 7, device=dev)
for _ in range(reps):
    op, ap = F.roi_pool_forward(xp, rp, 7, 7, 1 / 16)
    F.roi_pool_backward(tp, ap, rp, xp.shape, 1 / 16)
# ---- proposal layer, 2 images ----
A = 12
prob, deltas = synth_rpn(2, A, 37, 75, 3)
info = torch.tensor([[600.0, 1200.0, 0.5859375]] * 2)
anchors = torch.from_numpy(generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
args = [t.to(dev) for t in (prob, deltas, info, anchors)]
for _ in range(reps):
    F.proposals(*args, 16, 12000, 2000, 0.7)
    F.proposals(*args, 16, 6000, 300, 0.7)
# ---- multi-level DA losses at cfg4 sizes ----
maps = [torch.randn(8, 2, 150, 300, device=dev), torch.randn(8, 2, 75, 150, device=dev),
        torch.randn(8, 2, 37, 75, device=dev), torch.randn(2048, 2, 1, 1, device=dev)]
for _ in range(reps):
    lo = F.da_image_loss_forward(maps, 1)
    F.da_image_loss_backward(maps, 1, lo)
torch.cuda.synchronize()
print("ok")
