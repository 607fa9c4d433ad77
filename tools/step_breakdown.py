"""Where one bench step's wall time goes: each captured graph alone, the host pieces alone, the step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import numpy as np
import torch
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step = bench.TlodStep(dev, seed=3, wl=bench.Workload("cfg2"), use_graph=True)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6


d = step.d
for g in ("src1", "src2", "src3", "tgt"):
    print("graph %-5s alone                  %.0f us" % (g, timed(lambda: step.graphs[g].replay())))
at_in = (d["src_prob"], d["src_gt"], d["src_im_info"], step.num_boxes)
print("anchor-target layer alone          %.0f us" % timed(lambda: step.anchor_target(at_in)))
pend = [None]


def begin_only():
    pend[0] = step.anchor_target.begin(at_in)
    pend[0]["copied"].synchronize()


print("  anchor begin (kernels + D2H)     %.0f us" % timed(begin_only))
from model.rpn.anchor_target_layer import subsample_labels
lab = pend[0]["labels_host"].numpy().copy()
t0 = time.perf_counter()
for _ in range(50):
    subsample_labels(lab.copy(), 128, 256)
print("  anchor host subsample            %.0f us" % ((time.perf_counter() - t0) / 50 * 1e6))
st = step.static["src1"][1]
st["copied"] = None
t0 = time.perf_counter()
for _ in range(50):
    step.proposal_target.sample(st)
print("  proposal-target host sampling    %.0f us" % ((time.perf_counter() - t0) / 50 * 1e6))
print("full step                          %.0f us" % timed(lambda: step.step()))
print("full step, eager                   %.0f us" % timed(lambda: step.step_eager()))
