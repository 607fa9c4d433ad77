"""Where one bench step's wall time goes: device part alone, anchor-target layer alone, both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import torch
import bench
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step = bench.TlodStep(dev, seed=3, use_graph=True)
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
print("device part (2 domain graphs)      %.0f us" % timed(lambda: step.device_part(d)))
print("anchor-target layer alone          %.0f us" % timed(lambda: step.anchor_part(d, {})))
print("full step                          %.0f us" % timed(lambda: step.step()))
def one(dom):
    step.graphs[dom].replay()
print("src graph alone                    %.0f us" % timed(lambda: one("src")))
print("tgt graph alone                    %.0f us" % timed(lambda: one("tgt")))
