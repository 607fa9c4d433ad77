"""Timeline of one graph-mode bench step: host timestamps (perf_counter) and device timestamps (CUDA
events) of the same step, relative to the step's start.  The body below is TlodStep.step() without the
e2e copies, with marks in between."""
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
for _ in range(5):
    step.step()
torch.cuda.synchronize()


def one():
    self = step
    host, devs = [], []

    def hmark(name):
        host.append((name, time.perf_counter()))

    def dmark(name, stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        devs.append((name, e))

    cur = torch.cuda.current_stream(self.dev)
    s_src, s_tgt, s_side = self.streams["src"], self.streams["tgt"], self.streams["side"]
    t0e = torch.cuda.Event(enable_timing=True)
    t0e.record(cur)
    hmark("start")
    for s in (s_src, s_tgt, s_side):
        s.wait_stream(cur)
    with torch.cuda.stream(s_side):
        self.graphs["at1"].replay()
        pending = self.static["at1"]
        pending["copied"] = self.events["at1"]
        pending["copied"].record(s_side)
        pending["stream"] = s_side
        dmark("at1 done", s_side)
    hmark("at1 queued")
    with torch.cuda.stream(s_src):
        self.graphs["src1"].replay()
        copied = self.events["src1"]
        copied.record(s_src)
        dmark("src1 done", s_src)
    hmark("src1 queued")
    with torch.cuda.stream(s_tgt):
        self.graphs["tgt"].replay()
        dmark("tgt done", s_tgt)
    hmark("tgt queued")
    pending = self.anchor_target.finish_host(pending)
    hmark("anchor subsample done")
    st = self.static["src1"][1]
    st["copied"] = copied
    self.proposal_target.sample(st, out=(self.keep_np, self.fg_np))
    hmark("proposal sampling done")
    with torch.cuda.stream(s_src):
        dmark("src2 start", s_src)
        self.graphs["src2"].replay()
        dmark("src2 done", s_src)
    hmark("src2 queued")
    with torch.cuda.stream(s_side):
        self.at_weights_np[...] = self.anchor_target.weights(pending["num_examples"])
        self.graphs["src3"].replay()
        dmark("src3 done", s_side)
    hmark("src3 queued")
    self.anchor_target.prefetch_stream(self.at_host.numel())
    hmark("next key blocks generated")
    for s in (s_src, s_tgt, s_side):
        cur.wait_stream(s)
    torch.cuda.synchronize()
    hmark("synchronized")
    h0 = host[0][1]
    return ([(n, (t - h0) * 1e6) for n, t in host], [(n, t0e.elapsed_time(e) * 1e3) for n, e in devs])


runs = []
for _ in range(30):
    flush.zero_()
    torch.cuda.synchronize()
    runs.append(one())
print("median over 30 steps, us after the step's start")
print("host:")
for i, (name, _) in enumerate(runs[0][0]):
    print("  %-26s %7.0f" % (name, float(np.median([r[0][i][1] for r in runs]))))
print("device (event on its stream):")
for i, (name, _) in enumerate(runs[0][1]):
    print("  %-26s %7.0f" % (name, float(np.median([r[1][i][1] for r in runs]))))
