#!/usr/bin/env python
"""Benchmark of the RoI / proposal hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path (oracle port)

Workload (BASELINE.json configs[1]), per GPU: the operator sequence of one DAF
training step on 2 source + 2 target synthetic 600x1200 images (VGG16 conv5 maps 512x37x75, A = 12):
  source: proposal layer TRAIN (12000 -> 2000, NMS 0.7) -> _ProposalTargetLayer (IoU, host-RNG
          fg/bg sampling of 256 RoIs/image, regression targets) -> RoIAlignAvg 7x7 forward +
          backward, anchor-target layer -> RPN cls / box losses forward + backward,
          image / instance / consistency DA losses forward + backward, GRL;
  target: proposal layer TEST (6000 -> 300) -> RoIAlignAvg forward + backward on 300 RoIs/image,
          DA losses forward + backward, GRL.
`value` = RoIs pooled forward+backward per second over the whole step, all GPUs (weak scaling:
every rank runs its own 4 images; there is no collective on the path).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

C, H, W, A = 512, 37, 75, 12
ROIS_SRC, ROIS_TGT = 256, 300
METRIC = "RoIs/sec RoIAlign fwd+bwd & proposals/sec (12000->2000 NMS) vs HBM roofline"


class Workload(object):
    """cfg2: the DAF step (2 + 2 images).  cfg4: 4 + 4 images per GPU with the MAF multi-level domain
    classifiers (conv3 / conv4 / conv5 image heads, instance cross-entropy, weighted GRL)."""

    def __init__(self, name):
        self.name = name
        self.maf = name == "cfg4"
        self.n_src = self.n_tgt = 4 if self.maf else 2
        self.rois_per_step = self.n_src * ROIS_SRC + self.n_tgt * ROIS_TGT
        self.proposals_per_step = self.n_src * 2000 + self.n_tgt * 300
        if self.maf:
            self.text = ("cfg4: MAF step ops, 4 src + 4 tgt 600x1200 images/GPU, conv5 512x37x75, RPN 12000->2000 / "
                         "6000->300 NMS 0.7, proposal targets, RoIAlignAvg 7x7 fwd+bwd on 256/300 RoIs per image, "
                         "anchor targets + RPN losses, image DA heads at conv3 (2x150x300) / conv4 (2x75x150) / "
                         "conv5 (2x37x75), instance cross-entropy, weighted GRL")
        else:
            self.text = ("cfg2: DAF VGG16 step ops, 2 src + 2 tgt 600x1200 images/GPU, conv5 512x37x75, "
                         "RPN 12000->2000 (src) / 6000->300 (tgt) NMS 0.7, proposal targets (256 RoIs/image), "
                         "RoIAlignAvg 7x7 fwd+bwd on 256/300 RoIs per image, anchor targets + RPN losses, "
                         "image+instance DA losses, GRL")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_inputs(seed, wl):
    """Host tensors of one rank's step (SURVEY.md 8d generators)."""
    from tools.synth import synth_gt, synth_rpn
    g = torch.Generator().manual_seed(seed)
    d = {}
    for dom, n in (("src", wl.n_src), ("tgt", wl.n_tgt)):
        d[dom + "_feat"] = torch.relu(torch.randn(n, C, H, W, generator=g))
        prob, deltas = synth_rpn(n, A, H, W, seed + (1 if dom == "src" else 2))
        d[dom + "_prob"], d[dom + "_deltas"] = prob, deltas
        d[dom + "_im_info"] = torch.tensor([[600.0, 1200.0, 0.5859375]] * n)
        d[dom + "_img_score"] = torch.randn(n, 2, H, W, generator=g)
        r = n * (ROIS_SRC if dom == "src" else ROIS_TGT)
        d[dom + "_ins_prob"] = torch.sigmoid(torch.randn(r, 1, generator=g))
        if wl.maf:
            d[dom + "_img_score3"] = torch.randn(n, 2, 150, 300, generator=g)
            d[dom + "_img_score4"] = torch.randn(n, 2, 75, 150, generator=g)
            d[dom + "_ins_logit"] = torch.randn(r, 2, generator=g)
    # RPN head outputs of the source images (logits; the proposal layer gets their softmax above)
    d["src_rpn_cls"] = 2.0 * torch.randn(wl.n_src, 2 * A, H, W, generator=g)
    d["src_rpn_box"] = 0.2 * torch.randn(wl.n_src, 4 * A, H, W, generator=g)
    d["src_gt"] = synth_gt(wl.n_src, 20, 50, seed + 50)
    return d


class Arena(object):
    """One pinned host buffer and one device buffer holding a group of tensors back to back: a
    whole group crosses PCIe with ONE copy (bench e2e mode)."""

    def __init__(self, tensors, dev, to_device):
        self.names = list(tensors)
        sizes = [(tensors[k].numel() * 4 + 255) // 256 * 256 for k in self.names]
        total = sum(sizes)
        self.host = torch.empty(total, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(total, dtype=torch.uint8, device=dev)
        self.hviews, self.dviews = {}, {}
        off = 0
        for k, sz in zip(self.names, sizes):
            t = tensors[k]
            n = t.numel() * 4
            self.hviews[k] = self.host[off:off + n].view(torch.float32).view(t.shape)
            self.dviews[k] = self.dev[off:off + n].view(torch.float32).view(t.shape)
            if to_device:
                self.hviews[k].copy_(t)
            off += sz
        self.nbytes = total
        if to_device:
            self.dev.copy_(self.host)

    def h2d(self):
        self.dev.copy_(self.host, non_blocking=True)

    def d2h(self):
        self.host.copy_(self.dev, non_blocking=True)


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
class TlodStep(object):
    """One rank's step.  Device work without a host dependency is captured in three CUDA graphs:
      src1: proposal layer TRAIN, candidate assembly, IoU/assignment, pinned D2H of the overlaps
      src2: proposal targets, RoIAlignAvg fwd + bwd, GRL                 (after the host's fg / bg sampling)
      src3: DA losses fwd + bwd, RPN losses fwd + bwd                    (after the anchor-target layer)
      tgt : proposal layer TEST, RoIAlignAvg fwd + bwd, DA losses fwd + bwd, GRL
    The host part -- the anchor-target layer's subsampling, then the fg / bg sampling of the proposal
    targets, in the reference's order on numpy's RNG stream -- runs while the GPU works through
    src1 / tgt; it is the longest chain of the step."""

    def __init__(self, dev, seed, wl, use_graph=True):
        from model.roi_align.modules.roi_align import RoIAlignAvg
        from model.rpn.anchor_target_layer import _AnchorTargetLayer
        from model.rpn.proposal_layer import _ProposalLayer
        from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
        from model.utils.config import cfg
        import tlod_b200
        self.tlod = tlod_b200
        self.dev = dev
        self.wl = wl
        cfg.TRAIN.BATCH_SIZE = ROIS_SRC  # cfgs/vgg16.yml
        host = synth_inputs(seed, wl)
        # inputs in two arenas (one H2D copy per domain in e2e mode)
        self.arena = {dom: Arena({k: v for k, v in host.items() if k.startswith(dom)}, dev, True)
                      for dom in ("src", "tgt")}
        self.d = {}
        for a in self.arena.values():
            self.d.update(a.dviews)
        self.proposal = _ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
        self.anchor_target = _AnchorTargetLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
        self.proposal_target = _ProposalTargetLayer(9)
        self.roi_align = RoIAlignAvg(7, 7, 1.0 / 16.0)
        g = torch.Generator().manual_seed(seed + 99)
        # gradient arriving from the detection head: device resident (it never exists on the host)
        self.top = {"src": torch.randn(wl.n_src * ROIS_SRC, C, 7, 7, generator=g).to(dev),
                    "tgt": torch.randn(wl.n_tgt * ROIS_TGT, C, 7, 7, generator=g).to(dev)}
        if wl.maf:
            self.cls_prob = {dom: torch.softmax(torch.randn(r, 9, generator=g), 1).to(dev)
                             for dom, r in (("src", wl.n_src * ROIS_SRC), ("tgt", wl.n_tgt * ROIS_TGT))}
            self.pooled_grad = {dom: torch.randn(r, 4096, generator=g).to(dev)
                                for dom, r in (("src", wl.n_src * ROIS_SRC), ("tgt", wl.n_tgt * ROIS_TGT))}
        self.num_boxes = torch.full((wl.n_src,), 20, dtype=torch.long)
        # the source chain (labels -> host sampling -> RoI work) is the critical path: its streams, and the
        # graphs captured for them, get the higher priority; the target domain fills the gaps
        prio = {"src": -1, "side": -1, "tgt": 0}
        if os.environ.get("TLOD_BENCH_FLAT_PRIORITY"):
            prio = {k: 0 for k in prio}
        self.streams = {k: torch.cuda.Stream(dev, priority=p) for k, p in prio.items()}
        self.capture_streams = {k: torch.cuda.Stream(dev, priority=p) for k, p in prio.items()}
        # static buffers of the proposal-target hand-over (graph mode replays into / out of them)
        self.pt_host = torch.empty((wl.n_src, 2000 + 50), dtype=torch.float32).pin_memory()
        self.keep_d = torch.zeros((wl.n_src, ROIS_SRC), dtype=torch.int32, device=dev)
        self.fg_d = torch.zeros((wl.n_src,), dtype=torch.int32, device=dev)
        self.keep_h = torch.zeros((wl.n_src, ROIS_SRC), dtype=torch.int32).pin_memory()
        self.fg_h = torch.zeros((wl.n_src,), dtype=torch.int32).pin_memory()
        self.at_static = None
        self.at_host = None
        self.graphs = None
        self.static = {}
        self.launches_per_replay = 0
        self.out_arena = None
        np.random.seed(3)
        if use_graph:
            self.capture()

    # ---- the pieces (each only enqueues work on the current stream) ----
    def da_part(self, dom, d, pooled_rois):
        """Domain-classifier losses forward + backward and the GRLs in front of the heads."""
        dl = 1 if dom == "src" else 0
        score = d[dom + "_img_score"].requires_grad_(True)
        if self.wl.maf:
            s3 = d[dom + "_img_score3"].requires_grad_(True)
            s4 = d[dom + "_img_score4"].requires_grad_(True)
            ins = d[dom + "_ins_logit"].requires_grad_(True)
            losses = self.tlod.image_da_losses([s3, s4, score, ins.view(-1, 2, 1, 1)], dl)
            (0.1 * losses.sum()).backward()
            # weighted GRL in front of the instance head (lib/MAF/DA.py:34-53): -0.2 * g * cls_prob[:, dc_label]
            g_ins = self.tlod.functional.grl_backward(self.pooled_grad[dom], 0.2, self.cls_prob[dom][:, dl].contiguous())
            out = (losses.detach(), score.grad, s3.grad, s4.grad, ins.grad, g_ins)
            for t in (score, s3, s4, ins):
                t.grad = None
            return out
        prob = d[dom + "_ins_prob"].requires_grad_(True)
        img, ins, cst = self.tlod.da_losses(score, prob, dl)
        (0.1 * (img + ins + cst)).backward()
        out = (torch.stack([img, ins, cst]).detach(), score.grad, prob.grad)
        score.grad = None
        prob.grad = None
        return out

    def pool_part(self, dom, d, rois_flat):
        feat = d[dom + "_feat"].requires_grad_(True)
        pooled = self.roi_align(feat, rois_flat)
        pooled.backward(self.top[dom])
        g_feat = self.tlod.functional.grl_backward(feat.grad, 0.1)  # GRL in front of the image DA head
        out = (feat.grad, g_feat)
        feat.grad = None
        return out

    def src1(self, d):
        rois = self.proposal((d["src_prob"], d["src_deltas"], d["src_im_info"], "TRAIN"))
        return rois, self.proposal_target.begin(rois, d["src_gt"], self.num_boxes, host=self.pt_host)

    def src2(self, d, pt_state, keep, fg):
        rois, labels, targets, inside, outside = self.proposal_target.finish(pt_state, keep, fg)
        return {"rois": rois, "labels": labels, "targets": targets, "pool": self.pool_part("src", d, rois.view(-1, 5))}

    def src3(self, d, at):
        """Source-domain work that does not wait for the sampled RoIs: DA losses, RPN losses."""
        da = self.da_part("src", d, None)
        cls = d["src_rpn_cls"].requires_grad_(True)
        box = d["src_rpn_box"].requires_grad_(True)
        l_cls, l_box = self.tlod.rpn_losses(cls, box, at[0], at[1], at[2], at[3])
        (l_cls + l_box).backward()
        out = {"da": da, "rpn": (torch.stack([l_cls, l_box]).detach(), cls.grad, box.grad)}
        cls.grad = None
        box.grad = None
        return out

    def tgt(self, d):
        rois = self.proposal((d["tgt_prob"], d["tgt_deltas"], d["tgt_im_info"], "TEST"))
        flat = rois.reshape(-1, 5)
        return {"rois": rois, "pool": self.pool_part("tgt", d, flat), "da": self.da_part("tgt", d, rois)}

    def capture(self):
        """Warm up eagerly (allocator, lazy module state), then capture the three graphs."""
        try:
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self.step_eager()
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            at = self.anchor_target((self.d["src_prob"], self.d["src_gt"], self.d["src_im_info"], self.num_boxes))
            self.at_static = [t.clone() for t in at]
            n_inside = self.anchor_target._inside(H, W, 1200, 600, self.dev)[0].size(0)
            self.at_host = torch.empty((self.wl.n_src, n_inside), dtype=torch.float32).pin_memory()
            graphs = {}
            n0 = self.tlod.launch_count()
            graphs["at1"] = torch.cuda.CUDAGraph()   # anchor labels + their pinned D2H copy
            with torch.cuda.graph(graphs["at1"], stream=self.capture_streams["side"]):
                self.static["at1"] = self.anchor_target.launch_labels(self.d["src_gt"], H, W, (600, 1200),
                                                                      labels_host=self.at_host)
            graphs["src1"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graphs["src1"], stream=self.capture_streams["src"]):
                self.static["src1"] = self.src1(self.d)
            graphs["src2"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graphs["src2"], stream=self.capture_streams["src"]):
                self.keep_d.copy_(self.keep_h, non_blocking=True)
                self.fg_d.copy_(self.fg_h, non_blocking=True)
                self.static["src2"] = self.src2(self.d, self.static["src1"][1], self.keep_d, self.fg_d)
            # src3 starts with the anchor targets' device half: upload of the subsampled labels and of
            # the three weights (written into this pinned tensor before every replay), finalize kernel
            self.at_weights = torch.zeros(3, dtype=torch.float32).pin_memory()
            self.at_weights_np, self.keep_np, self.fg_np = self.at_weights.numpy(), self.keep_h.numpy(), self.fg_h.numpy()
            self.events = {"src1": torch.cuda.Event(), "at1": torch.cuda.Event()}
            graphs["src3"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graphs["src3"], stream=self.capture_streams["side"]):
                self.at_static = self.anchor_target.launch_finalize(self.static["at1"], self.at_weights)
                self.static["src3"] = self.src3(self.d, self.at_static)
            graphs["tgt"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graphs["tgt"], stream=self.capture_streams["tgt"]):
                self.static["tgt"] = self.tgt(self.d)
            self.launches_per_replay = int(self.tlod.launch_count() - n0)
            torch.cuda.synchronize(self.dev)
            self.graphs = graphs
        except Exception as e:  # noqa: BLE001 -- eager is always available
            sys.stderr.write("bench.py: CUDA graph capture failed (%r); running eagerly\n" % (e,))
            self.graphs = None

    def step_eager(self):
        d = self.d
        res = {}
        res["tgt"] = self.tgt(d)
        pending = self.anchor_target.begin((d["src_prob"], d["src_gt"], d["src_im_info"], self.num_boxes),
                                           im_hw=(600, 1200))
        rois, st = self.src1(d)
        # numpy RNG order of the reference step: anchor-target layer (inside the RPN), then proposal targets
        at = self.anchor_target.finish(pending)
        keep, fg = self.proposal_target.sample(st)
        res["src"] = dict(self.src2(d, st, keep, fg), **self.src3(d, at))
        res["anchor_targets"] = at
        return res

    def step(self, copy_in=None, copy_out=None):
        """copy_in / copy_out (e2e mode) put each domain's H2D / D2H copy on that domain's stream, so
        that the target domain's results travel back while the source domain is still computing."""
        if self.graphs is None:
            if copy_in is not None:
                copy_in("src"), copy_in("tgt")
            r = self.step_eager()
            if copy_out is not None:
                copy_out("tgt", r), copy_out("src", r)
            return r
        d = self.d
        cur = torch.cuda.current_stream(self.dev)
        s_src, s_tgt, s_side = self.streams["src"], self.streams["tgt"], self.streams["side"]
        for s in (s_src, s_tgt, s_side):
            s.wait_stream(cur)
        with torch.cuda.stream(s_src):
            if copy_in is not None:
                copy_in("src")
                arrived = torch.cuda.Event()
                arrived.record(s_src)
                s_side.wait_event(arrived)  # the anchor-target layer reads inputs this copy delivers
        # two chains feed the host sampling: the anchor labels (at1, ~35 us of device time, then ~90 us of
        # host subsampling) and proposals + IoU (src1, ~125 us of device time): the one with the host part
        # behind it is queued first
        with torch.cuda.stream(s_side):
            self.graphs["at1"].replay()
            pending = self.static["at1"]
            pending["copied"] = self.events["at1"]
            pending["copied"].record(s_side)
            pending["stream"] = s_side
        with torch.cuda.stream(s_src):
            self.graphs["src1"].replay()
            copied = self.events["src1"]
            copied.record(s_src)
        result = {"src": dict(self.static["src2"], **self.static["src3"]), "tgt": self.static["tgt"],
                  "anchor_targets": self.at_static}
        with torch.cuda.stream(s_tgt):
            if copy_in is not None:
                copy_in("tgt")
            self.graphs["tgt"].replay()
            if copy_out is not None:
                copy_out("tgt", result)
        # host, in the reference's RNG order: anchor subsampling (inside the RPN), then the fg / bg sampling.
        # The longest device chain (src2: RoIAlignAvg forward + backward on the sampled RoIs) is queued
        # before the anchor targets' upload / finalize / RPN losses, which are short and run beside it.
        pending = self.anchor_target.finish_host(pending)
        st = self.static["src1"][1]
        st["copied"] = copied
        self.proposal_target.sample(st, out=(self.keep_np, self.fg_np))  # straight into the pinned upload buffers
        with torch.cuda.stream(s_src):
            self.graphs["src2"].replay()  # starts with the upload of keep / fg from their pinned buffers
        with torch.cuda.stream(s_side):
            self.at_weights_np[...] = self.anchor_target.weights(pending["num_examples"])
            self.graphs["src3"].replay()
        with torch.cuda.stream(s_src):
            if copy_out is not None:
                s_src.wait_stream(s_side)  # the source arena also carries the RPN / DA losses and anchor labels
                copy_out("src", result)
        # numpy's stream is where the next step will find it: generate the key blocks of the next anchor
        # subsampling now, while the device works on src2
        self.anchor_target.prefetch_stream(self.at_host.numel())
        for s in (s_src, s_tgt, s_side):
            cur.wait_stream(s)
        return result

    # ---- end to end through host buffers ----
    def e2e_outputs(self, r, dom=None):
        src = {"src_rois": r["src"]["rois"], "src_labels": r["src"]["labels"], "src_targets": r["src"]["targets"],
               "src_gfeat": r["src"]["pool"][1], "src_losses": r["src"]["da"][0], "src_rpn_losses": r["src"]["rpn"][0],
               "anchor_labels": r["anchor_targets"][0]}
        tgt = {"tgt_rois": r["tgt"]["rois"], "tgt_gfeat": r["tgt"]["pool"][1], "tgt_losses": r["tgt"]["da"][0]}
        if dom == "src":
            return src
        if dom == "tgt":
            return tgt
        return dict(src, **tgt)

    def step_e2e(self):
        """Same step through host buffers: each domain's inputs arrive with one H2D copy from a pinned
        arena on that domain's stream; each domain's results leave packed in one pinned arena with one
        D2H copy on the same stream, as soon as that domain is done."""
        if self.out_arena is None:
            self.out_arena = {}

        def copy_out(dom, r):
            outs = self.e2e_outputs(r, dom)
            if dom not in self.out_arena:
                self.out_arena[dom] = Arena(outs, self.dev, False)
            a = self.out_arena[dom]
            for k, v in outs.items():
                a.dviews[k].copy_(v.reshape(a.dviews[k].shape), non_blocking=True)
            a.d2h()

        self.step(copy_in=lambda dom: self.arena[dom].h2d(), copy_out=copy_out)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.out_arena

    def e2e_bytes(self):
        h2d = sum(a.nbytes for a in self.arena.values())
        d2h = sum(a.nbytes for a in self.out_arena.values()) if self.out_arena else 0
        return h2d, d2h

    def check_graph_against_eager(self):
        """The captured graphs must produce what the eager modules produce (same numpy RNG state)."""
        if self.graphs is None:
            return "eager (no graphs)"
        state = np.random.get_state()
        a = self.step()
        torch.cuda.synchronize(self.dev)
        keys = ("src_rois", "src_labels", "src_targets", "src_gfeat", "src_losses", "src_rpn_losses", "tgt_rois",
                "tgt_gfeat", "tgt_losses", "anchor_labels")
        got = {k: v.clone() for k, v in self.e2e_outputs(a).items()}
        np.random.set_state(state)
        b = self.e2e_outputs(self.step_eager())
        torch.cuda.synchronize(self.dev)
        for k in keys:
            x, y = got[k], b[k].reshape(got[k].shape)
            if not torch.allclose(x, y, rtol=1e-5, atol=1e-7, equal_nan=True):
                raise SystemExit("bench.py: CUDA-graph replay differs from the eager step in %s (max |d| = %g)"
                                 % (k, float((x - y).abs().max())))
        return "graph replay == eager step on %d outputs" % len(keys)


def clocks_sampler(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100",
                              "-i", "0"], stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None
    try:
        # the sampler gets the last core of this rank's slice and the step's host thread keeps off it:
        # with 4 cores per rank (8 ranks on a 32-core box) rank 0 was otherwise the straggler
        cores = sorted(os.sched_getaffinity(0))
        if len(cores) > 1:
            os.sched_setaffinity(p.pid, {cores[-1]})
            os.sched_setaffinity(0, set(cores[:-1]))
    except Exception:  # noqa: BLE001
        pass
    return p


def parse_clocks(path):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            sm.append(float(f[1]))
            mx.append(float(f[2]))
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons)}


def per_call_us(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def load_traffic(section):
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(section, {})
    except Exception:  # noqa: BLE001
        return {}


def roi_align_sizes(dev):
    """RoIAlign(8, 8) and RoIAlignAvg(7, 7) forward / backward alone at the BASELINE shapes: cfg1 and cfg2
    (L2 resident), cfg3 (630 MB of algorithmic traffic per direction: larger than L2, no flush needed)."""
    from tools.synth import synth_rois
    from tlod_b200 import functional as F
    peak, peak_src = peaks()
    out = {"peak_GBps": peak, "peak_source": peak_src}
    g = torch.Generator().manual_seed(7)
    for name, (B, Cc, Hh, Ww, R) in (("cfg1", (1, 512, 37, 75, 128)), ("cfg2", (2, 512, 37, 75, 512)),
                                     ("cfg3", (8, 1024, 38, 75, 2048))):
        x = torch.relu(torch.randn(B, Cc, Hh, Ww, generator=g)).to(dev)
        rois = synth_rois(R, B, 41)
        rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)  # (B, R/B, 5).view(-1, 5) order
        top8 = torch.randn(R, Cc, 8, 8, device=dev)
        top7 = torch.randn(R, Cc, 7, 7, device=dev)
        feat_b = B * Cc * Hh * Ww * 4 + R * 20
        alg8, alg7 = feat_b + R * Cc * 64 * 4, feat_b + R * Cc * 49 * 4
        plan = F.roi_align_plan(rois, x.shape, 8, 8, 1.0 / 16)
        it = 20
        t = {"plan": per_call_us(lambda: F.roi_align_plan(rois, x.shape, 8, 8, 1.0 / 16), it),
             "fwd": per_call_us(lambda: F.roi_align_forward(x, rois, 8, 8, 1.0 / 16, plan=plan), it),
             "bwd": per_call_us(lambda: F.roi_align_backward(top8, rois, x.shape, 1.0 / 16, plan=plan), it),
             "avg_fwd": per_call_us(lambda: F.roi_align_avg_forward(x, rois, 7, 7, 1.0 / 16, plan=plan), it),
             "avg_bwd": per_call_us(lambda: F.roi_align_avg_backward(top7, rois, x.shape, 1.0 / 16, plan=plan), it)}

        def frac(nbytes, us):
            return nbytes / (us * 1e-6) / 1e9 / peak

        # the same three launches (plan, forward, backward) captured in ONE CUDA graph: device time of the
        # sequence without the host's launch path (Python + ctypes + torch.empty per call, ~8 us each, is
        # what bounds the back-to-back eager `plan` figure above)
        def chain(avg):
            pl = F.roi_align_plan(rois, x.shape, 8, 8, 1.0 / 16)
            if avg:
                F.roi_align_avg_forward(x, rois, 7, 7, 1.0 / 16, plan=pl)
                return F.roi_align_avg_backward(top7, rois, x.shape, 1.0 / 16, plan=pl)
            F.roi_align_forward(x, rois, 8, 8, 1.0 / 16, plan=pl)
            return F.roi_align_backward(top8, rois, x.shape, 1.0 / 16, plan=pl)
        graph_us = {}
        for key, avg in (("graph_plan_fwd_bwd", False), ("graph_plan_avg_fwd_bwd", True)):
            try:
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    chain(avg)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    keep_alive = chain(avg)
                graph_us[key] = per_call_us(gr.replay, it)
                del gr, keep_alive
            except Exception as e:  # noqa: BLE001
                graph_us[key] = None
                sys.stderr.write("bench.py: graph timing of %s failed: %r\n" % (key, e))
        traffic = load_traffic(name)
        e = {"shape": [B, Cc, Hh, Ww, R], "l2_resident": name != "cfg3", "plan": {"us": t["plan"]},
             "fwd": {"us": t["fwd"], "kernel": "roi_align_fwd8_kernel", "algorithmic_bytes": alg8,
                     "achieved_GBps": alg8 / t["fwd"] / 1e3, "frac_of_hbm_peak": frac(alg8, t["fwd"]),
                     "traffic": traffic.get("roi_align_fwd8_kernel")},
             "bwd": {"us": t["bwd"], "kernel": "roi_align_bwd_rows_kernel", "algorithmic_bytes": alg8,
                     "achieved_GBps": alg8 / t["bwd"] / 1e3, "frac_of_hbm_peak": frac(alg8, t["bwd"]),
                     "traffic": traffic.get("roi_align_bwd_rows_kernel")},
             "avg_fwd": {"us": t["avg_fwd"], "kernel": "roi_align_avg_fwd8_kernel (fused 2x2 average)",
                         "algorithmic_bytes_fused": alg7, "frac_of_hbm_peak_fused_bytes": frac(alg7, t["avg_fwd"]),
                         "frac_of_hbm_peak_op_surface_bytes": frac(alg8 + R * Cc * (64 + 49) * 4, t["avg_fwd"]),
                         "traffic": traffic.get("roi_align_avg_fwd8_kernel")},
             "avg_bwd": {"us": t["avg_bwd"], "kernel": "avgpool2x2_bwd_kernel + roi_align_bwd_rows_kernel",
                         "algorithmic_bytes_fused": alg7, "frac_of_hbm_peak_fused_bytes": frac(alg7, t["avg_bwd"]),
                         "frac_of_hbm_peak_op_surface_bytes": frac(alg8 + R * Cc * (64 + 49) * 4, t["avg_bwd"])},
             "rois_per_s_fwd_bwd": R / ((t["plan"] + t["fwd"] + t["bwd"]) * 1e-6),
             "frac_fwd_bwd": frac(2 * alg8, t["plan"] + t["fwd"] + t["bwd"]),
             "graph_plan_fwd_bwd_us": graph_us["graph_plan_fwd_bwd"],
             "frac_fwd_bwd_graph": None if not graph_us["graph_plan_fwd_bwd"] else
             frac(2 * alg8, graph_us["graph_plan_fwd_bwd"]),
             "graph_plan_avg_fwd_bwd_us": graph_us["graph_plan_avg_fwd_bwd"],
             "frac_avg_fwd_bwd_fused_bytes_graph": None if not graph_us["graph_plan_avg_fwd_bwd"] else
             frac(2 * alg7, graph_us["graph_plan_avg_fwd_bwd"]),
             "rois_per_s_avg_fwd_bwd": R / ((t["plan"] + t["avg_fwd"] + t["avg_bwd"]) * 1e-6),
             "frac_avg_fwd_bwd_fused_bytes": frac(2 * alg7, t["plan"] + t["avg_fwd"] + t["avg_bwd"])}
        out[name] = e
        del x, top8, top7
    out["note"] = ("back-to-back launches, CUDA events around the loop; each call includes its torch.empty "
                   "(caching allocator); the plan is built once per rois tensor and shared by fwd and bwd; "
                   "'op surface' bytes count the (R,C,8,8) tensor written and re-read by the unfused reference path")
    return out


def secondary_metrics(dev, iters=20):
    """The other metrics SURVEY 8(d) names, measured alone on rank 0: the two proposal pipelines
    (fraction of the bound 8(d) defines: max(bytes / HBM rate, 16 flop per upper-triangle pair / fp32
    SIMT rate)), the cfg5 NMS sweep, plain NMS, RoIPool, RoICrop (unfused and fused with its max-pool)
    and the multi-level DA losses at cfg4 map sizes."""
    from tools.synth import synth_rois, synth_rpn
    from model.rpn.generate_anchors import generate_anchors
    from model.utils.net_utils import _affine_grid_gen
    from tlod_b200 import functional as F
    import tlod_b200
    peak, _ = peaks()
    fp32_tflops = 148 * 128 * 2 * 1.965e9 / 1e12  # nominal SIMT rate
    anchors = torch.from_numpy(generate_anchors(scales=np.array([4, 8, 16, 32]),
                                                ratios=np.array([0.5, 1, 2]))).float().to(dev)
    out = {"fp32_simt_tflops_nominal": fp32_tflops, "proposal_layer": {}, "nms_sweep_test_6000_300": []}
    inputs = {}

    def rpn(batch):
        if batch not in inputs:
            prob, deltas = synth_rpn(batch, A, H, W, 3)
            info = torch.tensor([[600.0, 1200.0, 0.5859375]] * batch)
            inputs[batch] = [t.to(dev) for t in (prob, deltas, info)]
        return inputs[batch]

    for name, pre, post in (("TRAIN_12000_2000", 12000, 2000), ("TEST_6000_300", 6000, 300)):
        prob, deltas, info = rpn(2)
        rois, order, boxes, num = F.proposals(prob, deltas, info, anchors, 16, pre, post, 0.7, return_debug=True)
        us = per_call_us(lambda: F.proposals(prob, deltas, info, anchors, 16, pre, post, 0.7), iters)
        nbytes = 2 * (A * H * W * 4 * 5 + post * 20)
        flops = 2 * (pre * (pre - 1) / 2) * 16
        bound_us = max(nbytes / (peak * 1e9), flops / (fp32_tflops * 1e12)) * 1e6
        out["proposal_layer"][name] = {"images": 2, "us_per_call": us, "proposal_sets_per_s": 2 / (us * 1e-6),
                                       "proposals_per_s": 2 * post / (us * 1e-6), "bound_us": bound_us,
                                       "frac_of_bound": bound_us / us, "kept": [int(v) for v in num.tolist()],
                                       "note": "bound = full upper-triangle IoU work on the nominal fp32 SIMT "
                                               "rate; the phased NMS computes only the triangle the scan reaches"}
    for batch in (1, 8, 64):
        prob, deltas, info = rpn(batch)
        for thr in (0.3, 0.5, 0.7):
            us = per_call_us(lambda: F.proposals(prob, deltas, info, anchors, 16, 6000, 300, thr), iters)
            out["nms_sweep_test_6000_300"].append({"batch": batch, "iou": thr, "us_per_call": us,
                                                   "proposal_sets_per_s": batch / (us * 1e-6)})
    # plain NMS on score-sorted boxes (the reference's nms_cuda_compute workload)
    from tools.reference_bench import sorted_boxes
    for n in (12000, 6000):
        dets = sorted_boxes(n, 500 + n).to(dev)
        out["nms_%d" % n] = {"us": per_call_us(lambda: F.nms_device(dets, 0.7), iters), "thresh": 0.7,
                             "note": "device-resident keep + count, no host synchronisation"}
    # RoIPool 7x7 forward + backward at cfg2 scale (VGG16 conv5, 2 images, 512 RoIs)
    g = torch.Generator().manual_seed(5)
    feat = torch.relu(torch.randn(2, C, H, W, generator=g)).to(dev)
    rois = synth_rois(512, 2, 41)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
    top = torch.randn(512, C, 7, 7, device=dev)
    pooled, argmax = F.roi_pool_forward(feat, rois, 7, 7, 1.0 / 16)
    us_f = per_call_us(lambda: F.roi_pool_forward(feat, rois, 7, 7, 1.0 / 16), iters)
    us_b = per_call_us(lambda: F.roi_pool_backward(top, argmax, rois, feat.shape, 1.0 / 16), iters)
    alg_f = feat.numel() * 4 + 512 * 20 + 512 * C * 49 * 8
    out["roi_pool_cfg2"] = {
        "fwd_us": us_f, "bwd_us": us_b, "rois_per_s_fwd_bwd": 512 / ((us_f + us_b) * 1e-6),
        "frac_of_hbm_peak_fwd": alg_f / (us_f * 1e-6) / 1e9 / peak,
        "frac_of_hbm_peak_bwd": alg_f / (us_b * 1e-6) / 1e9 / peak,
        "note": "L2-warm back-to-back launches (working set 63 MB < L2); bound by the instruction count of the "
                "strict-> window scan (~10 cell visits per output), see profiles/r02_rejected_roi_pool_planes_ncu.txt"}
    # BASELINE cfg3 (ii): the crop path -- RoICrop 14x14 on conv4 (8, 1024, 38, 75), 2048 RoIs, then max-pool 7x7
    Bc, Cc, Hc, Wc, Rc, G = 8, 1024, 38, 75, 2048, 14
    featc = torch.relu(torch.randn(Bc, Cc, Hc, Wc, generator=g)).to(dev)
    roisc = synth_rois(Rc, Bc, 41)
    roisc = roisc[torch.argsort(roisc[:, 0], stable=True)].contiguous().to(dev)
    grid_xy = _affine_grid_gen(roisc, (Hc, Wc), G)
    grid_yx = torch.stack([grid_xy[..., 1], grid_xy[..., 0]], 3).contiguous()
    gy, gx = grid_xy[:, :, 0, 1].contiguous(), grid_xy[:, 0, :, 0].contiguous()
    topc = torch.randn(Rc, Cc, G, G, device=dev)
    top7 = torch.randn(Rc, Cc, 7, 7, device=dev)
    us_cf = per_call_us(lambda: F.roi_crop_forward(featc, grid_yx), 5, 2)
    us_cb = per_call_us(lambda: F.roi_crop_backward(topc, grid_yx, featc.shape), 5, 2)
    o7, a7 = F.roi_crop_pool_forward(featc, gy, gx)
    us_pf = per_call_us(lambda: F.roi_crop_pool_forward(featc, gy, gx), 5, 2)
    us_pb = per_call_us(lambda: F.roi_crop_pool_backward(top7, a7, gy, gx, featc.shape), 5, 2)
    alg_c = featc.numel() * 4 + grid_yx.numel() * 4 + Rc * Cc * G * G * 4
    alg_p = featc.numel() * 4 + Rc * 28 * 4 + Rc * Cc * 49 * 5
    out["roi_crop_cfg3_14x14"] = {
        "fwd_us": us_cf, "bwd_us": us_cb, "algorithmic_bytes_per_direction": alg_c,
        "frac_of_hbm_peak_fwd": alg_c / (us_cf * 1e-6) / 1e9 / peak,
        "frac_of_hbm_peak_bwd": alg_c / (us_cb * 1e-6) / 1e9 / peak,
        "fused_with_max_pool": {
            "fwd_us": us_pf, "bwd_us": us_pb, "algorithmic_bytes_fused": alg_p,
            "frac_of_hbm_peak_fwd_fused_bytes": alg_p / (us_pf * 1e-6) / 1e9 / peak,
            "frac_of_hbm_peak_bwd_fused_bytes": alg_p / (us_pb * 1e-6) / 1e9 / peak,
            "frac_of_hbm_peak_fwd_op_surface_bytes": alg_c / (us_pf * 1e-6) / 1e9 / peak,
            "frac_of_hbm_peak_bwd_op_surface_bytes": alg_c / (us_pb * 1e-6) / 1e9 / peak,
            "rois_per_s_fwd_bwd": Rc / ((us_pf + us_pb) * 1e-6),
            "note": "crop 14x14 -> max_pool2d(2,2) as one kernel each way: the 1.64 GB sample tensor is never written"}}
    del featc, topc, top7, o7, a7
    # multi-level image DA losses at cfg4 map sizes (8 images): conv3 / conv4 / conv5 + instance CE
    maps = [torch.randn(8, 2, 150, 300, device=dev), torch.randn(8, 2, 75, 150, device=dev),
            torch.randn(8, 2, 37, 75, device=dev), torch.randn(2048, 2, 1, 1, device=dev)]
    lo = F.da_image_loss_forward(maps, 1)
    out["da_image_losses_cfg4"] = {
        "fwd_us": per_call_us(lambda: F.da_image_loss_forward(maps, 1), iters),
        "bwd_us": per_call_us(lambda: F.da_image_loss_backward(maps, 1, lo), iters),
        "levels": [list(m.shape) for m in maps], "launches": "1 forward (+ an 80-byte memset), 1 backward"}
    return out


def reference_subprocess(cpu_steps, timeout=900):
    """tools/reference_bench.py in its own process: the reference's recompiled CUDA kernels, its CPU
    RoIAlign as shipped and the oracle port -- nothing under oracle/ is mapped into this process."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_bench.py"), "--cpu-steps",
                            str(cpu_steps)], capture_output=True, text=True, timeout=timeout)
        lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not lines:
            return {"error": "reference_bench.py rc=%d: %s" % (p.returncode, p.stderr[-400:])}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def speedups(ref, mine_align, secondary):
    """Same-box GPU-vs-GPU ratios: reference kernel time / this library's time."""
    out = {}
    g = ref.get("gpu_reference", {}) if isinstance(ref, dict) else {}
    for cfg_name in ("cfg1", "cfg2", "cfg3"):
        r, m = g.get("roi_align_" + cfg_name), mine_align.get(cfg_name) if mine_align else None
        if r and m:
            out["roi_align_%s_fwd" % cfg_name] = r["fwd_us"] / m["fwd"]["us"]
            out["roi_align_%s_bwd" % cfg_name] = r["bwd_us"] / m["bwd"]["us"]
    if secondary:
        r = g.get("roi_pool_cfg2")
        if r:
            out["roi_pool_cfg2_fwd"] = r["fwd_us"] / secondary["roi_pool_cfg2"]["fwd_us"]
            out["roi_pool_cfg2_bwd"] = r["bwd_us"] / secondary["roi_pool_cfg2"]["bwd_us"]
        r = g.get("roi_crop_cfg3")
        if r:
            c = secondary["roi_crop_cfg3_14x14"]
            out["roi_crop_cfg3_fwd"] = r["fwd_us"] / c["fwd_us"]
            out["roi_crop_cfg3_bwd"] = r["bwd_us"] / c["bwd_us"]
        for n in (12000, 6000):
            r = g.get("nms_%d" % n)
            if r:
                out["nms_%d" % n] = r["us"] / secondary["nms_%d" % n]["us"]
    return out


def bind_rank_to_cores(local, world_local):
    """Give every rank its own slice of the host cores before any pinned allocation (first touch
    then places the pinned pages near those cores)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        # whole physical cores per rank: with a plain slice of the logical CPU list, rank r and rank r + N/2
        # end up on the two hyper-threads of the same cores (8 ranks on 32 logical CPUs: rank 4 sat on the
        # siblings of rank 0's cores and ran 44 % slower than the other seven)
        groups = {}
        for c in cores:
            try:
                with open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c) as f:
                    key = f.read().strip()
            except OSError:
                key = str(c)
            groups.setdefault(key, []).append(c)
        phys = sorted(groups.values(), key=lambda g: g[0])
        per = max(len(phys) // max(world_local, 1), 1)
        mine = [c for g in phys[local * per:(local + 1) * per] for c in g] or cores
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(len(mine), 4)))
        return len(mine)
    except Exception:  # noqa: BLE001
        return None


def run_tlod(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    cores_bound = bind_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import tlod_b200
    from tlod_b200 import _lib

    wl = Workload(args.workload)
    step = TlodStep(dev, seed=3 + rank, wl=wl, use_graph=not args.no_graph)
    used_graph = step.graphs is not None
    graph_check = "skipped (--no-graph-check)" if args.no_graph_check else step.check_graph_against_eager()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations, outside the event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - wall0
        ms_own = sum(a.elapsed_time(b) for a, b in evs)
        ranks = [ms_own]
        if world > 1:
            t = torch.tensor([ms_own], device=dev, dtype=torch.float64)
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            ranks = [float(v.item()) for v in allv]
        return max(ranks), wall, ranks

    def pcie_probe():
        """Host <-> device copy rate of this box with ALL ranks copying at once (64 MB pinned buffers):
        the ceiling the end-to-end number sits under when every rank moves its inputs and results."""
        nb = 64 * 1024 * 1024
        hbuf = torch.empty(nb, dtype=torch.uint8).pin_memory()
        dbuf = torch.empty(nb, dtype=torch.uint8, device=dev)
        out = {}
        for name, fn in (("h2d", lambda: dbuf.copy_(hbuf, non_blocking=True)),
                         ("d2h", lambda: hbuf.copy_(dbuf, non_blocking=True))):
            fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(4):
                fn()
            b.record()
            torch.cuda.synchronize()
            rate = 4 * nb / (a.elapsed_time(b) * 1e-3) / 1e9
            vals = [rate]
            if world > 1:
                t = torch.tensor([rate], device=dev, dtype=torch.float64)
                allv = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allv, t)
                vals = [float(v.item()) for v in allv]
            out[name + "_GBps_per_rank_min"] = min(vals)
            out[name + "_GBps_aggregate"] = sum(vals)
        return out

    clock_file = os.path.join(ROOT, "gpurun_out", "bench_clocks_rank0.csv")
    os.makedirs(os.path.dirname(clock_file), exist_ok=True)
    sampler = clocks_sampler(clock_file) if rank == 0 else None
    launches0 = tlod_b200.launch_count()
    ms_dev, wall, ranks_dev = timed(step.step, args.steps, args.warmup)
    # kernels of this library inside the timed region: eager launches are counted by the library;
    # a graph replay re-issues the launches counted once at capture time
    eager_per_step = (tlod_b200.launch_count() - launches0) // (args.steps + args.warmup)
    launches = (eager_per_step + (step.launches_per_replay if used_graph else 0)) * args.steps
    ms_e2e, _, ranks_e2e = timed(step.step_e2e, args.steps, max(3, args.warmup // 2))
    pcie = pcie_probe()

    # BASELINE config 4 companion: the same step while the training loop's data-parallel gradient all-reduce
    # (a VGG16-DAF sized fp32 buffer, ~570 MB, NCCL over NVLink) is in flight.  It is NOT part of the
    # path (no data-path collective); reported beside `value`, never inside it.
    ms_ar = None
    if world > 1:
        grads = torch.zeros(142 * 1000 * 1000, dtype=torch.float32, device=dev)

        def step_with_allreduce():
            work = dist.all_reduce(grads, async_op=True)
            step.step()
            work.wait()
        ms_ar, _, _ = timed(step_with_allreduce, args.steps, 3)
        del grads
    if sampler is not None:
        sampler.terminate()
        sampler.wait()

    # per-kernel device time (CUDA events on the launch stream, inside the library), eager step
    _lib.profile_reset()
    _lib.profile(True)
    for _ in range(max(5, min(args.steps, 20))):
        flush.zero_()
        step.step_eager()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile(False)

    # cfg4 (8 images / GPU, MAF multi-level DA heads) measured after the headline, every N
    cfg4 = None
    if args.workload == "cfg2" and not args.no_cfg4:
        wl4 = Workload("cfg4")
        step.graphs = None  # release the cfg2 graphs' memory pool
        step4 = TlodStep(dev, seed=13 + rank, wl=wl4, use_graph=not args.no_graph)
        ms4, _, ranks4 = timed(step4.step, min(args.steps, 20), 3)
        ms4_ar = None
        if world > 1:
            grads = torch.zeros(142 * 1000 * 1000, dtype=torch.float32, device=dev)

            def step4_ar():
                work = dist.all_reduce(grads, async_op=True)
                step4.step()
                work.wait()
            ms4_ar, _, _ = timed(step4_ar, min(args.steps, 20), 3)
            del grads
        n4 = min(args.steps, 20)
        cfg4 = {"workload": wl4.text, "images_per_gpu": 8, "rois_per_step_per_gpu": wl4.rois_per_step,
                "value": world * wl4.rois_per_step * n4 / (ms4 * 1e-3), "unit": "RoIs/s", "ms_per_step": ms4 / n4,
                "per_rank_ms_per_step": [v / n4 for v in ranks4],
                "with_grad_allreduce": None if ms4_ar is None else {
                    "value": world * wl4.rois_per_step * n4 / (ms4_ar * 1e-3), "unit": "RoIs/s",
                    "ms_per_step": ms4_ar / n4, "allreduce_bytes": 142 * 1000 * 1000 * 4}}
        del step4

    # HBM-bound evidence and the other 8(d) metrics, rank 0 only
    align = secondary = None
    if rank == 0 and not args.no_cfg3:
        align = roi_align_sizes(dev)
        secondary = secondary_metrics(dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    kernels = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share": v[0] / total_ms}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}

    # algorithmic bytes per launch inside the step; src and tgt launches
    # alternate, so the per-launch mean is used
    def roi_bytes(n_img, n_roi, cells):
        return n_img * C * H * W * 4 + n_roi * 20 + n_roi * C * cells * 4
    r_src, r_tgt = wl.n_src * ROIS_SRC, wl.n_tgt * ROIS_TGT
    alg = {"roi_align_avg_fwd8_kernel": (roi_bytes(wl.n_src, r_src, 49) + roi_bytes(wl.n_tgt, r_tgt, 49)) / 2.0,
           "roi_align_bwd_rows_kernel": (roi_bytes(wl.n_src, r_src, 64) + roi_bytes(wl.n_tgt, r_tgt, 64)) / 2.0,
           "avgpool2x2_bwd_kernel": (r_src + r_tgt) / 2.0 * C * (64 + 49) * 4}
    traffic = load_traffic("bench_cfg2")
    in_step = {}
    for name, nbytes in alg.items():
        if name in kernels:
            t = kernels[name]["ms_per_launch"] * 1e-3
            in_step[name] = {"kernel": name, "bound": "hbm", "achieved": nbytes / t / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": nbytes / t / 1e9 / peak, "traffic": traffic.get(name),
                             "algorithmic_bytes_per_launch": nbytes, "share_of_step": kernels[name]["share"],
                             "l2_resident": True}
    dominant = next(iter(kernels)) if kernels else None
    # `roofline`: the step's dominant HBM-bound kernel, quoted on its HBM-sized launch (cfg3: 630 MB per
    # direction, larger than the 126 MB L2), measured live above with CUDA events; the same kernel inside the
    # cfg2 step (L2 resident, launch latency inside the event pair) is `roofline_in_step`
    roofline = None
    dom_roi = max(in_step.values(), key=lambda r: r["share_of_step"]) if in_step else None
    if align is not None and "cfg3" in align:
        key = "bwd" if (dom_roi and "bwd" in dom_roi["kernel"]) else "avg_fwd"
        c3 = align["cfg3"][key]
        nbytes = c3.get("algorithmic_bytes", c3.get("algorithmic_bytes_fused"))
        roofline = {"kernel": c3["kernel"], "bound": "hbm", "achieved": nbytes / (c3["us"] * 1e-6) / 1e9,
                    "peak": peak, "unit": "GB/s", "frac": nbytes / (c3["us"] * 1e-6) / 1e9 / peak,
                    "traffic": c3.get("traffic"), "algorithmic_bytes_per_launch": nbytes, "l2_resident": False,
                    "config": "cfg3 (8x1024x38x75, 2048 RoIs)", "peak_source": peak_src,
                    "note": "HBM-sized launch of the step's dominant RoI kernel, CUDA events around back-to-back "
                            "launches in this run; roofline_in_step has the same kernel inside the L2-resident cfg2 step"}
    elif dom_roi is not None:
        roofline = dict(dom_roi, peak_source=peak_src)
    h2d, d2h = step.e2e_bytes()
    n = args.steps
    line = {
        "metric": METRIC,
        "value": world * wl.rois_per_step * n / (ms_dev * 1e-3),
        "unit": "RoIs/s",
        "n_gpus": world, "steps": n, "warmup": args.warmup,
        "ms_per_step": ms_dev / n,
        "per_rank_ms_per_step": {"min": min(ranks_dev) / n, "median": statistics.median(ranks_dev) / n,
                                 "max": max(ranks_dev) / n, "ranks": [v / n for v in ranks_dev]},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.text, "rois_per_step_per_gpu": wl.rois_per_step,
                   "proposals_per_step_per_gpu": wl.proposals_per_step, "images_per_step_per_gpu": wl.n_src + wl.n_tgt,
                   "l2": "256 MB buffer zeroed between timed iterations (outside the event pairs)",
                   "parallelism": "image-sharded, %d rank(s), no data-path collective" % world},
        "proposals_per_s": world * wl.proposals_per_step * n / (ms_dev * 1e-3),
        "e2e": {"value": world * wl.rois_per_step * n / (ms_e2e * 1e-3), "unit": "RoIs/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / n,
                "per_rank_ms_per_step": {"min": min(ranks_e2e) / n, "median": statistics.median(ranks_e2e) / n,
                                         "max": max(ranks_e2e) / n, "ranks": [v / n for v in ranks_e2e]},
                "pcie": dict(pcie, copy_floor_ms_per_step=max(h2d / (pcie["h2d_GBps_per_rank_min"] * 1e6),
                                                               d2h / (pcie["d2h_GBps_per_rank_min"] * 1e6)),
                             note="64 MB pinned copies, all ranks at once; copy_floor = the larger of this rank's "
                                  "H2D and D2H bytes per step at the slowest rank's rate (the two directions overlap)"),
                "note": "H2D per step (one packed pinned arena per domain): feature maps, RPN softmax + deltas + "
                        "head logits, im_info, gt boxes, DA-head outputs.  D2H per step (one packed arena): rois, "
                        "sampled labels + regression targets, GRL'd feature gradients, DA + RPN losses, anchor "
                        "labels.  NOT copied: RoIAlign's pooled output and the detection-head gradient `top` (both "
                        "live only on the device in training), anchor bbox targets / weights (consumed by the RPN "
                        "loss on the device)"},
        "with_grad_allreduce": None if ms_ar is None else {
            "value": world * wl.rois_per_step * n / (ms_ar * 1e-3), "unit": "RoIs/s",
            "ms_per_step": ms_ar / n, "allreduce_bytes": 142 * 1000 * 1000 * 4,
            "note": "same step with a 568 MB fp32 NCCL all-reduce (the DP gradient exchange of the training "
                    "loop) overlapped; outside the path"},
        "cfg4": cfg4,
        "gpu_launches": int(launches),
        "cuda_graph": used_graph,
        "graph_check": graph_check,
        "clocks": parse_clocks(clock_file),
        "roofline": roofline, "roofline_in_step": in_step,
        "dominant_kernel": dominant, "kernels": kernels, "roi_align": align,
        "roi_align_cfg3": None if align is None else align.get("cfg3"), "secondary": secondary,
        "scaling_limiter": ("none on the device path (image-sharded, no collective); e2e is bounded by the host: "
                            "every rank moves %.1f MB H2D + %.1f MB D2H per step through the box's shared PCIe "
                            "root / host memory, and runs its RNG-exact host sampling on its own core slice%s"
                            % (h2d / 1e6, d2h / 1e6,
                               "" if cores_bound is None else " (%d cores per rank)" % cores_bound)),
        "wall_s": wall,
    }
    if not args.no_reference:
        ref = reference_subprocess(0 if args.no_cpu_baseline else 2)
        if "cpu_baseline" in ref:
            line["cpu_baseline"] = ref.pop("cpu_baseline")
        line["gpu_reference"] = ref.get("gpu_reference", ref)
        line["cpu_as_shipped"] = ref.get("cpu_as_shipped")
        line["speedup_vs_reference_kernels"] = speedups(ref, align, secondary)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's kernels + numpy host logic)
# ---------------------------------------------------------------------------
def cpu_step(inp, anchors, orc, half, wl):
    """One (half) step on the CPU.  half=True: 1 source + 1 target image (bounded sample)."""
    n_rois = 0
    for dom, key, per, pre, post in (("src", "TRAIN", ROIS_SRC, 12000, 2000), ("tgt", "TEST", ROIS_TGT, 6000, 300)):
        n = 1 if half else (wl.n_src if dom == "src" else wl.n_tgt)
        feat = inp[dom + "_feat"][:n].numpy()
        rois = orc.proposal_layer(inp[dom + "_prob"][:n].numpy(), inp[dom + "_deltas"][:n].numpy(),
                                  inp[dom + "_im_info"][:n].numpy(), anchors, 16, pre, post, 0.7)
        if dom == "src":
            orc.anchor_target_layer(H, W, inp["src_gt"][:n].numpy(), inp["src_im_info"][:n].numpy(), anchors, 16)
        sel = np.ascontiguousarray(rois[:, :per, :].reshape(-1, 5))
        out8 = orc.roi_align_forward(feat, sel, 8, 8, 1.0 / 16)
        pooled = torch.nn.functional.avg_pool2d(torch.from_numpy(out8), 2, 1)
        top8 = np.ascontiguousarray(np.random.RandomState(0).randn(*out8.shape).astype(np.float32))
        orc.roi_align_backward(top8, sel, feat.shape, 1.0 / 16)
        r = sel.shape[0]
        orc.da_losses(inp[dom + "_img_score"][:n].numpy(), inp[dom + "_ins_prob"][:r].numpy(), 1 if dom == "src" else 0)
        orc.da_losses_grad(inp[dom + "_img_score"][:n].numpy(), inp[dom + "_ins_prob"][:r].numpy(),
                           1 if dom == "src" else 0)
        _ = pooled.sum()
        n_rois += r
    return n_rois


def cpu_reference(steps, warmup, threads):
    from oracle import oracle as orc
    orc.build()
    orc.set_num_threads(threads)
    torch.set_num_threads(threads)
    wl = Workload("cfg2")
    anchors = orc.generate_anchors(scales=[4, 8, 16, 32], ratios=[0.5, 1, 2]).astype(np.float32)
    inp = synth_inputs(3, wl)
    np.random.seed(3)
    for _ in range(warmup):
        cpu_step(inp, anchors, orc, True, wl)
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        n += cpu_step(inp, anchors, orc, True, wl)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "RoIs/s", "cores": int(orc.num_threads()), "kind": "port",
            "sample": "%d x half step (1 src + 1 tgt image, %d RoIs): oracle C port of the reference kernels "
                      "(OpenMP over RoIs/images/channels) + numpy anchor targets; %.1f s" % (steps, n // max(steps, 1), dt),
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, min(args.warmup, 1)
    # bounded: ~1-2 s per half step -> cap the number of timed steps so the run ends in minutes
    steps_run = min(steps, 20)
    base = cpu_reference(steps=steps_run, warmup=warmup, threads=os.cpu_count())
    wl = Workload("cfg2")
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"], "unit": "RoIs/s", "n_gpus": world, "steps": steps_run, "warmup": warmup,
        "ms_per_step": base["seconds"] / steps_run * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.text, "sample_per_step": "half step: 1 src + 1 tgt image, 556 RoIs"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference's CUDA path cannot load on torch 2.x (torch.utils.ffi) and its CPU RoIAlign "
                "backward / nms_cpu are wrong (SURVEY.md 8c), so the CPU arm is the oracle port of its kernels; "
                "the reference's own recompiled CUDA kernels are timed in the product arm's `gpu_reference`",
    }
    emit(line)


_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints on fd 1
    (e.g. NCCL's version banner) has been diverted to stderr by main()."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tlod", choices=["tlod", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference", action="store_true", help="skip the reference-kernel subprocess")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the stand-alone roofline / secondary sections")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the cfg4 (8 images/GPU, MAF heads) companion")
    ap.add_argument("--no-graph", action="store_true", help="run the device part eagerly instead of CUDA graphs")
    ap.add_argument("--no-graph-check", action="store_true",
                    help="skip the graph-replay == eager-step comparison (profiling runs: keeps torch's compare "
                         "kernels out of the launch list)")
    ap.add_argument("--watchdog", type=int, default=1500,
                    help="seconds after which a stuck run dumps its Python stacks and exits (0 = off)")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True, file=sys.stderr)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_tlod(args)


if __name__ == "__main__":
    main()
