#!/usr/bin/env python
"""Benchmark of the RoI / proposal hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path (oracle port)

Workload (BASELINE.json configs[1]), per GPU: one DAF-style
training step's operator sequence on 2 source + 2 target synthetic 600x1200 images
(VGG16 conv5 maps 512x37x75, A = 12 anchors):
  source: proposal layer TRAIN (12000 -> 2000, NMS 0.7), anchor targets, RoIAlignAvg 7x7
          forward + backward on 256 RoIs/image, image/instance DA losses fwd+bwd, GRL;
  target: proposal layer TEST (6000 -> 300), RoIAlignAvg forward + backward on 300
          RoIs/image, DA losses fwd+bwd, GRL.
`value` = RoIs pooled forward+backward per second over the whole step, all GPUs
(weak scaling: every rank runs its own 4 images; there is no collective on the path).
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
N_SRC, N_TGT = 2, 2
ROIS_SRC, ROIS_TGT = 256, 300
ROIS_PER_STEP = N_SRC * ROIS_SRC + N_TGT * ROIS_TGT
PROPOSALS_PER_STEP = N_SRC * 2000 + N_TGT * 300
WORKLOAD = ("cfg2: DAF VGG16 step ops, 2 src + 2 tgt 600x1200 images/GPU, conv5 512x37x75, "
            "RPN 12000->2000 (src) / 6000->300 (tgt) NMS 0.7, RoIAlignAvg 7x7 fwd+bwd on 256/300 RoIs per image, "
            "anchor targets, image+instance DA losses, GRL")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_inputs(seed, pin=False):
    """Host tensors of one rank's step (SURVEY.md 8d generators)."""
    from tools.synth import synth_gt, synth_rpn
    g = torch.Generator().manual_seed(seed)
    d = {}
    for dom, n in (("src", N_SRC), ("tgt", N_TGT)):
        d[dom + "_feat"] = torch.relu(torch.randn(n, C, H, W, generator=g))
        prob, deltas = synth_rpn(n, A, H, W, seed + (1 if dom == "src" else 2))
        d[dom + "_prob"], d[dom + "_deltas"] = prob, deltas
        d[dom + "_im_info"] = torch.tensor([[600.0, 1200.0, 0.5859375]] * n)
        d[dom + "_img_score"] = torch.randn(n, 2, H, W, generator=g)
        r = n * (ROIS_SRC if dom == "src" else ROIS_TGT)
        d[dom + "_ins_prob"] = torch.sigmoid(torch.randn(r, 1, generator=g))
    d["src_gt"] = synth_gt(N_SRC, 20, 50, seed + 50)
    if pin:
        d = {k: v.pin_memory() for k, v in d.items()}
    return d


# ---------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------
class TlodStep(object):
    """One rank's step.  The device-only part (proposal layers, RoIAlignAvg forward + backward, DA
    losses, GRL for both domains: ~45 launches, static shapes, no host synchronisation anywhere)
    is captured once into a CUDA graph and replayed; the anchor-target layer follows eagerly
    because its subsampling consumes numpy's host RNG exactly like the reference."""

    def __init__(self, dev, seed, use_graph=True):
        from model.roi_align.modules.roi_align import RoIAlignAvg
        from model.rpn.anchor_target_layer import _AnchorTargetLayer
        from model.rpn.proposal_layer import _ProposalLayer
        import tlod_b200
        self.tlod = tlod_b200
        self.dev = dev
        self.host = synth_inputs(seed, pin=True)
        self.d = {k: v.to(dev) for k, v in self.host.items()}
        self.proposal = _ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
        self.anchor_target = _AnchorTargetLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
        self.roi_align = RoIAlignAvg(7, 7, 1.0 / 16.0)
        g = torch.Generator().manual_seed(seed + 99)
        # gradient arriving from the detection head: device resident (it never exists on the host)
        self.top = {"src": torch.randn(N_SRC * ROIS_SRC, C, 7, 7, generator=g).to(dev),
                    "tgt": torch.randn(N_TGT * ROIS_TGT, C, 7, 7, generator=g).to(dev)}
        self.num_boxes = torch.full((N_SRC,), 20, dtype=torch.long)
        np.random.seed(3)
        self.graph = None
        self.graphs = None
        self.static = None
        self.streams = None
        self.launches_per_replay = 0
        self.host_out = None
        self.side = None
        if use_graph:
            self.capture()

    def domain(self, dom, d, results):
        key = "TRAIN" if dom == "src" else "TEST"
        per = ROIS_SRC if dom == "src" else ROIS_TGT
        feat = d[dom + "_feat"].requires_grad_(True)
        rois = self.proposal((d[dom + "_prob"], d[dom + "_deltas"], d[dom + "_im_info"], key))
        sel = rois[:, :per, :].reshape(-1, 5)  # stand-in for _ProposalTargetLayer's sampling
        pooled = self.roi_align(feat, sel)
        score = d[dom + "_img_score"].requires_grad_(True)
        prob = d[dom + "_ins_prob"].requires_grad_(True)
        img, ins, cst = self.tlod.da_losses(score, prob, 1 if dom == "src" else 0)
        loss = 0.1 * (img + ins + cst)
        torch.autograd.backward([pooled, loss], [self.top[dom], None])
        # GRL in front of the DA heads: base_feat gradient and pooled-feature gradient
        g_feat = self.tlod.functional.grl_backward(feat.grad, 0.1)
        results[dom] = (rois, feat.grad, g_feat, torch.stack([img, ins, cst]).detach(), score.grad, prob.grad)
        feat.grad = None
        score.grad = None
        prob.grad = None

    def device_part(self, d, copy_in=None, copy_out=None):
        """Source and target domain are independent until the losses are summed: each runs on its
        own stream (one CUDA graph per domain when captured), so the latency-bound kernels of one
        domain can fill SMs the other leaves idle.  copy_in / copy_out (e2e mode) put each
        domain's H2D / D2H copies on the same stream, overlapping the other domain's compute."""
        results = {}
        cur = torch.cuda.current_stream(self.dev)
        if self.streams is None:
            self.streams = {"src": torch.cuda.Stream(self.dev), "tgt": torch.cuda.Stream(self.dev)}
        for dom in ("src", "tgt"):
            st = self.streams[dom]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                if copy_in is not None:
                    copy_in(dom)
                if self.graphs is not None and d is self.d:
                    self.graphs[dom].replay()
                    results[dom] = self.static[dom]
                else:
                    self.domain(dom, d, results)
                if copy_out is not None:
                    copy_out(dom, results)
        for dom in ("src", "tgt"):
            cur.wait_stream(self.streams[dom])
        return results

    def anchor_part(self, d, results):
        results["anchor_targets"] = self.anchor_target((d["src_prob"], d["src_gt"], d["src_im_info"],
                                                        self.num_boxes))
        return results

    def capture(self):
        """Warm up on a side stream, then capture one CUDA graph per domain over the static inputs."""
        try:
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    r = {}
                    self.domain("src", self.d, r)
                    self.domain("tgt", self.d, r)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            graphs, static = {}, {}
            n0 = self.tlod.launch_count()
            for dom in ("src", "tgt"):
                graphs[dom] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graphs[dom]):
                    self.domain(dom, self.d, static)
            self.launches_per_replay = int(self.tlod.launch_count() - n0)
            for dom in ("src", "tgt"):
                graphs[dom].replay()
            torch.cuda.synchronize(self.dev)
            self.graphs, self.static = graphs, static
            self.graph = graphs
        except Exception as e:  # noqa: BLE001 -- eager is always available
            sys.stderr.write("bench.py: CUDA graph capture failed (%s); running eagerly\n" % (e,))
            self.graph = self.graphs = self.static = None

    def step(self, d=None, copy_in=None, copy_out=None):
        d = self.d if d is None else d
        # Queue the two domains first (two graph launches), then start the anchor-target layer on a
        # side stream that only depends on the step's inputs: its label kernels run beside the
        # domain graphs, and its host-side subsampling (numpy's RNG stream, like the reference)
        # runs while the GPU works through the queue.
        cur = torch.cuda.current_stream(self.dev)
        if self.side is None:
            self.side = torch.cuda.Stream(self.dev)
        self.side.wait_stream(cur)
        results = self.device_part(d, copy_in, copy_out)
        with torch.cuda.stream(self.side):
            pending = self.anchor_target.begin((d["src_prob"], d["src_gt"], d["src_im_info"], self.num_boxes))
        results["anchor_targets"] = self.anchor_target.finish(pending)
        return results

    def step_e2e(self):
        """Same step through host buffers: H2D of the inputs from pinned memory into the static
        device buffers, the step, D2H of the results into pinned host buffers; each domain's
        copies ride on that domain's stream."""
        d = self.d

        def copy_in(dom):
            for k, v in self.host.items():
                if k.startswith(dom) and k not in ("src_gt", "src_im_info"):
                    d[k].detach().copy_(v, non_blocking=True)

        def copy_out(dom, results):
            rois, gfeat, _, losses, _, _ = results[dom]
            outs = [rois, gfeat, losses]
            if self.host_out is None:
                self.host_out = {}
            if dom not in self.host_out:
                self.host_out[dom] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
            for h, o in zip(self.host_out[dom], outs):
                h.copy_(o, non_blocking=True)

        d["src_gt"].copy_(self.host["src_gt"], non_blocking=True)  # the anchor-target layer reads these
        d["src_im_info"].copy_(self.host["src_im_info"], non_blocking=True)
        d["src_prob"].detach().copy_(self.host["src_prob"], non_blocking=True)
        r = self.step(None, copy_in, copy_out)
        at = r["anchor_targets"]
        if "at" not in self.host_out:
            self.host_out["at"] = torch.empty(at[0].shape, dtype=at[0].dtype).pin_memory()
        self.host_out["at"].copy_(at[0], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.host_out

    def e2e_bytes(self):
        h2d = sum(v.numel() * v.element_size() for v in self.host.values())
        d2h = 0
        for dom, n, per in (("src", N_SRC, 2000), ("tgt", N_TGT, 300)):
            d2h += n * per * 5 * 4 + n * C * H * W * 4 + 3 * 4
        d2h += N_SRC * A * H * W * 4
        return h2d, d2h


def clocks_sampler(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100",
                                 "-i", "0"], stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


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


def roi_align_cfg3(dev, iters=20):
    """RoIAlign 8x8 forward and backward at BASELINE cfg3 (ResNet-101 conv4, batch 8, 256 RoIs/image):
    630 MB of algorithmic traffic per direction, larger than L2, so no flush is needed."""
    from tools.synth import synth_rois
    from tlod_b200 import functional as F
    peak, peak_src = peaks()
    B, Cc, Hh, Ww, R = 8, 1024, 38, 75, 2048
    g = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn(B, Cc, Hh, Ww, generator=g)).to(dev)
    rois = synth_rois(R, B, 41).to(dev)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous()  # (B, 256, 5).view(-1, 5) order
    top = torch.randn(R, Cc, 8, 8, device=dev)
    alg = B * Cc * Hh * Ww * 4 + R * 20 + R * Cc * 64 * 4
    plan = F.roi_align_plan(rois, x.shape, 8, 8, 1.0 / 16)
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get("cfg3", {})
    except Exception:  # noqa: BLE001
        pass
    out = {}
    for name, kern, fn in (("plan", None, lambda: F.roi_align_plan(rois, x.shape, 8, 8, 1.0 / 16)),
                           ("fwd", "roi_align_fwd_planes_kernel",
                            lambda: F.roi_align_forward(x, rois, 8, 8, 1.0 / 16, plan=plan)),
                           ("bwd", "roi_align_bwd_rows_kernel",
                            lambda: F.roi_align_backward(top, rois, x.shape, 1.0 / 16, plan=plan))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        if kern is None:
            out[name] = {"ms": ms}
        else:
            out[name] = {"ms": ms, "achieved_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peak,
                         "traffic": traffic.get(kern)}
    out["algorithmic_bytes_per_direction"] = alg
    total_ms = out["plan"]["ms"] + out["fwd"]["ms"] + out["bwd"]["ms"]
    out["rois_per_s_fwd_bwd"] = R / (total_ms * 1e-3)
    out["frac_fwd_bwd"] = 2 * alg / (total_ms * 1e6) / peak
    out["peak_GBps"] = peak
    out["peak_source"] = peak_src
    out["note"] = ("back-to-back launches, CUDA events around the loop; each call includes its torch.empty "
                   "(caching allocator); the plan is built once per rois tensor and shared by fwd and bwd")
    return out


def secondary_metrics(dev, iters=20):
    """The other metrics SURVEY 8(d) names, measured alone on rank 0: the two proposal pipelines
    (12000 -> 2000 TRAIN, 6000 -> 300 TEST; proposal sets/s, proposals/s, fraction of the bound
    8(d) defines: max(bytes / HBM rate, 16 flop per upper-triangle pair / fp32 SIMT rate)), the
    cfg5 NMS sweep (batch x IoU threshold) and RoIPool forward + backward."""
    from tools.synth import synth_rois, synth_rpn
    from model.rpn.generate_anchors import generate_anchors
    from tlod_b200 import functional as F
    peak, _ = peaks()
    fp32_tflops = 148 * 128 * 2 * 1.965e9 / 1e12  # nominal SIMT rate

    def per_call_us(fn):
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
        us = per_call_us(lambda: F.proposals(prob, deltas, info, anchors, 16, pre, post, 0.7))
        nbytes = 2 * (A * H * W * 4 * 5 + post * 20)
        flops = 2 * (pre * (pre - 1) / 2) * 16
        bound_us = max(nbytes / (peak * 1e9), flops / (fp32_tflops * 1e12)) * 1e6
        out["proposal_layer"][name] = {"images": 2, "us_per_call": us, "proposal_sets_per_s": 2 / (us * 1e-6),
                                       "proposals_per_s": 2 * post / (us * 1e-6), "bound_us": bound_us,
                                       "frac_of_bound": bound_us / us,
                                       "note": "bound = full upper-triangle IoU work on the nominal fp32 SIMT "
                                               "rate; the phased NMS computes only the triangle the scan reaches"}
    for batch in (1, 8, 64):
        prob, deltas, info = rpn(batch)
        for thr in (0.3, 0.5, 0.7):
            us = per_call_us(lambda: F.proposals(prob, deltas, info, anchors, 16, 6000, 300, thr))
            out["nms_sweep_test_6000_300"].append({"batch": batch, "iou": thr, "us_per_call": us,
                                                   "proposal_sets_per_s": batch / (us * 1e-6)})
    # RoIPool 7x7 forward + backward at cfg1 / cfg2 scale (VGG16 conv5, 2 images, 512 RoIs)
    g = torch.Generator().manual_seed(5)
    feat = torch.relu(torch.randn(2, C, H, W, generator=g)).to(dev)
    rois = synth_rois(512, 2, 41).to(dev)
    top = torch.randn(512, C, 7, 7, device=dev)
    pooled, argmax = F.roi_pool_forward(feat, rois, 7, 7, 1.0 / 16)
    us_f = per_call_us(lambda: F.roi_pool_forward(feat, rois, 7, 7, 1.0 / 16))
    us_b = per_call_us(lambda: F.roi_pool_backward(top, argmax, rois, feat.shape, 1.0 / 16))
    alg_f = feat.numel() * 4 + 512 * 20 + 512 * C * 49 * 8
    out["roi_pool_2x512x37x75_512rois"] = {
        "fwd_us": us_f, "bwd_us": us_b, "rois_per_s_fwd_bwd": 512 / ((us_f + us_b) * 1e-6),
        "frac_of_hbm_peak_fwd": alg_f / (us_f * 1e-6) / 1e9 / peak,
        "frac_of_hbm_peak_bwd": alg_f / (us_b * 1e-6) / 1e9 / peak,
        "note": "L2-warm back-to-back launches (working set 63 MB < L2)"}
    # BASELINE cfg3 (ii): the crop path -- RoICrop 14x14 on conv4 (8, 1024, 38, 75), 2048 RoIs
    from model.utils.net_utils import _affine_grid_gen
    Bc, Cc, Hc, Wc, Rc, G = 8, 1024, 38, 75, 2048, 14
    featc = torch.relu(torch.randn(Bc, Cc, Hc, Wc, generator=g)).to(dev)
    roisc = synth_rois(Rc, Bc, 41)
    roisc = roisc[torch.argsort(roisc[:, 0], stable=True)].contiguous().to(dev)
    grid_xy = _affine_grid_gen(roisc, (Hc, Wc), G)
    grid_yx = torch.stack([grid_xy[..., 1], grid_xy[..., 0]], 3).contiguous()
    topc = torch.randn(Rc, Cc, G, G, device=dev)
    us_cf = per_call_us(lambda: F.roi_crop_forward(featc, grid_yx))
    us_cb = per_call_us(lambda: F.roi_crop_backward(topc, grid_yx, featc.shape))
    alg_c = featc.numel() * 4 + grid_yx.numel() * 4 + Rc * Cc * G * G * 4
    out["roi_crop_cfg3_14x14"] = {
        "fwd_us": us_cf, "bwd_us": us_cb, "rois_per_s_fwd_bwd": Rc / ((us_cf + us_cb) * 1e-6),
        "algorithmic_bytes_per_direction": alg_c,
        "frac_of_hbm_peak_fwd": alg_c / (us_cf * 1e-6) / 1e9 / peak,
        "frac_of_hbm_peak_bwd": alg_c / (us_cb * 1e-6) / 1e9 / peak}
    return out


def run_tlod(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import tlod_b200
    from tlod_b200 import _lib

    step = TlodStep(dev, seed=3 + rank, use_graph=not args.no_graph)
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
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    clock_file = os.path.join(ROOT, "gpurun_out", "bench_clocks_rank0.csv")
    os.makedirs(os.path.dirname(clock_file), exist_ok=True)
    sampler = clocks_sampler(clock_file) if rank == 0 else None
    launches0 = tlod_b200.launch_count()
    ms_dev, wall = timed(step.step, args.steps, args.warmup)
    # kernels of this library inside the timed region: eager launches are counted by the library;
    # a graph replay re-issues the launches counted once at capture time
    eager_per_step = (tlod_b200.launch_count() - launches0) // (args.steps + args.warmup)
    launches = (eager_per_step + (step.launches_per_replay if step.graph is not None else 0)) * args.steps
    ms_e2e, _ = timed(step.step_e2e, args.steps, max(3, args.warmup // 2))

    # cfg4 companion number (N > 1 only): the same step while the training loop's data-parallel
    # gradient all-reduce (a VGG16-DAF sized fp32 buffer, ~570 MB, NCCL over NVLink) is in flight.
    # It is NOT part of the path (no data-path collective); reported beside `value`, never inside it.
    ms_ar = None
    if world > 1:
        grads = torch.zeros(142 * 1000 * 1000, dtype=torch.float32, device=dev)

        def step_with_allreduce():
            work = dist.all_reduce(grads, async_op=True)
            step.step()
            work.wait()
        ms_ar, _ = timed(step_with_allreduce, args.steps, 3)
        del grads
    if sampler is not None:
        sampler.terminate()
        sampler.wait()

    # per-kernel device time (CUDA events on the launch stream, inside the library)
    _lib.profile_reset()
    _lib.profile(True)
    graphs, step.graphs = step.graphs, None  # eager: the library brackets each launch with events
    for _ in range(max(5, min(args.steps, 20))):
        flush.zero_()
        step.anchor_part(step.d, step.device_part(step.d))
    step.graphs = graphs
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile(False)

    # HBM-bound evidence: RoIAlign forward / backward alone at cfg3 scale (8x1024x38x75, 2048 RoIs,
    # 630 MB of algorithmic traffic per direction: larger than L2, so no flush is needed)
    cfg3 = None
    secondary = None
    if rank == 0 and not args.no_cfg3:
        cfg3 = roi_align_cfg3(dev)
        secondary = secondary_metrics(dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    total_ms = sum(v[0] for v in prof.values()) or 1.0
    kernels = {k: {"ms_per_launch": v[0] / v[1], "launches": v[1], "share": v[0] / total_ms}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    # algorithmic bytes per launch at this workload.  src and tgt
    # launches alternate, so the per-launch mean is used.
    #   RoIAlign fwd / bwd: feature (gradient) map once + rois + the (R, C, 8, 8) tensor once
    #   avg-pool fwd / bwd: the (R, C, 8, 8) and the (R, C, 7, 7) tensor once each
    def roi_bytes(n_img, n_roi):
        return n_img * C * H * W * 4 + n_roi * 20 + n_roi * C * 64 * 4

    def pool_bytes(n_roi):
        return n_roi * C * (64 + 49) * 4
    r_src, r_tgt = N_SRC * ROIS_SRC, N_TGT * ROIS_TGT
    alg = {"roi_align_fwd_planes_kernel": (roi_bytes(N_SRC, r_src) + roi_bytes(N_TGT, r_tgt)) / 2.0,
           "roi_align_bwd_rows_kernel": (roi_bytes(N_SRC, r_src) + roi_bytes(N_TGT, r_tgt)) / 2.0,
           "avgpool2x2_fwd_kernel": (pool_bytes(r_src) + pool_bytes(r_tgt)) / 2.0,
           "avgpool2x2_bwd_kernel": (pool_bytes(r_src) + pool_bytes(r_tgt)) / 2.0}
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get("bench_cfg2", {})
    except Exception:  # noqa: BLE001
        pass
    by_kernel = {}
    for name, nbytes in alg.items():
        if name in kernels:
            t = kernels[name]["ms_per_launch"] * 1e-3
            by_kernel[name] = {"kernel": name, "bound": "hbm", "achieved": nbytes / t / 1e9, "peak": peak,
                               "unit": "GB/s", "frac": nbytes / t / 1e9 / peak, "traffic": traffic.get(name),
                               "algorithmic_bytes_per_launch": nbytes, "share_of_step": kernels[name]["share"]}
    dominant = next(iter(kernels)) if kernels else None
    # the roofline object describes the dominant kernel of the step (largest share of kernel time)
    roofline = by_kernel.get(dominant)
    if roofline is None and by_kernel:
        roofline = max(by_kernel.values(), key=lambda r: r["share_of_step"])
    if roofline is not None:
        roofline = dict(roofline, peak_source=peak_src,
                        note="cfg2 working set (~80 MB per launch) fits the 126 MB L2; L2 is flushed between steps; "
                             "launch time measured with CUDA events around each launch (includes launch latency); "
                             "the HBM-sized measurement is roi_align_cfg3")
    h2d, d2h = step.e2e_bytes()
    line = {
        "metric": "RoIs/sec RoIAlign fwd+bwd & proposals/sec (12000->2000 NMS) vs HBM roofline",
        "value": world * ROIS_PER_STEP * args.steps / (ms_dev * 1e-3),
        "unit": "RoIs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rois_per_step_per_gpu": ROIS_PER_STEP,
                   "proposals_per_step_per_gpu": PROPOSALS_PER_STEP, "images_per_step_per_gpu": N_SRC + N_TGT,
                   "l2": "256 MB buffer zeroed between timed iterations (outside the event pairs)",
                   "parallelism": "image-sharded, %d rank(s), no data-path collective" % world},
        "proposals_per_s": world * PROPOSALS_PER_STEP * args.steps / (ms_dev * 1e-3),
        "e2e": {"value": world * ROIS_PER_STEP * args.steps / (ms_e2e * 1e-3), "unit": "RoIs/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "with_grad_allreduce": None if ms_ar is None else {
            "value": world * ROIS_PER_STEP * args.steps / (ms_ar * 1e-3), "unit": "RoIs/s",
            "ms_per_step": ms_ar / args.steps, "allreduce_bytes": 142 * 1000 * 1000 * 4,
            "note": "same step with a 568 MB fp32 NCCL all-reduce (the DP gradient exchange of the training "
                    "loop) overlapped; outside the path, reported for BASELINE config 4"},
        "gpu_launches": int(launches),
        "cuda_graph": step.graph is not None,
        "clocks": parse_clocks(clock_file),
        "roofline": roofline, "roofline_by_kernel": by_kernel,
        "dominant_kernel": dominant, "kernels": kernels, "roi_align_cfg3": cfg3, "secondary": secondary,
        "wall_s": wall,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(steps=2, warmup=0, threads=os.cpu_count())
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's kernels + numpy host logic)
# ---------------------------------------------------------------------------
def cpu_step(inp, anchors, orc, half):
    """One (half) step on the CPU.  half=True: 1 source + 1 target image (bounded sample)."""
    n_rois = 0
    for dom, key, per, pre, post in (("src", "TRAIN", ROIS_SRC, 12000, 2000), ("tgt", "TEST", ROIS_TGT, 6000, 300)):
        n = 1 if half else (N_SRC if dom == "src" else N_TGT)
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
    anchors = orc.generate_anchors(scales=[4, 8, 16, 32], ratios=[0.5, 1, 2]).astype(np.float32)
    inp = synth_inputs(3)
    np.random.seed(3)
    for _ in range(warmup):
        cpu_step(inp, anchors, orc, half=True)
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        n += cpu_step(inp, anchors, orc, half=True)
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
    line = {
        "impl": "reference",
        "metric": "RoIs/sec RoIAlign fwd+bwd & proposals/sec (12000->2000 NMS) vs HBM roofline",
        "value": base["value"], "unit": "RoIs/s", "n_gpus": world, "steps": steps_run, "warmup": warmup,
        "ms_per_step": base["seconds"] / steps_run * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_per_step": "half step: 1 src + 1 tgt image, 556 RoIs"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference's CUDA path cannot load on torch 2.x (torch.utils.ffi) and its CPU RoIAlign "
                "backward / nms_cpu are wrong (SURVEY.md 8c), so the CPU arm is the oracle port of its kernels",
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg3-scale RoIAlign roofline section")
    ap.add_argument("--no-graph", action="store_true", help="run the device part eagerly instead of a CUDA graph")
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
