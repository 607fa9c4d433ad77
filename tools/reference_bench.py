#!/usr/bin/env python
"""The reference's OWN kernels timed on this box (bench.py runs this in a subprocess, so that the
product process never maps oracle/): BASELINE.md section 3's "kernel to beat".

  * GPU: lib/model/*/src/*.cu recompiled unmodified for sm_100a (oracle/_ref/libref_cuda.so, built by
    oracle/Makefile where /root/reference exists; the .so travels to the GPU box):
    ROIAlignForward/BackwardLaucher at cfg1 / cfg2 / cfg3, ROIPoolForward/BackwardLaucher,
    nms_cuda_compute at 12000 and 6000 boxes, BilinearSamplerBHWD at cfg3.
    The op-surface cost is timed: the caller-side zero fill the reference's Python does before a
    backward launch (functions/roi_align.py:42) is inside the timed region, as it is inside ours.
  * CPU as shipped: ROIAlignForwardCpu from lib/model/roi_align/src/roi_align.c
    (oracle/_ref/libref_cpu.so) at cfg1, one thread, and sharded over all host cores by RoI.
  * CPU port: the oracle (C restatement, OpenMP) on a bounded half step -- bench.py's `cpu_baseline`.

Prints ONE JSON object on stdout.  Usage: python tools/reference_bench.py [--cpu-steps N] [--no-gpu]
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools.synth import synth_rois  # noqa: E402

REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
REF_CPU = os.path.join(ROOT, "oracle", "_ref", "libref_cpu.so")
P, F32, I = ctypes.c_void_p, ctypes.c_float, ctypes.c_int


def _events(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def _wall(fn, iters, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


def sorted_boxes(n, seed):
    """n score-sorted boxes of proposal-like geometry on a 600 x 1200 image."""
    g = torch.Generator().manual_seed(seed)
    cx, cy = torch.rand(n, generator=g) * 1200, torch.rand(n, generator=g) * 600
    w, h = 16 + torch.rand(n, generator=g) * 400, 16 + torch.rand(n, generator=g) * 300
    sc = torch.sort(torch.rand(n, generator=g), descending=True).values
    return torch.stack([(cx - w / 2).clamp(0, 1199), (cy - h / 2).clamp(0, 599), (cx + w / 2).clamp(0, 1199),
                        (cy + h / 2).clamp(0, 599), sc], 1).contiguous()


def gpu_reference(dev):
    lib = ctypes.CDLL(REF_CUDA)
    lib.ROIAlignForwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, P, P, P]
    lib.ROIAlignBackwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, I, P, P, P]
    lib.ROIPoolForwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, P, P, P, P]
    lib.ROIPoolBackwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, I, P, P, P, P]
    lib.nms_cuda_compute.argtypes = [P, P, P, I, I, F32]
    lib.nms_cuda_compute.restype = None
    st = torch.cuda.current_stream(dev).cuda_stream
    out = {}
    g = torch.Generator().manual_seed(7)
    for name, (B, C, H, W, R) in (("cfg1", (1, 512, 37, 75, 128)), ("cfg2", (2, 512, 37, 75, 512)),
                                  ("cfg3", (8, 1024, 38, 75, 2048))):
        x = torch.relu(torch.randn(B, C, H, W, generator=g)).to(dev)
        rois = synth_rois(R, B, 41)
        rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(dev)
        top = torch.randn(R, C, 8, 8, device=dev)
        y = torch.empty(R, C, 8, 8, device=dev)
        gx = torch.empty_like(x)

        def fwd():
            y.zero_()  # functions/roi_align.py:22
            lib.ROIAlignForwardLaucher(x.data_ptr(), 1 / 16, R, H, W, C, 8, 8, rois.data_ptr(), y.data_ptr(), st)

        def bwd():
            gx.zero_()  # functions/roi_align.py:42
            lib.ROIAlignBackwardLaucher(top.data_ptr(), 1 / 16, B, R, H, W, C, 8, 8, rois.data_ptr(),
                                        gx.data_ptr(), st)
        it = 20 if name != "cfg3" else 5
        out["roi_align_" + name] = {"fwd_us": _events(fwd, it), "bwd_us": _events(bwd, it),
                                    "shape": [B, C, H, W, R], "aligned": 8}
        if name == "cfg2":
            o7 = torch.empty(R, C, 7, 7, device=dev)
            a7 = torch.empty(R, C, 7, 7, dtype=torch.int32, device=dev)
            t7 = torch.randn(R, C, 7, 7, device=dev)

            def pfwd():
                lib.ROIPoolForwardLaucher(x.data_ptr(), 1 / 16, R, H, W, C, 7, 7, rois.data_ptr(), o7.data_ptr(),
                                          a7.data_ptr(), st)

            def pbwd():
                gx.zero_()
                lib.ROIPoolBackwardLaucher(t7.data_ptr(), 1 / 16, B, R, H, W, C, 7, 7, rois.data_ptr(),
                                           gx.data_ptr(), a7.data_ptr(), st)
            out["roi_pool_cfg2"] = {"fwd_us": _events(pfwd, 20), "bwd_us": _events(pbwd, 5),
                                    "shape": [B, C, H, W, R], "pooled": 7}
        if name == "cfg3" and hasattr(lib, "BilinearSamplerBHWD_updateOutput_cuda_kernel"):
            from model.utils.net_utils import _affine_grid_gen
            G = 14
            grid_xy = _affine_grid_gen(rois, (H, W), G)
            gyx = torch.stack([grid_xy[..., 1], grid_xy[..., 0]], 3).contiguous()
            oc = torch.empty(R, C, G, G, device=dev)
            tc = torch.randn(R, C, G, G, device=dev)
            gg = torch.zeros_like(gyx)
            fu, bu = lib.BilinearSamplerBHWD_updateOutput_cuda_kernel, lib.BilinearSamplerBHWD_updateGradInput_cuda_kernel

            def cfwd():
                fu(I(C), I(G), I(G), I(R), I(C), I(H), I(W), I(B), P(x.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
                   P(gyx.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
                   P(oc.data_ptr()), I(C * G * G), I(G * G), I(G), I(1), P(st))

            def cbwd():
                gx.zero_()  # functions/roi_crop.py:16-17
                bu(I(C), I(G), I(G), I(R), I(C), I(H), I(W), I(B), P(x.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
                   P(gyx.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
                   P(gx.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
                   P(gg.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
                   P(tc.data_ptr()), I(C * G * G), I(G * G), I(G), I(1), P(st))
            out["roi_crop_cfg3"] = {"fwd_us": _events(cfwd, 3), "bwd_us": _events(cbwd, 3),
                                    "shape": [B, C, H, W, R], "grid": G}
            del oc, tc
        del x, top, y, gx
        torch.cuda.empty_cache()
    for n in (12000, 6000):
        dets = sorted_boxes(n, 500 + n).to(dev)
        keep = torch.zeros(n, dtype=torch.int32, device=dev)
        num = torch.zeros(1, dtype=torch.int32, device=dev)
        # host-synchronous (blocking memcpy + host scan + cudaMalloc/cudaFree inside): wall clock
        us = _wall(lambda: lib.nms_cuda_compute(keep.data_ptr(), num.data_ptr(), dets.data_ptr(), n, 5, 0.7), 10)
        out["nms_%d" % n] = {"us": us, "kept": int(num.item()), "thresh": 0.7,
                             "note": "wall clock: the call blocks the host (18 MB mask D2H + host scan)"}
    return out


def cpu_as_shipped():
    """ROIAlignForwardCpu (roi_align.c:80-136) at cfg1: one thread as shipped, then sharded by RoI."""
    import multiprocessing as mp
    B, C, H, W, R = 1, 512, 37, 75, 128
    g = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn(B, C, H, W, generator=g)).numpy()
    rois = synth_rois(R, B, 41).numpy()
    t1 = _cpu_roi_align(x, rois)
    cores = os.cpu_count() or 1
    shards = [rois[i::cores] for i in range(cores) if len(rois[i::cores])]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(shards)) as pool:  # forked workers share x; each times its own call
        tn = max(pool.starmap(_cpu_roi_align, [(x, s) for s in shards]))
    return {"roi_align_fwd_cfg1_1thread_ms": t1 * 1e3, "rois_per_s_1thread": R / t1,
            "roi_align_fwd_cfg1_sharded_ms": tn * 1e3, "rois_per_s_sharded": R / tn, "cores": cores,
            "note": "reference roi_align.c compiled unmodified (gcc -O2); one process per core, RoIs sharded, time = slowest "
                    "shard (process start-up excluded); "
                    "its ROIAlignBackwardCpu is wrong (roi_align.c:175) and is not timed"}


def _cpu_roi_align(x, rois):
    lib = ctypes.CDLL(REF_CPU)
    lib.ROIAlignForwardCpu.argtypes = [P, F32, I, I, I, I, I, I, P, P]
    lib.ROIAlignForwardCpu.restype = None
    out = np.zeros((rois.shape[0], x.shape[1], 8, 8), np.float32)
    rois = np.ascontiguousarray(rois)
    t0 = time.perf_counter()
    lib.ROIAlignForwardCpu(x.ctypes.data, 1 / 16, rois.shape[0], x.shape[2], x.shape[3], x.shape[1], 8, 8,
                           rois.ctypes.data, out.ctypes.data)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed half steps of the oracle port (0 = skip)")
    args = ap.parse_args()
    out = {}
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not args.no_gpu:
        if not os.path.exists(REF_CUDA):
            out["gpu_reference"] = {"unavailable": "oracle/_ref/libref_cuda.so not shipped (built where "
                                                   "/root/reference exists: make -C oracle ref)"}
        elif not torch.cuda.is_available():
            out["gpu_reference"] = {"unavailable": "no CUDA device"}
        else:
            out["gpu_reference"] = gpu_reference(torch.device("cuda:0"))
    if os.path.exists(REF_CPU):
        out["cpu_as_shipped"] = cpu_as_shipped()
    if args.cpu_steps > 0:
        import bench
        out["cpu_baseline"] = bench.cpu_reference(steps=args.cpu_steps, warmup=0, threads=os.cpu_count())
    real_stdout.write(json.dumps(out) + "\n")
    real_stdout.flush()


if __name__ == "__main__":
    main()
