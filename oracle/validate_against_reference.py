#!/usr/bin/env python
"""Pin the oracle against the reference's OWN code and write tests/golden/*.npz.

Runs only in the build container (needs /root/reference).  It

1. calls the reference's C ``ROIAlignForwardCpu`` (lib/model/roi_align/src/roi_align.c:80,
   compiled unmodified into oracle/_ref/libref_cpu.so) and requires bit-equality
   with ``oracle.roi_align_forward``;
2. imports the reference's Python RPN code (lib/model/rpn/{generate_anchors,
   bbox_transform,proposal_layer,anchor_target_layer}.py) with an ``easydict`` shim
   and requires bit-equality of ``generate_anchors`` (incl. the MATLAB known-answer
   table at generate_anchors.py:19-37), ``bbox_transform_inv``+``clip_boxes``,
   ``bbox_overlaps_batch``, ``bbox_transform_batch``, ``_ProposalLayer`` (TRAIN and
   TEST keys) and ``_AnchorTargetLayer`` outputs.
   The reference's ``nms`` needs either its cffi CUDA extension (unloadable on
   torch 2.x) or the buggy ``nms_cpu``, so ``_ProposalLayer`` is
   run with ``nms`` replaced by ``oracle.nms``; everything else in that layer is
   the reference's code.  The NMS / RoIPool / RoIAlign-backward restatements are
   pinned on the GPU box against the reference CUDA kernels recompiled unmodified
   (oracle/_ref/libref_cuda*.so; tests/test_gpu_reference_cuda.py).
3. runs the reference's ``_ProposalTargetLayer`` (proposal_target_layer_cascade.py, with a
   one-line ``Tensor.index`` shim for torch 2.x) under the same numpy seed and requires equality
   of the sampled rois, labels, weights (bit-exact) and regression targets (log() within 2e-6);
4. saves small golden vectors (inputs are regenerated from the seed in the tests;
   outputs are stored) under tests/golden/.

Usage: python oracle/validate_against_reference.py [--no-write]
"""
import argparse
import ctypes
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

REF_LIB = "/root/reference/lib"
GOLD = os.path.join(ROOT, "tests", "golden")

c_float_p = ctypes.POINTER(ctypes.c_float)


from oracle.synth import synth_gt, synth_rois, synth_rpn  # noqa: E402


def check(name, a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    ok = a.shape == b.shape and np.array_equal(a.view(np.uint8) if a.dtype.kind == "f" else a,
                                               b.view(np.uint8) if b.dtype.kind == "f" else b)
    if not ok and a.shape == b.shape and a.dtype.kind == "f":
        # +0.0 / -0.0 are the only permitted byte differences
        ok = np.array_equal(a, b)
    print(("PASS " if ok else "FAIL ") + name, a.shape)
    if not ok:
        if a.shape == b.shape:
            d = np.nonzero(a != b)
            print("   first mismatches:", [tuple(int(x[i]) for x in d) for i in range(min(5, len(d[0])))])
        raise SystemExit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-write", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    orc.build()

    # ---------------- 1. reference C RoIAlign forward --------------------
    ref_cpu = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_cpu.so"))
    gold_roi = {}
    for tag, (B, C, H, W, R, AH, AW, scale, seed) in {
        "cfg1_small": (1, 8, 37, 75, 128, 8, 8, 1.0 / 16, 3),
        "multi_img": (3, 5, 20, 31, 64, 8, 8, 1.0 / 16, 4),
        "odd_grid": (2, 3, 13, 17, 40, 3, 5, 1.0 / 8, 5),
    }.items():
        g = torch.Generator().manual_seed(seed)
        feat = torch.relu(torch.randn(B, C, H, W, generator=g)).numpy()
        rois = synth_rois(R, B, seed + 100, im_h=int(H / scale), im_w=int(W / scale)).numpy()
        # edge cases: box ending on the last pixel (extrapolation branch), a
        # degenerate box, a box partly outside, an inverted box
        rois[0, 1:] = [W / scale - 40, H / scale - 30, W / scale - 1, H / scale - 1]
        rois[1, 1:] = [10, 10, 10, 10]
        rois[2, 1:] = [-30, -20, 50, 40]
        rois[3, 1:] = [100, 80, 60, 40]
        rois[4, 1:] = [W / scale - 8, 0, W / scale + 30, H / scale + 20]
        ref_out = np.zeros((R, C, AH, AW), np.float32)
        ref_cpu.ROIAlignForwardCpu(feat.ctypes.data_as(c_float_p), ctypes.c_float(scale), R, H, W, C,
                                   AH, AW, rois.ctypes.data_as(c_float_p),
                                   ref_out.ctypes.data_as(c_float_p))
        mine = orc.roi_align_forward(feat, rois, AH, AW, scale)
        check("roi_align_fwd vs reference roi_align.c [%s]" % tag, mine, ref_out)
        gold_roi[tag + "_rois"] = rois
        gold_roi[tag + "_out"] = ref_out
        gold_roi[tag + "_meta"] = np.array([B, C, H, W, R, AH, AW, seed], np.int64)
        gold_roi[tag + "_scale"] = np.array([scale], np.float64)

    # ---------------- 2. reference Python RPN layers ----------------------
    sys.path[:0] = [os.path.join(HERE, "stub"), REF_LIB]
    from model.utils.config import cfg, cfg_from_list
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    cfg.USE_GPU_NMS = False
    from model.rpn.generate_anchors import generate_anchors as ref_generate_anchors
    from model.rpn import bbox_transform as ref_bt
    import model.rpn.proposal_layer as ref_pl
    import model.rpn.anchor_target_layer as ref_atl

    # known-answer test from generate_anchors.py:19-37
    kat = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                    [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                    [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]], np.float64)
    # the table is MATLAB (1-based pixel coordinates); the Python code is 0-based
    check("generate_anchors default vs MATLAB table - 1", orc.generate_anchors(), kat - 1)
    check("generate_anchors default vs reference", orc.generate_anchors(), ref_generate_anchors())
    sc, ra = np.array([4, 8, 16, 32]), np.array([0.5, 1, 2])
    anchors12 = ref_generate_anchors(scales=sc, ratios=ra)
    check("generate_anchors A=12 vs reference", orc.generate_anchors(scales=sc, ratios=ra), anchors12)

    def oracle_nms_as_reference(dets, thresh, force_cpu=False):
        if dets.shape[0] == 0:
            return []
        keep = orc.nms(dets.numpy(), float(thresh))
        return torch.from_numpy(keep.astype(np.int32)).view(-1, 1)

    ref_pl.nms = oracle_nms_as_reference

    # Tie order of `torch.sort(scores, 1, True)` (proposal_layer.py:125) is unpinned by the
    # reference (torch 0.4 THC sort).  On this torch the CPU default puts equal scores in
    # REVERSE index order while stable=True (and the CUDA radix path) keeps the lower index
    # first.  The contract of this repo is the stable order, so the
    # reference layer is run with torch.sort forced stable; nothing else is altered.
    class _StableSortTorch(object):
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def sort(x, dim=-1, descending=False):
            return torch.sort(x, dim=dim, descending=descending, stable=True)

    ref_pl.torch = _StableSortTorch()

    gold_rpn = {"anchors12": anchors12}
    for tag, (B, H, W, seed) in {"vgg_600x1200": (1, 37, 75, 3), "batch3_small": (3, 19, 25, 7)}.items():
        A = 12
        prob, deltas = synth_rpn(B, A, H, W, seed)
        im_info = torch.tensor([[H * 16.0 + 8, W * 16.0, 0.5859375]] * B)
        if B > 1:
            im_info[1, 0] -= 40  # per-image clip bounds differ
            im_info[1, 1] -= 24
        # decode + clip
        layer = ref_pl._ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
        exp_d = torch.exp(deltas)
        for key in ("TRAIN", "TEST"):
            if B > 1 and key == "TRAIN":
                cfg.TRAIN.RPN_PRE_NMS_TOP_N, cfg.TRAIN.RPN_POST_NMS_TOP_N = 3000, 500
            ref_out = layer((prob, deltas, im_info, key)).numpy()
            c = cfg[key]
            mine, order, boxes, num = orc.proposal_layer(
                prob.numpy(), deltas.numpy(), im_info.numpy(), anchors12, 16, c.RPN_PRE_NMS_TOP_N,
                c.RPN_POST_NMS_TOP_N, c.RPN_NMS_THRESH, exp_deltas=exp_d.numpy(), return_debug=True)
            check("proposal_layer %s [%s]" % (key, tag), mine, ref_out)
            gold_rpn["%s_%s_rois" % (tag, key)] = ref_out
            gold_rpn["%s_%s_num" % (tag, key)] = num
            gold_rpn["%s_%s_cfg" % (tag, key)] = np.array(
                [c.RPN_PRE_NMS_TOP_N, c.RPN_POST_NMS_TOP_N], np.int64)
            # reference sort order: torch.sort(scores, 1, True)
            s_flat = prob[:, A:].permute(0, 2, 3, 1).contiguous().view(B, -1)
            _, ref_order = torch.sort(s_flat, dim=1, descending=True, stable=True)
            check("top-k order %s [%s]" % (key, tag), order, ref_order[:, :order.shape[1]].numpy().astype(np.int32))
        cfg.TRAIN.RPN_PRE_NMS_TOP_N, cfg.TRAIN.RPN_POST_NMS_TOP_N = 12000, 2000
        gold_rpn[tag + "_meta"] = np.array([B, A, H, W, seed], np.int64)
        gold_rpn[tag + "_im_info"] = im_info.numpy()
        # exp values the CPU torch used (the one non-IEEE op on the path), dw/dh channels only
        gold_rpn[tag + "_exp_dwdh"] = exp_d.view(B, A, 4, H, W)[:, :, 2:].contiguous().numpy()

        # bbox_transform_inv + clip_boxes directly
        anc = torch.from_numpy(orc.shifted_anchors(anchors12, H, W, 16))
        dl = deltas.permute(0, 2, 3, 1).contiguous().view(B, -1, 4)
        ref_boxes = ref_bt.clip_boxes(ref_bt.bbox_transform_inv(anc.view(1, -1, 4).expand(B, -1, 4), dl, B),
                                      im_info, B).numpy()
        full = orc.proposal_layer(prob.numpy(), deltas.numpy(), im_info.numpy(), anchors12, 16, 0, 1, 2.0,
                                  exp_deltas=exp_d.numpy(), return_debug=True)
        inv = np.empty_like(ref_boxes)
        for b in range(B):
            inv[b, full[1][b]] = full[2][b]
        check("bbox_transform_inv+clip_boxes [%s]" % tag, inv, ref_boxes)

        # anchor targets
        gt = synth_gt(B, 20, 50, seed + 50, im_h=int(im_info[0, 0]), im_w=int(im_info[0, 1]))
        num_boxes = torch.full((B,), 20, dtype=torch.long)
        atl = ref_atl._AnchorTargetLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
        np.random.seed(3)
        ref_t = atl((torch.zeros(B, 2 * A, H, W), gt, im_info, num_boxes))
        np.random.seed(3)
        mine_t = orc.anchor_target_layer(H, W, gt.numpy(), im_info.numpy(), anchors12, 16)
        for nm, a, b in zip(("labels", "bbox_targets", "inside_w", "outside_w"), mine_t, ref_t):
            if nm == "bbox_targets":
                # dx, dy are IEEE-exact; dw, dh go through log(), which differs by an ulp
                # between numpy and torch (library-dependent transcendental)
                av = a.reshape(B, A, 4, H, W)
                bv = b.numpy().reshape(B, A, 4, H, W)
                check("anchor_target bbox_targets dx,dy [%s]" % tag, av[:, :, :2], bv[:, :, :2])
                ok = np.allclose(av[:, :, 2:], bv[:, :, 2:], rtol=2e-6, atol=2e-7, equal_nan=True)
                print(("PASS " if ok else "FAIL ") + "anchor_target bbox_targets dw,dh (2e-6) [%s]" % tag)
                if not ok:
                    raise SystemExit(1)
            else:
                check("anchor_target %s [%s]" % (nm, tag), a, b.numpy())
            gold_rpn["%s_at_%s" % (tag, nm)] = b.numpy()
        gold_rpn[tag + "_gt"] = gt.numpy()
        # IoU matrix
        inside = orc.shifted_anchors(anchors12, H, W, 16)[:2000]
        ref_ov = ref_bt.bbox_overlaps_batch(torch.from_numpy(inside), gt).numpy()
        check("bbox_overlaps_batch 2-D anchors [%s]" % tag, orc.bbox_overlaps_batch(inside, gt.numpy()), ref_ov)
        rois3 = torch.cat([torch.zeros(B, 300, 1), torch.from_numpy(inside[:300]).expand(B, 300, 4) + 3.0], 2)
        ref_ov3 = ref_bt.bbox_overlaps_batch(rois3, gt).numpy()
        check("bbox_overlaps_batch 3-D rois(5) [%s]" % tag, orc.bbox_overlaps_batch(rois3.numpy(), gt.numpy()), ref_ov3)
        gold_rpn[tag + "_iou2000"] = ref_ov

    # adversarial ties: duplicated scores must keep lower index first
    B, A, H, W = 1, 12, 6, 7
    prob, deltas = synth_rpn(B, A, H, W, 11)
    prob[:, A:] = torch.round(prob[:, A:] * 8) / 8  # many exact ties
    im_info = torch.tensor([[H * 16.0, W * 16.0, 1.0]])
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = 200, 50
    layer = ref_pl._ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
    ref_out = layer((prob, deltas, im_info, "TEST")).numpy()
    mine = orc.proposal_layer(prob.numpy(), deltas.numpy(), im_info.numpy(), anchors12, 16, 200, 50, 0.7,
                              exp_deltas=torch.exp(deltas).numpy())
    check("proposal_layer with tied scores", mine, ref_out)
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = 6000, 300
    gold_rpn["ties_rois"] = ref_out
    gold_rpn["ties_prob"] = prob.numpy()
    gold_rpn["ties_deltas"] = deltas.numpy()
    gold_rpn["ties_exp"] = torch.exp(deltas).numpy()

    # ---------------- 3. reference _ProposalTargetLayer ----------------------
    # proposal_target_layer_cascade.py:133 calls Tensor.index((idx,)), which torch 2.x no longer has;
    # the shim below is the only change (the reference source is used as it lies).
    _orig_index = getattr(torch.Tensor, "index", None)  # torch 2.11 has an unrelated Tensor.index(idx, dims)
    torch.Tensor.index = lambda self, idx, *a: self[idx]
    import model.rpn.proposal_target_layer_cascade as ref_ptl
    gold_pt = {}
    for tag, (B, seed, batch, bg_lo) in {"default_cfg": (2, 21, 128, 0.1), "cityscape_yml": (3, 22, 256, 0.0)}.items():
        A, H, W = 12, 37, 75
        prob, deltas = synth_rpn(B, A, H, W, seed)
        im_info = torch.tensor([[600.0, 1200.0, 0.5859375]] * B)
        gt = synth_gt(B, 20, 50, seed + 50, im_h=600, im_w=1200)
        rois = orc.proposal_layer(prob.numpy(), deltas.numpy(), im_info.numpy(), anchors12, 16, 12000, 2000, 0.7,
                                  exp_deltas=torch.exp(deltas).numpy())
        cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.BG_THRESH_LO = batch, bg_lo
        layer = ref_ptl._ProposalTargetLayer(9)
        np.random.seed(3)
        ref_o = layer(torch.from_numpy(rois), gt, torch.full((B,), 20, dtype=torch.long))
        np.random.seed(3)
        mine_o = orc.proposal_target_layer(rois, gt.numpy(), batch_size=batch, bg_thresh_lo=bg_lo)
        for nm, a, b in zip(("rois", "labels", "bbox_targets", "inside_w", "outside_w"), mine_o, ref_o):
            b = b.numpy()
            if nm == "bbox_targets":
                check("proposal_target bbox_targets dx,dy [%s]" % tag, a[:, :, :2], b[:, :, :2])
                ok = np.allclose(a[:, :, 2:], b[:, :, 2:], rtol=2e-6, atol=2e-6)
                print(("PASS " if ok else "FAIL ") + "proposal_target bbox_targets dw,dh (2e-6) [%s]" % tag)
                if not ok:
                    raise SystemExit(1)
            else:
                check("proposal_target %s [%s]" % (nm, tag), a, b)
            gold_pt["%s_%s" % (tag, nm)] = b
        gold_pt[tag + "_all_rois"] = rois
        gold_pt[tag + "_gt"] = gt.numpy()
        gold_pt[tag + "_cfg"] = np.array([batch, bg_lo], np.float64)
    cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.BG_THRESH_LO = 128, 0.1
    if _orig_index is not None:
        torch.Tensor.index = _orig_index

    if not args.no_write:
        np.savez_compressed(os.path.join(GOLD, "proposal_target_ref_py.npz"), **gold_pt)
        np.savez_compressed(os.path.join(GOLD, "roi_align_ref_cpu.npz"), **gold_roi)
        np.savez_compressed(os.path.join(GOLD, "rpn_layers_ref_py.npz"), **gold_rpn)
        print("wrote", GOLD)
    print("ALL REFERENCE CHECKS PASSED")


if __name__ == "__main__":
    main()
