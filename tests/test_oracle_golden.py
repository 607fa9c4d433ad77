"""CPU: the oracle against the golden vectors produced by the REFERENCE's own code
(oracle/validate_against_reference.py: roi_align.c and lib/model/rpn/*.py)."""
import os

import numpy as np
import torch

from oracle import oracle as orc
from oracle.synth import synth_gt, synth_rpn
from util import ANCHOR_RATIOS, ANCHOR_SCALES, GOLD, bits_equal, features


def test_generate_anchors_known_answer():
    # lib/model/rpn/generate_anchors.py:19-37 (MATLAB, 1-based) minus one
    kat = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                    [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                    [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]], np.float64) - 1
    assert np.array_equal(orc.generate_anchors(), kat)
    g = np.load(os.path.join(GOLD, "rpn_layers_ref_py.npz"))
    assert np.array_equal(orc.generate_anchors(scales=ANCHOR_SCALES, ratios=ANCHOR_RATIOS), g["anchors12"])


def test_roi_align_forward_vs_reference_c():
    g = np.load(os.path.join(GOLD, "roi_align_ref_cpu.npz"))
    for tag in ("cfg1_small", "multi_img", "odd_grid"):
        B, C, H, W, R, AH, AW, seed = [int(v) for v in g[tag + "_meta"]]
        scale = float(g[tag + "_scale"][0])
        feat = features(B, C, H, W, seed).numpy()
        out = orc.roi_align_forward(feat, g[tag + "_rois"], AH, AW, scale)
        assert bits_equal(out, g[tag + "_out"]), tag


def test_roi_align_backward_is_adjoint_of_forward():
    feat = features(2, 4, 11, 13, 0).numpy()
    rois = np.array([[0, 5, 7, 120, 90], [1, 30, 20, 200, 170], [1, 0, 0, 207, 175], [0, 150, 100, 207, 175]],
                    np.float32)
    out = orc.roi_align_forward(feat, rois, 8, 8, 1 / 16)
    g = np.random.RandomState(0).randn(*out.shape).astype(np.float32)
    gin = orc.roi_align_backward(g, rois, feat.shape, 1 / 16, accumulate_double=True)
    lhs = float((out.astype(np.float64) * g).sum())
    rhs = float((feat.astype(np.float64) * gin).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_roi_pool_oracle_properties():
    feat = features(2, 3, 12, 15, 1).numpy()
    rois = np.array([[0, 0, 0, 239, 191], [1, 33, 20, 150, 100], [1, 100, 80, 60, 40], [0, 230, 180, 260, 200]],
                    np.float32)
    out, arg = orc.roi_pool_forward(feat, rois, 7, 7, 1 / 16)
    flat = feat.reshape(-1)
    ok = arg >= 0
    assert np.array_equal(out[ok], flat[arg[ok]])
    assert np.all(out[~ok] == 0)
    # whole-map RoI: the global max of each plane is among the pooled values
    assert np.allclose(out[0].reshape(3, -1).max(1), feat[0].reshape(3, -1).max(1))
    g = np.ones_like(out)
    gin = orc.roi_pool_backward(g, arg, rois, feat.shape, 1 / 16)
    # inverted RoI (row 2) is forced to 1x1 in the forward but fails the backward's in_roi test
    assert gin.sum() <= ok.sum()
    assert gin.sum() >= ok[[0, 1, 3]].sum() - 1e-3


def test_nms_oracle_small_known_case():
    dets = np.array([[0, 0, 99, 99, 0.9], [5, 5, 104, 104, 0.8], [200, 200, 299, 299, 0.7],
                     [0, 0, 99, 99, 0.6], [205, 200, 304, 299, 0.5]], np.float32)
    # IoU(0,1) = 95*95/(2*100*100-95*95) = 0.822 > 0.7; (0,3) identical; (2,4) = 95*100/(2e4-9500)=0.905
    assert orc.nms(dets, 0.7).tolist() == [0, 2]
    assert orc.nms(dets, 0.95).tolist() == [0, 1, 2, 4]
    assert orc.nms(dets, 0.7, max_keep=1).tolist() == [0]
    assert orc.nms(dets[:0], 0.7).tolist() == []


def test_proposal_layer_vs_reference_python():
    g = np.load(os.path.join(GOLD, "rpn_layers_ref_py.npz"))
    for tag in ("vgg_600x1200", "batch3_small"):
        B, A, H, W, seed = [int(v) for v in g[tag + "_meta"]]
        prob, deltas = synth_rpn(B, A, H, W, seed)
        exp_d = torch.exp(deltas)
        # the golden file stores the exp values the reference run used for dw/dh
        assert bits_equal(exp_d.view(B, A, 4, H, W)[:, :, 2:].contiguous().numpy(), g[tag + "_exp_dwdh"])
        for key in ("TRAIN", "TEST"):
            pre, post = [int(v) for v in g["%s_%s_cfg" % (tag, key)]]
            rois, order, boxes, num = orc.proposal_layer(prob.numpy(), deltas.numpy(), g[tag + "_im_info"],
                                                         g["anchors12"], 16, pre, post, 0.7,
                                                         exp_deltas=exp_d.numpy(), return_debug=True)
            assert bits_equal(rois, g["%s_%s_rois" % (tag, key)]), (tag, key)
            assert np.array_equal(num, g["%s_%s_num" % (tag, key)])


def test_proposal_layer_ties_keep_lower_index_first():
    g = np.load(os.path.join(GOLD, "rpn_layers_ref_py.npz"))
    rois = orc.proposal_layer(g["ties_prob"], g["ties_deltas"], np.array([[96.0, 112.0, 1.0]], np.float32),
                              g["anchors12"], 16, 200, 50, 0.7, exp_deltas=g["ties_exp"])
    assert bits_equal(rois, g["ties_rois"])


def test_anchor_targets_vs_reference_python():
    g = np.load(os.path.join(GOLD, "rpn_layers_ref_py.npz"))
    for tag in ("vgg_600x1200", "batch3_small"):
        B, A, H, W, seed = [int(v) for v in g[tag + "_meta"]]
        gt = g[tag + "_gt"]
        assert bits_equal(gt, synth_gt(B, 20, 50, seed + 50, im_h=int(g[tag + "_im_info"][0, 0]),
                                       im_w=int(g[tag + "_im_info"][0, 1])).numpy())
        np.random.seed(3)
        out = orc.anchor_target_layer(H, W, gt, g[tag + "_im_info"], g["anchors12"], 16)
        assert bits_equal(out[0], g[tag + "_at_labels"])
        assert bits_equal(out[2], g[tag + "_at_inside_w"])
        assert bits_equal(out[3], g[tag + "_at_outside_w"])
        mine = out[1].reshape(B, A, 4, H, W)
        ref = g[tag + "_at_bbox_targets"].reshape(B, A, 4, H, W)
        assert bits_equal(mine[:, :, :2], ref[:, :, :2])
        # log() is library dependent (numpy vs torch): 2e-6
        assert np.allclose(mine[:, :, 2:], ref[:, :, 2:], rtol=2e-6, atol=2e-7, equal_nan=True)
        inside = orc.shifted_anchors(g["anchors12"], H, W, 16)[:2000]
        assert bits_equal(orc.bbox_overlaps_batch(inside, gt), g[tag + "_iou2000"])


def test_da_losses_match_torch_autograd():
    torch.manual_seed(0)
    s = torch.randn(2, 2, 5, 7, dtype=torch.float64, requires_grad=True)
    p = torch.sigmoid(torch.randn(9, 1, dtype=torch.float64)).requires_grad_(True)
    for d in (0, 1):
        lab = torch.full((2, 5, 7), d, dtype=torch.long)
        img = torch.nn.functional.nll_loss(torch.log_softmax(s, 1), lab)
        ins = torch.nn.BCELoss()(p, torch.full_like(p, float(d)))
        cons = torch.softmax(s, 1)[:, d].mean().detach()
        cst = torch.nn.MSELoss(reduction="sum")(p, cons.repeat(p.size()))
        o = orc.da_losses(s.detach().numpy(), p.detach().numpy(), d)
        assert abs(o["img_loss"] - img.item()) < 1e-12
        assert abs(o["ins_loss"] - ins.item()) < 1e-12
        assert abs(o["cst_loss"] - cst.item()) < 1e-12
        gi, gp = torch.autograd.grad(0.3 * img + 0.5 * ins + 0.7 * cst, [s, p])
        oi, op = orc.da_losses_grad(s.detach().numpy(), p.detach().numpy(), d, None, 0.3, 0.5, 0.7)
        assert np.allclose(oi, gi.numpy(), atol=1e-12)
        assert np.allclose(op, gp.numpy().reshape(-1), atol=1e-10)
