"""Drop-in proof.

CPU, where /root/reference exists (the build container): the reference's OWN lib/model/rpn/rpn.py
is imported unchanged with this repo's mirror package in front of the reference's lib/ on sys.path;
its _RPN must construct, hold THIS repo's _ProposalLayer / _AnchorTargetLayer, and its forward()
must reach them (they refuse CPU tensors: there is no CPU path).  The reference's remaining modules
(model.utils.blob, the full config) stay importable behind the overlay.

GPU: (i) the reference's launcher ABI (ROIAlign*Laucher, ROIPool*Laucher, nms_cuda_compute) exported
by libtlod_b200.so, driven through the same ctypes harness as the reference's recompiled kernels and
compared with them; (ii) the RPN glue of rpn.py:58-110 (proposal layer, anchor targets, index_select
cross-entropy, _smooth_l1_loss) over this repo's layers, train and eval mode, against the oracle."""
import ctypes
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from util import ROOT, bits_equal, edge_rois, features, rel_err

REF_LIB = "/root/reference/lib"
PKG = os.path.join(ROOT, "transfer-learning-library-for-object-detection_b200")
DEV = "cuda:0"


def test_compat_header_symbols_are_exported():
    import re
    from tlod_b200 import _lib
    text = open(os.path.join(ROOT, "include", "tlod_b200_compat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(ROI\w+Laucher|nms_cuda_compute)\s*\(", text)
    assert sorted(names) == ["ROIAlignBackwardLaucher", "ROIAlignForwardLaucher", "ROIPoolBackwardLaucher",
                             "ROIPoolForwardLaucher", "nms_cuda_compute"]
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), n


@pytest.mark.skipif(not os.path.isdir(REF_LIB), reason="needs the reference checkout (build container only)")
def test_reference_rpn_module_runs_on_top_of_this_package():
    # a fresh interpreter: sys.path order is the point of the test
    code = textwrap.dedent("""
        import sys
        sys.path[:0] = [%r, %r, %r]          # mirror package, easydict stand-in, reference lib/
        import torch
        import model
        from model.rpn.rpn import _RPN                       # the reference's file, unchanged
        import model.rpn.rpn as ref_rpn
        assert ref_rpn.__file__.startswith(%r), ref_rpn.__file__
        from model.rpn.proposal_layer import _ProposalLayer
        from model.rpn.anchor_target_layer import _AnchorTargetLayer
        import model.rpn.proposal_layer as pl
        assert pl.__file__.startswith(%r), pl.__file__      # ... resolved to THIS repo
        from model.utils.config import cfg, cfg_from_list    # the reference's full config, adopted
        assert "POOLING_MODE" in cfg and "RESNET" in cfg and cfg.TRAIN.RPN_BATCHSIZE == 256
        cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]"])
        import model.utils.blob                              # reference-only module behind the overlay
        rpn = _RPN(512)
        assert type(rpn.RPN_proposal) is _ProposalLayer and type(rpn.RPN_anchor_target) is _AnchorTargetLayer
        assert rpn.RPN_proposal._num_anchors == 12
        rpn.eval()
        try:
            rpn(torch.zeros(1, 512, 8, 10), torch.tensor([[128., 160., 1.]]), None, None)
        except RuntimeError as e:
            assert "no CPU implementation" in str(e), e      # reached this repo's proposal layer
        else:
            raise AssertionError("the CPU call should have been refused by tlod_b200")
        print("DROPIN-OK")
    """) % (PKG, os.path.join(ROOT, "oracle", "stub"), REF_LIB, REF_LIB, PKG)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "DROPIN-OK" in p.stdout, p.stdout + p.stderr


# ---------------------------------------------------------------------------------------------
P, F32, I = ctypes.c_void_p, ctypes.c_float, ctypes.c_int


def _bind(path):
    lib = ctypes.CDLL(path)
    lib.ROIAlignForwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, P, P, P]
    lib.ROIAlignBackwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, I, P, P, P]
    lib.ROIPoolForwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, P, P, P, P]
    lib.ROIPoolBackwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, I, P, P, P, P]
    lib.nms_cuda_compute.argtypes = [P, P, P, I, I, F32]
    lib.nms_cuda_compute.restype = None
    return lib


@pytest.mark.gpu
def test_reference_launcher_abi_served_by_this_library():
    from oracle.synth import synth_rois
    from tlod_b200 import _lib
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_cuda_nofma.so")
    if not os.path.exists(ref_so):
        pytest.fail("oracle/_ref/libref_cuda_nofma.so missing: run `make -C oracle ref` where /root/reference exists")
    mine, ref = _bind(_lib.LIB_PATH), _bind(ref_so)
    B, C, H, W, R, scale = 2, 64, 37, 75, 160, 1 / 16
    feat = features(B, C, H, W, 1).to(DEV)
    rois = edge_rois(synth_rois(R, B, 2), H, W, scale).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for lib in (mine, ref):
        y = torch.zeros(R, C, 8, 8, device=DEV)
        assert lib.ROIAlignForwardLaucher(feat.data_ptr(), scale, R, H, W, C, 8, 8, rois.data_ptr(), y.data_ptr(), st) == 1
        top = torch.randn(R, C, 8, 8, generator=torch.Generator().manual_seed(3)).to(DEV)
        g = torch.zeros(B, C, H, W, device=DEV)
        assert lib.ROIAlignBackwardLaucher(top.data_ptr(), scale, B, R, H, W, C, 8, 8, rois.data_ptr(), g.data_ptr(), st) == 1
        o7 = torch.zeros(R, C, 7, 7, device=DEV)
        a7 = torch.zeros(R, C, 7, 7, dtype=torch.int32, device=DEV)
        assert lib.ROIPoolForwardLaucher(feat.data_ptr(), scale, R, H, W, C, 7, 7, rois.data_ptr(), o7.data_ptr(),
                                         a7.data_ptr(), st) == 1
        t7 = torch.randn(R, C, 7, 7, generator=torch.Generator().manual_seed(4)).to(DEV)
        g7 = torch.zeros(B, C, H, W, device=DEV)
        assert lib.ROIPoolBackwardLaucher(t7.data_ptr(), scale, B, R, H, W, C, 7, 7, rois.data_ptr(), g7.data_ptr(),
                                          a7.data_ptr(), st) == 1
        torch.cuda.synchronize()
        outs.append((y, g, o7, a7, g7))
    (y, g, o7, a7, g7), (yr, gr, o7r, a7r, g7r) = outs
    assert rel_err(y.cpu().numpy(), yr.cpu().numpy()) <= 1e-5
    assert rel_err(g.cpu().numpy(), gr.cpu().numpy()) <= 1e-4
    assert torch.equal(o7, o7r) and torch.equal(a7, a7r)
    assert rel_err(g7.cpu().numpy(), g7r.cpu().numpy()) <= 1e-4
    from test_gpu_proposals import _sorted_dets
    for n, thr in ((300, 0.3), (6000, 0.7), (12000, 0.7)):
        dets = _sorted_dets(n, 900 + n, jitter=True).to(DEV)
        keeps = []
        for lib in (mine, ref):
            keep = torch.zeros(n, dtype=torch.int32, device=DEV)
            num = torch.zeros(1, dtype=torch.int32, device=DEV)
            torch.cuda.synchronize()
            lib.nms_cuda_compute(keep.data_ptr(), num.data_ptr(), dets.data_ptr(), n, 5, thr)
            keeps.append(keep[:int(num.item())].cpu().numpy())  # no extra sync: the call has completed
        assert np.array_equal(keeps[0], keeps[1]), (n, thr)


@pytest.mark.gpu
@pytest.mark.parametrize("training", [True, False])
def test_rpn_glue_over_this_repos_layers(training):
    """rpn.py:58-110 with this repo's layers: proposal layer -> (training) anchor targets ->
    index_select cross-entropy + _smooth_l1_loss, against the oracle's layers and the fused
    tlod_b200.rpn_losses on the same inputs."""
    import tlod_b200
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    from model.rpn.proposal_layer import _ProposalLayer
    from model.utils.config import cfg
    from model.utils.net_utils import _smooth_l1_loss
    from oracle import oracle as orc
    from oracle.synth import synth_gt
    from util import ANCHOR_RATIOS, ANCHOR_SCALES
    Bn, A, Hh, Ww = 2, 12, 37, 75
    g = torch.Generator().manual_seed(5)
    cls_score = (2 * torch.randn(Bn, 2 * A, Hh, Ww, generator=g)).to(DEV).requires_grad_(True)
    bbox_pred = (0.2 * torch.randn(Bn, 4 * A, Hh, Ww, generator=g)).to(DEV).requires_grad_(True)
    im_info = torch.tensor([[600.0, 1200.0, 0.5859375]] * Bn).to(DEV)
    gt = synth_gt(Bn, 20, 50, 55).to(DEV)
    num_boxes = torch.full((Bn,), 20, dtype=torch.long)
    proposal = _ProposalLayer(cfg.FEAT_STRIDE[0], ANCHOR_SCALES, ANCHOR_RATIOS)
    anchor_target = _AnchorTargetLayer(cfg.FEAT_STRIDE[0], ANCHOR_SCALES, ANCHOR_RATIOS)
    # rpn.py:66-72: softmax over (bg, fg) on the (B, 2, A*H, W) view
    prob = torch.softmax(cls_score.view(Bn, 2, A * Hh, Ww), 1).view(Bn, 2 * A, Hh, Ww)
    key = "TRAIN" if training else "TEST"
    rois = proposal((prob.data, bbox_pred.data, im_info, key))
    anchors = orc.generate_anchors(scales=ANCHOR_SCALES, ratios=ANCHOR_RATIOS).astype(np.float32)
    pre, post = (12000, 2000) if training else (6000, 300)
    exp_d = torch.exp(bbox_pred.data).cpu().numpy()
    ref_rois = orc.proposal_layer(prob.data.cpu().numpy(), bbox_pred.data.cpu().numpy(), im_info.cpu().numpy(), anchors,
                                  16, pre, post, 0.7, exp_deltas=exp_d)
    assert bits_equal(rois.cpu().numpy(), ref_rois)
    if not training:
        return
    np.random.seed(3)
    rpn_data = anchor_target((cls_score.data, gt, im_info, num_boxes))
    np.random.seed(3)
    ref = orc.anchor_target_layer(Hh, Ww, gt.cpu().numpy(), im_info.cpu().numpy(), anchors, 16)
    for k in (0, 2, 3):  # labels and weights: bit-exact
        assert bits_equal(rpn_data[k].cpu().numpy(), ref[k])
    mine_t = rpn_data[1].cpu().numpy().reshape(Bn, A, 4, Hh, Ww)
    ref_t = ref[1].reshape(Bn, A, 4, Hh, Ww)
    assert bits_equal(mine_t[:, :, :2], ref_t[:, :, :2])  # dx, dy exact; the log terms to the library's ulp
    assert np.allclose(mine_t[:, :, 2:], ref_t[:, :, 2:], rtol=2e-6, atol=2e-7, equal_nan=True)
    # rpn.py:90-108
    score = cls_score.view(Bn, 2, A * Hh, Ww).permute(0, 2, 3, 1).contiguous().view(Bn, -1, 2)
    label = rpn_data[0].view(Bn, -1)
    keep = label.view(-1).ne(-1).nonzero().view(-1)
    loss_cls = torch.nn.functional.cross_entropy(torch.index_select(score.view(-1, 2), 0, keep),
                                                 torch.index_select(label.view(-1), 0, keep).long())
    loss_box = _smooth_l1_loss(bbox_pred, rpn_data[1], rpn_data[2], rpn_data[3], sigma=3, dim=[1, 2, 3])
    (loss_cls + loss_box).backward()
    c2 = cls_score.detach().clone().requires_grad_(True)
    b2 = bbox_pred.detach().clone().requires_grad_(True)
    f_cls, f_box = tlod_b200.rpn_losses(c2, b2, *rpn_data)
    (f_cls + f_box).backward()
    assert torch.allclose(f_cls, loss_cls, rtol=1e-5) and torch.allclose(f_box, loss_box, rtol=1e-5)
    assert torch.allclose(c2.grad, cls_score.grad, rtol=1e-4, atol=1e-9)
    assert torch.allclose(b2.grad, bbox_pred.grad, rtol=1e-4, atol=1e-9)
