"""_ProposalTargetLayer.  CPU: the oracle against the golden vectors the
reference's own class produced (oracle/validate_against_reference.py section 3).  GPU: this
repo's layer against the oracle under the same numpy seed -- sampled rois, labels and weights
bit-exact, regression targets exact in dx, dy and within 2e-6 in the log() terms."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from util import GOLD, bits_equal

TAGS = ["default_cfg", "cityscape_yml"]
NAMES = ("rois", "labels", "bbox_targets", "inside_w", "outside_w")


def _gold():
    return np.load(os.path.join(GOLD, "proposal_target_ref_py.npz"))


def _compare(mine, ref):
    for nm, a, b in zip(NAMES, mine, ref):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, nm
        if nm == "bbox_targets":
            assert bits_equal(a[:, :, :2], b[:, :, :2]), nm
            assert np.allclose(a[:, :, 2:], b[:, :, 2:], rtol=2e-6, atol=2e-6), nm
        else:
            assert bits_equal(a, b), nm


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_proposal_target_vs_reference_golden(tag):
    g = _gold()
    batch, bg_lo = g[tag + "_cfg"]
    np.random.seed(3)
    mine = orc.proposal_target_layer(g[tag + "_all_rois"], g[tag + "_gt"], batch_size=int(batch),
                                     bg_thresh_lo=float(bg_lo))
    _compare(mine, [g["%s_%s" % (tag, nm)] for nm in NAMES])
    # structure: foreground first, labels zero behind it, weights only on foreground
    rois, labels, targets, inside, outside = mine
    assert ((labels > 0) == (inside[:, :, 0] > 0)).all()
    assert (targets[labels == 0] == 0).all()
    for i in range(rois.shape[0]):
        assert (rois[i, :, 0] == i).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_proposal_target_layer_cuda_vs_reference_golden(tag):
    from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    from model.utils.config import cfg
    g = _gold()
    batch, bg_lo = g[tag + "_cfg"]
    old = cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.BG_THRESH_LO
    cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.BG_THRESH_LO = int(batch), float(bg_lo)
    try:
        layer = _ProposalTargetLayer(9)
        rois = torch.from_numpy(g[tag + "_all_rois"]).cuda()
        gt = torch.from_numpy(g[tag + "_gt"]).cuda()
        np.random.seed(3)
        out = layer(rois, gt, torch.full((gt.size(0),), 20, dtype=torch.long))
    finally:
        cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.BG_THRESH_LO = old
    _compare([o.cpu().numpy() for o in out], [g["%s_%s" % (tag, nm)] for nm in NAMES])


@pytest.mark.gpu
def test_roi_gt_assign_matches_oracle_iou():
    """tlod_roi_gt_assign == row max / argmax (lowest index on ties) of the oracle IoU matrix."""
    from tlod_b200 import functional as F
    g = _gold()
    rois, gt = g["default_cfg_all_rois"], g["default_cfg_gt"]
    ov = orc.bbox_overlaps_batch(rois, gt)
    mx, asg, lab = F.roi_gt_assign(torch.from_numpy(rois).cuda(), torch.from_numpy(gt).cuda())
    assert bits_equal(mx.cpu().numpy(), ov.max(2))
    assert np.array_equal(asg.cpu().numpy(), ov.argmax(2).astype(np.int32))
    assert bits_equal(lab.cpu().numpy(), np.take_along_axis(gt[:, :, 4], ov.argmax(2), axis=1))
