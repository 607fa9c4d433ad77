"""CPU: the C-ABI boundary and the host-side mirror of the reference's operator surface
(no kernel is launched here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from util import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tlod_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tlod_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tlod_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libtlod_b200.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib.tlod_version() == 100


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "tlod_b200.h")).read()
    assert "Tensor" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert 'extern "C"' in text


def test_argument_errors_are_returned_not_fatal():
    from tlod_b200._lib import lib
    assert lib.tlod_roi_align_forward(None, None, None, 1, 16, 8, 8, 4, 8, 8, 0.0625, None, 0, None) == -1
    assert lib.tlod_roi_align_plan(None, 1, 8, 8, 4, 8, 8, 0.0625, None, 0, None) == -1
    assert lib.tlod_roi_align_plan_bytes(8, 2048) >= 2048 * (512 + 176 + 32 + 4)
    assert lib.tlod_nms(None, 5, 3, 0.7, 0, None, None, None, 0, None) == -1
    assert b"null" in lib.tlod_error_string(-1)
    assert lib.tlod_nms_workspace_bytes(12000) >= 12000 * 188 * 8
    assert lib.tlod_proposals_n_sorted(1, 12, 37, 75, 12000) == 12000
    # the reference's whole-batch numel test (proposal_layer.py:138): no truncation here
    assert lib.tlod_proposals_n_sorted(1, 12, 10, 10, 12000) == 1200
    assert lib.tlod_proposals_n_sorted(20, 12, 10, 10, 12000) == 1200
    assert lib.tlod_proposals_n_sorted(1, 12, 37, 75, 0) == 33300


def test_cpu_tensors_are_rejected_loudly():
    from tlod_b200 import functional as F
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        F.roi_align_forward(torch.zeros(1, 16, 8, 8), torch.zeros(2, 5), 8, 8, 1 / 16)
    from model.nms.nms_wrapper import nms
    assert nms(torch.zeros(0, 5), 0.7) == []  # nms_wrapper.py:15-16
    with pytest.raises(RuntimeError):
        nms(torch.zeros(3, 5), 0.7, force_cpu=True)


def test_operator_surface_names_and_constructors():
    from model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    from model.roi_pooling.modules.roi_pool import _RoIPooling
    from model.roi_layers import ROIAlign, ROIPool
    from model.rpn.proposal_layer import _ProposalLayer
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    from model.rpn.bbox_transform import bbox_overlaps_batch, bbox_transform_batch, bbox_transform_inv, clip_boxes  # noqa
    from model.nms.nms_gpu import nms_gpu  # noqa
    from DAF.DA import grad_reverse  # noqa
    for cls in (RoIAlign, RoIAlignAvg, RoIAlignMax):
        m = cls(7, 7, 1.0 / 16)
        assert (m.aligned_height, m.aligned_width, m.spatial_scale) == (7, 7, 0.0625)
    p = _RoIPooling(7, 7, 1.0 / 16)
    assert (p.pooled_height, p.pooled_width) == (7, 7)
    assert ROIAlign((7, 7), 1.0 / 16, 0).output_size == (7, 7)
    with pytest.raises(NotImplementedError):
        ROIAlign((7, 7), 1.0 / 16, 2)
    assert ROIPool(7, 1.0 / 16).output_size == (7, 7)
    pl = _ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
    assert pl._num_anchors == 12 and pl._anchors.dtype == torch.float32
    at = _AnchorTargetLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
    assert at._num_anchors == 12


def test_generate_anchors_matches_known_answer_and_oracle():
    from model.rpn.generate_anchors import generate_anchors
    from oracle import oracle as orc
    kat = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                    [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                    [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]], np.float64) - 1
    assert np.array_equal(generate_anchors(), kat)
    a12 = generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2]))
    assert np.array_equal(a12, orc.generate_anchors(scales=[4, 8, 16, 32], ratios=[0.5, 1, 2]))
    assert a12[0].tolist() == [-38, -16, 53, 31] and a12[-1].tolist() == [-168, -344, 183, 359]


def test_inside_anchor_index_matches_oracle_layout():
    """The cached inside-anchor tables of _AnchorTargetLayer (host numpy) follow
    anchor_target_layer.py:66-91, including the first-image / int-truncation quirk."""
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    from oracle import oracle as orc
    at = _AnchorTargetLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
    anchors, inds, inv = at._inside(37, 75, 1200, 600, torch.device("cpu"))
    all_a = orc.shifted_anchors(at._anchors_np, 37, 75, 16)
    keep = (all_a[:, 0] >= 0) & (all_a[:, 1] >= 0) & (all_a[:, 2] < 1200) & (all_a[:, 3] < 600)
    assert int(keep.sum()) == 17434 == anchors.shape[0]
    assert np.array_equal(inds.numpy(), np.nonzero(keep)[0])
    assert np.array_equal(anchors.numpy(), all_a[keep])
    assert (inv.numpy() >= 0).sum() == 17434 and inv.numpy()[inds.numpy()[5]] == 5


def test_config_defaults_match_reference():
    from model.utils.config import cfg
    assert cfg.TRAIN.RPN_PRE_NMS_TOP_N == 12000 and cfg.TRAIN.RPN_POST_NMS_TOP_N == 2000
    assert cfg.TEST.RPN_PRE_NMS_TOP_N == 6000 and cfg.TEST.RPN_POST_NMS_TOP_N == 300
    assert cfg.TRAIN.RPN_NMS_THRESH == 0.7 and cfg["TEST"].RPN_NMS_THRESH == 0.7
    assert cfg.TRAIN.RPN_BATCHSIZE == 256 and cfg.TRAIN.RPN_FG_FRACTION == 0.5
