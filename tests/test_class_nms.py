"""Test-time per-class NMS (methods/DAF/DAF_test.py:302-320)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from util import bits_equal


def _inputs(R, K, seed, agnostic=False):
    g = torch.Generator().manual_seed(seed)
    scores = torch.softmax(3 * torch.randn(R, K, generator=g), 1)
    ctr = torch.rand(R, 2, generator=g) * torch.tensor([1100.0, 500.0])
    cols = 1 if agnostic else K
    jit = torch.randn(R, cols, 4, generator=g) * 6
    wh = (torch.rand(R, 1, 2, generator=g) * 120 + 20).expand(R, cols, 2)
    x1y1 = ctr[:, None, :] - wh / 2 + jit[:, :, :2]
    boxes = torch.cat([x1y1, x1y1 + wh + jit[:, :, 2:]], 2).reshape(R, cols * 4)
    # clusters of near-duplicates so that NMS at 0.3 really suppresses
    boxes[1::3] = boxes[0::3][: boxes[1::3].shape[0]] + 2.0
    return scores.contiguous(), boxes.contiguous()


def _reference_loop(scores, pred_boxes, thresh, nms_thresh, agnostic):
    """Literal transcription of DAF_test.py:302-320 on CPU tensors (nms = the oracle's)."""
    out = []
    for j in range(1, scores.size(1)):
        inds = torch.nonzero(scores[:, j] > thresh).view(-1)
        if inds.numel() > 0:
            cls_scores = scores[:, j][inds]
            _, order = torch.sort(cls_scores, dim=0, descending=True, stable=True)
            cls_boxes = pred_boxes[inds, :] if agnostic else pred_boxes[inds][:, j * 4:(j + 1) * 4]
            cls_dets = torch.cat((cls_boxes, cls_scores.unsqueeze(1)), 1)
            cls_dets = cls_dets[order]
            keep = torch.from_numpy(orc.nms(cls_dets.numpy(), nms_thresh).astype(np.int64))
            out.append(cls_dets[keep.view(-1).long()].numpy())
        else:
            out.append(np.zeros((0, 5), np.float32))
    return out


@pytest.mark.parametrize("R,K,agn,thresh", [(300, 9, False, 0.05), (300, 9, True, 0.0), (77, 4, False, 0.3)])
def test_oracle_per_class_nms_is_the_reference_loop(R, K, agn, thresh):
    scores, boxes = _inputs(R, K, 5, agn)
    ref = _reference_loop(scores, boxes, thresh, 0.3, agn)
    mine = orc.per_class_nms(scores.numpy(), boxes.numpy(), thresh, 0.3)
    assert len(ref) == len(mine) == K - 1
    for a, b in zip(mine, ref):
        assert bits_equal(a, b)
    assert any(a.shape[0] > 0 for a in mine)


@pytest.mark.gpu
@pytest.mark.parametrize("R,K,agn,thresh", [(300, 9, False, 0.05), (300, 9, True, 0.0), (77, 4, False, 0.3),
                                            (1000, 21, False, 0.01), (64, 3, False, 0.99)])
def test_class_nms_cuda_bit_exact(R, K, agn, thresh):
    from tlod_b200 import functional as F
    scores, boxes = _inputs(R, K, 6, agn)
    ref = orc.per_class_nms(scores.numpy(), boxes.numpy(), thresh, 0.3)
    out = F.class_nms(scores.cuda(), boxes.cuda(), thresh, 0.3)
    assert len(out) == K - 1
    for a, b in zip(out, ref):
        assert bits_equal(a.cpu().numpy(), b)


@pytest.mark.gpu
def test_class_nms_ties_keep_lower_row_first():
    from tlod_b200 import functional as F
    scores, boxes = _inputs(200, 5, 7)
    scores = torch.round(scores * 16) / 16  # many exact ties
    ref = orc.per_class_nms(scores.numpy(), boxes.numpy(), 0.05, 0.3)
    out = F.class_nms(scores.cuda(), boxes.cuda(), 0.05, 0.3)
    for a, b in zip(out, ref):
        assert bits_equal(a.cpu().numpy(), b)
