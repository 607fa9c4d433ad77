"""RPN head losses against a literal torch transcription of
lib/model/rpn/rpn.py:90-108 and lib/model/utils/net_utils.py:72-86 (tolerance 1e-5 relative:
floating-point reductions in a different order)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu


def _smooth_l1_loss(bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma=1.0, dim=[1]):
    # net_utils.py:72-86, verbatim semantics
    sigma_2 = sigma ** 2
    box_diff = bbox_pred - bbox_targets
    in_box_diff = bbox_inside_weights * box_diff
    abs_in_box_diff = torch.abs(in_box_diff)
    smoothL1_sign = (abs_in_box_diff < 1. / sigma_2).detach().float()
    in_loss_box = torch.pow(in_box_diff, 2) * (sigma_2 / 2.) * smoothL1_sign \
        + (abs_in_box_diff - (0.5 / sigma_2)) * (1. - smoothL1_sign)
    loss_box = bbox_outside_weights * in_loss_box
    for i in sorted(dim, reverse=True):
        loss_box = loss_box.sum(i)
    return loss_box.mean()


def _reference(score, pred, labels, targets, inside, outside):
    B = score.size(0)
    A = score.size(1) // 2
    reshaped = score.view(B, 2, A * score.size(2), score.size(3))          # rpn.py:63 reshape(x, 2)
    cls = reshaped.permute(0, 2, 3, 1).contiguous().view(B, -1, 2)          # :93
    lab = labels.view(B, -1)
    keep = lab.view(-1).ne(-1).nonzero().view(-1)                           # :96
    cls = torch.index_select(cls.view(-1, 2), 0, keep)
    lab = torch.index_select(lab.view(-1), 0, keep).long()
    loss_cls = TF.cross_entropy(cls, lab)
    loss_box = _smooth_l1_loss(pred, targets, inside, outside, sigma=3, dim=[1, 2, 3])
    return loss_cls, loss_box


@pytest.mark.parametrize("B,A,H,W", [(2, 12, 37, 75), (1, 9, 5, 7), (3, 12, 20, 31)])
def test_rpn_losses_forward_backward(B, A, H, W):
    import tlod_b200
    g = torch.Generator().manual_seed(B * 100 + A)
    dev = "cuda:0"
    score = (2 * torch.randn(B, 2 * A, H, W, generator=g)).to(dev)
    pred = (0.3 * torch.randn(B, 4 * A, H, W, generator=g)).to(dev)
    labels = torch.full((B, 1, A * H, W), -1.0)
    flat = labels.view(B, -1)
    for b in range(B):
        idx = torch.randperm(flat.size(1), generator=g)[:256]
        flat[b, idx[:100]] = 1.0
        flat[b, idx[100:]] = 0.0
    labels = labels.to(dev)
    targets = (0.5 * torch.randn(B, 4 * A, H, W, generator=g)).to(dev)
    fgmask = (labels.view(B, A, H, W) == 1).repeat_interleave(4, dim=1)
    # anchor-target layout: (B, 4A, H, W) from (B, H, W, 4A); any mask works for the formula
    inside = fgmask.float()
    outside = ((labels.view(B, A, H, W) >= 0).repeat_interleave(4, dim=1)).float() / 256.0

    s1, p1 = score.clone().requires_grad_(True), pred.clone().requires_grad_(True)
    rc, rb = _reference(s1, p1, labels, targets, inside, outside)
    (1.5 * rc + 0.7 * rb).backward()
    s2, p2 = score.clone().requires_grad_(True), pred.clone().requires_grad_(True)
    mc, mb = tlod_b200.rpn_losses(s2, p2, labels, targets, inside, outside, 3.0)
    (1.5 * mc + 0.7 * mb).backward()
    mc, mb, rc, rb = mc.detach(), mb.detach(), rc.detach(), rb.detach()
    assert abs(float(mc) - float(rc)) <= 1e-5 * max(1.0, abs(float(rc)))
    assert abs(float(mb) - float(rb)) <= 1e-5 * max(1.0, abs(float(rb)))
    for a, b in ((s2.grad, s1.grad), (p2.grad, p1.grad)):
        den = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-5 * max(den, 1e-12)
    # deterministic: the same call twice gives the same bits
    again = tlod_b200.functional.rpn_loss_forward(score, labels, pred, targets, inside, outside)
    again2 = tlod_b200.functional.rpn_loss_forward(score, labels, pred, targets, inside, outside)
    assert torch.equal(again, again2)
    assert int(again[2]) == 256 * B and int(again[3]) == 100 * B


def test_rpn_losses_all_ignored():
    import tlod_b200
    dev = "cuda:0"
    B, A, H, W = 1, 3, 4, 5
    z = torch.zeros(B, 4 * A, H, W, device=dev)
    out = tlod_b200.functional.rpn_loss_forward(torch.randn(B, 2 * A, H, W, device=dev),
                                                torch.full((B, 1, A * H, W), -1.0, device=dev), z, z, z, z)
    assert out.tolist() == [0.0, 0.0, 0.0, 0.0]
