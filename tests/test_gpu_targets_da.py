"""GPU parity: box arithmetic, anchor-target assignment (labels bit-exact), GRL and the
domain-classifier loss reduction."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle.synth import synth_gt, synth_rpn
from util import ANCHOR_RATIOS, ANCHOR_SCALES, bits_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _anchors():
    return orc.generate_anchors(scales=ANCHOR_SCALES, ratios=ANCHOR_RATIOS).astype(np.float32)


@pytest.mark.parametrize("B,H,W", [(1, 37, 75), (3, 19, 25), (8, 38, 75)])
def test_bbox_overlaps_batch_bit_exact(B, H, W):
    from model.rpn.bbox_transform import bbox_overlaps_batch
    gt = synth_gt(B, 20, 50, 60 + B, im_h=H * 16, im_w=W * 16)
    gt[0, 3, :4] = torch.tensor([7.0, 9.0, 7.0, 9.0])  # w == h == 1 real box -> 0
    anchors = torch.from_numpy(orc.shifted_anchors(_anchors(), H, W, 16)[::3].copy())
    anchors[5] = torch.tensor([4.0, 4.0, 4.0, 4.0])  # degenerate anchor -> -1
    ov = bbox_overlaps_batch(anchors.to(DEV), gt.to(DEV)).cpu().numpy()
    assert bits_equal(ov, orc.bbox_overlaps_batch(anchors.numpy(), gt.numpy()))
    # 3-D anchors with a leading batch column (proposal-target call site)
    rois = torch.cat([torch.zeros(B, 300, 1), anchors[:300].expand(B, 300, 4) + 3.0], 2).contiguous()
    ov3 = bbox_overlaps_batch(rois.to(DEV), gt.to(DEV)).cpu().numpy()
    assert bits_equal(ov3, orc.bbox_overlaps_batch(rois.numpy(), gt.numpy()))
    # the same numbers from torch's own eager ops on this device (what lib/model executes)
    a = anchors.to(DEV)
    g = gt.to(DEV)[:, :, :4]
    aw, ah = a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1
    gw, gh = g[:, :, 2] - g[:, :, 0] + 1, g[:, :, 3] - g[:, :, 1] + 1
    iw = (torch.min(a[None, :, None, 2], g[:, None, :, 2]) - torch.max(a[None, :, None, 0], g[:, None, :, 0]) + 1).clamp(min=0)
    ih = (torch.min(a[None, :, None, 3], g[:, None, :, 3]) - torch.max(a[None, :, None, 1], g[:, None, :, 1]) + 1).clamp(min=0)
    ua = (aw * ah).view(1, -1, 1) + (gw * gh).view(B, 1, -1) - iw * ih
    eager = iw * ih / ua
    eager.masked_fill_(((gw == 1) & (gh == 1)).view(B, 1, -1).expand_as(eager), 0)
    eager.masked_fill_(((aw == 1) & (ah == 1)).view(1, -1, 1).expand_as(eager), -1)
    assert bits_equal(ov, eager.cpu().numpy())


def test_bbox_transform_inv_clip_and_batch():
    from model.rpn.bbox_transform import bbox_transform_batch, bbox_transform_inv, clip_boxes
    B, A, H, W = 2, 12, 19, 25
    _, deltas = synth_rpn(B, A, H, W, 5)
    dl = deltas.permute(0, 2, 3, 1).contiguous().view(B, -1, 4)
    anc = torch.from_numpy(orc.shifted_anchors(_anchors(), H, W, 16)).view(1, -1, 4).expand(B, -1, 4).contiguous()
    im_info = torch.tensor([[H * 16.0, W * 16.0, 1.0], [H * 16.0 - 30, W * 16.0 - 50, 1.0]])
    pred = bbox_transform_inv(anc.to(DEV), dl.to(DEV), B)
    pred = clip_boxes(pred, im_info.to(DEV), B)
    # torch eager restatement on the same device (bbox_transform.py:77-103, :125-133)
    a, d = anc.to(DEV), dl.to(DEV)
    w = a[:, :, 2] - a[:, :, 0] + 1.0
    h = a[:, :, 3] - a[:, :, 1] + 1.0
    cx = a[:, :, 0] + 0.5 * w
    cy = a[:, :, 1] + 0.5 * h
    pcx, pcy = d[:, :, 0] * w + cx, d[:, :, 1] * h + cy
    pw, ph = torch.exp(d[:, :, 2]) * w, torch.exp(d[:, :, 3]) * h
    ref = torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], 2)
    for i in range(B):
        ref[i, :, 0::2].clamp_(0, im_info[i, 1].item() - 1)
        ref[i, :, 1::2].clamp_(0, im_info[i, 0].item() - 1)
    assert torch.equal(pred, ref)
    # encode: dx, dy exact against the oracle; log terms equal torch's CUDA log
    gt = (anc + 5.0).to(DEV)
    t = bbox_transform_batch(anc.to(DEV), gt)
    ref_t = orc.bbox_transform_batch(anc.numpy(), (anc + 5.0).numpy())
    assert bits_equal(t[:, :, :2].cpu().numpy(), ref_t[:, :, :2])
    ew = a[:, :, 2] - a[:, :, 0] + 1.0
    gw = gt[:, :, 2] - gt[:, :, 0] + 1.0
    assert torch.equal(t[:, :, 2], torch.log(gw / ew))


@pytest.mark.parametrize("B,H,W,seed", [(1, 37, 75, 3), (3, 19, 25, 7), (2, 38, 75, 9)])
def test_anchor_target_layer(B, H, W, seed):
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    im_info = torch.tensor([[H * 16.0 + 8, W * 16.0, 0.5859375]] * B)
    if B > 1:
        im_info[1, 0] -= 40
    gt = synth_gt(B, 20, 50, seed + 50, im_h=int(im_info[0, 0]), im_w=int(im_info[0, 1]))
    layer = _AnchorTargetLayer(16, ANCHOR_SCALES, ANCHOR_RATIOS)
    np.random.seed(3)
    out = layer((torch.zeros(B, 24, H, W, device=DEV), gt.to(DEV), im_info.to(DEV), torch.full((B,), 20)))
    state_after = np.random.get_state()[1][:8].copy()
    np.random.seed(3)
    ref = orc.anchor_target_layer(H, W, gt.numpy(), im_info.numpy(), _anchors(), 16)
    assert np.array_equal(state_after, np.random.get_state()[1][:8])  # same RNG consumption
    assert out[0].shape == (B, 1, 12 * H, W) and out[1].shape == (B, 48, H, W)
    assert bits_equal(out[0].cpu().numpy(), ref[0])  # labels: bit-exact
    assert bits_equal(out[2].cpu().numpy(), ref[2])
    assert bits_equal(out[3].cpu().numpy(), ref[3])
    mine = out[1].cpu().numpy().reshape(B, 12, 4, H, W)
    r = ref[1].reshape(B, 12, 4, H, W)
    assert bits_equal(mine[:, :, :2], r[:, :, :2])
    assert np.allclose(mine[:, :, 2:], r[:, :, 2:], rtol=2e-6, atol=2e-7, equal_nan=True)  # log(): library ulp


def test_anchor_labels_kernel_vs_oracle_presampling():
    from tlod_b200 import functional as F
    B, H, W = 4, 37, 75
    gt = synth_gt(B, 20, 50, 77)
    all_a = orc.shifted_anchors(_anchors(), H, W, 16)
    keep = (all_a[:, 0] >= 0) & (all_a[:, 1] >= 0) & (all_a[:, 2] < 1200) & (all_a[:, 3] < 600)
    anc = all_a[keep]
    labels, argmax, mx = F.anchor_labels(torch.from_numpy(anc).to(DEV), gt.to(DEV), 0.3, 0.7, False, want_max=True)
    ov = orc.bbox_overlaps_batch(anc, gt.numpy())
    rl, ra, rm = orc.anchor_labels(ov, 0.3, 0.7, False)
    assert bits_equal(labels.cpu().numpy(), rl)
    assert np.array_equal(argmax.cpu().numpy(), ra)
    assert bits_equal(mx.cpu().numpy(), rm)
    labels_c, _ = F.anchor_labels(torch.from_numpy(anc).to(DEV), gt.to(DEV), 0.3, 0.7, True)
    assert bits_equal(labels_c.cpu().numpy(), orc.anchor_labels(ov, 0.3, 0.7, True)[0])


def test_grad_reverse():
    from DAF.DA import GRLayer, grad_reverse
    x = torch.randn(1, 512, 37, 75, device=DEV, requires_grad=True)
    y = GRLayer.apply(x)
    assert torch.equal(y, x)
    g = torch.randn_like(x)
    y.backward(g)
    assert torch.equal(x.grad, g.neg() * 0.1)  # lib/DAF/DA.py:27-29
    # odd length + weighted variant (lib/MAF/DA.py:34-53)
    p = torch.randn(301, 4096, device=DEV, requires_grad=True)
    w = torch.rand(301, device=DEV)
    gp = torch.randn_like(p)
    grad_reverse(p, 0.2, w).backward(gp)
    assert torch.allclose(p.grad, gp.neg() * 0.2 * w.view(-1, 1), rtol=1e-6, atol=0)
    v = torch.randn(1237, device=DEV, requires_grad=True)
    gv = torch.randn_like(v)
    grad_reverse(v, 0.1).backward(gv)
    assert torch.equal(v.grad, gv.neg() * 0.1)


@pytest.mark.parametrize("d", [0, 1])
def test_da_losses_forward_backward(d):
    import tlod_b200
    g = torch.Generator().manual_seed(13 + d)
    score = torch.randn(2, 2, 37, 75, generator=g)
    prob = torch.sigmoid(torch.randn(556, 1, generator=g))
    s = score.to(DEV).requires_grad_(True)
    p = prob.to(DEV).requires_grad_(True)
    img, ins, cst = tlod_b200.da_losses(s, p, d)
    o = orc.da_losses(score.numpy(), prob.numpy(), d)
    assert abs(img.item() - o["img_loss"]) <= 1e-5 * abs(o["img_loss"])
    assert abs(ins.item() - o["ins_loss"]) <= 1e-5 * abs(o["ins_loss"])
    assert abs(cst.item() - o["cst_loss"]) <= 1e-5 * abs(o["cst_loss"])
    (0.1 * (img + ins + cst)).backward()  # DAF total, methods/DAF/DAF_train.py:397-400
    gi, gp = orc.da_losses_grad(score.numpy(), prob.numpy(), d, None, 0.1, 0.1, 0.1)
    assert np.abs(s.grad.cpu().numpy() - gi).max() <= 1e-5 * np.abs(gi).max()
    assert np.abs(p.grad.cpu().numpy().reshape(-1) - gp).max() <= 1e-5 * np.abs(gp).max()
    # and against the reference's own torch expressions on this device (faster_rcnn.py:181-196)
    s2 = score.to(DEV).requires_grad_(True)
    p2 = prob.to(DEV).requires_grad_(True)
    lab = torch.full((2, 37, 75), d, dtype=torch.long, device=DEV)
    ref_img = torch.nn.functional.nll_loss(torch.log_softmax(s2, 1), lab)
    ref_ins = torch.nn.BCELoss()(p2, torch.full_like(p2, float(d)))
    cons = torch.softmax(s2, 1)[:, d].mean().detach()
    ref_cst = torch.nn.MSELoss(reduction="sum")(p2, cons.repeat(p2.size()))
    (0.1 * (ref_img + ref_ins + ref_cst)).backward()
    assert torch.allclose(img, ref_img, rtol=1e-5) and torch.allclose(ins, ref_ins, rtol=1e-5)
    assert torch.allclose(cst, ref_cst, rtol=1e-5)
    assert torch.allclose(s.grad, s2.grad, rtol=1e-4, atol=1e-9)
    assert torch.allclose(p.grad, p2.grad, rtol=1e-4, atol=1e-7)


def test_da_losses_large_map_takes_the_multi_cta_image_reduction():
    """Above 2^16 cells the image head is reduced over many CTAs before the single-CTA instance
    part; the three losses and both gradients still follow faster_rcnn.py:181-196."""
    import tlod_b200
    g = torch.Generator().manual_seed(41)
    score = torch.randn(4, 2, 150, 300, generator=g)  # 180 000 cells
    prob = torch.sigmoid(torch.randn(1024, 1, generator=g))
    for d in (0, 1):
        s = score.to(DEV).requires_grad_(True)
        p = prob.to(DEV).requires_grad_(True)
        img, ins, cst = tlod_b200.da_losses(s, p, d)
        (0.1 * (img + ins + cst)).backward()
        s2 = score.to(DEV).requires_grad_(True)
        p2 = prob.to(DEV).requires_grad_(True)
        lab = torch.full((4, 150, 300), d, dtype=torch.long, device=DEV)
        ref_img = torch.nn.functional.nll_loss(torch.log_softmax(s2, 1), lab)
        ref_ins = torch.nn.BCELoss()(p2, torch.full_like(p2, float(d)))
        cons = torch.softmax(s2, 1)[:, d].mean().detach()
        ref_cst = torch.nn.MSELoss(reduction="sum")(p2, cons.repeat(p2.size()))
        (0.1 * (ref_img + ref_ins + ref_cst)).backward()
        assert torch.allclose(img, ref_img, rtol=1e-5) and torch.allclose(ins, ref_ins, rtol=1e-5)
        assert torch.allclose(cst, ref_cst, rtol=1e-4)
        assert torch.allclose(s.grad, s2.grad, rtol=1e-4, atol=1e-10)
        assert torch.allclose(p.grad, p2.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("d,batch", [(1, 1), (0, 8)])
def test_image_da_losses_maf_three_levels(d, batch):
    """MAF / PT-MAF: conv3 (B,2,150,300), conv4 (B,2,75,150), conv5 (B,2,37,75) heads in one launch
    each way against the literal torch expressions of lib/MAF/faster_rcnn.py:188-205, including
    the instance-level CrossEntropyLoss over (R, 2) logits (:207-213) as a fourth 'level'."""
    import tlod_b200
    g = torch.Generator().manual_seed(31 + d)
    shapes = [(batch, 2, 150, 300), (batch, 2, 75, 150), (batch, 2, 37, 75), (256 * batch, 2, 1, 1)]
    maps = [torch.randn(*s, generator=g) * 3 for s in shapes]
    mine = [m.to(DEV).requires_grad_(True) for m in maps]
    ref = [m.to(DEV).requires_grad_(True) for m in maps]
    losses = tlod_b200.image_da_losses(mine, d)
    ref_losses = []
    for r in ref[:3]:
        lab = torch.full((r.size(0), r.size(2), r.size(3)), d, dtype=torch.long, device=DEV)
        ref_losses.append(torch.nn.functional.nll_loss(torch.log_softmax(r, 1), lab))
    lab_ins = torch.full((ref[3].size(0),), d, dtype=torch.long, device=DEV)
    ref_losses.append(torch.nn.CrossEntropyLoss()(ref[3].view(-1, 2), lab_ins))
    for a, b in zip(losses, ref_losses):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    w = torch.tensor([1.0, 0.5, 2.0, 0.1], device=DEV)
    (losses * w).sum().backward()
    sum(l * wi for l, wi in zip(ref_losses, w)).backward()
    for a, b in zip(mine, ref):
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-10)


def test_image_da_losses_atf_ignore_index():
    """ATF: F.nll_loss(log_softmax(score, 1), label, ignore_index=-1) with label maps that mix
    0 / 1 / -1 (lib/ATF/faster_rcnn.py:303-321); a level whose cells are all ignored gives NaN like torch."""
    import tlod_b200
    g = torch.Generator().manual_seed(41)
    shapes = [(2, 2, 150, 300), (2, 2, 75, 150), (2, 2, 37, 75)]
    maps = [torch.randn(*s, generator=g) for s in shapes]
    labels = [torch.randint(-1, 2, (s[0], s[2], s[3]), generator=g) for s in shapes]
    mine = [m.to(DEV).requires_grad_(True) for m in maps]
    ref = [m.to(DEV).requires_grad_(True) for m in maps]
    labs = [t.to(DEV) for t in labels]
    losses = tlod_b200.image_da_losses(mine, 1, labs, ignore_index=-1)
    ref_losses = [torch.nn.functional.nll_loss(torch.log_softmax(r, 1), t, ignore_index=-1) for r, t in zip(ref, labs)]
    for a, b in zip(losses, ref_losses):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    losses.sum().backward()
    sum(ref_losses).backward()
    for a, b in zip(mine, ref):
        assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-10)
        assert torch.equal(a.grad == 0, b.grad == 0)  # ignored cells: exactly zero in both
    # mixed: one level without a label map (domain label everywhere), one fully ignored
    from tlod_b200 import functional as F
    out = F.da_image_loss_forward([mine[2].detach(), mine[1].detach()], 0,
                                  [None, torch.full_like(labs[1], -1)], ignore_index=-1)
    lab0 = torch.zeros_like(labs[2])
    assert torch.allclose(out[0, 0], torch.nn.functional.nll_loss(torch.log_softmax(ref[2].detach(), 1), lab0), rtol=1e-5)
    assert torch.isnan(out[1, 0]) and out[1, 2].item() == 0
