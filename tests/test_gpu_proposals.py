"""GPU parity: NMS and the fused proposal layer through the C ABI against the oracle.
Keep indices, top-k order and the padded output must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle.synth import synth_rpn
from util import ANCHOR_RATIOS, ANCHOR_SCALES, bits_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _anchors():
    return orc.generate_anchors(scales=ANCHOR_SCALES, ratios=ANCHOR_RATIOS).astype(np.float32)


def _sorted_dets(n, seed, spread=900.0, jitter=False):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(n, generator=g) * spread
    y1 = torch.rand(n, generator=g) * spread * 0.5
    w = 16 + torch.rand(n, generator=g) * 300
    h = 16 + torch.rand(n, generator=g) * 200
    if jitter:  # many near-duplicates so that IoUs crowd the threshold
        base = torch.randint(0, max(n // 20, 1), (n,), generator=g)
        x1 = x1[base] + torch.rand(n, generator=g) * 8
        y1 = y1[base] + torch.rand(n, generator=g) * 8
        w, h = w[base], h[base]
    s, _ = torch.sort(torch.rand(n, generator=g), descending=True)
    return torch.stack([x1, y1, x1 + w, y1 + h, s], 1).contiguous()


@pytest.mark.parametrize("n,thresh,jitter", [(1, 0.7, False), (63, 0.7, False), (64, 0.5, True), (65, 0.3, True),
                                             (300, 0.3, True), (1000, 0.5, False), (6000, 0.7, True),
                                             (12000, 0.7, False), (12000, 0.7, True), (4097, 0.0, False),
                                             # 313 chunks, every phase, helper batches of up to 20 rounds; few
                                             # survivors per chunk at a low threshold
                                             (20000, 0.7, True), (8000, 0.1, True)])
def test_nms_keep_indices_bit_exact(n, thresh, jitter):
    from model.nms.nms_wrapper import nms
    dets = _sorted_dets(n, 100 + n, jitter=jitter)
    keep = nms(dets.to(DEV), thresh)
    assert keep.dtype == torch.int32 and keep.dim() == 2 and keep.size(1) == 1  # nms_gpu.py:8-12
    ref = orc.nms(dets.numpy(), thresh)
    assert np.array_equal(keep.view(-1).cpu().numpy(), ref)


def test_nms_threshold_boundary_cases():
    """Pairs whose IoU is within an ulp of the threshold, duplicates, degenerate boxes."""
    from tlod_b200 import functional as F
    rows = []
    for k in range(200):
        # two 100x100 boxes shifted by d: IoU = (100-d)*100 / (2e4 - (100-d)*100)
        d = k * 0.37
        rows.append([10.0, 20.0 + 3 * k * 0, 109.0, 119.0])
        rows.append([10.0 + d, 20.0, 109.0 + d, 119.0])
    rows.append([5.0, 5.0, 5.0, 5.0])
    rows.append([5.0, 5.0, 5.0, 5.0])
    rows.append([50.0, 50.0, 40.0, 40.0])  # inverted
    dets = torch.tensor(rows, dtype=torch.float32)
    dets = torch.cat([dets, torch.linspace(1, 0, dets.size(0)).view(-1, 1)], 1).contiguous()
    for thresh in (0.3, 0.5, 0.7, 0.8181818, 1.0 / 3.0, 0.9999):
        keep, num = F.nms_device(dets.to(DEV), thresh)
        k = int(num.item())
        assert np.array_equal(keep[:k].cpu().numpy(), orc.nms(dets.numpy(), thresh)), thresh
    # exact-boundary: feed the oracle's own IoU as the threshold (suppress iff IoU > thresh is false)
    a, b = dets[0:1, :4].numpy(), dets[5:6, :4].numpy()
    iou = orc.bbox_overlaps_batch(a, b.reshape(1, 1, 4))[0, 0, 0]
    pair = torch.cat([dets[0:1], dets[5:6]], 0).contiguous()
    for t in (np.nextafter(iou, 0, dtype=np.float32), iou, np.nextafter(iou, 1, dtype=np.float32)):
        keep, num = F.nms_device(pair.to(DEV), float(t))
        assert np.array_equal(keep[:int(num.item())].cpu().numpy(), orc.nms(pair.numpy(), float(t)))


def test_nms_max_keep_and_4col_stride():
    from tlod_b200 import functional as F
    dets = _sorted_dets(3000, 7, jitter=True)
    keep, num = F.nms_device(dets[:, :4].contiguous().to(DEV), 0.7, max_keep=100)
    ref = orc.nms(dets[:, :4].numpy(), 0.7, max_keep=100)
    assert int(num.item()) == len(ref) == 100
    assert np.array_equal(keep[:100].cpu().numpy(), ref)


def _run_proposals(B, H, W, seed, pre, post, thresh, im_info=None):
    from tlod_b200 import functional as F
    A = 12
    prob, deltas = synth_rpn(B, A, H, W, seed)
    if im_info is None:
        im_info = torch.tensor([[H * 16.0, W * 16.0, 1.0]] * B)
    anchors = torch.from_numpy(_anchors())
    rois, order, boxes, num = F.proposals(prob.to(DEV), deltas.to(DEV), im_info.to(DEV), anchors.to(DEV), 16, pre,
                                          post, thresh, return_debug=True)
    # exp as the reference's torch ops compute it on this device
    exp_d = torch.exp(deltas.to(DEV)).cpu().numpy()
    ref = orc.proposal_layer(prob.numpy(), deltas.numpy(), im_info.numpy(), anchors.numpy(), 16, pre, post, thresh,
                             exp_deltas=exp_d, return_debug=True)
    return (rois.cpu().numpy(), order.cpu().numpy(), boxes.cpu().numpy(), num.cpu().numpy()), ref


@pytest.mark.parametrize("B,H,W,pre,post,thresh", [
    (1, 37, 75, 12000, 2000, 0.7),   # cfg1 TRAIN
    (1, 37, 75, 6000, 300, 0.7),     # TEST
    (2, 38, 75, 12000, 2000, 0.7),   # ResNet-101 conv4 map
    (3, 19, 25, 3000, 500, 0.7),
    (4, 37, 75, 6000, 300, 0.3),     # cfg5 sweep
    (8, 37, 75, 6000, 300, 0.5),
    (1, 10, 10, 12000, 2000, 0.7),   # pre_nms_topN >= numel: no truncation (proposal_layer.py:138)
    (2, 6, 7, 200, 50, 0.7),
])
def test_proposal_layer_bit_exact(B, H, W, pre, post, thresh):
    mine, ref = _run_proposals(B, H, W, 3 + B, pre, post, thresh)
    assert np.array_equal(mine[1], ref[1]), "top-k order"
    assert bits_equal(mine[2], ref[2]), "decoded boxes"
    assert np.array_equal(mine[3], ref[3]), "survivor counts"
    assert bits_equal(mine[0], ref[0]), "rois"


def test_proposal_layer_per_image_clip_and_module():
    """_ProposalLayer module with cfg read at call time; per-image im_info."""
    from model.rpn.proposal_layer import _ProposalLayer
    from model.utils.config import cfg
    B, A, H, W = 3, 12, 19, 25
    prob, deltas = synth_rpn(B, A, H, W, 7)
    im_info = torch.tensor([[H * 16.0 + 8, W * 16.0, 0.5859375]] * B)
    im_info[1, 0] -= 40
    im_info[1, 1] -= 24
    layer = _ProposalLayer(16, ANCHOR_SCALES, ANCHOR_RATIOS)
    old = cfg.TEST.RPN_POST_NMS_TOP_N
    try:
        for post in (300, 128):  # ATF mutates this between calls (lib/ATF/faster_rcnn.py:260)
            cfg.TEST.RPN_POST_NMS_TOP_N = post
            rois = layer((prob.to(DEV), deltas.to(DEV), im_info.to(DEV), "TEST"))
            assert rois.shape == (B, post, 5)
            exp_d = torch.exp(deltas.to(DEV)).cpu().numpy()
            ref = orc.proposal_layer(prob.numpy(), deltas.numpy(), im_info.numpy(), _anchors(), 16, 6000, post,
                                     0.7, exp_deltas=exp_d)
            assert bits_equal(rois.cpu().numpy(), ref)
    finally:
        cfg.TEST.RPN_POST_NMS_TOP_N = old


def test_proposal_layer_tied_scores():
    """Duplicated scores: the lower anchor index comes first (stable descending order)."""
    from tlod_b200 import functional as F
    B, A, H, W = 2, 12, 12, 14
    prob, deltas = synth_rpn(B, A, H, W, 11)
    prob[:, A:] = torch.round(prob[:, A:] * 8) / 8
    im_info = torch.tensor([[H * 16.0, W * 16.0, 1.0]] * B)
    anchors = torch.from_numpy(_anchors())
    for pre in (200, 1000, 5000):
        rois, order, boxes, num = F.proposals(prob.to(DEV), deltas.to(DEV), im_info.to(DEV), anchors.to(DEV), 16, pre,
                                              50, 0.7, return_debug=True)
        exp_d = torch.exp(deltas.to(DEV)).cpu().numpy()
        ref = orc.proposal_layer(prob.numpy(), deltas.numpy(), im_info.numpy(), anchors.numpy(), 16, pre, 50, 0.7,
                                 exp_deltas=exp_d, return_debug=True)
        assert np.array_equal(order.cpu().numpy(), ref[1])
        assert bits_equal(rois.cpu().numpy(), ref[0])


def test_torch_cuda_sort_tie_order_is_the_contract():
    """The reference's torch.sort(scores, 1, True) on CUDA: equal scores keep the lower index first."""
    x = (torch.randint(0, 50, (2, 33300), generator=torch.Generator().manual_seed(0)).float() / 64).to(DEV)
    _, o = torch.sort(x, 1, True)
    _, os_ = torch.sort(x, dim=1, descending=True, stable=True)
    assert torch.equal(o, os_)


def test_kernel_expf_equals_torch_cuda_exp():
    """decode uses CUDA expf; torch.exp on CUDA (what lib/model runs) must give the same bits."""
    from tlod_b200 import functional as F
    g = torch.Generator().manual_seed(9)
    boxes = torch.tensor([[0.0, 0.0, 15.0, 15.0]]).repeat(4096, 1)
    deltas = torch.zeros(1, 4096, 4)
    deltas[0, :, 2] = torch.randn(4096, generator=g) * 2
    deltas[0, :, 3] = torch.randn(4096, generator=g) * 0.2
    out = F.bbox_transform_inv(boxes.to(DEV), deltas.to(DEV))
    d = deltas.to(DEV)
    w = torch.full((4096,), 16.0, device=DEV)
    pred_w = torch.exp(d[0, :, 2]) * w
    ref_x2 = (d[0, :, 0] * w + 8.0) + 0.5 * pred_w
    assert torch.equal(out[0, :, 2], ref_x2)


def test_full_size_properties_cfg5_batch64():
    """cfg5 at full size (64 images, 6000 -> 300): properties that do not need the CPU oracle."""
    from tlod_b200 import functional as F
    B, A, H, W = 64, 12, 37, 75
    prob, deltas = synth_rpn(B, A, H, W, 5)
    im_info = torch.tensor([[600.0, 1200.0, 0.5859375]] * B)
    anchors = torch.from_numpy(_anchors())
    rois, order, boxes, num = F.proposals(prob.to(DEV), deltas.to(DEV), im_info.to(DEV), anchors.to(DEV), 16, 6000,
                                          300, 0.7, return_debug=True)
    rois, order, boxes, num = rois.cpu(), order.cpu(), boxes.cpu(), num.cpu()
    assert torch.all(rois[:, :, 0] == torch.arange(B).view(B, 1).float())
    s = prob[:, A:].permute(0, 2, 3, 1).reshape(B, -1)
    picked = torch.gather(s, 1, order.long())
    assert torch.all(picked[:, :-1] >= picked[:, 1:])  # sortedness
    assert torch.all(picked[:, -1:] >= torch.kthvalue(s, s.size(1) - 6000 + 1, 1).values.view(B, 1))
    assert torch.all(num == 300)
    # survivors are mutually non-overlapping above the threshold: NMS is idempotent
    for b in (0, 31, 63):
        kept = torch.cat([rois[b, :, 1:], torch.linspace(1, 0, 300).view(-1, 1)], 1).contiguous()
        k, n = F.nms_device(kept.to(DEV), 0.7)
        assert int(n.item()) == 300 and torch.equal(k.cpu(), torch.arange(300, dtype=torch.int32))
    # two images against the oracle
    exp_d = torch.exp(deltas[:2].to(DEV)).cpu().numpy()
    ref = orc.proposal_layer(prob[:2].numpy(), deltas[:2].numpy(), im_info[:2].numpy(), anchors.numpy(), 16, 6000,
                             300, 0.7, exp_deltas=exp_d)
    # pre_nms_topN < numel holds for both batch sizes, so image 0/1 results are batch independent
    assert bits_equal(rois[:2].numpy(), ref)
