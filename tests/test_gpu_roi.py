"""GPU parity: RoIAlign / RoIPool through the C ABI against the oracle.
Tolerances are BASELINE.json's: forward 1e-5, atomics/scatter backward 1e-4 (relative to
the largest reference magnitude); RoIPool values and argmax are exact."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle.synth import synth_rois
from util import bits_equal, edge_rois, features, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (B, C, H, W, R, AH, AW, scale): plane-resident forward (AW == 8 fast path, TMA-stored 8x8
# tiles, and general sampling 7x7 / 3x5 / 14x14), row-resident backward (AW == 8, C % 32 == 0,
# B * H <= 8192 row lists; 1-4 warps per row depending on the grid), and shapes that fall back
# to the generic kernels (channels % 16, planes too big for shared memory, too many row bins)
ALIGN_CASES = {
    "cfg1_vgg_conv5": (1, 64, 37, 75, 128, 8, 8, 1 / 16),
    "multi_image_res_conv4": (4, 32, 38, 75, 200, 8, 8, 1 / 16),
    "raw_7x7": (2, 16, 20, 31, 64, 7, 7, 1 / 16),
    "odd_grid_3x5": (2, 16, 13, 17, 40, 3, 5, 1 / 8),
    "channels_not_x16": (2, 5, 20, 31, 64, 8, 8, 1 / 16),
    "plane_too_big_for_smem": (1, 16, 150, 300, 24, 8, 8, 1 / 4),
    "aligned_14x14": (1, 16, 38, 75, 32, 14, 14, 1 / 16),
    "many_rois_one_image": (1, 16, 37, 75, 1300, 8, 8, 1 / 16),   # > one RoI-list refill
    "aligned_16x8": (2, 32, 38, 75, 48, 16, 8, 1 / 16),           # tallest tile of the planes path
    "full_planes_no_bands": (20, 256, 20, 31, 400, 8, 8, 1 / 16), # many (image, slab) pairs
    "bwd_partial_channel_group": (2, 160, 20, 31, 64, 8, 8, 1 / 16),  # 128 + 32 channels
    "bwd_one_warp_group": (3, 32, 38, 75, 96, 8, 8, 1 / 16),
    "tall_tile_6x8": (2, 64, 37, 75, 80, 6, 8, 1 / 16),
    "bwd_wide_rows": (1, 32, 60, 300, 40, 8, 8, 1 / 4),              # 38 KB of row per warp, 4 warps per row
    "bwd_too_many_row_bins": (300, 32, 30, 20, 400, 8, 8, 1 / 16),   # B * H > 8192: generic backward
    "bwd_one_warp_per_row": (8, 256, 38, 75, 300, 8, 8, 1 / 16),     # grid large enough for K = 1
    "bwd_long_row_lists": (1, 32, 12, 20, 900, 8, 8, 1 / 16),        # > 32-entry chunks, every ph per row
    "bwd_two_channels_per_lane": (16, 256, 38, 75, 200, 8, 8, 1 / 16),  # >= 2 waves of 64-channel warps
}


def _case(tag):
    B, C, H, W, R, AH, AW, scale = ALIGN_CASES[tag]
    feat = features(B, C, H, W, 11)
    rois = edge_rois(synth_rois(R, B, 12, im_h=int(H / scale), im_w=int(W / scale)), H, W, scale)
    return feat, rois, AH, AW, scale


@pytest.mark.parametrize("tag", list(ALIGN_CASES))
def test_roi_align_forward(tag):
    from tlod_b200 import functional as F
    feat, rois, AH, AW, scale = _case(tag)
    out = F.roi_align_forward(feat.to(DEV), rois.to(DEV), AH, AW, scale).cpu().numpy()
    ref = orc.roi_align_forward(feat.numpy(), rois.numpy(), AH, AW, scale)
    assert rel_err(out, ref) <= 1e-5
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-5)
    # exactly-zero pattern (out-of-range samples) must match
    assert np.array_equal(ref == 0, out == 0) or np.abs(out[ref == 0]).max() <= 1e-6


@pytest.mark.parametrize("tag", list(ALIGN_CASES))
def test_roi_align_backward(tag):
    from tlod_b200 import functional as F
    feat, rois, AH, AW, scale = _case(tag)
    g = torch.Generator().manual_seed(5)
    top = torch.randn(rois.size(0), feat.size(1), AH, AW, generator=g)
    grad = F.roi_align_backward(top.to(DEV), rois.to(DEV), feat.shape, scale).cpu().numpy()
    ref = orc.roi_align_backward(top.numpy(), rois.numpy(), feat.shape, scale, accumulate_double=True)
    assert rel_err(grad, ref) <= 1e-4
    ref32 = orc.roi_align_backward(top.numpy(), rois.numpy(), feat.shape, scale, accumulate_double=False)
    assert rel_err(grad, ref32) <= 1e-4


@pytest.mark.parametrize("tag", ["cfg1_vgg_conv5", "odd_grid_3x5", "bwd_partial_channel_group"])
def test_roi_align_generic_kernels_without_plan(tag):
    """plan == NULL: the generic (per-RoI CTA, atomics in the backward) kernels."""
    from tlod_b200 import functional as F
    feat, rois, AH, AW, scale = _case(tag)
    out = F.roi_align_forward(feat.to(DEV), rois.to(DEV), AH, AW, scale, use_plan=False).cpu().numpy()
    ref = orc.roi_align_forward(feat.numpy(), rois.numpy(), AH, AW, scale)
    assert rel_err(out, ref) <= 1e-5
    g = torch.Generator().manual_seed(5)
    top = torch.randn(rois.size(0), feat.size(1), AH, AW, generator=g)
    grad = F.roi_align_backward(top.to(DEV), rois.to(DEV), feat.shape, scale, use_plan=False).cpu().numpy()
    refg = orc.roi_align_backward(top.numpy(), rois.numpy(), feat.shape, scale, accumulate_double=True)
    assert rel_err(grad, refg) <= 1e-4


def test_roi_align_small_and_huge_rois_column_chain():
    """RoIs from sub-cell to whole-map size: every code of the backward column chain (same cell,
    shift by one, jump) and the all-jump fast path; one plan shared by forward and backward."""
    from tlod_b200 import functional as F
    B, C, H, W, scale = 2, 64, 37, 75, 1 / 16
    g = torch.Generator().manual_seed(21)
    R = 600
    wpx = torch.cat([torch.rand(200, generator=g) * 40, torch.rand(200, generator=g) * 200,
                     torch.rand(200, generator=g) * 1300])
    hpx = torch.cat([torch.rand(200, generator=g) * 30, torch.rand(200, generator=g) * 150,
                     torch.rand(200, generator=g) * 700])[torch.randperm(R, generator=g)]
    x1 = torch.rand(R, generator=g) * 1250 - 40
    y1 = torch.rand(R, generator=g) * 640 - 30
    rois = torch.stack([torch.randint(0, B, (R,), generator=g).float(), x1, y1, x1 + wpx, y1 + hpx], 1)
    feat = features(B, C, H, W, 22)
    top = torch.randn(R, C, 8, 8, generator=g)
    rd = rois.to(DEV)
    plan = F.roi_align_plan(rd, feat.shape, 8, 8, scale)
    out = F.roi_align_forward(feat.to(DEV), rd, 8, 8, scale, plan=plan).cpu().numpy()
    ref = orc.roi_align_forward(feat.numpy(), rois.numpy(), 8, 8, scale)
    assert rel_err(out, ref) <= 1e-5
    grad = F.roi_align_backward(top.to(DEV), rd, feat.shape, scale, plan=plan).cpu().numpy()
    refg = orc.roi_align_backward(top.numpy(), rois.numpy(), feat.shape, scale, accumulate_double=True)
    assert rel_err(grad, refg) <= 1e-4
    # bitwise reproducible: no atomics, fixed summation order
    grad2 = F.roi_align_backward(top.to(DEV), rd, feat.shape, scale, plan=plan).cpu().numpy()
    assert np.array_equal(grad, grad2)


@pytest.mark.parametrize("shape", [(37, 16, 8, 8), (1, 1, 8, 8), (1030, 3, 8, 8), (5, 7, 15, 15), (9, 4, 3, 6)])
def test_avgpool2x2_matches_torch(shape):
    """tlod_avgpool2x2_forward/backward == avg_pool2d(kernel_size=2, stride=1) and its adjoint."""
    from tlod_b200 import functional as F
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    xd = x.to(DEV)
    y = F.avgpool2x2_forward(xd)
    ref = torch.nn.functional.avg_pool2d(x, kernel_size=2, stride=1)
    assert y.shape == ref.shape
    assert rel_err(y.cpu().numpy(), ref.numpy()) <= 1e-6
    top = torch.randn(*ref.shape, generator=g)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.avg_pool2d(xr, kernel_size=2, stride=1).backward(top)
    gx = F.avgpool2x2_backward(top.to(DEV))
    assert rel_err(gx.cpu().numpy(), xr.grad.numpy()) <= 1e-6


AVG_CASES = ["cfg1_vgg_conv5", "multi_image_res_conv4", "channels_not_x16", "plane_too_big_for_smem",
             "many_rois_one_image", "full_planes_no_bands", "bwd_one_warp_per_row", "raw_7x7", "odd_grid_3x5"]


@pytest.mark.parametrize("tag", AVG_CASES)
def test_roi_align_avg_forward_backward(tag):
    """RoIAlignAvg in one call (fused 8x8 -> 7x7 kernel: bulk-staged and register-staged planes,
    run-time and compile-time width; composed path for every other shape) against
    avg_pool2d(oracle RoIAlign, 2, 1) and the adjoint chain."""
    from tlod_b200 import functional as F
    feat, rois, AH, AW, scale = _case(tag)
    fd, rd = feat.to(DEV), rois.to(DEV)
    plan = F.roi_align_plan(rd, feat.shape, AH, AW, scale)
    out = F.roi_align_avg_forward(fd, rd, AH - 1, AW - 1, scale, plan=plan)
    ref_s = torch.from_numpy(orc.roi_align_forward(feat.numpy(), rois.numpy(), AH, AW, scale))
    ref = torch.nn.functional.avg_pool2d(ref_s, 2, 1)
    assert out.shape == ref.shape
    assert rel_err(out.cpu().numpy(), ref.numpy()) <= 1e-5
    assert np.allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)
    top = torch.randn(ref.shape, generator=torch.Generator().manual_seed(6))
    grad = F.roi_align_avg_backward(top.to(DEV), rd, feat.shape, scale, plan=plan).cpu().numpy()
    r = ref_s.clone().requires_grad_(True)
    torch.nn.functional.avg_pool2d(r, 2, 1).backward(top)
    refg = orc.roi_align_backward(r.grad.numpy(), rois.numpy(), feat.shape, scale, accumulate_double=True)
    assert rel_err(grad, refg) <= 1e-4
    # without a plan the same entry point composes the generic kernels
    out2 = F.roi_align_avg_forward(fd, rd, AH - 1, AW - 1, scale, plan=None)
    assert rel_err(out2.cpu().numpy(), ref.numpy()) <= 1e-5


def test_roi_align_avg_fused_edge_geometry_and_guard_band():
    """Fused RoIAlignAvg on sub-cell to whole-map RoIs (every row-reuse code: same pair, shifted by
    one, jump; clamped last rows; samples outside the map), an invalid image index, and the bytes
    around the caller's output untouched by the bulk stores."""
    from tlod_b200 import functional as F
    from tlod_b200._lib import check, lib
    B, C, H, W, scale = 3, 48, 38, 75, 1 / 16
    g = torch.Generator().manual_seed(77)
    R = 700
    wpx = torch.cat([torch.rand(250, generator=g) * 40, torch.rand(250, generator=g) * 250,
                     torch.rand(200, generator=g) * 1400])
    hpx = torch.cat([torch.rand(250, generator=g) * 30, torch.rand(250, generator=g) * 160,
                     torch.rand(200, generator=g) * 800])[torch.randperm(R, generator=g)]
    x1 = torch.rand(R, generator=g) * 1250 - 40
    y1 = torch.rand(R, generator=g) * 650 - 30
    rois = torch.stack([torch.randint(0, B, (R,), generator=g).float(), x1, y1, x1 + wpx, y1 + hpx], 1)
    rois[5, 0] = 7.0    # invalid image index -> zeros
    rois[6, 0] = -2.0
    feat = features(B, C, H, W, 78)
    fd, rd = feat.to(DEV), rois.to(DEV)
    PAD, SENT = 4096, 4321.0
    n_out = R * C * 49
    big = torch.full((n_out + 2 * PAD,), SENT, dtype=torch.float32, device=DEV)
    out = big[PAD:PAD + n_out]
    plan = F.roi_align_plan(rd, feat.shape, 8, 8, scale)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.tlod_roi_align_avg_forward(fd.data_ptr(), rd.data_ptr(), out.data_ptr(), B, C, H, W, R, 7, 7, scale,
                                         plan.data_ptr(), plan.numel(), None, 0, st), "avg forward (fused, no scratch)")
    torch.cuda.synchronize()
    assert bool((big[:PAD] == SENT).all()) and bool((big[PAD + n_out:] == SENT).all())
    out = out.view(R, C, 7, 7).cpu().numpy()
    assert np.all(out[5] == 0) and np.all(out[6] == 0)
    keep = [i for i in range(R) if i not in (5, 6)]
    ref_s = torch.from_numpy(orc.roi_align_forward(feat.numpy(), rois.numpy()[keep], 8, 8, scale))
    ref = torch.nn.functional.avg_pool2d(ref_s, 2, 1).numpy()
    assert rel_err(out[keep], ref) <= 1e-5
    # 30 back-to-back launches: bit-identical (tile reuse / bulk-store ordering races would show)
    first = F.roi_align_avg_forward(fd, rd, 7, 7, scale, plan=plan)
    for _ in range(30):
        assert torch.equal(first, F.roi_align_avg_forward(fd, rd, 7, 7, scale, plan=plan))


def test_roi_align_invalid_batch_index_gives_zeros():
    from tlod_b200 import functional as F
    feat, rois, AH, AW, scale = _case("multi_image_res_conv4")
    rois[7, 0] = 9.0
    rois[8, 0] = -1.0
    out = F.roi_align_forward(feat.to(DEV), rois.to(DEV), AH, AW, scale).cpu().numpy()
    assert np.all(out[7] == 0) and np.all(out[8] == 0)
    keep = [i for i in range(rois.size(0)) if i not in (7, 8)]
    ref = orc.roi_align_forward(feat.numpy(), rois.numpy()[keep], AH, AW, scale)
    assert rel_err(out[keep], ref) <= 1e-5


def test_roi_align_unsorted_and_empty_images():
    """RoIs in arbitrary image order, one image without any RoI."""
    from tlod_b200 import functional as F
    feat = features(5, 32, 37, 75, 3)
    rois = synth_rois(300, 5, 4)
    rois[rois[:, 0] == 2, 0] = 4.0  # image 2 gets nothing
    out = F.roi_align_forward(feat.to(DEV), rois.to(DEV), 8, 8, 1 / 16).cpu().numpy()
    ref = orc.roi_align_forward(feat.numpy(), rois.numpy(), 8, 8, 1 / 16)
    assert rel_err(out, ref) <= 1e-5
    top = torch.randn(300, 32, 8, 8, generator=torch.Generator().manual_seed(1))
    grad = F.roi_align_backward(top.to(DEV), rois.to(DEV), feat.shape, 1 / 16).cpu().numpy()
    refg = orc.roi_align_backward(top.numpy(), rois.numpy(), feat.shape, 1 / 16, accumulate_double=True)
    assert rel_err(grad, refg) <= 1e-4
    assert np.all(grad[2] == 0)


def test_roi_align_modules_and_autograd():
    """RoIAlignAvg / RoIAlignMax / ROIAlign alias through autograd (modules/roi_align.py:26-42)."""
    from model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    from model.roi_layers import ROIAlign
    feat, rois, _, _, scale = _case("cfg1_vgg_conv5")
    ref8 = torch.from_numpy(orc.roi_align_forward(feat.numpy(), rois.numpy(), 8, 8, scale))
    fd = feat.to(DEV).requires_grad_(True)
    avg = RoIAlignAvg(7, 7, scale)(fd, rois.to(DEV))
    assert avg.shape == (rois.size(0), feat.size(1), 7, 7)
    assert rel_err(avg.detach().cpu().numpy(), torch.nn.functional.avg_pool2d(ref8, 2, 1).numpy()) <= 1e-5
    mx = RoIAlignMax(7, 7, scale)(fd, rois.to(DEV))
    assert rel_err(mx.detach().cpu().numpy(), torch.nn.functional.max_pool2d(ref8, 2, 1).numpy()) <= 1e-5
    raw = RoIAlign(8, 8, scale)(fd, rois.to(DEV))
    assert rel_err(raw.detach().cpu().numpy(), ref8.numpy()) <= 1e-5
    alias = ROIAlign((7, 7), scale, 0)(fd, rois.to(DEV))
    assert torch.equal(alias, avg)
    top = torch.randn(avg.shape, generator=torch.Generator().manual_seed(2))
    avg.backward(top.to(DEV))
    # reference chain on CPU: avg_pool2d backward, then the oracle's RoIAlign backward
    r8 = ref8.clone().requires_grad_(True)
    torch.nn.functional.avg_pool2d(r8, 2, 1).backward(top)
    refg = orc.roi_align_backward(r8.grad.numpy(), rois.numpy(), feat.shape, scale, accumulate_double=True)
    assert rel_err(fd.grad.cpu().numpy(), refg) <= 1e-4


POOL_CASES = {
    "vgg_conv5_7x7": (2, 32, 37, 75, 96, 7, 7, 1 / 16),
    "res_conv4_7x7_many": (3, 64, 38, 75, 700, 7, 7, 1 / 16),      # plane-resident, several parts per unit
    "one_image_few_rois": (1, 16, 37, 75, 9, 7, 7, 1 / 16),
    "channels_x4_not_x16": (2, 20, 20, 31, 64, 7, 7, 1 / 16),       # generic forward, plane backward (NC = 4)
    "odd_channels": (2, 7, 20, 31, 64, 7, 7, 1 / 16),               # NC = 1 backward
    "cfg2_width": (2, 512, 37, 75, 512, 7, 7, 1 / 16),
    "pa_atf_conv3_stride4": (1, 8, 150, 300, 20, 7, 7, 1 / 4),
    "pa_atf_conv4_stride8": (1, 16, 75, 150, 20, 7, 7, 1 / 8),
    "odd_3x5": (3, 5, 13, 17, 40, 3, 5, 1 / 8),
}


@pytest.mark.parametrize("tag", list(POOL_CASES))
def test_roi_pool_forward_backward(tag):
    from tlod_b200 import functional as F
    B, C, H, W, R, PH, PW, scale = POOL_CASES[tag]
    feat = features(B, C, H, W, 21)
    rois = edge_rois(synth_rois(R, B, 22, im_h=int(H / scale), im_w=int(W / scale)), H, W, scale)
    out, arg = F.roi_pool_forward(feat.to(DEV), rois.to(DEV), PH, PW, scale)
    ref, ref_arg = orc.roi_pool_forward(feat.numpy(), rois.numpy(), PH, PW, scale)
    assert bits_equal(out.cpu().numpy(), ref)
    assert np.array_equal(arg.cpu().numpy(), ref_arg)
    top = torch.randn(out.shape, generator=torch.Generator().manual_seed(23))
    grad = F.roi_pool_backward(top.to(DEV), arg, rois.to(DEV), feat.shape, scale).cpu().numpy()
    refg = orc.roi_pool_backward(top.numpy(), ref_arg, rois.numpy(), feat.shape, scale)
    assert rel_err(grad, refg) <= 1e-4
    assert np.array_equal(grad == 0, refg == 0) or np.abs(grad[refg == 0]).max() < 1e-5


def test_roi_pool_invalid_image_index_and_ties():
    """RoIs with an image index outside [0, B) give 0 / -1; all-equal windows (ReLU zeros) pick the
    first cell in row-major order; repeated launches are bit-identical in the forward."""
    from tlod_b200 import functional as F
    B, C, H, W, R, scale = 2, 32, 37, 75, 80, 1 / 16
    feat = features(B, C, H, W, 91)
    feat[:, ::2] = 0.0                       # whole planes of ties
    feat[:, 1::4, 10:20, 30:50] = 1.0        # plateaus
    rois = edge_rois(synth_rois(R, B, 92), H, W, scale)
    rois[9, 0] = 5.0
    rois[10, 0] = -1.0
    out, arg = F.roi_pool_forward(feat.to(DEV), rois.to(DEV), 7, 7, scale)
    keep = [i for i in range(R) if i not in (9, 10)]
    ref, ref_arg = orc.roi_pool_forward(feat.numpy(), rois.numpy()[keep], 7, 7, scale)
    assert bits_equal(out.cpu().numpy()[keep], ref)
    assert np.array_equal(arg.cpu().numpy()[keep], ref_arg)
    assert np.all(out.cpu().numpy()[[9, 10]] == 0) and np.all(arg.cpu().numpy()[[9, 10]] == -1)
    out2, arg2 = F.roi_pool_forward(feat.to(DEV), rois.to(DEV), 7, 7, scale)
    assert torch.equal(out, out2) and torch.equal(arg, arg2)
    top = torch.randn(out.shape, generator=torch.Generator().manual_seed(93))
    grad = F.roi_pool_backward(top.to(DEV), arg, rois.to(DEV), feat.shape, scale).cpu().numpy()
    refg = orc.roi_pool_backward(top.numpy()[keep], ref_arg, rois.numpy()[keep], feat.shape, scale)
    assert rel_err(grad, refg) <= 1e-4


def test_roi_pool_special_values_and_kernel_variants(monkeypatch):
    """Signed zeros, -FLT_MAX, -inf and NaN cells (the reference's strict `>` from -FLT_MAX never
    takes a NaN, -inf or -FLT_MAX cell and keeps the first of +0 / -0), RoIs reaching over the map's
    border and bins from 1 to 11 columns wide: the plane-resident forward and the tabulated backward
    against the oracle, and against the generic kernels bit for bit (forward) / to rounding (backward)."""
    from tlod_b200 import functional as F
    B, C, H, W, R, scale = 2, 16, 37, 75, 160, 1 / 16
    g = torch.Generator().manual_seed(97)
    palette = torch.tensor([0.0, -0.0, 1.0, 1.0, 2.5, -3.0, -3.4028234663852886e38, float("-inf"), float("nan")])
    feat = palette[torch.randint(0, len(palette), (B, C, H, W), generator=g)]
    feat[:, :4] = torch.randn(B, 4, H, W, generator=g)
    rois = edge_rois(synth_rois(R, B, 98), H, W, scale)
    rois[0] = torch.tensor([0, -200.0, -100.0, 1500.0, 700.0])      # wider than the map: 11-column bins, clipped
    rois[1] = torch.tensor([1, 1100.0, 500.0, 1400.0, 800.0])       # mostly outside: empty bins
    rois[2] = torch.tensor([0, 300.0, 200.0, 310.0, 210.0])         # one cell: every bin the same cell
    rois[3] = torch.tensor([1, 0.0, 0.0, 1199.0, 599.0])            # the whole map
    fd, rd = feat.to(DEV), rois.to(DEV)
    out, arg = F.roi_pool_forward(fd, rd, 7, 7, scale)
    ref, ref_arg = orc.roi_pool_forward(feat.numpy(), rois.numpy(), 7, 7, scale)
    assert bits_equal(out.cpu().numpy(), ref)
    assert np.array_equal(arg.cpu().numpy(), ref_arg)
    top = torch.randn(out.shape, generator=g)
    grad = F.roi_pool_backward(top.to(DEV), arg, rd, feat.shape, scale)
    refg = orc.roi_pool_backward(top.numpy(), ref_arg, rois.numpy(), feat.shape, scale)
    assert rel_err(grad.cpu().numpy(), refg) <= 1e-4
    monkeypatch.setenv("TLOD_DISABLE_POOL_PLANES", "1")
    monkeypatch.setenv("TLOD_DISABLE_POOL_TAB", "1")
    out_g, arg_g = F.roi_pool_forward(fd, rd, 7, 7, scale)
    grad_g = F.roi_pool_backward(top.to(DEV), arg, rd, feat.shape, scale)
    assert torch.equal(out.view(torch.int32), out_g.view(torch.int32)) and torch.equal(arg, arg_g)
    assert rel_err(grad.cpu().numpy(), grad_g.cpu().numpy()) <= 1e-5


def test_roi_pool_module_autograd():
    from model.roi_pooling.modules.roi_pool import _RoIPooling
    B, C, H, W, R, PH, PW, scale = POOL_CASES["vgg_conv5_7x7"]
    feat = features(B, C, H, W, 31)
    rois = synth_rois(R, B, 32)
    fd = feat.to(DEV).requires_grad_(True)
    out = _RoIPooling(PH, PW, scale)(fd, rois.to(DEV))
    out.sum().backward()
    ref, ref_arg = orc.roi_pool_forward(feat.numpy(), rois.numpy(), PH, PW, scale)
    refg = orc.roi_pool_backward(np.ones_like(ref), ref_arg, rois.numpy(), feat.shape, scale)
    assert bits_equal(out.detach().cpu().numpy(), ref)
    assert rel_err(fd.grad.cpu().numpy(), refg) <= 1e-4


def test_full_size_adjointness_cfg3():
    """BASELINE cfg3 (8x1024x38x75, 2048 RoIs): too big for the CPU oracle in seconds, so check
    the size-independent property <fwd(x), g> == <x, bwd(g)> and partition of unity."""
    from tlod_b200 import functional as F
    torch.manual_seed(0)
    B, C, H, W, R = 8, 1024, 38, 75, 2048
    x = torch.relu(torch.randn(B, C, H, W, device=DEV))
    rois = synth_rois(R, B, 41).to(DEV)
    y = F.roi_align_forward(x, rois, 8, 8, 1 / 16)
    g = torch.randn_like(y)
    gx = F.roi_align_backward(g, rois, x.shape, 1 / 16)
    lhs = (y.double() * g.double()).sum().item()
    rhs = (x.double() * gx.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0) + 1e-3
    ones = F.roi_align_forward(torch.ones_like(x), rois, 8, 8, 1 / 16)
    nz = ones != 0
    assert (ones[nz] - 1).abs().max().item() <= 1e-5  # weights of every valid sample sum to 1
    # spot-check 16 RoIs x 8 channels of the big problem against the oracle
    idx = torch.arange(0, R, R // 16)[:16]
    sub = rois[idx].cpu()
    ref = orc.roi_align_forward(x[:, :8].cpu().numpy(), sub.numpy(), 8, 8, 1 / 16)
    assert rel_err(y[idx][:, :8].cpu().numpy(), ref) <= 1e-5


@pytest.mark.parametrize("tag", ["cfg1_vgg_conv5", "bwd_two_channels_per_lane", "bwd_wide_rows", "aligned_14x14"])
def test_roi_align_writes_stay_inside_the_callers_buffers(tag):
    """The C ABI works on caller-owned memory: plan, output and gradient are carved out of larger
    sentinel-filled allocations and the bytes around them must come back untouched (the TMA-stored
    forward tiles, the row-resident backward and the plan's row lists all compute their own
    addresses)."""
    from tlod_b200 import functional as F
    from tlod_b200._lib import check, lib
    B, C, H, W, R, AH, AW, scale = ALIGN_CASES[tag]
    feat, rois, AH, AW, scale = _case(tag)
    fd, rd = feat.to(DEV), rois.to(DEV)
    PAD = 4096  # floats; keeps the 128-byte alignment of the carved buffers
    SENT = 12345.0

    def carve(numel):
        big = torch.full((numel + 2 * PAD,), SENT, dtype=torch.float32, device=DEV)
        return big, big[PAD:PAD + numel]

    def untouched(big, numel):
        return bool((big[:PAD] == SENT).all()) and bool((big[PAD + numel:] == SENT).all())

    nplan = (lib.tlod_roi_align_plan_bytes(B, R) + 3) // 4
    plan_big, plan = carve(nplan)
    out_big, out = carve(R * C * AH * AW)
    grad_big, grad = carve(B * C * H * W)
    top = torch.randn(R, C, AH, AW, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    check(lib.tlod_roi_align_plan(rd.data_ptr(), B, H, W, R, AH, AW, scale, plan.data_ptr(), nplan * 4, st), "plan")
    check(lib.tlod_roi_align_forward(fd.data_ptr(), rd.data_ptr(), out.data_ptr(), B, C, H, W, R, AH, AW, scale,
                                     plan.data_ptr(), nplan * 4, st), "forward")
    check(lib.tlod_roi_align_backward(top.data_ptr(), rd.data_ptr(), grad.data_ptr(), B, C, H, W, R, AH, AW, scale,
                                      plan.data_ptr(), nplan * 4, st), "backward")
    torch.cuda.synchronize()
    assert untouched(plan_big, nplan) and untouched(out_big, out.numel()) and untouched(grad_big, grad.numel())
    ref = F.roi_align_forward(fd, rd, AH, AW, scale)
    assert torch.equal(out.view_as(ref), ref)
    refg = F.roi_align_backward(top, rd, feat.shape, scale)
    if AW == 8 and C % 32 == 0:  # row-resident backward: fixed summation order
        assert torch.equal(grad.view_as(refg), refg)
    else:                        # generic backward: fp32 atomics, order varies
        assert rel_err(grad.view_as(refg).cpu().numpy(), refg.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("shape", [(8, 1024, 38, 75, 2048), (2, 512, 37, 75, 512), (1, 64, 37, 75, 128)])
def test_roi_align_backward_is_bitwise_stable_over_repeated_launches(shape):
    """The row-resident backward has a fixed summation order and recycles its TMA stages right
    after reading them: thirty back-to-back launches must give bit-identical gradients (a stage
    refilled too early, or any other race, would show here)."""
    from tlod_b200 import functional as F
    B, C, H, W, R = shape
    rois = synth_rois(R, B, 41).to(DEV)
    g = torch.Generator().manual_seed(17)
    top = torch.randn(R, C, 8, 8, generator=g).to(DEV)
    plan = F.roi_align_plan(rois, (B, C, H, W), 8, 8, 1 / 16)
    first = F.roi_align_backward(top, rois, (B, C, H, W), 1 / 16, plan=plan)
    for _ in range(30):
        again = F.roi_align_backward(top, rois, (B, C, H, W), 1 / 16, plan=plan)
        assert torch.equal(first, again)
    fwd0 = None
    feat = features(B, C, H, W, 3).to(DEV)
    for _ in range(10):
        out = F.roi_align_forward(feat, rois, 8, 8, 1 / 16, plan=plan)
        fwd0 = out if fwd0 is None else fwd0
        assert torch.equal(fwd0, out)
