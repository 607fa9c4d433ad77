"""RoICrop: oracle vs torch grid_sample on the CPU; CUDA path vs oracle and
vs the reference's own CUDA kernel on the GPU."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle.synth import synth_rois
from util import ROOT, features, rel_err

DEV = "cuda:0"
CASES = {
    "cfg3_small": (2, 8, 38, 75, 24, 14),    # (ib, C, H, W, rois per image, grid)
    "pool7": (1, 5, 20, 31, 16, 7),
    "three_images": (3, 4, 13, 17, 10, 14),
}


def _case(tag):
    ib, C, H, W, per, G = CASES[tag]
    feat = features(ib, C, H, W, 31)
    rois = synth_rois(ib * per, ib, 32, im_h=H * 16, im_w=W * 16)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous()
    # a few boxes reaching outside the image: corners outside the map contribute zero
    rois[0, 1:] = torch.tensor([-60.0, -40.0, 90.0, 70.0])
    rois[1, 1:] = torch.tensor([W * 16 - 50.0, H * 16 - 40.0, W * 16 + 80.0, H * 16 + 60.0])
    grid_xy = orc.affine_grid_gen(rois.numpy(), (H, W), G)
    grid_yx = np.ascontiguousarray(grid_xy[..., ::-1])
    return feat, rois, grid_yx, G


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_roi_crop_vs_torch_grid_sample(tag):
    """Pins the oracle's sampler on torch's grid_sample(align_corners=True, zeros padding), which
    is the same formula ((x + 1) (W - 1) / 2, corners outside contribute zero)."""
    feat, rois, grid_yx, G = _case(tag)
    ib, per = feat.shape[0], rois.shape[0] // feat.shape[0]
    out = orc.roi_crop_forward(feat.numpy(), grid_yx)
    grid_xy = torch.from_numpy(np.ascontiguousarray(grid_yx[..., ::-1]))
    rep = feat.unsqueeze(1).expand(ib, per, *feat.shape[1:]).reshape(-1, *feat.shape[1:])
    ref = torch.nn.functional.grid_sample(rep, grid_xy, mode="bilinear", padding_mode="zeros", align_corners=True)
    assert rel_err(out, ref.numpy()) <= 1e-5
    # the host-side affine grid (net_utils.py:142-164, torch-0.4 semantics = align_corners=True)
    from model.utils.net_utils import _affine_grid_gen
    g = _affine_grid_gen(rois, feat.shape[2:], G)
    assert np.allclose(g.numpy(), grid_xy.numpy(), atol=2e-6)
    # adjoint check of the oracle backward
    top = np.random.RandomState(1).randn(*out.shape).astype(np.float32)
    gin = orc.roi_crop_backward(top, grid_yx, feat.shape)
    lhs = float((out.astype(np.float64) * top).sum())
    rhs = float((gin.astype(np.float64) * feat.numpy()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_roi_crop_cuda_vs_oracle(tag):
    from tlod_b200 import functional as F
    feat, rois, grid_yx, G = _case(tag)
    gd = torch.from_numpy(grid_yx).to(DEV)
    out = F.roi_crop_forward(feat.to(DEV), gd)
    ref = orc.roi_crop_forward(feat.numpy(), grid_yx)
    assert rel_err(out.cpu().numpy(), ref) <= 1e-5
    top = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    grad = F.roi_crop_backward(top.to(DEV), gd, feat.shape)
    refg = orc.roi_crop_backward(top.numpy(), grid_yx, feat.shape)
    assert rel_err(grad.cpu().numpy(), refg) <= 1e-4


@pytest.mark.gpu
def test_roi_crop_module_autograd_and_crop_pool():
    """_RoICrop through autograd, and the crop branch of faster_rcnn.py:77-83 (grid 14 -> maxpool 7)."""
    from model.roi_crop.modules.roi_crop import _RoICrop
    from model.utils.net_utils import roi_crop_pool
    feat, rois, grid_yx, G = _case("cfg3_small")
    fd = feat.to(DEV).requires_grad_(True)
    crop = _RoICrop()
    y = roi_crop_pool(crop, fd, rois.to(DEV), G)
    pooled = torch.nn.functional.max_pool2d(y, 2, 2)
    assert pooled.shape == (rois.shape[0], feat.shape[1], G // 2, G // 2)
    ref = orc.roi_crop_forward(feat.numpy(), grid_yx)
    assert rel_err(y.detach().cpu().numpy(), ref) <= 1e-5
    top = torch.randn(y.shape, generator=torch.Generator().manual_seed(4))
    y.backward(top.to(DEV))
    refg = orc.roi_crop_backward(top.numpy(), grid_yx, feat.shape)
    assert rel_err(fd.grad.cpu().numpy(), refg) <= 1e-4


@pytest.mark.gpu
def test_reference_roi_crop_kernel_vs_oracle_and_tlod():
    """The reference's own roi_crop_cuda_kernel.cu (recompiled unmodified, oracle/_ref) on the B200."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
    if not os.path.exists(so):
        pytest.fail("oracle/_ref/libref_cuda.so missing: run `make -C oracle ref` where /root/reference exists "
                    "and ship oracle/_ref/ with the tree")
    ref_lib = ctypes.CDLL(so)
    if not hasattr(ref_lib, "BilinearSamplerBHWD_updateOutput_cuda_kernel"):
        pytest.fail("reference roi_crop kernel not in this libref_cuda.so: rebuild oracle/_ref")
    from tlod_b200 import functional as F
    feat, rois, grid_yx, G = _case("cfg3_small")
    ib, C, H, W = feat.shape
    ob = grid_yx.shape[0]
    fd, gd = feat.to(DEV), torch.from_numpy(grid_yx).to(DEV)
    out = torch.zeros(ob, C, G, G, device=DEV)
    I, P = ctypes.c_int, ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    # (oc, ow, oh, ob, ic, ih, iw, ib, input + 4 strides (b, c, h, w), grid + 4, output + 4, stream)
    ref_lib.BilinearSamplerBHWD_updateOutput_cuda_kernel(
        I(C), I(G), I(G), I(ob), I(C), I(H), I(W), I(ib),
        P(fd.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
        P(gd.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
        P(out.data_ptr()), I(C * G * G), I(G * G), I(G), I(1), P(st))
    torch.cuda.synchronize()
    mine = F.roi_crop_forward(fd, gd)
    orac = orc.roi_crop_forward(feat.numpy(), grid_yx)
    assert rel_err(out.cpu().numpy(), orac) <= 1e-5
    assert rel_err(mine.cpu().numpy(), out.cpu().numpy()) <= 1e-5
    top = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).to(DEV)
    gin = torch.zeros_like(fd)
    ggrid = torch.zeros_like(gd)
    ref_lib.BilinearSamplerBHWD_updateGradInput_cuda_kernel(
        I(C), I(G), I(G), I(ob), I(C), I(H), I(W), I(ib),
        P(fd.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
        P(gd.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
        P(gin.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
        P(ggrid.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
        P(top.data_ptr()), I(C * G * G), I(G * G), I(G), I(1), P(st))
    torch.cuda.synchronize()
    mine_g = F.roi_crop_backward(top, gd, feat.shape)
    assert rel_err(mine_g.cpu().numpy(), gin.cpu().numpy()) <= 1e-4
    assert float(ggrid.abs().max()) == 0.0  # the reference kernel never writes the grid gradient


POOL_FUSED_CASES = {
    "res_conv4": (2, 32, 38, 75, 40),      # bulk-staged planes (38 * 75 = 2 mod 4)
    "vgg_conv5": (3, 48, 37, 75, 30),      # register-staged planes, padded plane stride
    "small_map": (2, 16, 20, 31, 25),
    "many_rois": (1, 16, 38, 75, 600),
}


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(POOL_FUSED_CASES))
def test_roi_crop_max_pool_fused_vs_unfused_and_oracle(tag):
    """The fused 'crop' branch (affine grid -> RoICrop 14x14 -> max_pool2d(2, 2), one kernel each way)
    against (i) the oracle's RoICrop + torch's max_pool2d on the CPU and (ii) the unfused CUDA
    kernels; RoIs include sub-cell boxes, boxes larger than the map and boxes outside it."""
    from model.utils.net_utils import _affine_grid_gen, roi_crop_max_pool
    from tlod_b200 import functional as F
    ib, C, H, W, per = POOL_FUSED_CASES[tag]
    feat = features(ib, C, H, W, 51)
    g = torch.Generator().manual_seed(52)
    R = ib * per
    rois = synth_rois(R, ib, 53, im_h=H * 16, im_w=W * 16)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous()
    rois[0, 1:] = torch.tensor([-60.0, -40.0, 90.0, 70.0])
    rois[1, 1:] = torch.tensor([W * 16 - 50.0, H * 16 - 40.0, W * 16 + 80.0, H * 16 + 60.0])
    rois[2, 1:] = torch.tensor([100.0, 100.0, 103.0, 102.0])                   # sub-cell
    rois[3, 1:] = torch.tensor([-300.0, -200.0, W * 16 + 300.0, H * 16 + 200.0])  # larger than the map
    rois[4, 1:] = torch.tensor([W * 16 + 200.0, H * 16 + 100.0, W * 16 + 400.0, H * 16 + 300.0])  # outside
    fd = feat.to(DEV).requires_grad_(True)
    rd = rois.to(DEV)
    out = roi_crop_max_pool(fd, rd, 14)
    assert out.shape == (R, C, 7, 7)
    grid_xy = orc.affine_grid_gen(rois.numpy(), (H, W), 14)
    grid_yx = np.ascontiguousarray(grid_xy[..., ::-1])
    ref14 = torch.from_numpy(orc.roi_crop_forward(feat.numpy(), grid_yx))
    ref = torch.nn.functional.max_pool2d(ref14, 2, 2)
    assert rel_err(out.detach().cpu().numpy(), ref.numpy()) <= 1e-5
    # unfused CUDA kernels on the same device grid
    gxy = _affine_grid_gen(rd, (H, W), 14)
    gyx = torch.stack([gxy[..., 1], gxy[..., 0]], 3).contiguous()
    f2 = feat.to(DEV).requires_grad_(True)
    from tlod_b200.autograd import RoICropFunction
    out2 = torch.nn.functional.max_pool2d(RoICropFunction.apply(f2, gyx), 2, 2)
    assert rel_err(out.detach().cpu().numpy(), out2.detach().cpu().numpy()) <= 1e-5
    top = torch.randn(out.shape, generator=g).to(DEV)
    out.backward(top)
    out2.backward(top)
    # ties between the four samples of a window can route the gradient differently only where the
    # forward values tie; the synthetic maps (ReLU of a Gaussian) tie at exact zeros, which both
    # paths resolve to the first sample of the window
    assert rel_err(fd.grad.cpu().numpy(), f2.grad.cpu().numpy()) <= 1e-4
    # bit-identical forward over repeated launches (tile reuse / bulk-store ordering)
    gy, gx = gxy[:, :, 0, 1].contiguous(), gxy[:, 0, :, 0].contiguous()
    first, arg = F.roi_crop_pool_forward(fd.detach(), gy, gx)
    for _ in range(10):
        again, arg2 = F.roi_crop_pool_forward(fd.detach(), gy, gx)
        assert torch.equal(first, again) and torch.equal(arg, arg2)
    assert int(arg.max()) <= 3
