"""GPU: the reference's OWN CUDA kernels (lib/model/*/src/*.cu recompiled unmodified for
sm_100a into oracle/_ref/libref_cuda*.so by oracle/Makefile) against the oracle and
against this library.  This is what pins the oracle's NMS / RoIPool / RoIAlign-backward
restatements.  oracle/_ref/ is git-ignored but travels to the GPU box with the tree (built by
`make -C oracle ref` where /root/reference exists); its absence on a CUDA box is a FAILURE, not a
skip: these are the strongest parity tests."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle.synth import synth_rois
from util import ROOT, bits_equal, edge_rois, features, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
REF_SO_NOFMA = os.path.join(ROOT, "oracle", "_ref", "libref_cuda_nofma.so")


def needs_ref(fn):
    """Fail (never skip) when the recompiled reference kernels were not shipped with the tree."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if not os.path.exists(REF_SO) or not os.path.exists(REF_SO_NOFMA):
            pytest.fail("oracle/_ref/libref_cuda{,_nofma}.so missing: run `make -C oracle ref` where "
                        "/root/reference exists and ship oracle/_ref/ with the tree")
        return fn(*a, **k)
    return wrapper


P = ctypes.c_void_p
F32 = ctypes.c_float
I = ctypes.c_int


def _ref(path=REF_SO):
    lib = ctypes.CDLL(path)
    lib.ROIAlignForwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, P, P, P]
    lib.ROIAlignBackwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, I, P, P, P]
    lib.ROIPoolForwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, P, P, P, P]
    lib.ROIPoolBackwardLaucher.argtypes = [P, F32, I, I, I, I, I, I, I, P, P, P, P]
    lib.nms_cuda_compute.argtypes = [P, P, P, I, I, F32]
    lib.nms_cuda_compute.restype = None
    return lib


def _ref_nms(lib, dets, thresh):
    n = dets.size(0)
    keep = torch.zeros(n, dtype=torch.int32, device=DEV)
    num = torch.zeros(1, dtype=torch.int32, device=DEV)
    torch.cuda.synchronize()
    lib.nms_cuda_compute(keep.data_ptr(), num.data_ptr(), dets.data_ptr(), n, 5, thresh)
    torch.cuda.synchronize()
    return keep[:int(num.item())].cpu().numpy()


@needs_ref
def test_reference_roi_align_kernels_vs_oracle_and_tlod():
    from tlod_b200 import functional as F
    lib = _ref()
    B, C, H, W, R, scale = 2, 32, 37, 75, 96, 1 / 16
    feat = features(B, C, H, W, 1)
    rois = edge_rois(synth_rois(R, B, 2), H, W, scale)
    fd, rd = feat.to(DEV), rois.to(DEV)
    out_ref = torch.zeros(R, C, 8, 8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    lib.ROIAlignForwardLaucher(fd.data_ptr(), scale, R, H, W, C, 8, 8, rd.data_ptr(), out_ref.data_ptr(), st)
    torch.cuda.synchronize()
    oracle_out = orc.roi_align_forward(feat.numpy(), rois.numpy(), 8, 8, scale)
    # the default-flag reference binary contracts p*bin+start into an FMA (the oracle follows the
    # source); the -fmad=false build of the same file is the source-level contract
    assert rel_err(out_ref.cpu().numpy(), oracle_out) <= 1e-5
    if os.path.exists(REF_SO_NOFMA):
        out_nofma = torch.zeros(R, C, 8, 8, device=DEV)
        _ref(REF_SO_NOFMA).ROIAlignForwardLaucher(fd.data_ptr(), scale, R, H, W, C, 8, 8, rd.data_ptr(),
                                                  out_nofma.data_ptr(), st)
        torch.cuda.synchronize()
        e = rel_err(out_nofma.cpu().numpy(), oracle_out)
        print("reference RoIAlign fwd (-fmad=false) vs oracle: rel err %.3g, bit-identical: %s"
              % (e, bits_equal(out_nofma.cpu().numpy(), oracle_out)))
        assert e <= 1e-6
    mine = F.roi_align_forward(fd, rd, 8, 8, scale)
    assert rel_err(mine.cpu().numpy(), out_ref.cpu().numpy()) <= 1e-5
    top = torch.randn(R, C, 8, 8, device=DEV)
    g_ref = torch.zeros(B, C, H, W, device=DEV)
    lib.ROIAlignBackwardLaucher(top.data_ptr(), scale, B, R, H, W, C, 8, 8, rd.data_ptr(), g_ref.data_ptr(), st)
    torch.cuda.synchronize()
    g_or = orc.roi_align_backward(top.cpu().numpy(), rois.numpy(), feat.shape, scale, accumulate_double=True)
    assert rel_err(g_ref.cpu().numpy(), g_or) <= 1e-5
    g_mine = F.roi_align_backward(top, rd, feat.shape, scale)
    assert rel_err(g_mine.cpu().numpy(), g_ref.cpu().numpy()) <= 1e-4


@needs_ref
def test_reference_roi_pool_kernels_vs_oracle_and_tlod():
    from tlod_b200 import functional as F
    lib = _ref()
    B, C, H, W, R, scale = 2, 16, 37, 75, 64, 1 / 16
    feat = features(B, C, H, W, 3)
    rois = edge_rois(synth_rois(R, B, 4), H, W, scale)
    fd, rd = feat.to(DEV), rois.to(DEV)
    out_ref = torch.zeros(R, C, 7, 7, device=DEV)
    arg_ref = torch.zeros(R, C, 7, 7, dtype=torch.int32, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    lib.ROIPoolForwardLaucher(fd.data_ptr(), scale, R, H, W, C, 7, 7, rd.data_ptr(), out_ref.data_ptr(),
                              arg_ref.data_ptr(), st)
    torch.cuda.synchronize()
    o, a = orc.roi_pool_forward(feat.numpy(), rois.numpy(), 7, 7, scale)
    assert bits_equal(out_ref.cpu().numpy(), o) and np.array_equal(arg_ref.cpu().numpy(), a)
    mo, ma = F.roi_pool_forward(fd, rd, 7, 7, scale)
    assert torch.equal(mo, out_ref) and torch.equal(ma, arg_ref)
    top = torch.randn(R, C, 7, 7, device=DEV)
    g_ref = torch.zeros(B, C, H, W, device=DEV)
    lib.ROIPoolBackwardLaucher(top.data_ptr(), scale, B, R, H, W, C, 7, 7, rd.data_ptr(), g_ref.data_ptr(),
                               arg_ref.data_ptr(), st)
    torch.cuda.synchronize()
    g_or = orc.roi_pool_backward(top.cpu().numpy(), a, rois.numpy(), feat.shape, scale)
    assert bits_equal(g_ref.cpu().numpy(), g_or)  # same summation order as the reference gather
    g_mine = F.roi_pool_backward(top, ma, rd, feat.shape, scale)
    assert rel_err(g_mine.cpu().numpy(), g_ref.cpu().numpy()) <= 1e-4


@needs_ref
@pytest.mark.parametrize("n,thresh", [(300, 0.3), (2000, 0.5), (6000, 0.7), (12000, 0.7)])
def test_reference_nms_vs_oracle_and_tlod(n, thresh):
    """Source-level contract = the -fmad=false build; the default-flag
    build contracts Sa+Sb into an FMA and may flip a borderline pair: agreement is reported."""
    from test_gpu_proposals import _sorted_dets
    from tlod_b200 import functional as F
    dets = _sorted_dets(n, 500 + n, jitter=True)
    dd = dets.to(DEV)
    oracle_keep = orc.nms(dets.numpy(), thresh)
    k, num = F.nms_device(dd, thresh)
    mine = k[:int(num.item())].cpu().numpy()
    assert np.array_equal(mine, oracle_keep)
    if os.path.exists(REF_SO_NOFMA):
        ref_nofma = _ref_nms(_ref(REF_SO_NOFMA), dd, thresh)
        assert np.array_equal(ref_nofma, oracle_keep), "oracle differs from the reference kernel (-fmad=false)"
    ref_default = _ref_nms(_ref(), dd, thresh)
    same = np.array_equal(ref_default, oracle_keep)
    print("n=%d thresh=%.1f: default-flag reference build %s the source-level result (%d kept)"
          % (n, thresh, "matches" if same else "DIFFERS from", len(oracle_keep)))


# ---------------------------------------------------------------------------------------------
# BASELINE configs at their own full shapes, against the reference's kernels on the same device
# ---------------------------------------------------------------------------------------------
def _proposal_rois(B, per_image, seed):
    """RoIs the way cfg1 / cfg2 get them: this library's proposal layer on the synthetic RPN
    outputs of SURVEY 8(d), the first `per_image` of every image."""
    from oracle.synth import synth_rpn
    from tlod_b200 import functional as F
    from util import ANCHOR_RATIOS, ANCHOR_SCALES
    prob, deltas = synth_rpn(B, 12, 37, 75, seed)
    info = torch.tensor([[600.0, 1200.0, 0.5859375]] * B)
    anchors = torch.from_numpy(orc.generate_anchors(scales=ANCHOR_SCALES, ratios=ANCHOR_RATIOS).astype(np.float32))
    rois = F.proposals(prob.to(DEV), deltas.to(DEV), info.to(DEV), anchors.to(DEV), 16, 12000, 2000, 0.7)
    return rois[:, :per_image, :].reshape(-1, 5).contiguous()


@needs_ref
@pytest.mark.parametrize("tag,B,C,H,W,per_image", [("cfg1", 1, 512, 37, 75, 128), ("cfg2_src", 2, 512, 37, 75, 256),
                                                   ("cfg2_tgt", 2, 512, 37, 75, 300),
                                                   ("cfg3", 8, 1024, 38, 75, 256)])
def test_full_shape_roi_align_vs_reference_kernels(tag, B, C, H, W, per_image):
    """RoIAlign(8, 8) forward (1e-5) and backward (1e-4), and RoIAlignAvg(7, 7) fused forward /
    backward, at the BASELINE shapes against the reference's recompiled kernels."""
    from tlod_b200 import functional as F
    lib = _ref()
    R = B * per_image
    if tag == "cfg3":
        rois = synth_rois(R, B, 41)
        rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(DEV)
    else:
        rois = _proposal_rois(B, per_image, 3)
    x = torch.relu(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(7))).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    y_ref = torch.zeros(R, C, 8, 8, device=DEV)
    lib.ROIAlignForwardLaucher(x.data_ptr(), 1 / 16, R, H, W, C, 8, 8, rois.data_ptr(), y_ref.data_ptr(), st)
    plan = F.roi_align_plan(rois, x.shape, 8, 8, 1 / 16)
    y = F.roi_align_forward(x, rois, 8, 8, 1 / 16, plan=plan)
    den = y_ref.abs().max().item()
    assert (y - y_ref).abs().max().item() <= 1e-5 * den
    y7 = F.roi_align_avg_forward(x, rois, 7, 7, 1 / 16, plan=plan)
    y7_ref = torch.nn.functional.avg_pool2d(y_ref, 2, 1)
    assert (y7 - y7_ref).abs().max().item() <= 1e-5 * y7_ref.abs().max().item()
    top = torch.randn(R, C, 8, 8, device=DEV)
    g_ref = torch.zeros_like(x)
    lib.ROIAlignBackwardLaucher(top.data_ptr(), 1 / 16, B, R, H, W, C, 8, 8, rois.data_ptr(), g_ref.data_ptr(), st)
    g = F.roi_align_backward(top, rois, x.shape, 1 / 16, plan=plan)
    assert (g - g_ref).abs().max().item() <= 1e-4 * g_ref.abs().max().item()
    # RoIAlignAvg backward: the average's adjoint (torch autograd on the same device), then the reference kernel
    top7 = torch.randn(R, C, 7, 7, device=DEV)
    t8 = torch.zeros(R, C, 8, 8, device=DEV, requires_grad=True)
    torch.nn.functional.avg_pool2d(t8, 2, 1).backward(top7)
    g7_ref = torch.zeros_like(x)
    lib.ROIAlignBackwardLaucher(t8.grad.data_ptr(), 1 / 16, B, R, H, W, C, 8, 8, rois.data_ptr(),
                                g7_ref.data_ptr(), st)
    g7 = F.roi_align_avg_backward(top7, rois, x.shape, 1 / 16, plan=plan)
    assert (g7 - g7_ref).abs().max().item() <= 1e-4 * g7_ref.abs().max().item()


@needs_ref
def test_full_shape_roi_pool_cfg2_vs_reference_kernels():
    from tlod_b200 import functional as F
    lib = _ref()
    B, C, H, W, per_image = 2, 512, 37, 75, 256
    R = B * per_image
    rois = _proposal_rois(B, per_image, 3)
    x = torch.relu(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(8))).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    o_ref = torch.zeros(R, C, 7, 7, device=DEV)
    a_ref = torch.zeros(R, C, 7, 7, dtype=torch.int32, device=DEV)
    lib.ROIPoolForwardLaucher(x.data_ptr(), 1 / 16, R, H, W, C, 7, 7, rois.data_ptr(), o_ref.data_ptr(),
                              a_ref.data_ptr(), st)
    o, a = F.roi_pool_forward(x, rois, 7, 7, 1 / 16)
    assert torch.equal(o, o_ref) and torch.equal(a, a_ref)
    top = torch.randn(R, C, 7, 7, device=DEV)
    g_ref = torch.zeros_like(x)
    lib.ROIPoolBackwardLaucher(top.data_ptr(), 1 / 16, B, R, H, W, C, 7, 7, rois.data_ptr(), g_ref.data_ptr(),
                               a_ref.data_ptr(), st)
    g = F.roi_pool_backward(top, a, rois, x.shape, 1 / 16)
    assert (g - g_ref).abs().max().item() <= 1e-4 * g_ref.abs().max().item()


@needs_ref
def test_full_shape_roi_crop_cfg3_vs_reference_kernels():
    """BASELINE cfg3 (ii): RoICrop 14 x 14 on (8, 1024, 38, 75), 2048 RoIs, forward and backward."""
    from model.utils.net_utils import _affine_grid_gen
    from tlod_b200 import functional as F
    lib = ctypes.CDLL(REF_SO)
    B, C, H, W, R, G = 8, 1024, 38, 75, 2048, 14
    x = torch.relu(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(9))).to(DEV)
    rois = synth_rois(R, B, 41)
    rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous().to(DEV)
    grid_xy = _affine_grid_gen(rois, (H, W), G)
    gyx = torch.stack([grid_xy[..., 1], grid_xy[..., 0]], 3).contiguous()
    out = torch.zeros(R, C, G, G, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    lib.BilinearSamplerBHWD_updateOutput_cuda_kernel(
        I(C), I(G), I(G), I(R), I(C), I(H), I(W), I(B),
        P(x.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
        P(gyx.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
        P(out.data_ptr()), I(C * G * G), I(G * G), I(G), I(1), P(st))
    mine = F.roi_crop_forward(x, gyx)
    assert (mine - out).abs().max().item() <= 1e-5 * out.abs().max().item()
    del mine
    top = torch.randn(R, C, G, G, device=DEV)
    gin = torch.zeros_like(x)
    ggrid = torch.zeros_like(gyx)
    lib.BilinearSamplerBHWD_updateGradInput_cuda_kernel(
        I(C), I(G), I(G), I(R), I(C), I(H), I(W), I(B),
        P(x.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
        P(gyx.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
        P(gin.data_ptr()), I(C * H * W), I(H * W), I(W), I(1),
        P(ggrid.data_ptr()), I(G * G * 2), I(1), I(G * 2), I(2),
        P(top.data_ptr()), I(C * G * G), I(G * G), I(G), I(1), P(st))
    g = F.roi_crop_backward(top, gyx, x.shape)
    assert (g - gin).abs().max().item() <= 1e-4 * gin.abs().max().item()


@needs_ref
@pytest.mark.parametrize("batch", [1, 2, 4, 8, 16, 32, 64])
def test_cfg5_nms_sweep_every_image_vs_reference_kernel(batch):
    """BASELINE cfg5: TEST proposals 6000 -> 300 at IoU 0.3 / 0.5 / 0.7, EVERY image of the batch:
    the reference's nms_cuda_compute (-fmad=false build = the source-level formula) on the sorted,
    decoded boxes of each image must keep exactly the boxes tlod_proposals emits, in order."""
    from oracle.synth import synth_rpn
    from tlod_b200 import functional as F
    from util import ANCHOR_RATIOS, ANCHOR_SCALES
    lib = _ref(REF_SO_NOFMA)
    prob, deltas = synth_rpn(batch, 12, 37, 75, 3)
    info = torch.tensor([[600.0, 1200.0, 0.5859375]] * batch)
    anchors = torch.from_numpy(orc.generate_anchors(scales=ANCHOR_SCALES, ratios=ANCHOR_RATIOS).astype(np.float32))
    pd, dd, idv, ad = prob.to(DEV), deltas.to(DEV), info.to(DEV), anchors.to(DEV)
    for thresh in (0.3, 0.5, 0.7):
        rois, order, boxes, num = F.proposals(pd, dd, idv, ad, 16, 6000, 300, thresh, return_debug=True)
        rois, boxes, num = rois.cpu(), boxes.cpu(), num.cpu()
        for b in range(batch):
            dets = torch.cat([boxes[b], torch.ones(boxes.size(1), 1)], 1).contiguous().to(DEV)
            keep = _ref_nms(lib, dets, thresh)[:300]
            assert int(num[b]) == len(keep), (batch, thresh, b)
            expect = torch.zeros(300, 5)
            expect[:, 0] = b
            expect[:len(keep), 1:] = boxes[b][torch.from_numpy(keep.astype(np.int64))]
            assert torch.equal(rois[b], expect), (batch, thresh, b)
