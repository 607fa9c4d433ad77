"""GPU: the reference's OWN CUDA kernels (lib/model/*/src/*.cu recompiled unmodified for
sm_100a into oracle/_ref/libref_cuda*.so by oracle/Makefile) against the oracle and
against this library.  This is what pins the oracle's NMS / RoIPool / RoIAlign-backward
restatements.  Skipped when oracle/_ref was not shipped."""
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
needs_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_cuda.so not present")

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
