"""MAF DRM rearrangement and the label-resize layers against literal
transcriptions of lib/MAF/drm.py:21-42 and lib/DAF/LabelResizeLayer.py:25-58."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _drm_rearrange_reference(low_dim, scale):
    # drm.py:24-42, verbatim semantics
    h_num = int(low_dim.size(2) / scale)
    w_num = int(low_dim.size(3) / scale)
    low_dim = low_dim[:, :, :int(scale * h_num), :int(scale * w_num)]
    sp = list(torch.chunk(low_dim, h_num, dim=2))
    for i in range(len(sp)):
        sp[i] = list(torch.chunk(sp[i], w_num, dim=3))
    for i in range(len(sp)):
        for j in range(len(sp[i])):
            sp[i][j] = sp[i][j].reshape(sp[i][j].size(0), sp[i][j].size(1) * scale * scale, 1, 1)
    for i in range(len(sp)):
        sp[i] = torch.cat(sp[i], dim=3)
    return torch.cat(sp, dim=2)


@pytest.mark.parametrize("B,C,H,W,s", [(2, 8, 12, 16, 4), (1, 5, 37, 75, 2), (2, 3, 30, 41, 4), (1, 64, 150, 300, 4)])
def test_space_to_depth_matches_drm_loops(B, C, H, W, s):
    import tlod_b200
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(B, C, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = _drm_rearrange_reference(xr, s)
    top = torch.randn(ref.shape, generator=g)
    ref.backward(top)
    xd = x.to(DEV).requires_grad_(True)
    out = tlod_b200.space_to_depth(xd, s)
    assert out.shape == ref.shape
    assert torch.equal(out.detach().cpu(), ref.detach())
    out.backward(top.to(DEV))
    assert torch.equal(xd.grad.cpu(), xr.grad)


def test_drm_module_forward_shape_and_values():
    from MAF.drm import DRM
    m = DRM(16, 4, 4).to(DEV)
    x = torch.randn(2, 16, 37, 75, device=DEV)
    y = m(x)
    assert y.shape == (2, 4 * 16, 9, 18)
    low = torch.relu(m.conv_low_dim(x))
    assert torch.equal(y.detach().cpu(), _drm_rearrange_reference(low.detach().cpu(), 4))


def test_label_resize_layers():
    from DAF.LabelResizeLayer import ImageLabelResizeLayer, InstanceLabelResizeLayer
    x = torch.randn(2, 2, 37, 75, device=DEV)
    need = torch.tensor([1.0, 0.0])
    y = ImageLabelResizeLayer()(x, need)
    assert y.shape == (2, 37, 75) and y.dtype == torch.long
    assert bool((y[0] == 1).all()) and bool((y[1] == 0).all())
    # instance labels: blocks of 256 rows per image, rows beyond keep the fill (np.ones in DAF)
    feats = torch.randn(600, 1, device=DEV)
    lab = InstanceLabelResizeLayer()(feats, need).cpu().numpy()
    ref = np.ones((600, 1), np.float32)
    for i, v in enumerate(need.numpy()):
        ref[i * 256:(i + 1) * 256] = v
    assert np.array_equal(lab, ref)
    lab0 = InstanceLabelResizeLayer(fill=0.0)(feats, torch.tensor([1.0])).cpu().numpy()
    assert lab0[:256].min() == 1.0 and lab0[256:].max() == 0.0
