"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ANCHOR_SCALES = [4, 8, 16, 32]
ANCHOR_RATIOS = [0.5, 1, 2]


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind == "f":
        # identical up to the sign of zero
        return bool(np.array_equal(a, b, equal_nan=True))
    return bool(np.array_equal(a, b))


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative error' the north_star tolerances are stated in."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def features(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(B, C, H, W, generator=g))


def edge_rois(rois, H, W, scale):
    """Overwrite the first rows with the edge cases the reference's arithmetic has:
    last-pixel box (extrapolation branch), degenerate box, partly outside, inverted, beyond."""
    r = rois.clone()
    r[0, 1:] = torch.tensor([W / scale - 40, H / scale - 30, W / scale - 1, H / scale - 1])
    r[1, 1:] = torch.tensor([10.0, 10.0, 10.0, 10.0])
    r[2, 1:] = torch.tensor([-30.0, -20.0, 50.0, 40.0])
    r[3, 1:] = torch.tensor([100.0, 80.0, 60.0, 40.0])
    r[4, 1:] = torch.tensor([W / scale - 8, 0.0, W / scale + 30, H / scale + 20])
    return r
