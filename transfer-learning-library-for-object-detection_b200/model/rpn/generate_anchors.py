"""generate_anchors (lib/model/rpn/generate_anchors.py:45-104): enumerate aspect
ratios, then scales, around the (0, 0, base-1, base-1) window.  Host numpy, run once
at layer construction.  Known answer (generate_anchors.py:19-37, MATLAB 1-based):
the default 9 anchors are that table minus one."""
import numpy as np


def _whctrs(box):
    w = box[2] - box[0] + 1
    h = box[3] - box[1] + 1
    return w, h, box[0] + 0.5 * (w - 1), box[1] + 0.5 * (h - 1)


def _windows(ws, hs, x_ctr, y_ctr):
    ws = np.asarray(ws, dtype=np.float64).reshape(-1, 1)
    hs = np.asarray(hs, dtype=np.float64).reshape(-1, 1)
    half_w, half_h = 0.5 * (ws - 1), 0.5 * (hs - 1)
    return np.hstack((x_ctr - half_w, y_ctr - half_h, x_ctr + half_w, y_ctr + half_h))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    base = np.array([0, 0, base_size - 1, base_size - 1], dtype=np.float64)
    w, h, x_ctr, y_ctr = _whctrs(base)
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    per_ratio = _windows(ws, hs, x_ctr, y_ctr)
    blocks = []
    for row in per_ratio:
        rw, rh, rx, ry = _whctrs(row)
        blocks.append(_windows(rw * scales, rh * scales, rx, ry))
    return np.vstack(blocks)
