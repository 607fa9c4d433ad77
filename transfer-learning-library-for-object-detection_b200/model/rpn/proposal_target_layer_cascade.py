"""_ProposalTargetLayer (lib/model/rpn/proposal_target_layer_cascade.py:20-212).

Same constructor (nclasses) and forward(all_rois, gt_boxes, num_boxes) ->
(rois (B,P,5), labels (B,P), bbox_targets, bbox_inside_weights, bbox_outside_weights (B,P,4)).

Device work is two launches (tlod_roi_gt_assign: IoU + row max / argmax + class of the assigned gt,
no (B,N,K) tensor; tlod_proposal_targets: gather + label clamp + target encode / normalise +
weights).  The fg / bg sampling stays on the host with numpy's global RNG, consumed in exactly the
reference's order (:140-181: ``permutation`` for the foreground, ``rand`` for the background),
on one pinned copy of the (B, n) max-overlap array; the reference synchronises per image and
runs a Python double loop over the sampled RoIs (:83-91)."""
import numpy as np
import torch
import torch.nn as nn

import ctypes

from model.rpn.anchor_target_layer import _MT_WORDS, _numpy_mt19937_address
from model.utils.config import cfg
from tlod_b200 import functional as F
from tlod_b200._lib import lib


_native_ok = None


def _sample_numpy(mo, rois_per_image, fg_rois_per_image):
    """:140-181 with numpy's own calls (what the native sampler is checked against)."""
    B = mo.shape[0]
    keep = np.empty((B, rois_per_image), np.int32)
    fg_count = np.empty((B,), np.int32)
    fg_thresh = np.float32(cfg.TRAIN.FG_THRESH)
    bg_hi, bg_lo = np.float32(cfg.TRAIN.BG_THRESH_HI), np.float32(cfg.TRAIN.BG_THRESH_LO)
    for i in range(B):
        fg_inds = np.nonzero(mo[i] >= fg_thresh)[0]
        bg_inds = np.nonzero((mo[i] < bg_hi) & (mo[i] >= bg_lo))[0]
        fg_num, bg_num = fg_inds.shape[0], bg_inds.shape[0]
        if fg_num > 0 and bg_num > 0:
            fg_this = min(fg_rois_per_image, fg_num)
            rand_num = np.random.permutation(fg_num)
            fg_inds = fg_inds[rand_num[:fg_this]]
            bg_this = rois_per_image - fg_this
            rand_num = np.floor(np.random.rand(bg_this) * bg_num).astype(np.int64)
            bg_inds = bg_inds[rand_num]
        elif fg_num > 0 and bg_num == 0:
            rand_num = np.floor(np.random.rand(rois_per_image) * fg_num).astype(np.int64)
            fg_inds = fg_inds[rand_num]
            fg_this = rois_per_image
            bg_inds = bg_inds[:0]
        elif bg_num > 0 and fg_num == 0:
            rand_num = np.floor(np.random.rand(rois_per_image) * bg_num).astype(np.int64)
            bg_inds = bg_inds[rand_num]
            fg_this = 0
            fg_inds = fg_inds[:0]
        else:
            raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
        keep[i] = np.concatenate([fg_inds, bg_inds])
        fg_count[i] = fg_this
    return keep, fg_count


def _sample_native(mo, rois_per_image, fg_rois_per_image, out=None):
    """out: optional (keep (B, rois_per_image) int32, fg_count (B,) int32) C-contiguous numpy arrays
    the samples are written into (e.g. views of pinned upload buffers)."""
    B, n = mo.shape
    if out is None:
        keep = np.empty((B, rois_per_image), np.int32)
        fg_count = np.empty((B,), np.int32)
    else:
        keep, fg_count = out
        if keep.shape != (B, rois_per_image) or fg_count.shape != (B,) or keep.dtype != np.int32 or \
                fg_count.dtype != np.int32 or not keep.flags.c_contiguous or not fg_count.flags.c_contiguous:
            raise ValueError("out must be C-contiguous int32 arrays of shapes (B, rois_per_image) and (B,)")
    addr = _numpy_mt19937_address()
    with np.random.mtrand._rand._bit_generator.lock:
        rc = lib.tlod_proposal_sample_host(mo.ctypes.data, B, n, rois_per_image, fg_rois_per_image,
                                           float(cfg.TRAIN.FG_THRESH), float(cfg.TRAIN.BG_THRESH_HI),
                                           float(cfg.TRAIN.BG_THRESH_LO), addr,
                                           ctypes.cast(addr + 4 * _MT_WORDS, ctypes.POINTER(ctypes.c_int)),
                                           keep.ctypes.data, fg_count.ctypes.data)
    if rc != 0:
        raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
    return keep, fg_count


def _native_matches_numpy():
    """One-time self check on a private copy of the stream: same samples AND same stream position."""
    if _numpy_mt19937_address() is None:
        return False
    saved = np.random.get_state()
    try:
        rs = np.random.RandomState(4321)
        mo = (rs.rand(3, 700) ** 2).astype(np.float32)
        mo[2] = mo[2] * 0.4  # an image without foreground
        ok = True
        for rpi, fgr in ((256, 64), (128, 32)):
            np.random.seed(77)
            ka, fa = _sample_numpy(mo, rpi, fgr)
            sa = np.random.get_state()
            np.random.seed(77)
            kb, fb = _sample_native(mo, rpi, fgr)
            sb = np.random.get_state()
            ok = ok and np.array_equal(ka, kb) and np.array_equal(fa, fb) and sa[2] == sb[2] and \
                np.array_equal(sa[1], sb[1])
        return bool(ok)
    except Exception:  # noqa: BLE001
        return False
    finally:
        np.random.set_state(saved)


class _ProposalTargetLayer(nn.Module):
    def __init__(self, nclasses):
        super(_ProposalTargetLayer, self).__init__()
        self._num_classes = nclasses

    def forward(self, all_rois, gt_boxes, num_boxes):
        state = self.begin(all_rois, gt_boxes, num_boxes)
        keep, fg_count = self.sample(state)
        return self.finish(state, keep, fg_count)

    # forward() = finish(sample(begin())).  The three parts exist so that a caller can overlap other
    # GPU work with the host-side sampling, or capture the device parts in CUDA graphs:
    #   begin()  launches only: candidate assembly, tlod_roi_gt_assign, a pinned D2H copy + event;
    #   sample() waits for that event alone and draws the fg / bg samples on the host (numpy RNG);
    #   finish() uploads the sampled indices and launches tlod_proposal_targets.
    def begin(self, all_rois, gt_boxes, num_boxes, host=None):
        # :42-46 -- ground-truth boxes join the candidates (class column dropped, image index 0)
        gt_append = gt_boxes.new_zeros(gt_boxes.size())
        gt_append[:, :, 1:5] = gt_boxes[:, :, :4]
        all_rois = torch.cat([all_rois, gt_append], 1).contiguous()
        max_overlaps, assignment, labels = F.roi_gt_assign(all_rois, gt_boxes)
        if host is None:
            host = torch.empty(max_overlaps.shape, dtype=torch.float32).pin_memory()
        host.copy_(max_overlaps, non_blocking=True)
        copied = None
        if not torch.cuda.is_current_stream_capturing():
            copied = torch.cuda.Event()
            copied.record(torch.cuda.current_stream(gt_boxes.device))
        return {"all_rois": all_rois, "gt_boxes": gt_boxes, "assignment": assignment, "labels": labels,
                "host": host, "copied": copied}

    def sample(self, state, out=None):
        """Host-side fg / bg sampling, :140-181: same index order, same RNG calls.  Runs the native
        sampler (tlod_proposal_sample_host, on numpy's own MT19937 state) once it has been checked
        against the numpy transcription below on this installation.  out: optional (keep, fg_count)
        int32 numpy arrays to fill (see _sample_native)."""
        global _native_ok
        if state["copied"] is not None:
            state["copied"].synchronize()
        mo = state["host"].numpy()
        num_images = 1
        rois_per_image = int(cfg.TRAIN.BATCH_SIZE / num_images)
        fg_rois_per_image = int(np.round(cfg.TRAIN.FG_FRACTION * rois_per_image))
        fg_rois_per_image = 1 if fg_rois_per_image == 0 else fg_rois_per_image
        if _native_ok is None:
            _native_ok = _native_matches_numpy()
        if _native_ok and mo.dtype == np.float32 and mo.flags.c_contiguous:
            return _sample_native(mo, rois_per_image, fg_rois_per_image, out)
        keep, fg_count = _sample_numpy(mo, rois_per_image, fg_rois_per_image)
        if out is not None:
            out[0][...] = keep
            out[1][...] = fg_count
            return out
        return keep, fg_count

    def finish(self, state, keep, fg_count):
        """keep (B, rois_per_image) / fg_count (B,): numpy arrays from sample(), or int32 device tensors
        that already hold them (a caller replaying a captured graph uploads into static buffers)."""
        dev = state["gt_boxes"].device
        keep_d = keep if torch.is_tensor(keep) else torch.from_numpy(keep).to(dev, non_blocking=True)
        fg_d = fg_count if torch.is_tensor(fg_count) else torch.from_numpy(fg_count).to(dev, non_blocking=True)
        return F.proposal_targets(
            state["all_rois"], state["gt_boxes"], state["assignment"], state["labels"], keep_d, fg_d,
            cfg.TRAIN.BBOX_NORMALIZE_MEANS, cfg.TRAIN.BBOX_NORMALIZE_STDS, cfg.TRAIN.BBOX_INSIDE_WEIGHTS,
            cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED)

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""
        pass

    def reshape(self, bottom, top):
        pass
