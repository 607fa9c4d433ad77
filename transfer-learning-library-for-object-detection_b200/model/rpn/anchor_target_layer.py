"""_AnchorTargetLayer (lib/model/rpn/anchor_target_layer.py:31-219).

Same constructor and forward((rpn_cls_score, gt_boxes, im_info, num_boxes)) -> list of
[labels (B,1,A*H,W), bbox_targets, bbox_inside_weights, bbox_outside_weights (B,4A,H,W)].

Device work is two fused passes (tlod_anchor_labels: IoU + per-gt max + label rules, no
(B,N,K) tensor; tlod_anchor_targets_finalize: targets + weights + unmap + permute).  The
random subsampling stays on the host with numpy's global RNG, consumed in exactly the
reference's order (:123-145), because the RNG stream position is data dependent: the (B, n)
label array makes one pinned round trip (the reference synchronises 2 + 2B times)."""
import ctypes
import threading

import numpy as np
import torch
import torch.nn as nn

from model.utils.config import cfg
from tlod_b200 import functional as F
from tlod_b200._lib import check, lib

from .generate_anchors import generate_anchors


def subsample_labels_numpy(lab, num_fg, rpn_batchsize):
    """anchor_target_layer.py:118-145 with numpy's own calls (the transcription the native
    sampler is checked against).  lab: (B, n) fp32 in {-1, 0, 1}, edited in place; returns the
    number of labels >= 0 of the last image (:156)."""
    i = 0
    for i in range(lab.shape[0]):
        fg_inds = np.nonzero(lab[i] == 1)[0]
        if fg_inds.shape[0] > num_fg:
            rand_num = np.random.permutation(fg_inds.shape[0])
            lab[i][fg_inds[rand_num[:fg_inds.shape[0] - num_fg]]] = -1
            n_fg_i = num_fg
        else:
            n_fg_i = fg_inds.shape[0]
        num_bg = rpn_batchsize - n_fg_i
        bg_inds = np.nonzero(lab[i] == 0)[0]
        if bg_inds.shape[0] > num_bg:
            rand_num = np.random.permutation(bg_inds.shape[0])
            lab[i][bg_inds[rand_num[:bg_inds.shape[0] - num_bg]]] = -1
    return int((lab[i] >= 0).sum())


_MT_WORDS = 624
_native_ok = None


_mt_address = (None, None)  # (bit generator object, address of its state)


def _numpy_mt19937_address():
    """Address of numpy's global legacy MT19937 state (key[624], pos) or None.  The state lives
    inside the bit-generator object (np.random.seed / set_state rewrite it in place), so the
    address is looked up once per bit-generator object (building the ctypes interface costs ~20 us)."""
    global _mt_address
    try:
        bitgen = np.random.mtrand._rand._bit_generator
        if _mt_address[0] is bitgen:
            return _mt_address[1]
        if type(bitgen).__name__ != "MT19937":
            return None
        addr = int(bitgen.ctypes.state_address)
        _mt_address = (bitgen, addr)
        return addr
    except Exception:  # noqa: BLE001 -- a numpy without the ctypes interface
        return None


def _subsample_native(lab, num_fg, rpn_batchsize, ahead=None):
    addr = _numpy_mt19937_address()
    nex = ctypes.c_int(0)
    with np.random.mtrand._rand._bit_generator.lock:
        check(lib.tlod_anchor_subsample_host_ahead(
            lab.ctypes.data, lab.shape[0], lab.shape[1], int(num_fg), int(rpn_batchsize), addr,
            ctypes.cast(addr + 4 * _MT_WORDS, ctypes.POINTER(ctypes.c_int)),
            None if ahead is None else ahead.ctypes.data, 0 if ahead is None else ahead.size // _MT_WORDS - 1,
            ctypes.byref(nex)), "tlod_anchor_subsample_host_ahead")
    return int(nex.value)


def pregenerate_stream(words, out=None):
    """MT19937 key blocks for about `words` future draws of numpy's global stream, generated now
    (tlod_mt_pregen) so that subsample_labels(..., ahead=...) does not have to while the step waits
    for it.  numpy's state is not touched; the blocks are ignored if anything draws from the stream
    before they are used.  Returns a uint32 array ((blocks + 1) * 624) or None (no native sampler)."""
    global _native_ok
    if _native_ok is None:
        _native_ok = _native_sampler_matches_numpy()
    if not _native_ok:
        return None
    blocks = max(1, -(-int(words) // _MT_WORDS))
    if out is None or out.size < (blocks + 1) * _MT_WORDS:
        out = np.empty((blocks + 1) * _MT_WORDS, np.uint32)
    ahead = out[:(blocks + 1) * _MT_WORDS]
    with np.random.mtrand._rand._bit_generator.lock:
        check(lib.tlod_mt_pregen(_numpy_mt19937_address(), ahead.ctypes.data, blocks), "tlod_mt_pregen")
    return ahead


def _native_sampler_matches_numpy():
    """One-time self check: the native sampler, working on numpy's live state, must leave the same
    labels AND the same stream position as numpy's own calls.  The global stream is restored."""
    if _numpy_mt19937_address() is None:
        return False
    saved = np.random.get_state()
    try:
        rs = np.random.RandomState(1234)
        lab = rs.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=(2, 3000), p=[0.1, 0.8, 0.1])
        a, b = lab.copy(), lab.copy()
        np.random.seed(99)
        ra = subsample_labels_numpy(a, 128, 256)
        sa = np.random.get_state()
        np.random.seed(99)
        rb = _subsample_native(b, 128, 256)
        sb = np.random.get_state()
        return bool(ra == rb and np.array_equal(a, b) and sa[2] == sb[2] and np.array_equal(sa[1], sb[1]))
    except Exception:  # noqa: BLE001
        return False
    finally:
        np.random.set_state(saved)


def subsample_labels(lab, num_fg, rpn_batchsize, ahead=None):
    """The random subsampling of anchor_target_layer.py:118-145 on the host copy of the labels,
    consuming numpy's global stream exactly like the reference.  Runs the native sampler
    (tlod_anchor_subsample_host, ~4x faster than numpy's shuffle: this loop is the longest host
    chain of a training step) once it has been checked against numpy on this installation."""
    global _native_ok
    if _native_ok is None:
        _native_ok = _native_sampler_matches_numpy()
    if _native_ok and lab.dtype == np.float32 and lab.flags.c_contiguous:
        return _subsample_native(lab, num_fg, rpn_batchsize, ahead)
    return subsample_labels_numpy(lab, num_fg, rpn_batchsize)


class _AnchorTargetLayer(nn.Module):
    def __init__(self, feat_stride, scales, ratios):
        super(_AnchorTargetLayer, self).__init__()
        self._feat_stride = feat_stride
        self._scales = scales
        self._anchors_np = generate_anchors(scales=np.array(scales), ratios=np.array(ratios)).astype(np.float32)
        self._anchors = torch.from_numpy(self._anchors_np)
        self._num_anchors = self._anchors.size(0)
        self._allowed_border = 0
        self._cache = {}
        self._streams = {}      # device index -> side stream of this layer
        self._pinned_pool = {}  # (device index, shape, dtype) -> free pinned staging buffers
        self._pool_lock = threading.Lock()
        self._ahead = threading.local()  # scratch for pregenerate_stream

    def _inside(self, feat_h, feat_w, lim_w, lim_h, device):
        """Inside-image anchors for this map size (:66-91): (anchors (n,4), inds (n,), inverse (total,))."""
        key = (feat_h, feat_w, lim_w, lim_h, str(device))
        hit = self._cache.get(key)
        if hit is None:
            sx = (np.arange(0, feat_w) * self._feat_stride).astype(np.float32)
            sy = (np.arange(0, feat_h) * self._feat_stride).astype(np.float32)
            gx, gy = np.meshgrid(sx, sy)
            shifts = np.stack((gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()), axis=1)
            all_anchors = (self._anchors_np[None, :, :] + shifts[:, None, :]).reshape(-1, 4)
            b = self._allowed_border
            keep = ((all_anchors[:, 0] >= -b) & (all_anchors[:, 1] >= -b) &
                    (all_anchors[:, 2] < lim_w + b) & (all_anchors[:, 3] < lim_h + b))
            inds = np.nonzero(keep)[0]
            inv = np.full((all_anchors.shape[0],), -1, np.int32)
            inv[inds] = np.arange(inds.shape[0], dtype=np.int32)
            hit = (torch.from_numpy(np.ascontiguousarray(all_anchors[inds])).to(device),
                   torch.from_numpy(inds).to(device), torch.from_numpy(inv).to(device))
            if len(self._cache) > 16:
                self._cache.clear()
            self._cache[key] = hit
        return hit

    def forward(self, input):
        return self.finish(self.begin(input))

    # forward() = finish(begin()).  The two halves exist so that a caller can enqueue other GPU
    # work between them: begin() only launches (label kernel + a pinned D2H copy on this layer's
    # own stream), finish() waits for that copy alone -- not for whatever else the caller has
    # queued meanwhile -- runs the host-side subsampling and launches the finalize kernel.
    def begin(self, input, im_hw=None):
        """im_hw: (height, width) of the first image in pixels if the caller already knows it on
        the host -- the layer then does not read im_info[0] back (a blocking D2H of three floats,
        which is what the reference's `long(im_info[0][1])` does, anchor_target_layer.py:86-87)."""
        rpn_cls_score, gt_boxes, im_info, num_boxes = input[0], input[1], input[2], input[3]
        height, width = rpn_cls_score.size(2), rpn_cls_score.size(3)
        dev = gt_boxes.device
        with self._pool_lock:
            stream = self._streams.get(dev.index)
            if stream is None:
                stream = self._streams[dev.index] = torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)
        stream.wait_stream(cur)  # inputs produced on the caller's stream
        with torch.cuda.stream(stream):
            if im_hw is None:
                # :86-87 -- the FIRST image's size, truncated to int, is used for the whole batch
                info0 = im_info[0].tolist()
                im_hw = (info0[0], info0[1])
            state = self.launch_labels(gt_boxes, height, width, im_hw)
            state["copied"] = torch.cuda.Event()
            state["copied"].record(stream)
            state["stream"] = stream
        for t in (gt_boxes, im_info):
            t.record_stream(stream)
        return state

    def launch_labels(self, gt_boxes, height, width, im_hw, labels_host=None):
        """The device half of begin() as pure launches on the CURRENT stream (capturable in a CUDA
        graph): IoU + label rules, and the pinned D2H copy of the labels.  labels_host: a pinned
        (B, n) buffer owned by the caller (a graph replays into it); by default one is taken from
        the layer's pool and returned to it by finish().  The caller of this method sets
        state["copied"] (an event recorded after these launches) and state["stream"]."""
        dev = gt_boxes.device
        anchors, inds_inside, inv_index = self._inside(height, width, int(im_hw[1]), int(im_hw[0]), dev)
        labels, argmax = F.anchor_labels(anchors, gt_boxes, cfg.TRAIN.RPN_NEGATIVE_OVERLAP,
                                         cfg.TRAIN.RPN_POSITIVE_OVERLAP, cfg.TRAIN.RPN_CLOBBER_POSITIVES)
        pooled = labels_host is None
        if pooled:
            # the pinned staging buffer travels with the returned state (begin() may be called again
            # before finish(), e.g. by DataParallel replicas sharing this module's attributes)
            labels_host = self._take_pinned(dev.index, tuple(labels.shape), labels.dtype)
        labels_host.copy_(labels, non_blocking=True)
        return {"labels": labels, "argmax": argmax, "anchors": anchors, "inv_index": inv_index,
                "gt_boxes": gt_boxes, "height": height, "width": width, "labels_host": labels_host,
                "pooled": pooled, "copied": None, "stream": None}

    def _take_pinned(self, dev_index, shape, dtype):
        with self._pool_lock:
            free = self._pinned_pool.setdefault((dev_index, shape, dtype), [])
            if free:
                return free.pop()
        return torch.empty(shape, dtype=dtype).pin_memory()

    def _give_pinned(self, dev_index, buf):
        with self._pool_lock:
            free = self._pinned_pool.setdefault((dev_index, tuple(buf.shape), buf.dtype), [])
            if len(free) < 8:
                free.append(buf)

    def finish(self, state):
        return self.finish_device(self.finish_host(state))

    def finish_host(self, state):
        """The host half of finish(): waits for the labels and subsamples them on numpy's stream.
        A caller with more host work that has to follow in the reference's RNG order (the
        proposal-target sampling) can do it before finish_device() queues the upload."""
        copied, labels_host = state["copied"], state["labels_host"]
        # while the device still computes the labels: the MT19937 blocks the subsampling will draw
        # from (a permutation of the ~n background anchors per image, 1.33 words per draw on average)
        scratch = self._ahead  # per thread: DataParallel replicas share this module's attributes
        ahead = getattr(scratch, "ready", None)  # prefetch_stream(): generated in earlier idle time
        scratch.ready = None
        if ahead is None:
            ahead = pregenerate_stream(1.4 * labels_host.numel(), getattr(scratch, "buf", None))
            if ahead is not None:
                scratch.buf = ahead.base if ahead.base is not None else ahead
        copied.synchronize()
        lab = labels_host.numpy()  # (B, n) fp32 in {-1, 0, 1}: edited in place on the host

        # ---- host-side subsampling, :118-145: same index order, same RNG draws ----
        num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)
        state["num_examples"] = subsample_labels(lab, num_fg, cfg.TRAIN.RPN_BATCHSIZE, ahead)
        return state

    def prefetch_stream(self, labels_numel):
        """Generate the MT19937 key blocks of the NEXT finish_host() now, in host time that would
        otherwise be spent waiting for the device (e.g. after a step's last launch).  Safe at any time:
        the blocks are dropped if anything draws from numpy's stream before they are used."""
        scratch = self._ahead
        ahead = pregenerate_stream(1.4 * labels_numel, getattr(scratch, "buf", None))
        if ahead is not None:
            scratch.buf = ahead.base if ahead.base is not None else ahead
        scratch.ready = ahead

    @staticmethod
    def weights(num_examples):
        """(inside, positive, negative) weights of :153-164 for the sampled-anchor count finish_host() left."""
        inside_w = cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS[0]
        if cfg.TRAIN.RPN_POSITIVE_WEIGHT < 0:
            # :155-158 -- num_examples of the LAST image (stale loop variable) for every image
            positive_weights = 1.0 / num_examples if num_examples > 0 else float('inf')
            return inside_w, positive_weights, positive_weights
        assert ((cfg.TRAIN.RPN_POSITIVE_WEIGHT > 0) & (cfg.TRAIN.RPN_POSITIVE_WEIGHT < 1))
        raise NotImplementedError("RPN_POSITIVE_WEIGHT >= 0 leaves the weights undefined in the "
                                  "reference as well (anchor_target_layer.py:159-164)")

    def launch_finalize(self, state, weights_host):
        """The device half of finish() as pure launches on the CURRENT stream (capturable in a CUDA
        graph): upload of the subsampled labels from state["labels_host"] and of the three weights from
        the pinned (3,) fp32 tensor `weights_host` (the caller writes weights(num_examples) into it before
        every replay), then the finalize kernel reading the weights from device memory."""
        labels, argmax, anchors, inv_index = state["labels"], state["argmax"], state["anchors"], state["inv_index"]
        gt_boxes = state["gt_boxes"]
        weights_dev = torch.empty(3, dtype=torch.float32, device=gt_boxes.device)
        weights_dev.copy_(weights_host, non_blocking=True)
        labels.copy_(state["labels_host"], non_blocking=True)
        return list(F.anchor_targets_finalize(labels, argmax, anchors, gt_boxes, inv_index, self._num_anchors,
                                              state["height"], state["width"], 0.0, 0.0, 0.0,
                                              weights_dev=weights_dev))

    def finish_device(self, state):
        labels, argmax, anchors, inv_index = state["labels"], state["argmax"], state["anchors"], state["inv_index"]
        gt_boxes, height, width = state["gt_boxes"], state["height"], state["width"]
        stream, labels_host, num_examples = state["stream"], state["labels_host"], state["num_examples"]
        A = self._num_anchors
        inside_w, positive_weights, negative_weights = self.weights(num_examples)

        with torch.cuda.stream(stream):
            labels.copy_(labels_host, non_blocking=True)
            out = F.anchor_targets_finalize(labels, argmax, anchors, gt_boxes, inv_index, A, height, width,
                                            inside_w, positive_weights, negative_weights)
        torch.cuda.current_stream(gt_boxes.device).wait_stream(stream)
        # the upload reads the pinned buffer asynchronously: it returns to the pool only for work
        # queued on this device's side stream (begin() fills it there), which is ordered behind the upload
        if state["pooled"]:
            self._give_pinned(gt_boxes.device.index, labels_host)
        for t in out:
            t.record_stream(torch.cuda.current_stream(gt_boxes.device))
        return list(out)

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""
        pass

    def reshape(self, bottom, top):
        pass
