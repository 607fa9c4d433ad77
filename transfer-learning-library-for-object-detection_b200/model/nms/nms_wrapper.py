"""nms(dets, thresh, force_cpu=False) (lib/model/nms/nms_wrapper.py:13-21).

``force_cpu=True`` selected the reference's numpy ``nms_cpu`` (which is wrong:
nms_cpu.py:23-24 takes np.maximum of the far corners).  This package has no CPU
path, so force_cpu raises instead of silently running on the GPU."""
from model.nms.nms_gpu import nms_gpu


def nms(dets, thresh, force_cpu=False):
    """dets (n, 5) = [x1, y1, x2, y2, score] sorted by score descending."""
    if dets.shape[0] == 0:
        return []
    if force_cpu:
        raise RuntimeError("tlod_b200 has no CPU NMS (force_cpu=True); the reference's nms_cpu is "
                           "not a valid implementation either (nms_cpu.py:23-24)")
    return nms_gpu(dets, thresh)


def nms_per_class(scores, pred_boxes, thresh, nms_thresh=None, class_agnostic=False):
    """All classes of one image at once: the loop of methods/DAF/DAF_test.py:302-320
    (``for j in 1..num_classes-1: inds = scores[:, j] > thresh; sort; nms(cls_dets, cfg.TEST.NMS)``)
    as three launches and one host synchronisation.  Returns the list ``cls_dets`` for
    j = 1..num_classes-1, each (n_j, 5) = [x1, y1, x2, y2, score] on the inputs' device."""
    from model.utils.config import cfg
    from tlod_b200 import functional as F
    if nms_thresh is None:
        nms_thresh = cfg.TEST.NMS
    boxes = pred_boxes if not class_agnostic else pred_boxes[:, :4]
    return F.class_nms(scores, boxes.contiguous(), float(thresh), float(nms_thresh))
