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
