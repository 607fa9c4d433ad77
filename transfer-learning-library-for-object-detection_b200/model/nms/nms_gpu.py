"""nms_gpu(dets, thresh) (lib/model/nms/nms_gpu.py:7-12): returns keep of shape (k, 1)
int32 on dets' device.  The slice by the device-resident count is the one host
synchronisation (the reference's ``keep[:num_out[0]]`` does the same)."""
from tlod_b200 import functional as F


def nms_gpu(dets, thresh):
    keep, num = F.nms_device(dets, float(thresh))
    return keep[:int(num.item())].view(-1, 1)
