"""lib/model/nms/nms_wrapper.py:13-21.  The reference's CPU branch (nms_cpu.py) is not
equivalent to its GPU kernel and is never taken (USE_GPU_NMS=True); here it raises."""
from .nms_gpu import nms_gpu


def nms(dets, thresh, force_cpu=False):
    if dets.shape[0] == 0:
        return []
    if force_cpu:
        raise NotImplementedError("stereo_rcnn_b200 has no CPU NMS path (sm_100a kernels only)")
    return nms_gpu(dets, thresh)
