"""``nms.nms_cuda`` -- same name and argument order as the reference's cffi symbol
(lib/model/nms/src/nms_cuda.h:4-5, bound in lib/model/nms/_ext/nms/__init__.py)."""
from stereo_rcnn_b200 import ops as _ops

__all__ = ["nms_cuda"]


def nms_cuda(keep_out, boxes, num_out, nms_overlap_thresh):
    """keep_out int32 [N,1], boxes fp32 [N,5] sorted by score desc, num_out int32 [1]; returns 1"""
    return _ops.nms_into(keep_out, boxes, num_out, nms_overlap_thresh)
