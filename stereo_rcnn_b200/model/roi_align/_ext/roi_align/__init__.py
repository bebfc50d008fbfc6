"""``roi_align.roi_align_forward_cuda`` / ``roi_align_backward_cuda`` -- names and argument order
of lib/model/roi_align/src/roi_align_cuda.h:1-5."""
from stereo_rcnn_b200 import ops as _ops

__all__ = ["roi_align_forward_cuda", "roi_align_backward_cuda"]


def roi_align_forward_cuda(aligned_height, aligned_width, spatial_scale, features, rois, output):
    return _ops.roi_align_forward(aligned_height, aligned_width, spatial_scale, features, rois, output)


def roi_align_backward_cuda(aligned_height, aligned_width, spatial_scale, top_grad, rois, bottom_grad):
    return _ops.roi_align_backward(aligned_height, aligned_width, spatial_scale, top_grad, rois, bottom_grad)
