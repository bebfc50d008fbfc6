"""RoIAlignFunction with the reference's call pattern ``RoIAlignFunction(ah, aw, scale)(features, rois)``
(lib/model/roi_align/functions/roi_align.py:8-47), on modern autograd."""
import torch
from torch.autograd import Function

from .._ext import roi_align


class _RoIAlign(Function):
    @staticmethod
    def forward(ctx, features, rois, ah, aw, scale):
        if not features.is_cuda:
            raise NotImplementedError          # functions/roi_align.py:28-29
        ctx.save_for_backward(rois)
        ctx.meta = (ah, aw, scale, features.size())
        output = features.new_zeros((rois.size(0), features.size(1), ah, aw))
        roi_align.roi_align_forward_cuda(ah, aw, scale, features.contiguous(), rois.contiguous(), output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        ah, aw, scale, fsize = ctx.meta
        grad_input = rois.new_zeros(tuple(fsize))
        roi_align.roi_align_backward_cuda(ah, aw, scale, grad_output.contiguous(), rois, grad_input)
        return grad_input, None, None, None, None


class RoIAlignFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)

    def __call__(self, features, rois):
        return _RoIAlign.apply(features, rois, self.aligned_height, self.aligned_width, self.spatial_scale)
