"""RoIAlign / RoIAlignAvg / RoIAlignMax modules (lib/model/roi_align/modules/roi_align.py:6-42).
The 2x2/stride-1 pooling after the (P+1)^2 tap lattice is torch's here because this module is the
reference-layout (NCHW) boundary; the product forward uses the fused NHWC kernel instead."""
from torch.nn.functional import avg_pool2d, max_pool2d
from torch.nn.modules.module import Module

from ..functions.roi_align import RoIAlignFunction


class _Base(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale):
        super().__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)


class RoIAlign(_Base):
    def forward(self, features, rois, scale):
        return RoIAlignFunction(self.aligned_height, self.aligned_width, scale)(features, rois)


class RoIAlignAvg(_Base):
    def forward(self, features, rois, scale):
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1, scale)(features, rois)
        return avg_pool2d(x, kernel_size=2, stride=1)


class RoIAlignMax(_Base):
    def forward(self, features, rois, scale):
        x = RoIAlignFunction(self.aligned_height + 1, self.aligned_width + 1, scale)(features, rois)
        return max_pool2d(x, kernel_size=2, stride=1)
