"""``_AnchorTargetLayer`` with the reference's constructor and forward signature
(lib/model/rpn/anchor_target_layer.py:26-164); the layer is four stream-ordered kernels behind one C-ABI call.  The
random subsampling takes words from a torch generator on the device instead of numpy's global stream (see
stereo_rcnn_b200/train.py).  Host synchronisation: none if `im_info` is a host tensor (as the data loader produces
it); a CUDA `im_info` costs one 3-float read per call, because the image size is a launch argument -- callers that
know (H, W) use `train.anchor_targets` directly."""
import torch.nn as nn

from stereo_rcnn_b200 import train as _train
from ..utils.config import cfg


class _AnchorTargetLayer(nn.Module):
    def __init__(self, feat_stride, ratios):
        super().__init__()
        self._feat_stride = feat_stride
        self._anchor_ratios = ratios
        self._fpn_scales = list(cfg.FPN_ANCHOR_SCALES)
        self._fpn_feature_strides = list(cfg.FPN_FEAT_STRIDES)
        self._fpn_anchor_stride = cfg.FPN_ANCHOR_STRIDE
        self._allowed_border = 0
        self._anchors = {}
        self.generator = None                       # optional torch.Generator (cuda) for reproducible sampling

    def forward(self, input):
        """input = (scores, gt_boxes_left [B,K,5], gt_boxes_right, gt_boxes_merge, im_info [B,3], num_boxes,
        feat_shapes) -> [labels [B,A], bbox_targets_left [B,A,4], bbox_targets_right [B,A,4],
        bbox_inside_weights [B,A], bbox_outside_weights [B,A]]"""
        _scores, gt_left, gt_right, gt_merge, im_info, _num_boxes, feat_shapes = input
        dev = gt_left.device
        key = (tuple(tuple(int(v) for v in s) for s in feat_shapes), dev)
        if key not in self._anchors:
            self._anchors[key] = _train.generate_anchors(feat_shapes, dev, ratios=self._anchor_ratios,
                                                         scales=self._fpn_scales, strides=self._fpn_feature_strides)
        anchors = self._anchors[key]
        hw = im_info[0] if not hasattr(im_info, "is_cuda") or not im_info.is_cuda else im_info[0].tolist()
        keys = _train.random_words((gt_left.shape[0], anchors.shape[0]), dev, self.generator)
        return list(_train.anchor_targets(anchors, gt_left, gt_right, gt_merge, (int(hw[0]), int(hw[1])), keys))

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""
        pass
