"""``_ProposalLayer`` with the reference's constructor and forward signature
(lib/model/rpn/proposal_layer.py:26-145); the whole layer is one stream-ordered C-ABI call."""
import torch.nn as nn

from stereo_rcnn_b200 import ops as _ops
from ..utils.config import cfg


class _ProposalLayer(nn.Module):
    def __init__(self, feat_stride, ratios):
        super().__init__()
        self._feat_stride = feat_stride
        self._anchor_ratios = ratios
        self._fpn_scales = list(cfg.FPN_ANCHOR_SCALES)
        self._fpn_feature_strides = list(cfg.FPN_FEAT_STRIDES)
        self._fpn_anchor_stride = cfg.FPN_ANCHOR_STRIDE

    def forward(self, input):
        """input = (rpn_cls_prob [B,A,2], rpn_bbox_pred_left_right [B,A,6], im_info [B,3], cfg_key,
        feat_shapes) -> (rois_left [B,N,5], rois_right [B,N,5])"""
        cls_prob, bbox_lr, im_info, cfg_key, feat_shapes = input
        c = cfg[cfg_key]
        pc = _ops.make_proposal_cfg(cfg_key, [list(map(int, s)) for s in feat_shapes], ratios=self._anchor_ratios,
                                    scales=self._fpn_scales, strides=self._fpn_feature_strides,
                                    cfg={cfg_key: dict(RPN_PRE_NMS_TOP_N=c.RPN_PRE_NMS_TOP_N,
                                                       RPN_POST_NMS_TOP_N=c.RPN_POST_NMS_TOP_N,
                                                       RPN_NMS_THRESH=c.RPN_NMS_THRESH)})
        return _ops.proposal_layer(cls_prob, bbox_lr, im_info, cfg_key, feat_shapes, pc=pc)

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""
        pass
