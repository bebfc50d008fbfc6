"""``_ProposalTargetLayer`` with the reference's constructor and forward signature
(lib/model/rpn/proposal_target_layer.py:20-70); one kernel (one CTA per image) behind one C-ABI call, no host
synchronisation (the reference loops in Python with per-element indexing and a `.cpu()` union, :230)."""
import torch
import torch.nn as nn

from stereo_rcnn_b200 import train as _train


class _ProposalTargetLayer(nn.Module):
    def __init__(self, nclasses):
        super().__init__()
        self._num_classes = nclasses
        self.generator = None

    def forward(self, all_rois_left, all_rois_right, gt_boxes_left, gt_boxes_right, gt_dim_orien, gt_kpts, num_boxes):
        """-> (rois_left [B,S,5], rois_right, labels [B,S], bbox_targets_left [B,S,4], bbox_targets_right,
        dim_orien_targets [B,S,5], kpts_targets [B,S,3] (long), kpts_weight [B,S,3], bbox_inside_weights [B,S,4],
        bbox_outside_weights [B,S,4]) as proposal_target_layer.py:64-65"""
        dev = all_rois_left.device
        B, R = all_rois_left.shape[:2]
        K = gt_boxes_left.shape[1]
        S = int(_train.CFG["BATCH_SIZE"])
        keys = _train.random_words((B, R + K), dev, self.generator)
        words = _train.random_words((B, S), dev, self.generator)
        o = _train.proposal_targets(all_rois_left, all_rois_right, gt_boxes_left, gt_boxes_right, gt_dim_orien,
                                    gt_kpts, keys, words)
        self.status = o["status"]                   # device flag: 1 where the reference would raise (:267)
        return (o["rois_left"], o["rois_right"], o["labels"], o["bbox_targets_left"], o["bbox_targets_right"],
                o["dim_orien_targets"], o["kpts_targets"].to(torch.int64), o["kpts_weight"], o["inside_w"],
                o["outside_w"])

    def backward(self, top, propagate_down, bottom):
        """This layer does not propagate gradients."""
        pass
