"""Train-time target layers and losses of the reference on the device.

Host-side mirror of `_AnchorTargetLayer` (lib/model/rpn/anchor_target_layer.py:42-164), `_ProposalTargetLayer`
(lib/model/rpn/proposal_target_layer.py:36-333), the RPN / RCNN losses (stereo_rpn.py:114-140,
stereo_rcnn.py:201-311), the multi-task sum (trainval_net.py:214-219) and `clip_gradient` (net_utils.py:37-49) over
libstereo_b200's kernels (csrc/train_targets.cu, csrc/train_loss.cu).  No host synchronisation anywhere: counts,
sampled indices and losses stay on the device.  Backbone dgrad / wgrad is NOT part of this module.
"""
import ctypes

import numpy as np
import torch

from . import lib as _l
from .ops import check, ptr, stream_ptr, workspace, _f32c

CFG = dict(                                          # lib/model/utils/config.py:55-108,173
    RPN_NEGATIVE_OVERLAP=0.3, RPN_POSITIVE_OVERLAP=0.7, RPN_FG_FRACTION=0.5, RPN_BATCHSIZE=512,
    BATCH_SIZE=512, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.0,
    BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0), BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2),
    DIM_NORMALIZE_MEANS=(1.6, 1.5, 4.0, 0.0, 0.0), DIM_NORMALIZE_STDS=(0.5, 0.5, 0.5, 0.5, 0.5),
    KPTS_GRID=28,
)


def random_words(shape, device, generator=None):
    """uint32 random words for the samplers (stored as int32 bit patterns: torch has no uint32 arithmetic)"""
    return torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int32, device=device, generator=generator)


def _words(t):
    assert t.dtype == torch.int32 and t.is_contiguous()
    return t


def generate_anchors(feat_shapes, device, ratios=None, scales=None, strides=None):
    """generate_anchors_all_pyramids (generate_anchors.py:157-173) on the device -> [A,4] fp32"""
    from .ops import make_proposal_cfg
    L = _l.load()
    pc = make_proposal_cfg("TRAIN", [list(map(int, s)) for s in feat_shapes], ratios=ratios, scales=scales, strides=strides)
    A = sum(int(h) * int(w) for h, w in feat_shapes) * pc.n_ratios
    out = torch.empty(A, 4, dtype=torch.float32, device=device)
    check(L.sb_generate_anchors(ctypes.byref(pc), A, ptr(out), stream_ptr()), "sb_generate_anchors")
    return out


def anchor_targets(anchors, gt_left, gt_right, gt_merge, im_info, keys, cfg=CFG):
    """_AnchorTargetLayer.forward: anchors [A,4] fp32, gt_* [B,K,5], im_info = (H, W, ...) of image 0 as Python
    numbers, keys [B,A] int32 words -> labels [B,A], targets_left [B,A,4], targets_right [B,A,4], inside_w [B,A],
    outside_w [B,A]"""
    L = _l.load()
    A = anchors.shape[0]
    B, K = gt_left.shape[:2]
    dev = anchors.device
    assert keys.shape == (B, A)
    labels = torch.empty(B, A, dtype=torch.float32, device=dev)
    tl = torch.empty(B, A, 4, dtype=torch.float32, device=dev)
    tr = torch.empty(B, A, 4, dtype=torch.float32, device=dev)
    iw = torch.empty(B, A, dtype=torch.float32, device=dev)
    ow = torch.empty(B, A, dtype=torch.float32, device=dev)
    nbytes = L.sb_anchor_targets_workspace(B, A)
    ws = workspace(nbytes, dev, "anchor_targets")
    num_fg = int(cfg["RPN_FG_FRACTION"] * cfg["RPN_BATCHSIZE"])
    check(L.sb_anchor_targets(ptr(_f32c(anchors)), A, ptr(_f32c(gt_left)), ptr(_f32c(gt_right)), ptr(_f32c(gt_merge)),
                              B, K, int(im_info[0]), int(im_info[1]), ptr(_words(keys)),
                              float(cfg["RPN_NEGATIVE_OVERLAP"]), float(cfg["RPN_POSITIVE_OVERLAP"]),
                              int(cfg["RPN_BATCHSIZE"]), num_fg, ptr(ws), nbytes, ptr(labels), ptr(tl), ptr(tr),
                              ptr(iw), ptr(ow), stream_ptr()), "sb_anchor_targets")
    return labels, tl, tr, iw, ow


def _pt_cfg(cfg):
    c = _l.ProposalTargetCfg()
    c.rois_per_image = int(cfg["BATCH_SIZE"])
    c.fg_rois_per_image = int(np.round(cfg["FG_FRACTION"] * cfg["BATCH_SIZE"]))
    c.fg_thresh, c.bg_thresh_hi, c.bg_thresh_lo = cfg["FG_THRESH"], cfg["BG_THRESH_HI"], cfg["BG_THRESH_LO"]
    c.bbox_means = (ctypes.c_float * 4)(*cfg["BBOX_NORMALIZE_MEANS"])
    c.bbox_stds = (ctypes.c_float * 4)(*cfg["BBOX_NORMALIZE_STDS"])
    c.dim_means = (ctypes.c_float * 5)(*cfg["DIM_NORMALIZE_MEANS"])
    c.dim_stds = (ctypes.c_float * 5)(*cfg["DIM_NORMALIZE_STDS"])
    c.kpts_grid = int(cfg["KPTS_GRID"])
    return c


def proposal_targets(rois_left, rois_right, gt_left, gt_right, gt_dim_orien, gt_kpts, keys, words, cfg=CFG):
    """_ProposalTargetLayer.forward: rois_* [B,R,5], gt_* [B,K,5], gt_dim_orien [B,K,5], gt_kpts [B,K,6],
    keys [B,R+K] / words [B,S] int32 random words -> dict like oracle.train_targets.proposal_target_layer
    (kpts_targets int32, plus status [B])"""
    L = _l.load()
    B, R = rois_left.shape[:2]
    K = gt_left.shape[1]
    S = int(cfg["BATCH_SIZE"])
    dev = rois_left.device
    assert keys.shape == (B, R + K) and words.shape == (B, S)
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    o = dict(rois_left=f(B, S, 5), rois_right=f(B, S, 5), labels=f(B, S), bbox_targets_left=f(B, S, 4),
             bbox_targets_right=f(B, S, 4), dim_orien_targets=f(B, S, 5),
             kpts_targets=torch.empty(B, S, 3, dtype=torch.int32, device=dev), kpts_weight=f(B, S, 3),
             inside_w=f(B, S, 4), outside_w=f(B, S, 4),
             keep_inds=torch.empty(B, S, dtype=torch.int32, device=dev),
             status=torch.empty(B, dtype=torch.int32, device=dev))
    c = _pt_cfg(cfg)
    check(L.sb_proposal_targets(ptr(_f32c(rois_left)), ptr(_f32c(rois_right)), B, R, ptr(_f32c(gt_left)),
                                ptr(_f32c(gt_right)), ptr(_f32c(gt_dim_orien)), ptr(_f32c(gt_kpts)), K,
                                ptr(_words(keys)), ptr(_words(words)), ctypes.byref(c), ptr(o["rois_left"]),
                                ptr(o["rois_right"]), ptr(o["labels"]), ptr(o["bbox_targets_left"]),
                                ptr(o["bbox_targets_right"]), ptr(o["dim_orien_targets"]), ptr(o["kpts_targets"]),
                                ptr(o["kpts_weight"]), ptr(o["inside_w"]), ptr(o["outside_w"]), ptr(o["keep_inds"]),
                                ptr(o["status"]), stream_ptr()), "sb_proposal_targets")
    return o


def rpn_loss(rpn_cls_score, rpn_bbox_pred, labels, targets_left, targets_right, inside_w, outside_w, uncert=None,
             grads=True):
    """stereo_rpn.py:114-140: rpn_cls_score [B,A,2], rpn_bbox_pred [B,A,6] -> losses [2] (cls, box) and, with grads,
    d loss / d rpn_cls_score, d loss / d rpn_bbox_pred (times exp(-uncert[0..1]) if uncert is given)"""
    L = _l.load()
    B, A = labels.shape
    dev = labels.device
    losses = torch.empty(2, dtype=torch.float32, device=dev)
    d_cls = torch.empty_like(rpn_cls_score) if grads else None
    d_box = torch.empty_like(rpn_bbox_pred) if grads else None
    nbytes = L.sb_loss_workspace_bytes()
    ws = workspace(nbytes, dev, "loss")
    check(L.sb_rpn_loss(ptr(_f32c(rpn_cls_score)), ptr(_f32c(rpn_bbox_pred)), ptr(_f32c(labels)),
                        ptr(_f32c(targets_left)), ptr(_f32c(targets_right)), ptr(_f32c(inside_w)),
                        ptr(_f32c(outside_w)), B, A, ptr(uncert) if uncert is not None else None, ptr(ws), nbytes,
                        ptr(losses), ptr(d_cls) if grads else None, ptr(d_box) if grads else None, stream_ptr()),
          "sb_rpn_loss")
    return (losses, d_cls, d_box) if grads else losses


def rcnn_loss(cls_score, bbox_pred, dim_orien_pred, kpts_pred, left_border_pred, right_border_pred, tgt, uncert=None,
              grads=True, kpts_grid=28):
    """stereo_rcnn.py:201-311: predictions of the R = B*S sampled RoIs (bbox_pred [R,6C], dim_orien_pred [R,5C] per
    class), tgt = proposal_targets(...) -> losses [4] (cls, bbox, dim_orien, kpts) and, with grads, the six gradients"""
    L = _l.load()
    R, C = cls_score.shape
    dev = cls_score.device
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    preds = [_f32c(t) for t in (cls_score, bbox_pred, dim_orien_pred, kpts_pred, left_border_pred, right_border_pred)]
    g = [torch.empty_like(t) for t in preds] if grads else [None] * 6
    check(L.sb_rcnn_loss(*[ptr(t) for t in preds], ptr(_f32c(tgt["labels"])), ptr(_f32c(tgt["bbox_targets_left"])),
                         ptr(_f32c(tgt["bbox_targets_right"])), ptr(_f32c(tgt["dim_orien_targets"])),
                         ptr(tgt["kpts_targets"]), ptr(_f32c(tgt["kpts_weight"])), ptr(_f32c(tgt["inside_w"])),
                         ptr(_f32c(tgt["outside_w"])), R, C, kpts_grid, ptr(uncert) if uncert is not None else None,
                         ptr(losses), *[ptr(t) if t is not None else None for t in g], stream_ptr()), "sb_rcnn_loss")
    return (losses, g) if grads else losses


def multitask_loss(losses, uncert):
    """trainval_net.py:214-219: losses [n], uncert [n] -> total [1], d total / d uncert [n]"""
    L = _l.load()
    n = losses.numel()
    total = torch.empty(1, dtype=torch.float32, device=losses.device)
    d_u = torch.empty(n, dtype=torch.float32, device=losses.device)
    check(L.sb_multitask_loss(ptr(_f32c(losses)), ptr(_f32c(uncert)), n, ptr(total), ptr(d_u), stream_ptr()),
          "sb_multitask_loss")
    return total, d_u


def clip_gradient(grads, clip_norm):
    """net_utils.clip_gradient over a list of fp32 gradient tensors, in place -> norm [2] (total norm, factor)"""
    L = _l.load()
    n = len(grads)
    dev = grads[0].device
    assert all(g.dtype == torch.float32 and g.is_contiguous() for g in grads)
    ptrs = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads])
    counts = (ctypes.c_size_t * n)(*[g.numel() for g in grads])
    nbytes = L.sb_clip_gradient_workspace(n)
    ws = workspace(nbytes, dev, "clip")
    out = torch.empty(2, dtype=torch.float32, device=dev)
    check(L.sb_clip_gradient(ptrs, counts, n, float(clip_norm), ptr(ws), nbytes, ptr(out), stream_ptr()),
          "sb_clip_gradient")
    return out
