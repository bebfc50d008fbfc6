"""CPU restatement of the reference's train-time target layers and losses.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and tests/golden/make_golden.py); the product never imports it.

Every function follows the reference line by line in fp32 (numpy, same operation order), so that labels, sampled
indices and every compare against a threshold are bit-identical to the reference's own Python run on CPU
(tests/golden/train_targets.npz, minted by executing the reference through oracle/ref_shim.py):

  bbox_overlaps_batch    lib/model/rpn/bbox_transform.py:220-309
  bbox_transform_batch   lib/model/rpn/bbox_transform.py:38-77
  anchor_target_layer    lib/model/rpn/anchor_target_layer.py:42-164 (+ _unmap :174-184)
  proposal_target_layer  lib/model/rpn/proposal_target_layer.py:36-333
  smooth_l1_loss         lib/model/utils/net_utils.py:79-99
  rpn_losses             lib/model/rpn/stereo_rpn.py:114-140
  rcnn_losses            lib/model/stereo_rcnn/stereo_rcnn.py:201-311
  multitask_loss         trainval_net.py:214-219

The only degree of freedom is the random sampler.  The reference draws from numpy's global stream
(`np.random.permutation(n)`, `np.random.rand(k)`) with n, k depending on the data; `NumpySampler` does exactly that
(pins the oracle to the reference), `KeySampler` is the sampler of the device kernels (sb_anchor_targets /
sb_proposal_targets): explicit random words per candidate, no data-dependent consumption, hence no host round trip.
Both go through the same two calls below, nothing else differs.
"""
import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32

CFG = dict(                                         # lib/model/utils/config.py:55-108,173
    RPN_NEGATIVE_OVERLAP=0.3, RPN_POSITIVE_OVERLAP=0.7, RPN_FG_FRACTION=0.5, RPN_BATCHSIZE=512,
    BATCH_SIZE=512, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.0,
    BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0), BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2),
    DIM_NORMALIZE_MEANS=(1.6, 1.5, 4.0, 0.0, 0.0), DIM_NORMALIZE_STDS=(0.5, 0.5, 0.5, 0.5, 0.5),
    KPTS_GRID=28,
)


class NumpySampler(object):
    """the reference's own use of numpy's stream (anchor_target_layer.py:109,121; proposal_target_layer.py:236-262)"""

    def __init__(self, rng=np.random):
        self.rng = rng

    def permutation(self, image, cand):
        return self.rng.permutation(len(cand))

    def draws(self, image, k):
        return self.rng.rand(k)


class KeySampler(object):
    """the device sampler: `keys` [B, n_candidates_total] uint32, one word per anchor / RoI -- the "permutation" of a
    candidate list is its order by (key, index); `words` [B, rois_per_image] uint32 -- draw j is words[j] / 2^32"""

    def __init__(self, keys, words=None):
        self.keys = np.asarray(keys, np.uint32)
        self.words = None if words is None else np.asarray(words, np.uint32)

    def permutation(self, image, cand):
        k = self.keys[image][np.asarray(cand, np.int64)]
        return np.lexsort((np.arange(len(cand)), k))          # ascending key, ties by position

    def draws(self, image, k):
        return self.words[image][:k].astype(np.float64) / 4294967296.0


# ------------------------------------------------------------------------------------------------ geometry
def bbox_overlaps_batch(anchors, gt_boxes):
    """bbox_transform.py:220-309 -> [B, N, K] fp32; zero-area gt -> 0, zero-area anchor -> -1"""
    gt = np.asarray(gt_boxes, f32)[:, :, :4]
    B, K = gt.shape[:2]
    a = np.asarray(anchors, f32)
    if a.ndim == 2:
        a = np.broadcast_to(a[None], (B,) + a.shape)
    elif a.shape[2] != 4:
        a = a[:, :, 1:5]
    gx = gt[:, :, 2] - gt[:, :, 0] + 1
    gy = gt[:, :, 3] - gt[:, :, 1] + 1
    garea = (gx * gy)[:, None, :]
    ax = a[:, :, 2] - a[:, :, 0] + 1
    ay = a[:, :, 3] - a[:, :, 1] + 1
    aarea = (ax * ay)[:, :, None]
    gzero = (gx == 1) & (gy == 1)
    azero = (ax == 1) & (ay == 1)
    iw = np.minimum(a[:, :, None, 2], gt[:, None, :, 2]) - np.maximum(a[:, :, None, 0], gt[:, None, :, 0]) + 1
    iw[iw < 0] = 0
    ih = np.minimum(a[:, :, None, 3], gt[:, None, :, 3]) - np.maximum(a[:, :, None, 1], gt[:, None, :, 1]) + 1
    ih[ih < 0] = 0
    inter = iw * ih
    ua = aarea + garea - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        ov = inter / ua
    ov = np.where(np.broadcast_to(gzero[:, None, :], ov.shape), f32(0), ov)
    ov = np.where(np.broadcast_to(azero[:, :, None], ov.shape), f32(-1), ov)
    return ov.astype(f32)


def bbox_transform_batch(ex, gt):
    """bbox_transform.py:38-77; ex [N,4] or [B,N,4], gt [B,N,4] -> [B,N,4] (dx, dy, dw, dh)"""
    ex = np.asarray(ex, f32)
    gt = np.asarray(gt, f32)
    if ex.ndim == 2:
        ex = ex[None]
    ew = ex[..., 2] - ex[..., 0] + f32(1.0)
    eh = ex[..., 3] - ex[..., 1] + f32(1.0)
    ecx = ex[..., 0] + f32(0.5) * ew
    ecy = ex[..., 1] + f32(0.5) * eh
    gw = gt[..., 2] - gt[..., 0] + f32(1.0)
    gh = gt[..., 3] - gt[..., 1] + f32(1.0)
    gcx = gt[..., 0] + f32(0.5) * gw
    gcy = gt[..., 1] + f32(0.5) * gh
    with np.errstate(divide="ignore", invalid="ignore"):
        dx = (gcx - ecx) / ew
        dy = (gcy - ecy) / eh
        dw = np.log(gw / ew)
        dh = np.log(gh / eh)
    return np.stack((dx, dy, dw, dh), axis=2).astype(f32)


def _first_argmax(x, axis):
    """torch.max(x, axis) on CPU: value and the FIRST index of the maximum"""
    return x.max(axis=axis), x.argmax(axis=axis)


# ------------------------------------------------------------------------------------- anchor target layer
def anchor_target_layer(anchors_all, gt_left, gt_right, gt_merge, im_info, sampler, cfg=CFG):
    """anchor_target_layer.py:42-164.  anchors_all [A,4] fp32 (all pyramid levels, proposal-layer order),
    gt_* [B,K,5], im_info [B,3] -> labels [B,A] (1 / 0 / -1), targets_left [B,A,4], targets_right [B,A,4],
    inside_w [B,A], outside_w [B,A] -- all fp32 like the reference

    Kept quirks: the image bounds of image 0 filter the anchors of the whole batch (:70-73); `num_bg` uses the
    foreground count BEFORE its own subsampling (:114); the outside weight 1/num_examples is computed from the
    LAST image of the batch only (:136, the loop variable leaks) and applied to all."""
    anchors_all = np.asarray(anchors_all, f32)
    gt_left, gt_right, gt_merge = (np.asarray(g, f32) for g in (gt_left, gt_right, gt_merge))
    B = gt_left.shape[0]
    A = anchors_all.shape[0]
    keep = ((anchors_all[:, 0] >= 0) & (anchors_all[:, 1] >= 0) &
            (anchors_all[:, 2] < int(im_info[0][1])) & (anchors_all[:, 3] < int(im_info[0][0])))
    inside = np.nonzero(keep)[0]
    anchors = anchors_all[inside]
    n = inside.size
    labels = np.full((B, n), -1, f32)
    overlaps = bbox_overlaps_batch(anchors, gt_merge)
    max_ov, argmax_ov = _first_argmax(overlaps, 2)
    gt_max = overlaps.max(axis=1)
    labels[max_ov < f32(cfg["RPN_NEGATIVE_OVERLAP"])] = 0
    gt_max[gt_max == 0] = f32(1e-5)
    hit = (overlaps == gt_max[:, None, :]).sum(axis=2)
    labels[hit > 0] = 1
    labels[max_ov >= f32(cfg["RPN_POSITIVE_OVERLAP"])] = 1
    num_fg = int(cfg["RPN_FG_FRACTION"] * cfg["RPN_BATCHSIZE"])
    sum_fg = (labels == 1).sum(axis=1)
    sum_bg = (labels == 0).sum(axis=1)
    for i in range(B):
        if sum_fg[i] > num_fg:
            fg = np.nonzero(labels[i] == 1)[0]
            perm = sampler.permutation(i, inside[fg])
            labels[i][fg[perm[:fg.size - num_fg]]] = -1
        num_bg = cfg["RPN_BATCHSIZE"] - sum_fg[i]
        if sum_bg[i] > num_bg:
            bg = np.nonzero(labels[i] == 0)[0]
            perm = sampler.permutation(i, inside[bg])
            labels[i][bg[perm[:bg.size - num_bg]]] = -1
    gl = np.take_along_axis(gt_left[:, :, :4], argmax_ov[:, :, None].repeat(4, 2), axis=1)
    gr = np.take_along_axis(gt_right[:, :, :4], argmax_ov[:, :, None].repeat(4, 2), axis=1)
    tl = bbox_transform_batch(anchors, gl)
    tr = bbox_transform_batch(anchors, gr)
    inside_w = np.zeros((B, n), f32)
    inside_w[labels == 1] = 1.0
    num_examples = int((labels[B - 1] >= 0).sum())
    w = f32(1.0 / num_examples) if num_examples > 0 else f32(np.inf)
    outside_w = np.zeros((B, n), f32)
    outside_w[labels == 1] = w
    outside_w[labels == 0] = w

    def unmap(d, fill):
        out = np.full((B, A) + d.shape[2:], fill, f32)
        out[:, inside] = d
        return out
    return unmap(labels, -1), unmap(tl, 0), unmap(tr, 0), unmap(inside_w, 0), unmap(outside_w, 0)


# ----------------------------------------------------------------------------------- proposal target layer
def kpts_targets(ex_rois, gt_kpts, grid=28):
    """proposal_target_layer.py:158-182: [B,R,4], [B,R,6] -> target int64 [B,R,3], weight fp32 [B,R,3]"""
    ex_rois, gt_kpts = np.asarray(ex_rois, f32), np.asarray(gt_kpts, f32)
    start = ex_rois[:, :, 0:1]
    width = (ex_rois[:, :, 2] - ex_rois[:, :, 0] + 1)[:, :, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.rint((gt_kpts - start) * f32(grid) / width).astype(f32)
    t[t < 0] = -225
    t[t > grid - 1] = -225
    pos, typ = _first_argmax(t[:, :, :4], 2)
    tgt = np.concatenate(((typ.astype(f32) * f32(grid) + pos)[:, :, None], t[:, :, 4:]), axis=2)
    w = np.ones_like(tgt)
    w[tgt < 0] = 0
    tgt[tgt < 0] = 0
    return tgt.astype(np.int64), w


def proposal_target_layer(rois_left, rois_right, gt_left, gt_right, gt_dim_orien, gt_kpts, sampler, cfg=CFG):
    """proposal_target_layer.py:36-333.  rois_* [B,R,5] (batch index, x1, y1, x2, y2), gt_* [B,K,5] (x1, y1, x2, y2,
    class), gt_dim_orien [B,K,5], gt_kpts [B,K,6] ->
    dict(rois_left [B,S,5], rois_right, labels [B,S], bbox_targets_left [B,S,4], bbox_targets_right,
         dim_orien_targets [B,S,5], kpts_targets int64 [B,S,3], kpts_weight [B,S,3], inside_w [B,S,4],
         outside_w [B,S,4], keep_inds int64 [B,S]) with S = 512"""
    rois_left, rois_right, gt_left, gt_right, gt_dim_orien, gt_kpts = (
        np.asarray(t, f32) for t in (rois_left, rois_right, gt_left, gt_right, gt_dim_orien, gt_kpts))
    B, K = gt_left.shape[:2]

    def with_gt(rois, gt):
        app = np.zeros_like(gt)
        app[:, :, 1:5] = gt[:, :, :4]
        return np.concatenate((rois, app), axis=1)
    all_l, all_r = with_gt(rois_left, gt_left), with_gt(rois_right, gt_right)
    S = int(cfg["BATCH_SIZE"])
    fg_per_image = int(np.round(cfg["FG_FRACTION"] * S))
    ov_l = bbox_overlaps_batch(all_l, gt_left)
    ov_r = bbox_overlaps_batch(all_r, gt_right)
    max_l, asg_l = _first_argmax(ov_l, 2)
    max_r, asg_r = _first_argmax(ov_r, 2)
    labels_all = np.take_along_axis(gt_left[:, :, 4], asg_l, axis=1)
    labels = np.zeros((B, S), f32)
    out_l = np.zeros((B, S, 5), f32)
    out_r = np.zeros((B, S, 5), f32)
    gt_sel_l = np.zeros((B, S, 5), f32)
    gt_sel_r = np.zeros((B, S, 5), f32)
    dim_sel = np.zeros((B, S, 5), f32)
    kp_sel = np.zeros((B, S, 6), f32)
    keep_all = np.zeros((B, S), np.int64)
    fg_t, hi, lo = f32(cfg["FG_THRESH"]), f32(cfg["BG_THRESH_HI"]), f32(cfg["BG_THRESH_LO"])
    for i in range(B):
        fg = np.nonzero((max_l[i] >= fg_t) & (max_r[i] >= fg_t) & (asg_l[i] == asg_r[i]))[0]
        bg = np.nonzero(((max_l[i] < hi) & (max_l[i] >= lo)) | ((max_r[i] < hi) & (max_r[i] >= lo)))[0]
        if fg.size > 0 and bg.size > 0:
            n_fg = min(fg_per_image, fg.size)
            perm = sampler.permutation(i, fg)
            fg = fg[perm[:n_fg]]
            n_bg = S - n_fg
            bg = bg[np.floor(sampler.draws(i, n_bg) * bg.size).astype(np.int64)]
        elif fg.size > 0:
            fg = fg[np.floor(sampler.draws(i, S) * fg.size).astype(np.int64)]
            n_fg, bg = S, bg[:0]
        elif bg.size > 0:
            bg = bg[np.floor(sampler.draws(i, S) * bg.size).astype(np.int64)]
            n_fg, fg = 0, fg[:0]
        else:
            raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
        keep = np.concatenate((fg, bg))
        keep_all[i] = keep
        labels[i] = labels_all[i][keep]
        labels[i, n_fg:] = 0
        out_l[i] = all_l[i][keep]
        out_l[i, :, 0] = i
        out_r[i] = all_r[i][keep]
        out_r[i, :, 0] = i
        gt_sel_l[i] = gt_left[i][asg_l[i][keep]]
        gt_sel_r[i] = gt_right[i][asg_r[i][keep]]
        dim_sel[i] = gt_dim_orien[i][asg_l[i][keep]]
        kp_sel[i] = gt_kpts[i][asg_l[i][keep]]

    def box_targets(rois, gts):
        t = bbox_transform_batch(rois[:, :, 1:5], gts[:, :, :4])
        return ((t - np.asarray(cfg["BBOX_NORMALIZE_MEANS"], f32)) / np.asarray(cfg["BBOX_NORMALIZE_STDS"], f32)).astype(f32)
    tl, tr = box_targets(out_l, gt_sel_l), box_targets(out_r, gt_sel_r)
    td = ((dim_sel - np.asarray(cfg["DIM_NORMALIZE_MEANS"], f32)) / np.asarray(cfg["DIM_NORMALIZE_STDS"], f32)).astype(f32)
    tk, wk = kpts_targets(out_l[:, :, 1:5], kp_sel, cfg["KPTS_GRID"])
    pos = labels > 0
    tl = np.where(pos[:, :, None], tl, f32(0))
    tr = np.where(pos[:, :, None], tr, f32(0))
    td = np.where(pos[:, :, None], td, f32(0))
    inside_w = np.where(pos[:, :, None], f32(1), f32(0)) * np.ones((1, 1, 4), f32)
    car = labels == 1
    tk = np.where(car[:, :, None], tk, 0)
    wk = np.where(car[:, :, None], wk, f32(0))
    return dict(rois_left=out_l, rois_right=out_r, labels=labels, bbox_targets_left=tl, bbox_targets_right=tr,
                dim_orien_targets=td, kpts_targets=tk, kpts_weight=wk, inside_w=inside_w.astype(f32),
                outside_w=(inside_w > 0).astype(f32), keep_inds=keep_all)


# -------------------------------------------------------------------------------------------------- losses
def smooth_l1_loss(pred, target, inside_w=None, outside_w=None, sigma=1.0, dim=(1,)):
    """net_utils.py:79-99 (torch tensors; differentiable w.r.t. pred)"""
    s2 = sigma ** 2
    d = pred - target
    if inside_w is not None:
        d = inside_w * d
    a = torch.abs(d)
    sign = (a < 1.0 / s2).detach().float()
    loss = torch.pow(d, 2) * (s2 / 2.0) * sign + (a - (0.5 / s2)) * (1.0 - sign)
    if outside_w is not None:
        loss = outside_w * loss
    for i in sorted(dim, reverse=True):
        loss = loss.sum(i)
    return loss.mean()


def rpn_losses(rpn_cls_score, rpn_bbox_pred, labels, targets_left, targets_right, inside_w, outside_w):
    """stereo_rpn.py:114-140.  rpn_cls_score [B,A,2], rpn_bbox_pred [B,A,6], anchor_target_layer outputs ->
    (rpn_loss_cls, rpn_loss_box_left_right)"""
    B = labels.shape[0]
    lab = labels.reshape(-1)
    keep = torch.nonzero(lab != -1).view(-1)
    loss_cls = F.cross_entropy(rpn_cls_score.reshape(-1, 2)[keep], lab[keep].long())
    t = torch.zeros(B, labels.shape[1], 6, dtype=targets_left.dtype)
    t[:, :, :4] = targets_left
    t[:, :, 4] = targets_right[:, :, 0]
    t[:, :, 5] = targets_right[:, :, 2]
    iw = inside_w.unsqueeze(2).expand(B, inside_w.shape[1], 6)
    ow = outside_w.unsqueeze(2).expand(B, outside_w.shape[1], 6)
    return loss_cls, smooth_l1_loss(rpn_bbox_pred, t, iw, ow, sigma=3)


def _weighted_ce(pred, label, weight):
    ce = F.cross_entropy(pred, label, reduction="none")
    s = torch.sum(weight)
    return torch.sum(ce * weight) if float(s) < 1 else torch.sum(ce * weight) / s


def rcnn_losses(cls_score, bbox_pred_all, dim_orien_pred_all, kpts_pred, left_border_pred, right_border_pred, tgt):
    """stereo_rcnn.py:201-311.  cls_score [R,C], bbox_pred_all [R,6C], dim_orien_pred_all [R,5C], kpts_pred [R,4*28],
    left/right_border_pred [R,28]; tgt = proposal_target_layer(...) as torch tensors (R = B*S rows) ->
    (loss_cls, loss_bbox, loss_dim_orien, loss_kpts)"""
    lab = tgt["labels"].reshape(-1).long()
    R = lab.shape[0]
    t6 = torch.zeros(R, 6)
    t6[:, :4] = tgt["bbox_targets_left"].reshape(R, 4)
    t6[:, 4] = tgt["bbox_targets_right"].reshape(R, 4)[:, 0]
    t6[:, 5] = tgt["bbox_targets_right"].reshape(R, 4)[:, 2]
    iw4, ow4 = tgt["inside_w"].reshape(R, 4), tgt["outside_w"].reshape(R, 4)
    iw = torch.cat((iw4, iw4[:, 0:2]), 1)
    ow = torch.cat((ow4, ow4[:, 0:2]), 1)
    bp = torch.gather(bbox_pred_all.view(R, -1, 6), 1, lab.view(R, 1, 1).expand(R, 1, 6)).squeeze(1)
    dp = torch.gather(dim_orien_pred_all.view(R, -1, 5), 1, lab.view(R, 1, 1).expand(R, 1, 5)).squeeze(1)
    loss_cls = F.cross_entropy(cls_score, lab)
    loss_bbox = smooth_l1_loss(bp, t6, iw, ow)
    loss_dim = smooth_l1_loss(dp, tgt["dim_orien_targets"].reshape(R, 5))
    kl = tgt["kpts_targets"].reshape(R, 3).long()
    kw = tgt["kpts_weight"].reshape(R, 3)
    lk = _weighted_ce(kpts_pred, kl[:, 0], kw[:, 0])
    ll = _weighted_ce(left_border_pred, kl[:, 1], kw[:, 1])
    lr = _weighted_ce(right_border_pred, kl[:, 2], kw[:, 2])
    return loss_cls, loss_bbox, loss_dim, (lk + ll + lr) / 3.0


def multitask_loss(losses, uncert):
    """trainval_net.py:214-219: sum_i L_i * exp(-u_i) + u_i"""
    total = 0
    for i, l in enumerate(losses):
        total = total + l.mean() * torch.exp(-uncert[i]) + uncert[i]
    return total
