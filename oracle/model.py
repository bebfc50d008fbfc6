"""CPU oracle of the Stereo R-CNN forward graph -- TEST INFRASTRUCTURE ONLY.

A functional (state-dict driven) torch-CPU fp32 restatement of the reference's
test-mode forward:

* trunk ``RCNN_layer0..4``  : lib/model/stereo_rcnn/resnet.py:66-163,236-240 (Q1-Q3)
* FPN                      : resnet.py:243-253, stereo_rcnn.py:91-108,155-185 (Q4,Q5)
* stereo RPN head          : lib/model/rpn/stereo_rpn.py:52-95 (Q6-Q8)
* proposal layer           : oracle.ops.proposal_layer (proposal_layer.py:42-145)
* PyramidRoI_Feat/RoIAlign : oracle.ops.pyramid_roi_feat (stereo_rcnn.py:110-139)
* box / keypoint heads     : resnet.py:256-286,345-348, stereo_rcnn.py:248-271 (Q17,Q18)

Weights use the reference's ``state_dict`` key names (``RCNN_layer1.0.0.conv1.weight``
...), so a reference checkpoint loads unchanged.  ``F.interpolate`` is pinned to
``align_corners=True`` (torch-0.3.0 semantics of ``F.upsample``).
"""
import zlib
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

LAYERS = [3, 4, 23, 3]
PLANES = [64, 128, 256, 512]
STRIDES = [1, 2, 2, 2]
N_CLASSES = 2


from stereo_rcnn_b200.synth import make_state_dict, param_shapes  # noqa: E402,F401  (shared synthetic weights)


# --------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------
def _bn(x, sd, k):
    return F.batch_norm(x, sd[k + ".running_mean"], sd[k + ".running_var"],
                        sd[k + ".weight"], sd[k + ".bias"], False, 0.0, 1e-5)


def _conv(x, sd, k, stride=1, pad=0):
    return F.conv2d(x, sd[k + ".weight"], sd.get(k + ".bias"), stride=stride, padding=pad)


def bottleneck(x, sd, p, stride, has_ds):
    """resnet.py:66-102; stride sits on conv1 (Q1)"""
    out = F.relu(_bn(_conv(x, sd, p + ".conv1", stride=stride), sd, p + ".bn1"))
    out = F.relu(_bn(_conv(out, sd, p + ".conv2", pad=1), sd, p + ".bn2"))
    out = _bn(_conv(out, sd, p + ".conv3"), sd, p + ".bn3")
    res = x
    if has_ds:
        res = _bn(_conv(x, sd, p + ".downsample.0", stride=stride), sd, p + ".downsample.1")
    return F.relu(out + res)


def layer0(x, sd):
    x = F.relu(_bn(_conv(x, sd, "RCNN_layer0.0", stride=2, pad=3), sd, "RCNN_layer0.1"))
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=0, ceil_mode=True)   # Q2


def res_layer(x, sd, li):
    for b in range(LAYERS[li]):
        x = bottleneck(x, sd, "RCNN_layer%d.0.%d" % (li + 1, b), STRIDES[li] if b == 0 else 1, b == 0)
    return x


def upsample_add(x, y):
    """stereo_rcnn.py:91-108 with torch-0.3.0 align_corners=True semantics (Q4)"""
    return F.interpolate(x, size=y.shape[2:], mode="bilinear", align_corners=True) + y


def trunk_fpn(im, sd):
    """stereo_rcnn.py:155-168 -> dict of C2..C5, P2..P6"""
    c1 = layer0(im, sd)
    c2 = res_layer(c1, sd, 0)
    c3 = res_layer(c2, sd, 1)
    c4 = res_layer(c3, sd, 2)
    c5 = res_layer(c4, sd, 3)
    p5 = _conv(c5, sd, "RCNN_toplayer")
    p4 = _conv(upsample_add(p5, _conv(c4, sd, "RCNN_latlayer1")), sd, "RCNN_smooth1", pad=1)
    p3 = _conv(upsample_add(p4, _conv(c3, sd, "RCNN_latlayer2")), sd, "RCNN_smooth2", pad=1)
    p2 = _conv(upsample_add(p3, _conv(c2, sd, "RCNN_latlayer3")), sd, "RCNN_smooth3", pad=1)
    p6 = p5[:, :, ::2, ::2].contiguous()          # MaxPool2d(1, stride=2) (Q5)
    return dict(c1=c1, c2=c2, c3=c3, c4=c4, c5=c5, p2=p2, p3=p3, p4=p4, p5=p5, p6=p6)


def rpn_head(feats_l, feats_r, sd):
    """stereo_rpn.py:73-95.  Returns cls_prob [B,A,2], bbox_pred [B,A,6], shapes, raw scores."""
    probs, boxes, shapes, scores = [], [], [], []
    for fl, fr in zip(feats_l, feats_r):
        B = fl.shape[0]
        x = torch.cat((F.relu(_conv(fl, sd, "RCNN_rpn.RPN_Conv", pad=1)),
                       F.relu(_conv(fr, sd, "RCNN_rpn.RPN_Conv", pad=1))), 1)
        s = _conv(x, sd, "RCNN_rpn.RPN_cls_score")                 # B,6,H,W
        H, W = s.shape[2:]
        sr = s.view(B, 2, 3 * H, W)                                # reshape(x, 2)
        pr = F.softmax(sr, 1).view(B, 6, H, W)                     # Q7 pairing (c, c+3)
        bp = _conv(x, sd, "RCNN_rpn.RPN_bbox_pred_left_right")
        shapes.append([H, W])
        scores.append(s.permute(0, 2, 3, 1).contiguous().view(B, -1, 2))
        probs.append(pr.permute(0, 2, 3, 1).contiguous().view(B, -1, 2))
        boxes.append(bp.permute(0, 2, 3, 1).contiguous().view(B, -1, 6))
    return torch.cat(probs, 1), torch.cat(boxes, 1), shapes, torch.cat(scores, 1)


def box_head(pooled, sd):
    """resnet.py:256-263,345-348; stereo_rcnn.py:252-257 (eval: dropout is identity)"""
    x = F.relu(_conv(pooled, sd, "RCNN_top.0", stride=7))
    x = F.relu(_conv(x, sd, "RCNN_top.3"))
    fc7 = x.mean(3).mean(2)
    bbox = F.linear(fc7, sd["RCNN_bbox_pred.weight"], sd["RCNN_bbox_pred.bias"])
    dim = F.linear(fc7, sd["RCNN_dim_orien_pred.weight"], sd["RCNN_dim_orien_pred.bias"])
    cls = F.linear(fc7, sd["RCNN_cls_score.weight"], sd["RCNN_cls_score.bias"])
    return F.softmax(cls, 1), bbox, dim, fc7


def kpts_head(pooled, sd, chunk=64):
    """resnet.py:265-280,286; stereo_rcnn.py:260-271 (Q18)"""
    outs = []
    for i in range(0, pooled.shape[0], chunk):
        x = pooled[i:i + chunk]
        for k in range(0, 12, 2):
            x = F.relu(_conv(x, sd, "RCNN_kpts.%d" % k, pad=1))
        x = F.relu(F.conv_transpose2d(x, sd["RCNN_kpts.12.weight"], sd["RCNN_kpts.12.bias"], stride=2))
        x = _conv(x, sd, "kpts_class")                 # R,6,28,28
        outs.append(x.sum(2))                          # sum over height -> R,6,28
    ka = torch.cat(outs, 0)
    g = ka.shape[2]
    kp = F.softmax(ka[:, :4, :].contiguous().view(-1, 4 * g), 1)
    lb = F.softmax(ka[:, 4, :].contiguous().view(-1, g), 1)
    rb = F.softmax(ka[:, 5, :].contiguous().view(-1, g), 1)
    return kp, lb, rb, ka


@torch.no_grad()
def forward(sd, im_left, im_right, im_info, cfg_key="TEST", stop_after=None):
    """Test-mode forward (stereo_rcnn.py:141-324).  Returns a dict of every stage's tensors."""
    out = {}
    L = trunk_fpn(im_left, sd)
    R = trunk_fpn(im_right, sd)
    out["left"], out["right"] = L, R
    if stop_after == "fpn":
        return out
    lv = ("p2", "p3", "p4", "p5", "p6")
    cls_prob, bbox_pred, shapes, _ = rpn_head([L[k] for k in lv], [R[k] for k in lv], sd)
    out.update(rpn_cls_prob=cls_prob, rpn_bbox_pred=bbox_pred, rpn_shapes=shapes)
    if stop_after == "rpn":
        return out
    info = im_info.numpy() if torch.is_tensor(im_info) else np.asarray(im_info, np.float32)
    rl, rr = ops.proposal_layer(cls_prob.numpy(), bbox_pred.numpy(), info, cfg_key, shapes)
    out.update(rois_left=torch.from_numpy(rl), rois_right=torch.from_numpy(rr))
    if stop_after == "proposal":
        return out
    heads = heads_from_rois(sd, L, R, rl.reshape(-1, 5), rr.reshape(-1, 5), float(info[0, 0]))
    out.update(heads)
    return out


@torch.no_grad()
def heads_from_rois(sd, L, R, rois_l, rois_r, im_h):
    mk = ("p2", "p3", "p4", "p5")
    fl = [L[k].numpy() for k in mk]
    fr = [R[k].numpy() for k in mk]
    pool_l = ops.pyramid_roi_feat(fl, rois_l, im_h, 7)
    pool_r = ops.pyramid_roi_feat(fr, rois_r, im_h, 7)
    pooled = torch.from_numpy(np.concatenate([pool_l, pool_r], 1))
    cls_prob, bbox, dim, fc7 = box_head(pooled, sd)
    pool_k = torch.from_numpy(ops.pyramid_roi_feat(fl, rois_l, im_h, 14))
    kp, lb, rb, ka = kpts_head(pool_k, sd)
    return dict(pooled_box=pooled, pooled_kpts=pool_k, fc7=fc7, cls_prob=cls_prob, bbox_pred=bbox,
                dim_orien_pred=dim, kpts_prob=kp, left_border_prob=lb, right_border_prob=rb,
                kpts_pred_all=ka)
