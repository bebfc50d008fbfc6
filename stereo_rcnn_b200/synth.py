"""Seeded synthetic inputs (no dataset / checkpoint is reachable offline).

Shared by bench.py, the tests and tests/golden/make_golden.py so that every leg
sees byte-identical inputs.  Shapes follow BASELINE.md section 2 / SURVEY 8(d):
network-scale pair 1x3x600x1987 (a 1242x375 KITTI frame resized by 1.6), and the
config-5 pose sampler for dense_align RoIs.
"""
import zlib
from collections import OrderedDict

import numpy as np

# demo/calib.txt of the reference (KITTI 000xxx): P2 and P3, 3x4
DEMO_P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                    [0.0, 721.5377, 172.854, 0.2163791],
                    [0.0, 0.0, 1.0, 0.002745884]], np.float64)
DEMO_P3 = np.array([[721.5377, 0.0, 609.5593, -339.5242],
                    [0.0, 721.5377, 172.854, 2.199936],
                    [0.0, 0.0, 1.0, 0.002729905]], np.float64)


def _bicubic_up(base, H, W):
    import torch
    t = torch.from_numpy(base)[None]
    return torch.nn.functional.interpolate(t, size=(H, W), mode="bicubic", align_corners=True)[0].numpy()


def synth_pair(H, W, seed, shift=12):
    """smooth seeded network-scale pair (3,H,W) fp32; right = left shifted `shift` px"""
    rs = np.random.RandomState(seed)
    base = rs.randn(3, H // 4 + 2, W // 4 + 2).astype(np.float32)
    left = (_bicubic_up(base, H, W) * 40.0).astype(np.float32)
    right = np.empty_like(left)
    right[:, :, :-shift] = left[:, :, shift:]
    right[:, :, -shift:] = left[:, :, -1:]
    return left, right


def gen_rois(D, seed, p2=DEMO_P2, width=1242, height=375):
    """seeded 3D boxes projected with P2 -> box_left [D,4], keypoints [D,5], poses [D,7]"""
    r = np.random.RandomState(seed)
    out = []
    while len(out) < D:
        z = r.uniform(6, 60); x = r.uniform(-0.45 * z, 0.45 * z); y = 1.65
        w, h, l = r.normal([1.6, 1.5, 3.9], 0.1); th = r.uniform(-np.pi, np.pi)
        c, s = np.cos(th), np.sin(th)
        xs = np.array([-w / 2, -w / 2, w / 2, w / 2] * 2); ys = np.array([0, 0, 0, 0, -h, -h, -h, -h])
        zs = np.array([-l / 2, l / 2, l / 2, -l / 2] * 2)
        X = c * xs + s * zs + x; Y = ys + y; Z = -s * xs + c * zs + z
        if Z.min() < 1:
            continue
        u = p2[0, 0] * X / Z + p2[0, 2]; v = p2[1, 1] * Y / Z + p2[1, 2]
        x1, x2 = np.clip([u.min(), u.max()], 0, width - 1)
        y1, y2 = np.clip([v.min(), v.max()], 0, height - 1)
        if x2 - x1 < 10 or y2 - y1 < 10:
            continue
        out.append(([x1, y1, x2, y2], [(x1 + x2) / 2, 1, 0.9, x1, x2], [x, y, z, w, h, l, th]))
    f = np.float32
    return (np.array([o[0] for o in out], f), np.array([o[1] for o in out], f),
            np.array([o[2] for o in out], f))


LAYERS = [3, 4, 23, 3]
PLANES = [64, 128, 256, 512]


# --------------------------------------------------------------------------
# state dict layout
# --------------------------------------------------------------------------
def param_shapes(n_classes=2):
    """ordered {key: shape} of every tensor the forward reads (reference key names)"""
    s = OrderedDict()

    def conv(k, co, ci, kh, kw, bias):
        s[k + ".weight"] = (co, ci, kh, kw)
        if bias:
            s[k + ".bias"] = (co,)

    def bn(k, c):
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[k + "." + n] = (c,)

    conv("RCNN_layer0.0", 64, 3, 7, 7, False)
    bn("RCNN_layer0.1", 64)
    inpl = 64
    for li, (nb, pl) in enumerate(zip(LAYERS, PLANES)):
        for b in range(nb):
            p = "RCNN_layer%d.0.%d" % (li + 1, b)
            conv(p + ".conv1", pl, inpl, 1, 1, False); bn(p + ".bn1", pl)
            conv(p + ".conv2", pl, pl, 3, 3, False); bn(p + ".bn2", pl)
            conv(p + ".conv3", pl * 4, pl, 1, 1, False); bn(p + ".bn3", pl * 4)
            if b == 0:
                conv(p + ".downsample.0", pl * 4, inpl, 1, 1, False); bn(p + ".downsample.1", pl * 4)
            inpl = pl * 4
    conv("RCNN_toplayer", 256, 2048, 1, 1, True)
    for i in (1, 2, 3):
        conv("RCNN_smooth%d" % i, 256, 256, 3, 3, True)
    conv("RCNN_latlayer1", 256, 1024, 1, 1, True)
    conv("RCNN_latlayer2", 256, 512, 1, 1, True)
    conv("RCNN_latlayer3", 256, 256, 1, 1, True)
    conv("RCNN_rpn.RPN_Conv", 512, 256, 3, 3, True)
    conv("RCNN_rpn.RPN_cls_score", 6, 1024, 1, 1, True)
    conv("RCNN_rpn.RPN_bbox_pred_left_right", 18, 1024, 1, 1, True)
    conv("RCNN_top.0", 2048, 512, 7, 7, True)
    conv("RCNN_top.3", 2048, 2048, 1, 1, True)
    for i in range(0, 12, 2):
        conv("RCNN_kpts.%d" % i, 256, 256, 3, 3, True)
    s["RCNN_kpts.12.weight"] = (256, 256, 2, 2)      # ConvTranspose2d: (Cin, Cout, kh, kw)
    s["RCNN_kpts.12.bias"] = (256,)
    s["RCNN_cls_score.weight"] = (n_classes, 2048); s["RCNN_cls_score.bias"] = (n_classes,)
    s["RCNN_bbox_pred.weight"] = (6 * n_classes, 2048); s["RCNN_bbox_pred.bias"] = (6 * n_classes,)
    s["RCNN_dim_orien_pred.weight"] = (5 * n_classes, 2048); s["RCNN_dim_orien_pred.bias"] = (5 * n_classes,)
    conv("kpts_class", 6, 256, 1, 1, True)
    return s


def make_state_dict(seed=3, n_classes=2, head_gain=1.0):
    """Deterministic synthetic weights, one private RNG stream per key.

    "Variance-preserving" variant of the reference's random init (SURVEY 7,
    hard part "random-init weights saturate the RPN"): He-normal conv weights,
    non-trivial frozen-BN statistics, and a small ``bn3.weight`` so the residual
    trunk neither explodes nor collapses; head weights scaled so that RPN scores
    and box deltas are spread out (non-degenerate NMS / top-k work).
    """
    import torch
    sd = OrderedDict()
    for k, shp in param_shapes(n_classes).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        leaf = k.rsplit(".", 1)[1]
        is_bn = ".bn" in k or "downsample.1" in k or k.startswith("RCNN_layer0.1")
        if is_bn:
            n = shp[0]
            if leaf == "weight":
                lo, hi = (0.15, 0.35) if ".bn3." in k else (0.8, 1.2)
                t = torch.rand(n, generator=g) * (hi - lo) + lo
            elif leaf == "bias":
                t = torch.randn(n, generator=g) * 0.05
            elif leaf == "running_mean":
                t = torch.randn(n, generator=g) * 0.1
            else:
                t = torch.rand(n, generator=g) * 0.4 + 0.8
        elif leaf == "bias":
            t = torch.randn(shp, generator=g) * 0.02
        else:
            fan_in = int(np.prod(shp[1:]))
            if k.startswith("RCNN_kpts.12"):
                fan_in = shp[0]
            std = (2.0 / fan_in) ** 0.5
            if k.startswith(("RCNN_cls_score", "RCNN_bbox_pred", "RCNN_dim_orien_pred",
                             "RCNN_rpn.RPN_cls_score", "RCNN_rpn.RPN_bbox_pred", "kpts_class")):
                std = head_gain * (1.0 / fan_in) ** 0.5
            if k.startswith("RCNN_rpn.RPN_bbox_pred"):
                std *= 0.25
            # fixed gains (calibrated once on a 200x333 input so that every stage's
            # activations have rms ~1; see DESIGN.md "synthetic weights")
            for pre, gain in _GAINS:
                if k.startswith(pre):
                    std *= gain
            t = torch.randn(shp, generator=g) * std
        sd[k] = t.float().contiguous()
    return sd


_GAINS = (("RCNN_layer0.0", 1.0 / 64), ("RCNN_toplayer", 0.0625), ("RCNN_latlayer1", 0.075),
          ("RCNN_latlayer2", 0.22), ("RCNN_latlayer3", 0.32), ("RCNN_smooth", 0.5),
          ("RCNN_rpn.RPN_cls_score", 2.0), ("kpts_class", 0.1))




def make_reference_init_state_dict(seed=3, n_classes=2):
    """The reference's OWN random initialisation, restated with seeded per-key RNG streams: ResNet convs
    N(0, sqrt(2 / (kh*kw*Cout))) and BatchNorm weight 1 / bias 0 / mean 0 / var 1 (lib/model/stereo_rcnn/resnet.py:123-129);
    FPN / RPN / predictor layers N(0, 0.01) (bbox_pred, dim_orien_pred 0.001; kpts_class 0.1) with zero bias, RCNN_top and
    RCNN_kpts N(0, 0.02) with zero bias (stereo_rcnn.py:47-85).  Un-normalised: activations grow to ~1e6 by C4, far
    outside fp16 -- the tf32 precision mode is the one that runs such weights."""
    import torch
    sd = OrderedDict()
    for k, shp in param_shapes(n_classes).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode()) + 77) % (2 ** 31))
        leaf = k.rsplit(".", 1)[1]
        is_bn = ".bn" in k or "downsample.1" in k or k.startswith("RCNN_layer0.1")
        if is_bn:
            t = torch.ones(shp) if leaf in ("weight", "running_var") else torch.zeros(shp)
        elif leaf == "bias":
            t = torch.zeros(shp)
        elif k.startswith("RCNN_layer"):
            t = torch.randn(shp, generator=g) * (2.0 / (shp[2] * shp[3] * shp[0])) ** 0.5
        else:
            std = 0.01
            if k.startswith(("RCNN_bbox_pred", "RCNN_dim_orien_pred")):
                std = 0.001
            elif k.startswith("kpts_class"):
                std = 0.1
            elif k.startswith(("RCNN_top", "RCNN_kpts")):
                std = 0.02
            t = torch.randn(shp, generator=g) * std
        sd[k] = t.float().contiguous()
    return sd


def synth_train_gt(B, K, H, W, seed, n_boxes=None, disparity=6.0):
    """ground truth of a training batch in the layout of roibatchLoader.py:199-240: gt_left / gt_right / gt_merge
    [B,K,5] (x1, y1, x2, y2, class; zero rows pad to K = MAX_NUM_GT_BOXES), gt_dim_orien [B,K,5], gt_kpts [B,K,6]
    (four perspective keypoints, one visible, -1 otherwise; left / right border), all fp32"""
    rng = np.random.RandomState(seed)
    n_boxes = n_boxes or [min(K, 3 + 2 * b) for b in range(B)]
    gl = np.zeros((B, K, 5), np.float32)
    for b in range(B):
        for j in range(n_boxes[b]):
            w = rng.uniform(0.08 * W, 0.35 * W)
            h = rng.uniform(0.1 * H, 0.4 * H)
            x1 = rng.uniform(disparity + 2, W - w - 1)
            y1 = rng.uniform(0, H - h - 1)
            gl[b, j] = [x1, y1, x1 + w, y1 + h, 1]
    gr = gl.copy()
    gr[:, :, 0] -= disparity
    gr[:, :, 2] -= disparity
    gr[gl[:, :, 4] == 0] = 0
    gm = gl.copy()
    gm[:, :, 0] = np.minimum(gl[:, :, 0], gr[:, :, 0])
    gm[:, :, 2] = np.maximum(gl[:, :, 2], gr[:, :, 2])
    dim = rng.uniform(-1, 4, (B, K, 5)).astype(np.float32)
    kp = np.zeros((B, K, 6), np.float32)
    for b in range(B):
        for j in range(K):
            x1, x2 = gl[b, j, 0], gl[b, j, 2]
            kp[b, j, :4] = -1
            kp[b, j, rng.randint(4)] = rng.uniform(x1, x2)
            kp[b, j, 4] = rng.uniform(x1, x1 + 0.3 * (x2 - x1))
            kp[b, j, 5] = rng.uniform(x2 - 0.3 * (x2 - x1), x2)
    return gl, gr.astype(np.float32), gm.astype(np.float32), dim, kp, np.asarray(n_boxes, np.int64)


def synth_train_rois(gt_left, R, H, W, seed, near_gt=0.34, jitter=6.0, disparity=6.0):
    """proposals of a training step: a fraction `near_gt` jittered copies of ground-truth boxes (foreground
    candidates), the rest uniform boxes -> rois_left, rois_right [B,R,5] (batch index first)"""
    rng = np.random.RandomState(seed)
    B = gt_left.shape[0]
    out = np.zeros((B, R, 5), np.float32)
    for b in range(B):
        n = int((gt_left[b, :, 4] > 0).sum())
        for r in range(R):
            if n > 0 and rng.uniform() < near_gt:
                out[b, r, 1:] = gt_left[b, rng.randint(n), :4] + rng.uniform(-jitter, jitter, 4)
            else:
                w = rng.uniform(0.04 * W, 0.4 * W)
                h = rng.uniform(0.06 * H, 0.5 * H)
                x1 = rng.uniform(disparity, W - w - 1)
                y1 = rng.uniform(0, H - h - 1)
                out[b, r, 1:] = [x1, y1, x1 + w, y1 + h]
        out[b, :, 0] = b
    right = out.copy()
    right[:, :, 1] -= disparity
    right[:, :, 3] -= disparity
    return out, right
