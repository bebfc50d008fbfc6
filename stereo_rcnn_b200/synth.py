"""Seeded synthetic inputs (no dataset / checkpoint is reachable offline).

Shared by bench.py, the tests and tests/golden/make_golden.py so that every leg
sees byte-identical inputs.  Shapes follow BASELINE.md section 2 / SURVEY 8(d):
network-scale pair 1x3x600x1987 (a 1242x375 KITTI frame resized by 1.6), and the
config-5 pose sampler for dense_align RoIs.
"""
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
