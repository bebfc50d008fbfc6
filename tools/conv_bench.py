"""Micro-benchmark of sb_conv2d_tc on the layer shapes of the forward (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from stereo_rcnn_b200 import ops  # noqa: E402

SHAPES = [
    # name, N, H, W, Cin, Cout, k, residual
    ("tiny (fixed overhead)", 1, 8, 16, 32, 32, 1, False),
    ("tiny 3x3", 1, 8, 16, 64, 64, 3, False),
    ("L1 conv1 256->64", 2, 150, 497, 256, 64, 1, False),
    ("L1 conv2 3x3 64", 2, 150, 497, 64, 64, 3, False),
    ("L1 conv3 64->256 +res", 2, 150, 497, 64, 256, 1, True),
    ("L1 conv3 64->256 nores", 2, 150, 497, 64, 256, 1, False),
    ("L1 1x1 256->256 nores", 2, 150, 497, 256, 256, 1, False),
    ("L1 1x1 256->256 +res", 2, 150, 497, 256, 256, 1, True),
    ("L2 conv3 128->512 +res", 2, 75, 249, 128, 512, 1, True),
    ("L3 conv1 1024->256", 2, 38, 125, 1024, 256, 1, False),
    ("L3 conv2 3x3 256", 2, 38, 125, 256, 256, 3, False),
    ("L3 conv3 256->1024 +res", 2, 38, 125, 256, 1024, 1, True),
    ("L4 conv2 3x3 512", 2, 19, 63, 512, 512, 3, False),
    ("smooth3 3x3 256 P2", 2, 150, 497, 256, 256, 3, False),
    ("rpn P2 3x3 256->512", 1, 150, 497, 256, 512, 3, False),
    ("kpts 3x3 256 (300x14x14)", 300, 14, 14, 256, 256, 3, False),
    ("FC 25088->2048 (300)", 300, 1, 1, 25088, 2048, 1, False),
]


def main():
    dev = "cuda"
    reps = 20
    a = torch.randn(2, 150, 497, 256, device=dev)
    b = torch.empty_like(a)
    c = torch.randn_like(a)
    for name, fn in () if len(sys.argv) > 1 else (("torch copy 153MB", lambda: b.copy_(a)), ("torch add 153MB x2 -> 153MB", lambda: torch.add(a, c, out=b))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        print("%-28s %8.1f us" % (name, s.elapsed_time(e) * 1e3 / reps))
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, N, H, W, Ci, Co, k, res in SHAPES:
        if only and only not in name:
            continue
        half = os.environ.get("CB_HALF", "1") == "1" and Ci % 64 == 0
        x = torch.randn(N, H, W, Ci, device=dev)
        w = torch.randn(Co, k, k, Ci, device=dev) * 0.05
        if half:
            x, w = x.half(), w.half()
        sc = torch.rand(Co, device=dev) + 0.5
        sh = torch.randn(Co, device=dev)
        out = torch.empty(N, H, W, Co, device=dev)
        r = torch.randn(N, H, W, Co, device=dev) if res else None
        line = "%-28s" % name
        for bn in ("128", "256", "small"):
            if Co < 256 and bn == "256":
                continue
            if bn == "small":
                os.environ["SB_TC_SMALL"] = "1"
                if Co < 128:
                    continue
            else:
                os.environ["SB_TC_SMALL"] = "0"
                os.environ["SB_TC_BLOCK_N"] = bn
            d = ops.conv_desc(x, w, out, Ci, Co, k, k, 1, k // 2, H, W, scale=sc, shift=sh, residual=r, relu=True)
            for _ in range(3):
                ops.conv2d(d, "tc")
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                ops.conv2d(d, "tc")
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) * 1e3 / reps
            fl = 2.0 * N * H * W * Ci * Co * k * k
            line += "  %s %7.1f us %6.1f TF" % (bn, us, fl / us / 1e6)
        print(line)


if __name__ == "__main__":
    main()
