"""Which byte stream bounds the residual epilogue?  Times the layer-3 conv3 shape (256 -> 1024, one image) with
the fp32 output / fp16 twin / residual read switched on and off, alone on the GPU, warm L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from stereo_rcnn_b200 import ops  # noqa: E402


def timeit(d, reps=40):
    for _ in range(5):
        ops.conv2d(d, "tc")
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        ops.conv2d(d, "tc")
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / reps


def main():
    dev = "cuda"
    only_shape = sys.argv[1] if len(sys.argv) > 1 else None       # e.g. "L1 conv3"
    only_tag = sys.argv[2] if len(sys.argv) > 2 else None         # e.g. "o32+o16+res" (for ncu captures)
    for name, N, H, W, Ci, Co, k in (("L3 conv3", 1, 38, 125, 256, 1024, 1), ("L3 conv3 x2", 2, 38, 125, 256, 1024, 1),
                                     ("L3 conv1", 1, 38, 125, 1024, 256, 1), ("L3 conv2", 1, 38, 125, 256, 256, 3),
                                     ("L2 conv3", 2, 75, 249, 128, 512, 1), ("L1 conv3", 2, 150, 497, 64, 256, 1)):
        if only_shape and only_shape != name:
            continue
        x = torch.randn(N, H, W, Ci, device=dev).half()
        w = (torch.randn(Co, k, k, Ci, device=dev) * 0.05).half()
        sc, sh = torch.rand(Co, device=dev) + 0.5, torch.randn(Co, device=dev)
        o32 = torch.empty(N, H, W, Co, device=dev)
        o16 = torch.empty(N, H, W, Co, device=dev, dtype=torch.float16)
        r = torch.randn(N, H, W, Co, device=dev)
        mb = N * H * W * Co / 1e6
        for cap in (0, 74):
            line = "%-12s cap %3d |" % (name, cap)
            for tag, out, out16, res in (("o32+o16+res", o32, o16, r), ("o32+res", o32, None, r), ("o16+res", None, o16, r),
                                         ("o32+o16", o32, o16, None), ("o32", o32, None, None), ("o16", None, o16, None)):
                if (res is not None and k != 1) or (only_tag and (only_tag != tag or cap)):
                    continue
                d = ops.conv_desc(x, w, out, Ci, Co, k, k, 1, k // 2, H, W, scale=sc, shift=sh, residual=res, relu=True,
                                  out16=out16, max_ctas=cap)
                us = timeit(d)
                by = mb * ((4 if out is not None else 0) + (2 if out16 is not None else 0) + (4 if res is not None else 0))
                line += " %s %6.1f us %5.2f TB/s |" % (tag, us, by / us)
            print(line, flush=True)


if __name__ == "__main__":
    main()
