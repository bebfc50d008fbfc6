"""Time the train-time target layers / losses on the device (CUDA events, median of 20 after 5 warm-ups, 256 MB L2
flush between iterations) and the CPU oracle (= the reference's algorithm in numpy / torch on the host cores) on the
same inputs: one training image pair per GPU x2 (config 4: batch 16 over 8 GPUs), 600x1987, 298 476 anchors,
2000 proposals.  Writes gpurun_out/train_targets_bench.md.

    python tests/tools/train_targets_bench.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ops as O                     # noqa: E402
from oracle import train_targets as T           # noqa: E402
from stereo_rcnn_b200 import synth              # noqa: E402
from stereo_rcnn_b200 import train as G         # noqa: E402


def gpu_time(fn, iters=20, warm=5):
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    ts = []
    for i in range(warm + iters):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def cpu_time(fn, iters=3):
    fn()
    ts = []
    for _ in range(iters):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e6


def main():
    H, W, B, R = 600, 1987, 2, 2000
    fs = [[int(np.ceil(H / s)), int(np.ceil(W / s))] for s in (4, 8, 16, 32, 64)]
    anchors_np = O.anchors_all_pyramids(fs).astype(np.float32)
    A = anchors_np.shape[0]
    gl, gr, gm, dim, kp, nb = synth.synth_train_gt(B, 30, H, W, 4, n_boxes=[9, 14])
    rl, rr = synth.synth_train_rois(gl, R, H, W, 5)
    rs = np.random.RandomState(0)
    w32 = lambda shape: rs.randint(0, 2 ** 32, shape, dtype=np.uint64).astype(np.uint32)
    k_a, k_p, w_p = w32((B, A)), w32((B, R + 30)), w32((B, 512))
    im_info = np.array([[H, W, 1.6]] * B, np.float32)
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    i32 = lambda a: cu(a.view(np.int32))
    d = dict(anchors=cu(anchors_np), gl=cu(gl), gr=cu(gr), gm=cu(gm), dim=cu(dim), kp=cu(kp), rl=cu(rl), rr=cu(rr),
             k_a=i32(k_a), k_p=i32(k_p), w_p=i32(w_p))
    rows = []
    at = G.anchor_targets(d["anchors"], d["gl"], d["gr"], d["gm"], (H, W), d["k_a"])
    pt = G.proposal_targets(d["rl"], d["rr"], d["gl"], d["gr"], d["dim"], d["kp"], d["k_p"], d["w_p"])
    g = torch.Generator(device="cuda").manual_seed(0)
    score = torch.randn(B, A, 2, device="cuda", generator=g)
    pred = torch.randn(B, A, 6, device="cuda", generator=g) * 0.3
    Rr, C, Gd = B * 512, 2, 28
    preds = [torch.randn(Rr, n, device="cuda", generator=g) for n in (C, 6 * C, 5 * C, 4 * Gd, Gd, Gd)]
    uncert = torch.zeros(6, device="cuda")
    t_anchor = gpu_time(lambda: G.generate_anchors(fs, "cuda"))
    t_at = gpu_time(lambda: G.anchor_targets(d["anchors"], d["gl"], d["gr"], d["gm"], (H, W), d["k_a"]))
    t_pt = gpu_time(lambda: G.proposal_targets(d["rl"], d["rr"], d["gl"], d["gr"], d["dim"], d["kp"], d["k_p"], d["w_p"]))
    t_rl = gpu_time(lambda: G.rpn_loss(score, pred, *at, uncert=uncert))
    t_cl = gpu_time(lambda: G.rcnn_loss(*preds, pt, uncert=uncert))
    # CPU: the restated reference (numpy / torch, all host threads torch gives it)
    c_anchor = cpu_time(lambda: O.anchors_all_pyramids(fs).astype(np.float32))
    c_at = cpu_time(lambda: T.anchor_target_layer(anchors_np, gl, gr, gm, im_info, T.KeySampler(k_a)))
    c_pt = cpu_time(lambda: T.proposal_target_layer(rl, rr, gl, gr, dim, kp, T.KeySampler(k_p, w_p)))
    at_c = [x.cpu() for x in at]
    sc, pc = score.cpu().requires_grad_(), pred.cpu().requires_grad_()

    def cpu_rpn():
        sc.grad = pc.grad = None
        a, b = T.rpn_losses(sc, pc, *at_c)
        (a + b).backward()
    c_rl = cpu_time(cpu_rpn)
    pt_c = {k: (v.cpu().long() if v.dtype == torch.int32 else v.cpu()) for k, v in pt.items()}
    pr_c = [p.cpu().requires_grad_() for p in preds]

    def cpu_rcnn():
        for p in pr_c:
            p.grad = None
        sum(T.rcnn_losses(*pr_c, pt_c)).backward()
    c_cl = cpu_time(cpu_rcnn)
    # algorithmic bytes (compulsory traffic) per call
    by_at = A * 16 + B * A * (4 + 4 + 16 + 16 + 4 + 4)                    # anchors + keys + the five outputs
    by_rl = B * A * (8 + 24 + 4 + 16 + 16 + 4 + 4) + B * A * (8 + 24)      # inputs + the two gradients
    rows = [("generate_anchors (A = %d)" % A, t_anchor, c_anchor, A * 16),
            ("anchor_targets (B = 2, 4 kernels)", t_at, c_at, by_at),
            ("proposal_targets (B = 2, R = 2000)", t_pt, c_pt, None),
            ("rpn_loss + gradients (B = 2)", t_rl, c_rl, by_rl),
            ("rcnn_loss + gradients (R = 1024)", t_cl, c_cl, None)]
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/train_targets_bench.md", "w") as f:
        f.write("| stage | device us | CPU oracle us (%d threads) | ratio | algorithmic MB | GB/s |\n|---|---:|---:|---:|---:|---:|\n"
                % torch.get_num_threads())
        for n, tg, tc, by in rows:
            f.write("| %s | %.1f | %.0f | %.0fx | %s | %s |\n" % (n, tg, tc, tc / tg, "%.1f" % (by / 1e6) if by else "-",
                                                                 "%.0f" % (by / tg / 1e3) if by else "-"))
        f.write("\nsum device: %.1f us; sum CPU: %.0f us\n" % (sum(r[1] for r in rows), sum(r[2] for r in rows)))
    print(open("gpurun_out/train_targets_bench.md").read())


if __name__ == "__main__":
    main()
