"""BASELINE.md section 2, the same-box baselines beside the B200 numbers (run on the GPU box; writes
gpurun_out/baselines.json):

  (a) the reference's OWN CUDA kernels compiled unmodified for sm_100a (oracle/_ref/libref_ops.so: nms_cuda_compute,
      ROIAlignForwardLaucher) against ours on the same inputs, CUDA events, 20 warm-up + 100 timed;
  (b) PyTorch / cuDNN forward of the oracle module moved to the GPU (fp32 and TF32), batch 1, the benchmarked pair;
  (c) per-op CPU timings of the oracle (all host cores): NMS N=6000, RoIAlign R=300 7x7 / 14x14, proposal layer
      A=298 476, dense_align D=128/512/2048.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import model as OM  # noqa: E402
from oracle import ops as O  # noqa: E402
from stereo_rcnn_b200 import ops as G  # noqa: E402
from stereo_rcnn_b200.synth import DEMO_P2, DEMO_P3, gen_rois, make_state_dict, synth_pair  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_ops.so")


def ev_time(fn, warm=20, reps=100):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3          # us


def cpu_time(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.time()
        fn()
        ts.append(time.time() - t)
    return float(np.mean(ts))


def main():
    out = {"host_cores": os.cpu_count()}
    torch.set_num_threads(os.cpu_count() or 1)
    dev = "cuda"
    rs = np.random.RandomState(0)
    # ---------------- (a) reference CUDA kernels vs ours
    n = 6000
    xy = rs.rand(n, 2) * 1500
    wh = rs.rand(n, 2) * 200 + 8
    dets = np.concatenate([xy, xy + wh, np.sort(rs.rand(n))[::-1][:, None]], 1).astype(np.float32)
    dd = torch.from_numpy(dets).to(dev)
    keep = torch.zeros(n, 1, dtype=torch.int32, device=dev)
    num = torch.zeros(1, dtype=torch.int32, device=dev)
    feat = torch.randn(1, 256, 150, 497, device=dev)
    x1 = rs.rand(300) * 1700
    y1 = rs.rand(300) * 450
    rois = np.stack([np.zeros(300), x1, y1, x1 + 20 + rs.rand(300) * 250, y1 + 20 + rs.rand(300) * 120], 1).astype(np.float32)
    rd = torch.from_numpy(rois).to(dev)
    res = {}
    if os.path.exists(REF_SO):
        ref = ctypes.CDLL(REF_SO)
        res["nms_ref_us_N6000"] = ev_time(lambda: ref.nms_cuda_compute(
            ctypes.c_void_p(keep.data_ptr()), ctypes.c_void_p(num.data_ptr()), ctypes.c_void_p(dd.data_ptr()), n, 5,
            ctypes.c_float(0.7)), 5, 30)
        for a in (8, 15):
            o = torch.zeros(300, 256, a, a, device=dev)
            res["roialign_ref_us_R300_%dx%d" % (a - 1, a - 1)] = ev_time(lambda: ref.ROIAlignForwardLaucher(
                ctypes.c_void_p(feat.data_ptr()), ctypes.c_float(0.25), 300, 150, 497, 256, a, a,
                ctypes.c_void_p(rd.data_ptr()), ctypes.c_void_p(o.data_ptr()), ctypes.c_void_p(0)))
    res["nms_ours_us_N6000"] = ev_time(lambda: G.nms_into(keep, dd, num, 0.7))
    for a in (8, 15):
        o = torch.zeros(300, 256, a, a, device=dev)
        res["roialign_ours_nchw_us_R300_%dx%d" % (a - 1, a - 1)] = ev_time(lambda: G.roi_align_forward(a, a, 0.25, feat, rd, o))
    fl = [torch.randn(1, h, w, 256, device=dev) for h, w in ((150, 497), (75, 249), (38, 125), (19, 63))]
    for pp in (7, 14):
        res["roialign_ours_pyramid_nhwc_fused_avg_us_R300_%dx%d" % (pp, pp)] = ev_time(
            lambda: G.roi_align_pyramid_nhwc(fl, 600.0, rd, pp, half=True))
    out["reference_cuda_kernels_vs_ours"] = res
    # ---------------- (b) cuDNN forward of the oracle module on the GPU
    sd = {k: v.to(dev) for k, v in make_state_dict(3).items()}
    left, right = synth_pair(600, 1987, 3, 48)
    iml, imr = torch.from_numpy(left)[None].to(dev), torch.from_numpy(right)[None].to(dev)
    pooled = torch.randn(300, 512, 7, 7, device=dev)
    pk = torch.randn(300, 256, 14, 14, device=dev)

    @torch.no_grad()
    def torch_forward():
        L = OM.trunk_fpn(iml, sd)
        R = OM.trunk_fpn(imr, sd)
        lv = ("p2", "p3", "p4", "p5", "p6")
        OM.rpn_head([L[k] for k in lv], [R[k] for k in lv], sd)
        OM.box_head(pooled, sd)
        OM.kpts_head(pk, sd, chunk=300)
    cud = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        cud["torch_cudnn_%s_ms" % name] = ev_time(torch_forward, 5, 20) / 1e3
    cud["what"] = ("oracle module on the GPU: trunk+FPN (L, R), RPN convs + heads, box head and keypoint head on 300 RoIs "
                   "(conv / GEMM work only: no proposal layer, RoIAlign, decode, NMS, dense_align)")
    out["torch_cudnn_on_b200"] = cud
    # ---------------- (c) per-op CPU (oracle)
    cpu = {}
    cpu["nms_N6000_s"] = cpu_time(lambda: O.nms(dets, 0.7))
    fnp = feat[0].cpu().numpy()[None]
    cpu["roialign_R300_7x7_s"] = cpu_time(lambda: O.roi_align_avg(fnp, rois, 7, 7, np.float32(0.25)), 3)
    cpu["roialign_R300_14x14_s"] = cpu_time(lambda: O.roi_align_avg(fnp, rois, 14, 14, np.float32(0.25)), 3)
    shapes = [[150, 497], [75, 249], [38, 125], [19, 63], [10, 32]]
    A = 3 * sum(h * w for h, w in shapes)
    prob = rs.rand(1, A, 2).astype(np.float32)
    bbox = (rs.randn(1, A, 6) * 0.3).astype(np.float32)
    info = np.array([[600., 1987., 1.6]], np.float32)
    cpu["proposal_layer_A298476_s"] = cpu_time(lambda: O.proposal_layer(prob, bbox, info, "TEST", shapes), 3)
    c4 = O.calib_vec(DEMO_P2, DEMO_P3)
    for D in (128, 512, 2048):
        b, k, p = gen_rois(D, seed=3)
        cpu["dense_align_D%d_s" % D] = cpu_time(lambda: O.dense_align(c4, float(np.float32(1.6)), left, right, b, k, p), 2)
    out["oracle_cpu_per_op"] = cpu
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/baselines.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
