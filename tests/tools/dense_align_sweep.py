"""BASELINE.json configs[4]: dense_align-only throughput sweep, D = 128 / 512 / 2048 RoIs per image.

Reports per D: GPU time (CUDA events, 20 reps after warm-up, L2 flushed between reps), SAD evaluations/s
(70 hypotheses x valid pixels x D / t), achieved GB/s on the compulsory bytes of SURVEY 8(d)
(12*[P + rows*(W_span + D_span + 2)] + 64 per RoI + the image pair once), and the oracle's CPU time.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ops as O  # noqa: E402
from stereo_rcnn_b200 import ops  # noqa: E402
from stereo_rcnn_b200.synth import DEMO_P2, DEMO_P3, gen_rois, synth_pair  # noqa: E402


def main():
    """single process: 1 GPU.  Under torchrun (WORLD_SIZE = N): one image x D RoIs per rank (SURVEY 8d config 5,
    "8 GPUs = 8 images x D"), no exchange; the aggregate is N*D RoIs / max-over-ranks time."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W = 600, 1987
    left, right = synth_pair(H, W, 3 + rank, 48)
    c4 = ops.calib_vec(DEMO_P2, DEMO_P3)
    scale = float(np.float32(1.6))
    iml, imr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    rows = []
    for D in (128, 512, 2048):
        b, k, p = gen_rois(D, seed=3 + rank)
        bd, kd, pd = (torch.from_numpy(x).cuda() for x in (b, k, p))
        for _ in range(3):
            st, dis = ops.dense_align(c4, scale, iml, imr, bd, kd, pd)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            ops.l2_flush(flush)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            st, dis = ops.dense_align(c4, scale, iml, imr, bd, kd, pd)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = float(np.median(ts))
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_max = float(t[0])
            if rank != 0:
                continue
            rows.append(dict(D_per_gpu=D, n_gpus=world, gpu_ms_max_over_ranks=round(ms_max, 3),
                             rois_per_s_aggregate=round(world * D / (ms_max / 1e3), 1)))
            print(json.dumps(rows[-1]))
            continue
        t0 = time.time()
        st_o, dis_o, dg = O.dense_align(c4, scale, left, right, b, k, p, diagnostics=True)
        cpu_s = time.time() - t0
        P = dg["npix"].astype(np.float64)
        agree = float(np.mean(np.abs(dis.cpu().numpy() - dis_o) <= 1e-5 * np.abs(dis_o)))
        # compulsory bytes: per RoI 12*[P + rows*(W_span + D_span + 2)] + 64; spans from the lattice
        s2 = 3.2
        hh = (b[:, 3] - b[:, 1]) * s2
        ww = (k[:, 4] - k[:, 3]) * s2
        rows_n = np.maximum(0.4 * hh / np.maximum(np.floor(hh / 56), 1), 1)
        fb = DEMO_P2[0, 0] * s2 * ((DEMO_P2[0, 3] - DEMO_P3[0, 3]) / DEMO_P2[0, 0])
        z = p[:, 2]
        dspan = fb / np.maximum(z - 12.5, 1.5) - fb / (z + 12.0)
        comp = float(np.sum(12 * (P + rows_n * (ww + dspan + 2)) + 64) + 2 * 3 * H * W * 4)
        rows.append(dict(D=D, gpu_ms=round(ms, 3), sad_evals_per_s=round(70 * P.sum() / (ms / 1e3), 1),
                         compulsory_MB=round(comp / 1e6, 2), achieved_GBs=round(comp / (ms / 1e3) / 1e9, 1),
                         valid_px_mean=round(float(P.mean()), 1), oracle_cpu_s=round(cpu_s, 2),
                         argmin_agreement=round(agree, 4), rois_per_s=round(D / (ms / 1e3), 1)))
        print(json.dumps(rows[-1]))
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(rows, open("gpurun_out/dense_align_sweep%s.json" % ("_%dgpu" % world if world > 1 else ""), "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
