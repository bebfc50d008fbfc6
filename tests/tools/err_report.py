"""Per-stage error of the GPU forward against the CPU oracle (diagnostic; run on the GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import model as OM  # noqa: E402
from stereo_rcnn_b200 import engine as E  # noqa: E402
from stereo_rcnn_b200.synth import make_state_dict, synth_pair  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30), np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 256)
    impl = os.environ.get("SB_CONV_IMPL", "auto")
    left, right = synth_pair(H, W, 11, 7)
    sd = make_state_dict(3)
    iml, imr = torch.from_numpy(left)[None], torch.from_numpy(right)[None]
    info = torch.tensor([[float(H), float(W), 1.0]])
    o = OM.forward(sd, iml, imr, info)
    eng = E.StereoRCNNEngine(sd, "cuda", conv_impl=impl)
    r = eng.forward(iml.cuda(), imr.cuda(), info.cuda(), keep_features=True)
    torch.cuda.synchronize()
    print("impl", impl, eng.precision, sorted(set(eng.impl_used.values())))
    for k in ("c1", "c2", "c3", "c4", "c5", "p5", "p4", "p3", "p2"):
        if r["feats"].get(k) is None:
            continue
        got = r["feats"][k].float().permute(0, 3, 1, 2).cpu().numpy()
        print("%-6s max-rel %.2e  l2-rel %.2e" % ((k,) + rel(got[0:1], o["left"][k].numpy())))
    for k in ("rpn_cls_prob", "rpn_bbox_pred"):
        print("%-14s max-rel %.2e  l2-rel %.2e" % ((k,) + rel(r[k].cpu().numpy(), o[k].numpy())))
    a = {tuple(np.round(x, 1)) for x in r["rois_left"][0].cpu().numpy()}
    b = {tuple(np.round(x, 1)) for x in o["rois_left"][0].numpy()}
    print("end-to-end proposal set overlap %.3f" % (len(a & b) / float(len(b))))
    h = eng.heads(r["feats_raw"], 1, o["rois_left"].cuda().view(-1, 5), o["rois_right"].cuda().view(-1, 5), float(H))
    torch.cuda.synchronize()
    for k in ("pooled_box", "pooled_kpts"):
        print("%-14s max-rel %.2e  l2-rel %.2e" % ((k,) + rel(h[k].float().permute(0, 3, 1, 2).cpu().numpy(), o[k].numpy())))
    for k in ("fc7", "cls_prob", "bbox_pred", "dim_orien_pred", "kpts_pred_all", "kpts_prob", "left_border_prob"):
        print("%-14s max-rel %.2e  l2-rel %.2e" % ((k,) + rel(h[k].cpu().numpy().reshape(o[k].shape), o[k].numpy())))


if __name__ == "__main__":
    main()
