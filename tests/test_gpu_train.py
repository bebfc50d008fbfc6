"""GPU: the train-time target layers and losses (csrc/train_targets.cu, csrc/train_loss.cu) against the CPU oracle
(oracle/train_targets.py, pinned to the reference's own layers by tests/test_train_targets.py) on the same seeded
inputs and the same random words -- labels, sampled indices, weights and integer targets bit-exact, regression
targets to the last ulp of log(), losses / gradients against torch autograd of the restated reference formulas."""
import numpy as np
import pytest
import torch

from oracle import ops as O
from oracle import train_targets as T
from stereo_rcnn_b200 import synth
from stereo_rcnn_b200 import train as G

pytestmark = pytest.mark.gpu


def cu(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def words_np(shape, seed):
    return np.random.RandomState(seed).randint(0, 2 ** 32, shape, dtype=np.uint64).astype(np.uint32)


def as_i32(w):
    return cu(w.view(np.int32))


def feat_shapes(H, W):
    return [[int(np.ceil(H / s)), int(np.ceil(W / s))] for s in (4, 8, 16, 32, 64)]


@pytest.mark.parametrize("H,W,B,batch,seed", [(160, 256, 2, 512, 0), (160, 256, 3, 32, 5), (600, 1987, 2, 512, 7),
                                               (224, 400, 2, 16, 11)])
def test_anchor_targets_vs_oracle(H, W, B, batch, seed):
    """anchor_target_layer.py:42-164; 600x1987 is the benchmark image (A = 298 476 anchors); the small-batch cases force
    both subsampling branches (and, with 16, `num_bg` from a foreground count above the cap)"""
    anchors = O.anchors_all_pyramids(feat_shapes(H, W)).astype(np.float32)
    A = anchors.shape[0]
    gl, gr, gm, _dim, _kp, _nb = synth.synth_train_gt(B, 30, H, W, seed)
    im_info = np.array([[H, W, 1.6]] * B, np.float32)
    keys = words_np((B, A), seed + 1)
    keys[:, ::7] = keys[:, 3:4]                      # plenty of equal keys: ties resolve by index
    cfg = dict(T.CFG, RPN_BATCHSIZE=batch)
    ref = T.anchor_target_layer(anchors, gl, gr, gm, im_info, T.KeySampler(keys), cfg)
    got = G.anchor_targets(cu(anchors), cu(gl), cu(gr), cu(gm), (H, W), as_i32(keys), cfg)
    lab, tl, tr, iw, ow = [t.cpu().numpy() for t in got]
    np.testing.assert_array_equal(lab, ref[0])
    np.testing.assert_array_equal(iw, ref[3])
    np.testing.assert_array_equal(ow, ref[4])
    np.testing.assert_allclose(tl, ref[1], rtol=0, atol=4e-7)
    np.testing.assert_allclose(tr, ref[2], rtol=0, atol=4e-7)
    assert (lab == 1).sum() > 0 and (lab == 0).sum() > 0
    assert ((lab >= 0).sum(1) <= batch).all()


@pytest.mark.parametrize("R,near,seed", [(200, 0.34, 1), (700, 0.6, 2), (2000, 0.3, 3), (60, 1.0, 4), (300, 0.0, 6)])
def test_proposal_targets_vs_oracle(R, near, seed):
    """proposal_target_layer.py:36-333 incl. fewer / more foreground candidates than 128, TRAIN.RPN_POST_NMS_TOP_N =
    2000 proposals, and the one-sided cases (:249-265): near = 0 -> no proposal overlaps (the appended gt boxes are
    the only foreground), near = 1 with tight jitter -> hardly any background"""
    H, W, B = 160, 256, 2
    gl, gr, _gm, dim, kp, _nb = synth.synth_train_gt(B, 30, H, W, seed)
    rl, rr = synth.synth_train_rois(gl, R, H, W, seed + 100, near_gt=near, jitter=6.0 if near < 1 else 1.0)
    keys = words_np((B, R + 30), seed + 1)
    keys[:, ::5] = keys[:, 2:3]
    words = words_np((B, 512), seed + 2)
    ref = T.proposal_target_layer(rl, rr, gl, gr, dim, kp, T.KeySampler(keys, words))
    got = G.proposal_targets(cu(rl), cu(rr), cu(gl), cu(gr), cu(dim), cu(kp), as_i32(keys), as_i32(words))
    assert got["status"].cpu().tolist() == [0] * B
    np.testing.assert_array_equal(got["keep_inds"].cpu().numpy(), ref["keep_inds"])
    for n in ("rois_left", "rois_right", "labels", "dim_orien_targets", "kpts_weight", "inside_w", "outside_w"):
        np.testing.assert_array_equal(got[n].cpu().numpy(), ref[n], err_msg=n)
    np.testing.assert_array_equal(got["kpts_targets"].cpu().numpy().astype(np.int64), ref["kpts_targets"])
    for n in ("bbox_targets_left", "bbox_targets_right"):
        np.testing.assert_allclose(got[n].cpu().numpy(), ref[n], rtol=0, atol=4e-6, err_msg=n)


def test_anchor_targets_more_than_32_gt_boxes():
    """K = 40 ground-truth rows (the wide instance of the overlap kernel)"""
    H, W, B = 160, 256, 2
    anchors = O.anchors_all_pyramids(feat_shapes(H, W)).astype(np.float32)
    gl, gr, gm, _dim, _kp, _nb = synth.synth_train_gt(B, 40, H, W, 21, n_boxes=[37, 40])
    keys = words_np((B, anchors.shape[0]), 22)
    ref = T.anchor_target_layer(anchors, gl, gr, gm, np.array([[H, W, 1.6]] * B, np.float32), T.KeySampler(keys))
    got = G.anchor_targets(cu(anchors), cu(gl), cu(gr), cu(gm), (H, W), as_i32(keys))
    np.testing.assert_array_equal(got[0].cpu().numpy(), ref[0])
    np.testing.assert_array_equal(got[4].cpu().numpy(), ref[4])
    np.testing.assert_allclose(got[1].cpu().numpy(), ref[1], rtol=0, atol=4e-7)


def test_proposal_targets_no_candidates_sets_status():
    """no ground truth at all and zero-area proposals: neither foreground nor background (:267 raises)"""
    B, R = 1, 8
    z = np.zeros((B, 30, 5), np.float32)
    rl = np.zeros((B, R, 5), np.float32)
    got = G.proposal_targets(cu(rl), cu(rl), cu(z), cu(z), cu(z), cu(np.zeros((B, 30, 6), np.float32)),
                             as_i32(words_np((B, R + 30), 0)), as_i32(words_np((B, 512), 1)))
    assert got["status"].cpu().tolist() == [1]
    assert float(got["labels"].abs().sum()) == 0 and float(got["inside_w"].abs().sum()) == 0


def test_proposal_targets_background_only():
    """an image without ground truth: every proposal is background, drawn with replacement (:258-265)"""
    H, W, B, R = 160, 256, 1, 90
    z = np.zeros((B, 30, 5), np.float32)
    gl, _gr, _gm, dim, kp, _nb = synth.synth_train_gt(B, 30, H, W, 9)
    rl, rr = synth.synth_train_rois(gl, R, H, W, 109, near_gt=0.0)
    keys, words = words_np((B, R + 30), 1), words_np((B, 512), 2)
    ref = T.proposal_target_layer(rl, rr, z, z, dim, kp, T.KeySampler(keys, words))
    got = G.proposal_targets(cu(rl), cu(rr), cu(z), cu(z), cu(dim), cu(kp), as_i32(keys), as_i32(words))
    assert got["status"].cpu().tolist() == [0]
    np.testing.assert_array_equal(got["keep_inds"].cpu().numpy(), ref["keep_inds"])
    np.testing.assert_array_equal(got["rois_left"].cpu().numpy(), ref["rois_left"])
    assert float(got["labels"].abs().sum()) == 0 and float(ref["labels"].sum()) == 0


def _train_case(seed=0):
    H, W, B = 160, 256, 2
    anchors = O.anchors_all_pyramids(feat_shapes(H, W)).astype(np.float32)
    gl, gr, gm, dim, kp, _nb = synth.synth_train_gt(B, 30, H, W, seed)
    keys = words_np((B, anchors.shape[0]), seed + 1)
    at = T.anchor_target_layer(anchors, gl, gr, gm, np.array([[H, W, 1.6]] * B, np.float32), T.KeySampler(keys))
    rl, rr = synth.synth_train_rois(gl, 300, H, W, seed + 100)
    pt = T.proposal_target_layer(rl, rr, gl, gr, dim, kp,
                                 T.KeySampler(words_np((B, 330), seed + 2), words_np((B, 512), seed + 3)))
    return B, anchors.shape[0], at, pt


def test_rpn_loss_and_gradients_vs_autograd():
    """stereo_rpn.py:114-140 + net_utils.py:79-99 through torch autograd on the CPU"""
    B, A, at, _pt = _train_case()
    g = torch.Generator().manual_seed(0)
    score = (torch.randn(B, A, 2, generator=g) * 2).requires_grad_()
    pred = (torch.randn(B, A, 6, generator=g) * 0.3).requires_grad_()
    uncert = torch.tensor([0.3, -0.2, 0.1, 0.0, 0.5, -0.4])
    t = [torch.from_numpy(x) for x in at]
    lc, lb = T.rpn_losses(score, pred, *t)
    (lc * torch.exp(-uncert[0]) + lb * torch.exp(-uncert[1])).backward()
    losses, d_cls, d_box = G.rpn_loss(score.detach().cuda(), pred.detach().cuda(), *[x.cuda() for x in t],
                                      uncert=uncert.cuda())
    np.testing.assert_allclose(losses.cpu().numpy(), [float(lc.detach()), float(lb.detach())], rtol=2e-6)
    assert float((d_cls.cpu() - score.grad).abs().max()) <= 1e-6 * float(score.grad.abs().max()) + 1e-12
    assert float((d_box.cpu() - pred.grad).abs().max()) <= 1e-6 * float(pred.grad.abs().max()) + 1e-12
    assert float(score.grad.abs().max()) > 0 and float(pred.grad.abs().max()) > 0
    again = G.rpn_loss(score.detach().cuda(), pred.detach().cuda(), *[x.cuda() for x in t], uncert=uncert.cuda())
    assert torch.equal(again[0], losses) and torch.equal(again[1], d_cls)        # fixed-order reductions


def test_rcnn_loss_and_gradients_vs_autograd():
    """stereo_rcnn.py:201-311"""
    B, _A, _at, pt = _train_case(3)
    R, C, Gd = B * 512, 2, 28
    g = torch.Generator().manual_seed(1)
    preds = [(torch.randn(R, n, generator=g) * s).requires_grad_() for n, s in
             ((C, 2.0), (6 * C, 0.5), (5 * C, 0.5), (4 * Gd, 1.5), (Gd, 1.5), (Gd, 1.5))]
    uncert = torch.tensor([0.3, -0.2, 0.1, 0.25, 0.5, -0.4])
    tt = {k: torch.from_numpy(v) for k, v in pt.items()}
    ls = T.rcnn_losses(*preds, tt)
    sum(l * torch.exp(-uncert[2 + i]) for i, l in enumerate(ls)).backward()
    tg = {k: (v.cuda().to(torch.int32) if k in ("kpts_targets", "keep_inds") else v.cuda()) for k, v in tt.items()}
    losses, grads = G.rcnn_loss(*[p.detach().cuda() for p in preds], tg, uncert=uncert.cuda())
    np.testing.assert_allclose(losses.cpu().numpy(), [float(l.detach()) for l in ls], rtol=3e-6)
    assert float(tt["kpts_weight"].sum()) > 3
    for p, d, n in zip(preds, grads, ("cls", "bbox", "dim", "kpts", "left", "right")):
        assert float(p.grad.abs().max()) > 0, n
        assert float((d.cpu() - p.grad).abs().max()) <= 2e-6 * float(p.grad.abs().max()) + 1e-12, n


def test_multitask_loss_and_clip_gradient():
    """trainval_net.py:214-219 and net_utils.py:37-49"""
    losses = torch.tensor([0.7, 0.05, 0.4, 0.3, 1.2, 2.5], requires_grad=True)
    uncert = torch.tensor([0.3, -0.2, 0.1, 0.25, 0.5, -0.4], requires_grad=True)
    total = T.multitask_loss([l for l in losses], uncert)
    total.backward()
    tot, d_u = G.multitask_loss(losses.detach().cuda(), uncert.detach().cuda())
    assert abs(float(tot) - float(total.detach())) <= 2e-6 * abs(float(total.detach()))
    np.testing.assert_allclose(d_u.cpu().numpy(), uncert.grad.numpy(), rtol=2e-6, atol=1e-7)
    g = torch.Generator().manual_seed(2)
    shapes = [(64, 3, 7, 7), (2048,), (1, ), (1024, 25088 // 16)] + [(17, 13)] * 60        # > one launch chunk
    for scale, clipped in ((5.0, True), (1e-4, False)):
        grads = [(torch.randn(*s, generator=g) * scale) for s in shapes]
        norm = float(np.sqrt(sum(float(t.double().pow(2).sum()) for t in grads)))
        dev = [t.clone().cuda() for t in grads]
        out = G.clip_gradient(dev, 10.0).cpu().numpy()
        f = 10.0 / max(norm, 10.0)
        assert abs(out[0] - norm) <= 2e-6 * norm and abs(out[1] - f) <= 2e-6
        assert (f < 1) == clipped
        for a, b in zip(dev, grads):
            np.testing.assert_allclose(a.cpu().numpy(), (b * np.float32(out[1])).numpy(), rtol=1e-6, atol=0)


def test_drop_in_target_layers_and_device_anchors():
    """the reference-facing modules (model.rpn.anchor_target_layer / proposal_target_layer signatures) and the device
    anchor table (generate_anchors.py:112-173, bit-equal to the fp64 numpy table cast to fp32)"""
    from stereo_rcnn_b200.model.rpn.anchor_target_layer import _AnchorTargetLayer
    from stereo_rcnn_b200.model.rpn.proposal_target_layer import _ProposalTargetLayer
    H, W, B = 600, 1987, 2
    fs = feat_shapes(H, W)
    anchors = G.generate_anchors(fs, "cuda")
    np.testing.assert_array_equal(anchors.cpu().numpy(), O.anchors_all_pyramids(fs).astype(np.float32))
    assert anchors.shape[0] == 298476
    gl, gr, gm, dim, kp, nb = synth.synth_train_gt(B, 30, H, W, 4)
    layer = _AnchorTargetLayer(1, [0.5, 1, 2])
    layer.generator = torch.Generator(device="cuda").manual_seed(3)
    im_info = torch.tensor([[H, W, 1.6]] * B)
    out = layer((None, cu(gl), cu(gr), cu(gm), im_info, cu(nb), fs))
    lab = out[0]
    assert lab.shape == (B, 298476) and out[1].shape == (B, 298476, 4)
    n_fg, n_bg = (lab == 1).sum(1), (lab == 0).sum(1)
    assert (n_fg > 0).all() and (n_fg <= 256).all() and ((n_fg + n_bg) == 512).all()
    # everything except the sampled subset is independent of the words: compare with the oracle's pre-sampling sets
    keys0 = np.zeros((B, 298476), np.uint32)
    ref = T.anchor_target_layer(anchors.cpu().numpy(), gl, gr, gm, im_info.numpy(), T.KeySampler(keys0))
    for b in range(B):
        fg_dev, fg_ref = lab[b].cpu().numpy() == 1, ref[0][b] == 1
        assert (fg_dev == fg_ref).all() if fg_ref.sum() <= 256 else (fg_dev <= fg_ref).all()
    np.testing.assert_allclose(out[1].cpu().numpy(), ref[1], rtol=0, atol=4e-7)
    rl, rr = synth.synth_train_rois(gl, 2000, H, W, 5)
    pl = _ProposalTargetLayer(2)
    pl.generator = torch.Generator(device="cuda").manual_seed(4)
    res = pl(cu(rl), cu(rr), cu(gl), cu(gr), cu(dim), cu(kp), cu(nb))
    assert len(res) == 10 and res[0].shape == (B, 512, 5) and res[6].dtype == torch.int64
    assert pl.status.cpu().tolist() == [0, 0]
    assert ((res[2] > 0).sum(1) <= 128).all() and (res[2] > 0).sum() > 0
    assert torch.equal(res[9], (res[8] > 0).float())
