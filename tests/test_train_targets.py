"""CPU: the restated train-time target layers / losses (oracle/train_targets.py) against fixtures minted
from the reference's own `_AnchorTargetLayer`, `_ProposalTargetLayer` and `_smooth_l1_loss` (make_golden.py (11)),
with numpy's stream seeded like trainval_net.py -- labels, sampled indices and weights bit-exact, the log-space
regression targets to the last ulp of `log`; plus the device sampler's definition (KeySampler)."""
import os

import numpy as np
import pytest
import torch

from oracle import ops
from oracle import train_targets as T


def _load(golden_dir):
    return np.load(os.path.join(golden_dir, "train_targets.npz"))


@pytest.mark.parametrize("case", ["a", "b"])
def test_anchor_target_layer_matches_reference(golden_dir, case):
    g = _load(golden_dir)
    anchors = ops.anchors_all_pyramids(g["feat_shapes"].tolist()).astype(np.float32)
    cfg = dict(T.CFG, RPN_BATCHSIZE=int(g["at_%s_rpn_batchsize" % case]))
    np.random.seed(3)
    lab, tl, tr, iw, ow = T.anchor_target_layer(anchors, g["at_%s_gt_left" % case], g["at_%s_gt_right" % case],
                                                g["at_%s_gt_merge" % case], g["at_%s_im_info" % case],
                                                T.NumpySampler(), cfg)
    np.testing.assert_array_equal(lab, g["at_%s_labels" % case])
    np.testing.assert_array_equal(iw, g["at_%s_inside_w" % case])
    np.testing.assert_array_equal(ow, g["at_%s_outside_w" % case])
    np.testing.assert_allclose(tl, g["at_%s_targets_left" % case], rtol=0, atol=2.4e-7)
    np.testing.assert_allclose(tr, g["at_%s_targets_right" % case], rtol=0, atol=2.4e-7)
    n_fg = (lab == 1).sum(1)
    assert (n_fg <= cfg["RPN_BATCHSIZE"] // 2).all() and ((lab >= 0).sum(1) <= cfg["RPN_BATCHSIZE"]).all()
    if case == "b":
        assert n_fg.max() == cfg["RPN_BATCHSIZE"] // 2       # the fixture exercises the foreground subsampling


@pytest.mark.parametrize("case", ["a", "b"])
def test_proposal_target_layer_matches_reference(golden_dir, case):
    g = _load(golden_dir)
    np.random.seed(3)
    o = T.proposal_target_layer(g["pt_%s_in_rois_left" % case], g["pt_%s_in_rois_right" % case],
                                g["pt_%s_gt_left" % case], g["pt_%s_gt_right" % case],
                                g["pt_%s_gt_dim_orien" % case], g["pt_%s_gt_kpts" % case], T.NumpySampler())
    for n in ("rois_left", "rois_right", "labels", "dim_orien_targets", "kpts_targets", "kpts_weight", "inside_w",
              "outside_w"):
        np.testing.assert_array_equal(o[n], g["pt_%s_%s" % (case, n)], err_msg=n)
    for n in ("bbox_targets_left", "bbox_targets_right"):
        np.testing.assert_allclose(o[n], g["pt_%s_%s" % (case, n)], rtol=0, atol=2e-6, err_msg=n)
    assert o["kpts_weight"].sum() > 0 and (o["labels"] > 0).sum() > 0
    if case == "b":
        assert ((o["labels"] > 0).sum(1) == 128).all()        # more candidates than FG_FRACTION * BATCH_SIZE


def test_smooth_l1_matches_reference(golden_dir):
    g = _load(golden_dir)
    t = torch.from_numpy
    v = T.smooth_l1_loss(t(g["sl1_pred"]), t(g["sl1_target"]), t(g["sl1_inside"]), t(g["sl1_outside"]), sigma=3)
    assert abs(float(v) - float(g["sl1_sigma3"])) <= 1e-6 * abs(float(g["sl1_sigma3"]))
    v = T.smooth_l1_loss(t(g["sl1_pred"][0]), t(g["sl1_target"][0]))
    assert abs(float(v) - float(g["sl1_plain"])) <= 1e-6 * abs(float(g["sl1_plain"]))


def test_key_sampler_is_a_uniform_subset_sampler():
    """the device sampler: candidates ordered by (key, index); draws are words / 2^32 -- same call sites as numpy's"""
    rng = np.random.RandomState(0)
    keys = rng.randint(0, 2 ** 32, (1, 50), dtype=np.uint64).astype(np.uint32)
    keys[0, 7] = keys[0, 3]                                   # a tie resolves by position
    s = T.KeySampler(keys, rng.randint(0, 2 ** 32, (1, 16), dtype=np.uint64).astype(np.uint32))
    cand = np.array([3, 7, 11, 20, 41])
    perm = s.permutation(0, cand)
    assert sorted(perm.tolist()) == list(range(5))
    k = keys[0][cand][perm]
    assert (np.diff(k.astype(np.int64)) >= 0).all() and perm.tolist().index(0) < perm.tolist().index(1)
    d = s.draws(0, 16)
    assert d.dtype == np.float64 and (d >= 0).all() and (d < 1).all()
    n = 2030
    np.testing.assert_array_equal(np.floor(d * n).astype(np.int64),
                                  (s.words[0].astype(np.uint64) * np.uint64(n)) >> np.uint64(32))
