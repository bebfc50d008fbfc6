"""CPU: the oracle restatement against golden vectors minted from the REFERENCE's own
Python (tests/golden/make_golden.py), plus internal known-answer checks."""
import os

import numpy as np
import pytest
import torch

from oracle import model, ops
from stereo_rcnn_b200.synth import synth_pair


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_anchors_match_reference(golden_dir):
    g = _load(golden_dir, "anchors_decode.npz")
    a = ops.anchors_all_pyramids(g["shapes"].tolist())
    assert a.dtype == np.float64 and a.shape == (int(g["n_anchors"]), 4) == (298476, 4)
    np.testing.assert_array_equal(a[g["sel"]], g["anchors_sel"])          # bit-exact fp64
    np.testing.assert_allclose(a.sum(0), g["anchors_sum"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(np.abs(a).sum(0), g["anchors_abs_sum"], rtol=1e-12)


def test_decode_clip_matches_reference(golden_dir):
    g = _load(golden_dir, "anchors_decode.npz")
    out = ops.decode_clip(g["anchors_sel"].astype(np.float32), g["deltas"], 600, 1987)
    # identical op order; only exp differs (sb_expf vs torch.exp, <= 1 ulp each)
    np.testing.assert_allclose(out, g["decoded"], rtol=3e-7, atol=2e-4)
    assert (out == g["decoded"]).mean() > 0.95


def test_sb_expf_accuracy():
    x = np.linspace(-30, 30, 200001).astype(np.float32)
    y = ops.sb_expf(x).astype(np.float64)
    t = np.exp(x.astype(np.float64))
    assert np.max(np.abs(y - t) / t) < 1.0 * 2.0 ** -23
    e = ops.sb_expf(np.array([0.0, 100.0, -200.0], np.float32))
    assert e[0] == 1.0 and np.isinf(e[1]) and e[2] == 0.0


def test_proposal_layer_matches_reference(golden_dir):
    g = _load(golden_dir, "proposal_small.npz")
    rl, rr = ops.proposal_layer(g["cls_prob"], g["bbox_pred"], g["im_info"], "TEST", g["shapes"].tolist())
    assert rl.shape == g["rois_left"].shape == (1, 300, 5)
    # the reference leaves the order of tied scores to torch.sort; the oracle pins (score desc,
    # index asc).  Rows must agree except where a tie group straddles the comparison.
    same = np.all(np.abs(rl - g["rois_left"]) < 1e-3, axis=2) & np.all(np.abs(rr - g["rois_right"]) < 1e-3, axis=2)
    assert same.mean() > 0.97, same.mean()
    set_ref = {tuple(np.round(r, 2)) for r in g["rois_left"][0]}
    set_ora = {tuple(np.round(r, 2)) for r in rl[0]}
    assert len(set_ref & set_ora) >= 0.97 * len(set_ref)


def test_proposal_layer_properties():
    rs = np.random.RandomState(0)
    shapes = [[20, 32], [10, 16], [5, 8], [3, 4], [2, 2]]
    A = 3 * sum(h * w for h, w in shapes)
    prob = rs.rand(2, A, 2).astype(np.float32)
    bbox = (rs.randn(2, A, 6) * 0.2).astype(np.float32)
    info = np.array([[80, 128, 1.0], [80, 128, 1.0]], np.float32)
    rl, rr, dbg = ops.proposal_layer(prob, bbox, info, "TEST", shapes, return_debug=True)
    for b in range(2):
        k = dbg[b]["keep"]
        assert np.all(np.diff(k) > 0)                       # ascending = descending score
        assert np.all(rl[b, :, 0] == b) and np.all(rr[b, :, 0] == b)
        n = min(k.size, 300)
        assert np.all(rl[b, n:, 1:] == 0)                    # zero padding (Q11)
        # left/right share y (Q8)
        np.testing.assert_array_equal(rl[b, :n, 2], rr[b, :n, 2])
        np.testing.assert_array_equal(rl[b, :n, 4], rr[b, :n, 4])
        assert rl[b, :n, 1].min() >= 0 and rl[b, :n, 3].max() <= 127


def test_nms_known_answers():
    # two identical boxes, one disjoint, one overlapping at exactly IoU = thresh (strict >)
    d = np.array([[0, 0, 9, 9, .9], [0, 0, 9, 9, .8], [20, 20, 29, 29, .7], [0, 0, 9, 19, .6]], np.float32)
    assert ops.nms(d, 0.5).tolist() == [0, 2, 3]        # IoU(0,3) = 100/200 = 0.5 -> kept
    assert ops.nms(d, 0.49).tolist() == [0, 2]
    assert ops.nms(np.zeros((0, 5), np.float32), 0.5).size == 0
    rs = np.random.RandomState(1)
    xy = rs.rand(500, 2) * 100
    wh = rs.rand(500, 2) * 40 + 2
    dets = np.concatenate([xy, xy + wh, np.sort(rs.rand(500, 1), 0)[::-1]], 1).astype(np.float32)
    keep = ops.nms(dets, 0.7)
    m = ops.nms_mask(dets, 0.7)
    # greedy reduction over the bitmask (nms_cuda_kernel.cu:132-144) gives the same list
    remv = np.zeros(m.shape[1], np.uint64)
    k2 = []
    for i in range(500):
        if not (int(remv[i // 64]) >> (i % 64)) & 1:
            k2.append(i)
            remv |= m[i]
    assert keep.tolist() == k2


def test_roi_align_known_answers():
    # a linear ramp is reproduced exactly by bilinear taps inside the map
    H, W = 12, 16
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    feat = (2.0 * xx + 3.0 * yy).astype(np.float32)[None, None]
    rois = np.array([[0, 2, 1, 9, 8]], np.float32)
    out = ops.roi_align_forward(feat, rois, 8, 8, 1.0)
    bw = (9 - 2 + 1) / 7.0
    bh = (8 - 1 + 1) / 7.0
    exp = 2.0 * (2 + np.arange(8) * bw)[None, :] + 3.0 * (1 + np.arange(8) * bh)[:, None]
    np.testing.assert_allclose(out[0, 0], exp, rtol=1e-6)
    # taps with h >= H or w >= W are zero; [H-1,H) extrapolates (Q16)
    rois = np.array([[0, 10, 6, 15, 11]], np.float32)
    out = ops.roi_align_forward(feat, rois, 8, 8, 1.0)
    assert out[0, 0, -1, 0] == 0.0 or (6 + 7 * (6 / 7.0)) < H
    avg = ops.roi_align_avg(feat, np.array([[0, 2, 1, 9, 8]], np.float32), 7, 7, 1.0)
    assert avg.shape == (1, 1, 7, 7)


def test_roi_align_backward_is_adjoint():
    rs = np.random.RandomState(2)
    feat = rs.randn(1, 3, 10, 14).astype(np.float32)
    rois = np.array([[0, 1.3, 2.2, 10.7, 8.1], [0, 0, 0, 13, 9]], np.float32)
    top = rs.randn(2, 3, 5, 5).astype(np.float32)
    out = ops.roi_align_forward(feat, rois, 5, 5, 0.9)
    grad = ops.roi_align_backward(top, rois, feat.shape, 5, 5, 0.9)
    np.testing.assert_allclose((out * top).sum(), (grad * feat).sum(), rtol=1e-4)


def test_dense_align_matches_reference(golden_dir):
    g = _load(golden_dir, "dense_align.npz")
    left, right = synth_pair(int(g["H"]), int(g["W"]), int(g["seed"]), int(g["shift"]))
    calib = ops.calib_vec(g["p2"], g["p3"])
    st, dis, dg = ops.dense_align(calib, float(g["scale"]), left, right, g["box_left"], g["keypoints"],
                                  g["poses"], diagnostics=True)
    np.testing.assert_array_equal(st, g["status"])
    np.testing.assert_array_equal(dg["npix"], g["npix"])
    np.testing.assert_allclose(dis, g["best_dis"], rtol=1e-6)
    uvz, npix = ops.dense_sample(calib, float(g["scale"]), int(g["H"]), int(g["W"]), g["box_left"],
                                 g["keypoints"], g["poses"], maxp=64)
    for i in range(uvz.shape[0]):
        n = min(int(npix[i]), 64)
        np.testing.assert_allclose(uvz[i, :n], g["uvz_head"][i, :n], rtol=1e-6, atol=1e-6)
    # sanity (not parity): the synthetic pair is a pure 40 px shift at network scale = 25 px at
    # original scale; a box surface is not fronto-parallel (per-pixel dz), so RoIs whose search
    # window contains the truth land near it, not on it (+0.5 offset of dense_align.py:298)
    true_dis = 40 / 1.6 + 0.5
    z = g["poses"][:, 2]
    fb = 721.5377 * 0.53272  # f * baseline
    ok = np.abs(fb / z - 25.0) < 6.0   # only RoIs whose search window contains the truth
    if ok.any():
        assert np.median(np.abs(dis[ok] - true_dis)) < 4.0


def test_dense_align_edge_cases():
    left, right = synth_pair(120, 200, 5, 4)
    calib = ops.calib_vec(np.array([[100., 0, 100, 4], [0, 100., 60, 0], [0, 0, 1, 0]]),
                          np.array([[100., 0, 100, -50], [0, 100., 60, 0], [0, 0, 1, 0]]))
    # box far outside the 3D box's projection -> no valid pixel -> early-return branch
    b = np.array([[5, 5, 30, 30]], np.float32)
    k = np.array([[17, 1, .9, 5, 30]], np.float32)
    p = np.array([[30., 1.6, 20., 1.6, 1.5, 3.9, 0.3]], np.float32)
    st, dis = ops.dense_align(calib, 1.0, left, right, b, k, p)
    assert st[0] == 0.0
    fb32 = np.float32(np.float64(100.0 * 2) * (np.float64(54.0) * 2 / (100.0 * 2)))
    assert dis[0] == fb32 / np.float32(20.0)            # dis_init (dense_align.py:272-273)
    st, dis = ops.dense_align(calib, 1.0, left, right, b[:0], k[:0], p[:0])
    assert st.size == 0 and dis.size == 0


def test_roi_levels():
    rois = np.array([[0, 0, 0, 223, 223], [0, 0, 0, 20, 20], [0, 0, 0, 1000, 600], [0, 0, 0, 0, 0]], np.float32)
    assert ops.roi_levels(rois).tolist() == [4, 2, 5, 2]      # natural log, clamp [2,5] (Q14)


@pytest.mark.slow
def test_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "forward_small.npz")
    left, right = synth_pair(int(g["H"]), int(g["W"]), int(g["seed"]), int(g["shift"]))
    sd = model.make_state_dict(int(g["weight_seed"]))
    info = torch.tensor([[float(g["H"]), float(g["W"]), 1.0]])
    o = model.forward(sd, torch.from_numpy(left)[None], torch.from_numpy(right)[None], info)
    np.testing.assert_allclose(o["rois_left"].numpy(), g["rois_left"], rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(o["rois_right"].numpy(), g["rois_right"], rtol=1e-6, atol=1e-4)
    for n in ("cls_prob", "bbox_pred", "dim_orien_pred", "kpts_prob", "left_border_prob", "right_border_prob"):
        a = o[n].numpy().reshape(g[n].shape)
        np.testing.assert_allclose(a, g[n], rtol=1e-4, atol=1e-5, err_msg=n)


def test_proposal_layer_train_config_matches_reference(golden_dir):
    """cfg_key "TRAIN" (12000 pre-NMS / 2000 post-NMS, IoU 0.7) through the reference's own _ProposalLayer"""
    g = _load(golden_dir, "proposal_train.npz")
    rl, rr = ops.proposal_layer(g["cls_prob"], g["bbox_pred"], g["im_info"], "TRAIN", g["shapes"].tolist())
    assert rl.shape == g["rois_left"].shape and rr.shape == g["rois_right"].shape
    assert rl.shape[1] == 2000
    same = np.all(np.abs(rl - g["rois_left"]) < 1e-3, axis=2) & np.all(np.abs(rr - g["rois_right"]) < 1e-3, axis=2)
    assert same.mean() > 0.97, same.mean()
    nz = np.abs(g["rois_left"][0, :, 1:]).sum(1) > 0             # zero-padded tail (fewer than 2000 survivors)
    assert nz.sum() == (np.abs(rl[0, :, 1:]).sum(1) > 0).sum()


def test_test_time_decode_matches_reference_script_lines(golden_dir):
    """A12: the golden was produced by executing test_net.py's own decode lines (ref_lines) on synthetic head
    outputs; the oracle's restatement must reproduce them (float division kpts_type quirk included)"""
    g = _load(golden_dir, "test_decode.npz")
    sc, pbl, pbr, do, pk = ops.test_decode(g["rois_left"][0], g["rois_right"][0], g["cls_prob"][0], g["bbox_pred"][0],
                                           g["bbox_pred_dim"][0], g["kpts_prob"], g["left_prob"], g["right_prob"],
                                           g["im_info"][0])
    np.testing.assert_array_equal(sc, g["scores"])
    for a, b, name in ((pbl, g["pred_boxes_left"], "boxes_left"), (pbr, g["pred_boxes_right"], "boxes_right"),
                       (do, g["dim_orien"], "dim_orien"), (pk, g["pred_kpts"], "kpts")):
        assert a.shape == b.shape, name
        assert np.abs(a - b).max() <= 1e-4 * max(1.0, np.abs(b).max()), (name, np.abs(a - b).max())
    # integer-valued columns are exact: keypoint type = argmax / 28 as a float, and the argmax probability
    np.testing.assert_array_equal(pk[:, 1], g["pred_kpts"][:, 1])
    np.testing.assert_array_equal(pk[:, 2], g["pred_kpts"][:, 2])


def test_per_class_nms_matches_reference_script_lines(golden_dir):
    """A13: test_net.py:234-259 executed on the decode golden (threshold, descending sort, NMS 0.3, gather)"""
    g = _load(golden_dir, "class_nms.npz")
    keep = ops.per_class_nms(g["scores"], g["pred_boxes_left"], int(g["cls"]), float(g["score_thresh"]),
                             float(g["nms_thresh"]))
    np.testing.assert_array_equal(np.asarray(keep, np.int64), g["kept_rois"])
    j = int(g["cls"])
    np.testing.assert_array_equal(g["pred_boxes_left"][keep][:, 4 * j:4 * j + 4], g["cls_dets_left"][:, :4])


def test_prep_image_matches_cv2_golden(golden_dir):
    """8f-3: the restated prep_im_for_blob (blob.py:44-64) against blobs made by the reference's recipe with the
    cv2 of the build image (make_golden.py (8)), at the demo scale and at a non-dyadic KITTI scale"""
    g = _load(golden_dir, "prep_image.npz")
    for tag in ("a", "b"):
        got = ops.prep_image(g["crop_bgr"], float(g["scale_" + tag]))
        assert got.shape == g["blob_" + tag].shape
        np.testing.assert_allclose(got, g["blob_" + tag], rtol=0, atol=6.2e-5)      # <= 2 ulp at |x| <= 256


@pytest.mark.slow
def test_forward_on_reference_demo_crop_matches_reference(golden_dir):
    """the oracle on a 192x320 crop of the reference's demo/left.png|right.png (blob by the reference's recipe)
    against the reference's own forward on the same crop"""
    g = _load(golden_dir, "forward_demo.npz")
    sc = float(g["scale"])
    bl, br = ops.prep_image(g["crop_left_bgr"], sc), ops.prep_image(g["crop_right_bgr"], sc)
    np.testing.assert_allclose([bl.sum(dtype=np.float64), br.sum(dtype=np.float64)], g["blob_checksum"], rtol=1e-6)
    sd = model.make_state_dict(int(g["weight_seed"]))
    info = torch.tensor([[float(bl.shape[1]), float(bl.shape[2]), sc]])
    o = model.forward(sd, torch.from_numpy(bl)[None], torch.from_numpy(br)[None], info)
    same = (np.abs(o["rois_left"].numpy() - g["rois_left"]) < 1e-2).all(2) & \
           (np.abs(o["rois_right"].numpy() - g["rois_right"]) < 1e-2).all(2)
    assert same.mean() > 0.97, same.mean()          # the blobs differ by <= 2 ulp: a rare proposal may swap
    if same.all():
        for n in ("cls_prob", "bbox_pred", "dim_orien_pred", "kpts_prob", "left_border_prob", "right_border_prob"):
            np.testing.assert_allclose(o[n].numpy().reshape(g[n].shape), g[n], rtol=1e-3, atol=1e-4, err_msg=n)
