"""``align_parallel`` with the reference's signature (lib/model/dense_align/dense_align.py:240-300):
one C-ABI call (two kernel launches) instead of the per-RoI Python loop + grid_sample passes."""
from stereo_rcnn_b200 import ops as _ops


def align_parallel(calib, scale, im_left, im_right, box_left, keypoints, poses):
    """calib: object with .p2/.p3 (3x4); scale: H_im_left/H_origin (python float or 0-dim tensor);
    im_*: 1x3xHxW; box_left Dx4, keypoints Dx5, poses Dx7 -> (solve_status [D], best_dis [D])"""
    c4 = _ops.calib_vec(calib.p2, calib.p3)
    return _ops.dense_align(c4, float(scale), im_left, im_right, box_left[:, 0:4].contiguous(),
                            keypoints.contiguous(), poses[:, 0:7].contiguous())
