"""Multi-GPU plumbing of the hot path: independent stereo pairs shard over ranks (one process per
GPU, torch.distributed), and the only exchange is ONE all-gather of the fixed-size per-image
detection record.  No data-path collective exists anywhere else: BN is frozen, every
stage is per-image or per-RoI.
"""
import torch

REC_ROIS = 300
REC_COLS = 2 + 8 + 8 + 10 + 5        # cls scores, boxes L/R (per class), dim/orientation, keypoints


def shard_pairs(n_pairs, rank, world):
    """contiguous split of a batch of stereo pairs; returns the [begin, end) range of this rank"""
    base, rem = divmod(n_pairs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def detection_record(cls_prob, pred_boxes_left, pred_boxes_right, dim_orien, pred_kpts):
    """[R, 33] fp32 record of one image (what test_net.py:233-330 consumes downstream)"""
    return torch.cat((cls_prob, pred_boxes_left, pred_boxes_right, dim_orien, pred_kpts), 1)


def gather_records(rec, world, dist=None, out=None):
    """all-gather of one record per rank -> [world, R, C] (identity for world == 1)"""
    if world == 1:
        return rec.unsqueeze(0)
    if out is None:
        out = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
    # dim-0 concatenation layout (accepted by both NCCL and gloo)
    dist.all_gather_into_tensor(out.view((-1,) + tuple(rec.shape[1:])), rec.contiguous())
    return out


def max_over_ranks(value_ms, device, world, dist=None):
    """device-timed duration reduced with MAX over ranks (the bench contract)"""
    if world == 1:
        return float(value_ms)
    t = torch.tensor([float(value_ms)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
