"""CPU, world_size 2, gloo: the N>1 host path of bench.py (pair sharding, the single all-gather of
detection records, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereo_rcnn_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = P.shard_pairs(5, rank, world)
        g = torch.Generator().manual_seed(100 + rank)
        parts = [torch.rand(P.REC_ROIS, w, generator=g) for w in (2, 8, 8, 10, 5)]
        rec = P.detection_record(*parts) + rank
        allrec = P.gather_records(rec, world, dist)
        # the class bench.py uses (on CPU tensors its exchange is torch.distributed's collective, here over gloo):
        # two slots, several steps, every rank must see every rank's record of the same (step, slot)
        gat = P.RecordGather(world, rank, torch.device("cpu"), dist, n_slots=2)
        assert gat.mode == "nccl" and "ncclAllGather" in gat.describe()
        for step in range(3):
            for slot in range(2):
                mine = torch.full((P.REC_ROIS, P.REC_COLS), float(100 * step + 10 * slot + rank))
                got = gat(slot, mine)
                for r in range(world):
                    assert torch.equal(got[r], torch.full((P.REC_ROIS, P.REC_COLS), float(100 * step + 10 * slot + r)))
        t = P.max_over_ranks(10.0 + rank, torch.device("cpu"), world, dist)
        q.put((rank, (b, e), allrec.shape, float(allrec[0].mean()), float(allrec[1].mean()), float(rec.mean()), t))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_shard_gather_and_timing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, shp0, a0, a1, m0, t0), (r1, s1, shp1, b0, b1, m1, t1) = res
    assert s0 == (0, 3) and s1 == (3, 5)                        # contiguous, covering, remainder first
    assert tuple(shp0) == tuple(shp1) == (2, P.REC_ROIS, P.REC_COLS)
    assert abs(a0 - m0) < 1e-6 and abs(a1 - m1) < 1e-6          # every rank sees rank 0's and rank 1's record
    assert abs(b0 - m0) < 1e-6 and abs(b1 - m1) < 1e-6
    assert t0 == t1 == 11.0                                     # MAX over ranks


def test_shard_pairs_covers_everything():
    for n in (1, 7, 8, 16, 17):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                b, e = P.shard_pairs(n, r, w)
                seen += list(range(b, e))
            assert seen == list(range(n))


def test_world1_is_identity():
    rec = torch.rand(P.REC_ROIS, P.REC_COLS)
    assert torch.equal(P.gather_records(rec, 1)[0], rec)
    assert P.max_over_ranks(3.5, torch.device("cpu"), 1) == 3.5


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = [(64, 3, 7, 7), (64,), (5, 3), (1000, 33), (7,)]
        gb = P.GradientBuckets(shapes, "cpu", world, dist, bucket_bytes=40000)     # forces several buckets
        assert [tuple(g.shape) for g in gb.grads] == shapes and len(gb.buckets) >= 3
        assert all(g.data_ptr() % 16 == 0 for g in gb.grads)
        for i, g in enumerate(gb.grads):
            g.fill_(float(i + 1) * (rank + 1))                       # rank 0: i+1, rank 1: 2(i+1)
        gb.all_reduce()
        ok = all(torch.equal(g, torch.full_like(g, 1.5 * (i + 1))) for i, g in enumerate(gb.grads))
        norm = gb.clip(10.0)
        total = float(torch.sqrt(sum((torch.full(s, 1.5 * (i + 1)).double() ** 2).sum() for i, s in enumerate(shapes))))
        ok = ok and abs(float(norm[0]) - total) < 1e-3 * total and abs(float(norm[1]) - 10.0 / total) < 1e-6
        ok = ok and abs(float(gb.grads[0].flatten()[0]) - 1.5 * 10.0 / total) < 1e-6
        losses = P.average_losses(torch.tensor([1.0, 2.0]) * (rank + 1), world, dist)
        q.put((rank, ok, losses.tolist(), gb.describe()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_gradient_buckets():
    """the training step's exchange: gradients as views into flat buckets, one all-reduce per bucket, averaged; global
    norm clipping over the same views; loss averaging"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok, losses, desc in res:
        assert ok, (rank, desc)
        assert losses == [1.5, 3.0]
        assert "ncclAllReduce per bucket" in desc


def test_gradient_buckets_single_rank_identity():
    gb = P.GradientBuckets([(3, 4), (5,)], "cpu")
    gb.grads[0].fill_(2.0)
    gb.all_reduce()
    assert float(gb.grads[0].sum()) == 24.0 and "single rank" in gb.describe()
    gb.zero_()
    assert float(gb.buckets[0].abs().sum()) == 0.0
