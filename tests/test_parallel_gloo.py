"""world_size-2 gloo tests (CPU) of the data-parallel host logic: bucketed gradient all-reduce in backward order,
1/world averaging folded into the optimizer, volume sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from medicalseg_b200.parallel import DistributedGradReducer
    torch.manual_seed(rank)
    total = 1000
    flat = torch.randn(total)
    mine = flat.clone()
    red = DistributedGradReducer(flat, bucket_mb=300 * 4 / (1 << 20))  # 300-element buckets
    # ranges arrive from the end of the flat buffer, as VNet._backward fires them
    for lo, hi in ((900, 1000), (650, 900), (640, 650), (300, 640), (0, 300)):
        red.on_ready(lo, hi)
    launched = list(red.launched)
    red.wait()
    gathered = [torch.zeros(total) for _ in range(world)]
    dist.all_gather(gathered, mine)
    expect = sum(gathered)
    ok = torch.allclose(flat, expect, atol=1e-6)
    # volume sharding: every rank takes a disjoint slice of the global batch
    vols = list(range(8))
    shard = vols[rank::world]
    all_shards = [None] * world
    dist.all_gather_object(all_shards, shard)
    q.put((rank, ok, launched, red.grad_scale, sorted(sum(all_shards, []))))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, launched, scale, vols in res:
        assert ok
        assert launched == [(650, 1000), (300, 650), (0, 300)]
        assert scale == 0.5
        assert vols == list(range(8))


def _eval_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from medicalseg_b200.core import reduce_eval_metrics
    # 3 validation volumes over 2 ranks (core.evaluate shards indices[rank::world]): rank 0 scores volumes 0 and 2,
    # rank 1 volume 1; per-volume (mean Dice, loss, per-class Dice) are made up but distinct
    per_volume = {0: (0.8, 1.0, [0.9, 0.7]), 1: (0.6, 2.0, [0.5, 0.7]), 2: (0.4, 3.0, [0.3, 0.5])}
    mine = list(range(3))[rank::world]
    md = sum(per_volume[i][0] for i in mine)
    ls = sum(per_volume[i][1] for i in mine)
    cd = sum(np.asarray(per_volume[i][2]) for i in mine)
    out = reduce_eval_metrics(md, ls, cd, len(mine), torch.device("cpu"), num_classes=2)
    # a rank with an empty shard (1 volume, 2 ranks) must still take part in the collective
    mine1 = [0][rank::world]
    out1 = reduce_eval_metrics(0.8 if mine1 else 0.0, 1.0 if mine1 else 0.0, np.asarray([0.9, 0.7]) if mine1 else None,
                               len(mine1), torch.device("cpu"), num_classes=2)
    q.put((rank, out[0], out[1], out[2].tolist(), out1[0], out1[2].tolist()))
    dist.destroy_process_group()


def test_eval_metrics_are_global_means_world2():
    """core.evaluate: every rank reports the mean over ALL validation volumes (the reference reports per-rank shards,
    core/val.py:168), including when one rank's shard is empty"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, md, ls, cd, md1, cd1 in res:
        assert md == pytest.approx((0.8 + 0.6 + 0.4) / 3) and ls == pytest.approx(2.0)
        assert cd == pytest.approx([(0.9 + 0.5 + 0.3) / 3, (0.7 + 0.7 + 0.5) / 3])
        assert md1 == pytest.approx(0.8) and cd1 == pytest.approx([0.9, 0.7])
