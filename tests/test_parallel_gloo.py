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
