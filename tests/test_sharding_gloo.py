"""N>1 host logic on CPU: world_size-2 gloo run of the channel sharding + block broadcast."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cutesdr_b200.sharding import exchange_unique_id, shard_bounds, shard_channels


def test_shard_bounds_partition():
    for n in (1, 7, 256, 1024, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the one thing the host application does itself: hand rank 0's 128-byte communicator id to every rank
        # (the id is made by NCCL on a GPU box; any 128 bytes exercise the exchange here)
        uid = exchange_unique_id(lambda: bytes(range(128)), rank, world)
        block = torch.frombuffer(bytearray(uid), dtype=torch.uint8)
        ref = torch.arange(128, dtype=torch.uint8)
        chans = shard_channels(list(range(1000)), rank, world)
        # every rank reports its slice; rank 0 checks they tile the channel list
        gathered = [None] * world
        dist.all_gather_object(gathered, (chans[0], chans[-1], len(chans)))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bool(torch.equal(block, ref)), gathered, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, gathered, tmax in res:
        assert same, "rank %d did not receive the communicator id" % rank
        assert gathered == [(0, 499, 500), (500, 999, 500)]
        assert tmax == 2.0
