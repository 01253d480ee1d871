"""Channel sharding across the GPUs of one box: channels are independent given the
wideband stream, so each rank owns a contiguous slice of the channel list and the only exchange is
one broadcast of every wideband block from the rank that received it from the host. No torch import
at module level: the plan is plain Python so it can be tested anywhere."""


def shard_bounds(n_total, rank, world):
    """[lo, hi) slice of n_total channels owned by `rank`; slices differ by at most one channel."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_channels(items, rank, world):
    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def broadcast_block(block, src=0):
    """Broadcast one wideband block (a torch tensor, device or host) from `src` to every rank, in place.
    With the NCCL backend this is the NVLink/NVSwitch broadcast of the north star; with gloo it is the
    CPU stand-in used by the world_size-2 tests."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(block, src=src)
    return block
