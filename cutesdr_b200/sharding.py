"""Channel sharding across the GPUs of one box: channels are independent given the wideband stream, so
each rank owns a contiguous slice of the channel list and the only exchange is one broadcast of every wideband block
from the rank that received it from the host. The slicing policy lives in the C-ABI library
(cutesdr_mgpu_channel_slice) so C++ hosts and this Python mirror agree; the broadcast itself is
cutesdr_bank_process_async_bcast (NCCL, inside the library). What the host application has to do itself is hand rank
0's 128-byte communicator id to the other processes -- `exchange_unique_id` does that over torch.distributed (any
backend: gloo on CPU, nccl on GPUs); an MPI_Bcast or a socket would do as well."""
from .dsp import MultiGpu, channel_slice


def shard_bounds(n_total, rank, world):
    """[lo, hi) slice of n_total channels owned by `rank`; slices differ by at most one channel."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    first, count = channel_slice(n_total, rank, world)
    return first, first + count


def shard_channels(items, rank, world):
    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def exchange_unique_id(make_id, rank, world, device=None):
    """Rank 0 calls make_id() (e.g. MultiGpu.unique_id) and every rank receives the 128 bytes."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return make_id()
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def open_multi_gpu(rank, world, device):
    """This rank's MultiGpu handle; the id travels over the already initialised torch.distributed group."""
    import torch
    dev = torch.device("cuda", device) if world > 1 and torch.distributed.get_backend() == "nccl" else None
    uid = exchange_unique_id(MultiGpu.unique_id, rank, world, dev) if world > 1 else None
    return MultiGpu(uid, rank, world, device)
