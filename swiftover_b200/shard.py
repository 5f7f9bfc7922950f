"""Host-side sharding helpers shared by bench.py (NCCL, one rank per GPU) and the CPU tests
(gloo).  The liftover path needs no data-path collective: ranks own
contiguous byte ranges of the input, the chain table is replicated, and outputs are
concatenated in rank order.  torch.distributed is used for the rendezvous, the barrier
around the timed region and the max-over-ranks of the timings only.
"""
import os


def shard_bounds(data, rank, world):
    """Byte range [a, b) of `data` owned by `rank` of `world`: the lines that START inside the
    rank's equal-size byte range (same rule as swo_lift_bed_text's device sharder, swo_api.cu)."""
    n = len(data)

    def cut(k):
        if k <= 0:
            return 0
        if k >= world:
            return n
        e = n * k // world
        while 0 < e < n and data[e - 1] != 0x0A:
            e += 1
        return e

    a, b = cut(rank), cut(rank + 1)
    return a, max(a, b)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend, device_id=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_* (no-op for world 1)."""
    import torch.distributed as dist
    rank, local, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device_id} if device_id is not None else {}
        dist.init_process_group(backend, **kw)
    return rank, local, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (timings are reported as the
    slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_in_rank_order(obj):
    """All ranks' objects as a list ordered by rank (ordered concat of shard outputs)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def piece_offsets(lengths, device="cpu"):
    """Ordered concatenation without a copy: every rank contributes the byte lengths of its output
    pieces (e.g. [lifted, unmatched]); returned are this rank's offsets in the concatenated outputs
    (the sum of the lower ranks' lengths) and the totals.  One small all_gather: a rank then writes
    its piece at its offset (pwrite / its slot of a shared segment) and the result is the
    single-process output, byte for byte (outputs in shard order)."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(lengths), dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return [0] * len(lengths), mine.tolist()
    world, rank = dist.get_world_size(), dist.get_rank()
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    stack = torch.stack(every)
    return stack[:rank].sum(0).tolist(), stack.sum(0).tolist()
