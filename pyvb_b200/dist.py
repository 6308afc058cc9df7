"""The one exchange of a sweep: all-reduce(SUM) of the packed statistics buffer over torch.distributed.
NCCL over NVLink on the GPU box; the same call runs over gloo in the CPU tests."""


def allreduce_stats(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_rows(n_total, world, rank):
    """Contiguous row block [lo, hi) of `rank`: blocks differ by at most one row."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
