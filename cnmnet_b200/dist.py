"""Data-parallel plumbing: frame tuples are independent, so the batch is split contiguously by
rank (the reference's nn.DataParallel scatter, eval.py:206-207) and the only collectives are one
all-gather of the outputs and one all-reduce of scalar metrics.
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world_size):
    """Contiguous [start, stop) of `total` items for `rank`; the first `total % world_size` ranks get one more."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world_size %d" % (rank, world_size))
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_outputs(local, total, group=None):
    """All-gathers per-rank output blocks [n_local, ...] (sharded with shard_range) into [total, ...]."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    pad = local
    if local.shape[0] < nmax:
        pad = torch.cat([local, local.new_zeros((nmax - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * nmax,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(b - a == nmax for a, b in sizes):
        return out
    return torch.cat([out[r * nmax: r * nmax + (b - a)] for r, (a, b) in enumerate(sizes)], 0)


def reduce_metrics(values, op="sum", group=None):
    """All-reduces a small 1-D tensor of scalars (frame counts, error sums, elapsed time)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return values
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
    out = values.clone()
    dist.all_reduce(out, op=ops[op], group=group)
    return out
