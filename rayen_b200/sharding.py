"""Batch sharding of the layer across the GPUs of one box (one process per GPU, ``torch.distributed``).

The ray-shooting map is sample-wise, so the data-parallel layout needs no collective in forward or backward:
every rank holds a replica of the (KB-sized) plan and a contiguous slice of the batch.  The
only exchange is optional: ``all_gather_outputs`` for a downstream loss that couples samples across the
batch (NCCL all-gather forward; the backward is the matching reduce-scatter of the incoming gradient: this rank's
rows, summed over the ranks' copies of the loss).  The reference has no distributed code (single device).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world_size):
    """[lo, hi) of this rank's contiguous slice; the first ``batch % world_size`` ranks get one more sample."""
    base, extra = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, rank=None, world_size=None):
    """This rank's slice of a replicated batch tensor (a view, no copy)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


class _AllGatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_local, sizes, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        ctx.rank, ctx.sizes, ctx.group = rank, sizes, group
        y_local = y_local.contiguous()
        if len(set(sizes)) == 1:
            out = y_local.new_empty((sum(sizes),) + tuple(y_local.shape[1:]))
            dist.all_gather_into_tensor(out, y_local, group=group)
            return out
        # ragged shards: pad every shard to the largest one (equal-size gather works on every backend)
        top = max(sizes)
        padded = y_local.new_zeros((top,) + tuple(y_local.shape[1:]))
        padded[:y_local.shape[0]] = y_local
        parts = [torch.empty_like(padded) for _ in sizes]
        dist.all_gather(parts, padded, group=group)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)

    @staticmethod
    def backward(ctx, g):
        # every rank holds the gradient w.r.t. the full gathered tensor of ITS copy of the loss; the gradient of
        # this rank's rows is the sum over ranks of those copies' slices: a reduce-scatter (1/world of the bytes an
        # all-reduce of the whole [B, ...] gradient would move).  The incoming gradient is never modified in place:
        # autograd may share it with other consumers (retain_grad, hooks).
        sizes, rank, group = ctx.sizes, ctx.rank, ctx.group
        world = len(sizes)
        lo = sum(sizes[:rank])
        if dist.get_backend(group) == "nccl":
            if len(set(sizes)) == 1:
                out = g.new_empty((sizes[rank],) + tuple(g.shape[1:]))
                dist.reduce_scatter_tensor(out, g.contiguous(), op=dist.ReduceOp.SUM, group=group)
                return out, None, None
            top = max(sizes)
            padded = g.new_zeros((world * top,) + tuple(g.shape[1:]))
            off = 0
            for r, s in enumerate(sizes):
                padded[r * top:r * top + s] = g[off:off + s]
                off += s
            out = g.new_empty((top,) + tuple(g.shape[1:]))
            dist.reduce_scatter_tensor(out, padded, op=dist.ReduceOp.SUM, group=group)
            return out[:sizes[rank]], None, None
        # backends without reduce-scatter (gloo, the CPU tests): all-reduce a private copy, keep this rank's rows
        total = g.clone(memory_format=torch.contiguous_format)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return total[lo:lo + sizes[rank]], None, None


def all_gather_outputs(y_local, batch=None, group=None):
    """[B_local, ...] on every rank -> [B, ...] on every rank, differentiable.

    ``batch`` (the global batch size) is only needed when it does not divide evenly.  If the downstream loss
    is computed redundantly on every rank and then averaged by DDP-style gradient averaging, scale it by
    ``1 / world_size`` as usual.
    """
    world = dist.get_world_size(group)
    if batch is None:
        sizes = [int(y_local.shape[0])] * world
    else:
        sizes = [shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world)]
    return _AllGatherRows.apply(y_local, sizes, group)
