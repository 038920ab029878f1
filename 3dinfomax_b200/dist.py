"""Molecule-level data parallelism (new functionality: the reference is single-GPU, train.py:236).

Molecules are independent through both encoders, so each rank gets a contiguous slice of the batch and builds its
own CSR.  Two collectives per step (NCCL over NVLink/NVSwitch on the B200 box, gloo in CPU tests):

  * ``all_gather_rows``  — the 3-D projection-head embeddings become the GLOBAL negative set of the contrastive
    loss; its backward is a reduce-scatter(sum), because every rank's loss rows touch every rank's columns;
  * one all-reduce(sum) of the flat gradient buffer inside ``FusedAdam.step`` (the loss is already normalised by
    the global batch size, so gradients are summed, not averaged).

BatchNorm statistics stay LOCAL to each rank (north_star lists only these two collectives): an R-rank step equals
the single-GPU step on the same global batch only up to per-shard BN statistics — see DESIGN.md §multi-GPU.
"""
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class _AllGatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        ctx.rows = x.shape[0]
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty((ctx.rows,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None


def all_gather_rows(x, group=None):
    """[r, ...] on every rank -> [world*r, ...] in rank order; differentiable (backward = reduce-scatter sum).
    Every rank must hold the same number of rows."""
    if not is_distributed():
        return x
    return _AllGatherRows.apply(x, group)


def shard_bounds(n_items, rank, world):
    """Contiguous, equal-size molecule slices (the last partial slice is dropped, as DistributedSampler would)."""
    per = n_items // world
    return rank * per, (rank + 1) * per
