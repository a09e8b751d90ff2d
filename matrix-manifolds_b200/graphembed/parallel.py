"""Multi-GPU plumbing: one process per GPU (torchrun), the pair batch sharded across ranks (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Two ways to combine the per-rank dense node gradients:

* replicated update (`allreduce_step_buffers`): one all-reduce of the (N, ...) gradient, every rank applies the same
  optimizer update to its full replica;
* owner update (`RowShards`): reduce-scatter of the gradient so that rank r receives the summed rows [lo_r, hi_r) it
  owns, the optimizer (and its moment buffers) only ever touch those rows, then an all-gather of the updated rows.
  Same bytes on the wire as the all-reduce, but the optimizer kernel and its state shrink by the world size.

The reference's only multi-GPU mechanism is nn.DataParallel over node chunks (train.py:107-109,203-204), which
silently drops pairs that straddle two chunks; sharding the *pair list* keeps every pair."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) slice of n_items for `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_step_buffers(grad, acc, group=None):
    """Sum the dense (N, ...) gradient and the [loss, scale-grad] accumulator over the ranks of `group`."""
    if group is None or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(grad, group=group)
    dist.all_reduce(acc, group=group)


class RowShards:
    """Row ownership of an (N, ...) parameter over the ranks of `group` (equal shards; N must divide evenly)."""

    def __init__(self, n_rows, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if n_rows % self.world != 0:
            raise ValueError(f'{n_rows} rows do not split evenly over {self.world} ranks')
        self.rows = n_rows // self.world
        self.lo, self.hi = self.rank * self.rows, (self.rank + 1) * self.rows
        self._native_rs = dist.get_backend(group) == 'nccl'

    def own(self, t):
        """View of the rows of `t` this rank owns."""
        return t[self.lo:self.hi]

    def reduce_scatter(self, grad_full, out=None):
        """Sum `grad_full` over the ranks; returns the summed rows this rank owns (written to `out` if given)."""
        if out is None:
            out = torch.empty_like(self.own(grad_full))
        if self._native_rs:
            dist.reduce_scatter_tensor(out, grad_full, group=self.group)
        else:  # gloo has no reduce-scatter: all-reduce, keep the owned rows
            dist.all_reduce(grad_full, group=self.group)
            out.copy_(self.own(grad_full))
        return out

    def all_gather(self, full):
        """Publish this rank's rows of `full` to every rank (in place: the send buffer is the owned slice)."""
        dist.all_gather_into_tensor(full, self.own(full), group=self.group)
        return full
