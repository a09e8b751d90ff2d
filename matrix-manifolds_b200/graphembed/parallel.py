"""Multi-GPU plumbing: one process per GPU (torchrun), embeddings replicated, the pair batch sharded across ranks,
the dense node gradient combined with one all-reduce per step (NCCL over NVLink on GPUs, gloo in the CPU tests).

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
