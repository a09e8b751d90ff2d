"""Multi-GPU plumbing: one process per GPU (torchrun), the pair batch sharded across ranks (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Two ways to combine the per-rank dense node gradients:

* replicated update (`allreduce_step_buffers`): one all-reduce of the (N, ...) gradient, every rank applies the same
  optimizer update to its full replica;
* owner update (`RowShards`): reduce-scatter of the gradient so that rank r receives the summed rows [lo_r, hi_r) it
  owns, the optimizer (and its moment buffers) only ever touch those rows, then an all-gather of the updated rows.
  Same bytes on the wire as the all-reduce, but the optimizer kernel and its state shrink by the world size.

The reference's only multi-GPU mechanism is nn.DataParallel over node chunks (train.py:107-109,203-204), which
silently drops pairs that straddle two chunks; sharding the *pair list* keeps every pair."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib as L


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


# ---------------------------------------------------------------------------------------------------------------------
# NVLink peer memory: the owner update as ONE kernel (gm_optim_step_peer) instead of three collectives
# ---------------------------------------------------------------------------------------------------------------------
class _RawCuda:
    """A device allocation that is not torch's, exposed through __cuda_array_interface__ so torch can view it."""

    def __init__(self, address, nbytes):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (address, False), 'version': 2}


def _align(n, a=256):
    return (n + a - 1) // a * a


class PeerArena:
    """Per-rank device arena [points | partial gradients | step accumulator | accumulator sum | flag block] that every
    other rank of the process group maps through CUDA IPC, plus the table of mapped peer arenas that
    gm_optim_step_peer (include/gm_kernels.h) walks: the owner of a row block pulls the partial gradient rows from all
    ranks over NVLink, applies the optimizer update and pushes the new rows into every rank's point table.

    x : the (N, ...) CUDA tensor holding the points (its values are copied into the arena); N must divide evenly.
    """

    _need_even = True  # the (N, ...) table is cut into `world` equal row blocks

    def __init__(self, x, n_acc, group):
        """Runs all three construction phases back to back (every rank must succeed).  `try_peer_arena` drives the
        phases one at a time with a status vote in between, so that a failure on ONE rank makes ALL ranks fall back
        through the same sequence of collectives."""
        self._init_fields(x, n_acc, group)
        self.phase_alloc_export(x)
        self.phase_open(self.exchange_handles())
        self.phase_finish()

    def _init_fields(self, x, n_acc, group):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.base, self._opened, self.peer_base, self._handle = None, [], [], None
        self.rows = x.shape[0] // max(self.world, 1)
        self.lo, self.hi = self.rank * self.rows, (self.rank + 1) * self.rows
        self.n_acc = n_acc
        self.dev = x.device
        xb = _align(x.numel() * x.element_size())
        ab = _align(8 * n_acc)
        self.off_x, self.off_g, self.off_acc, self.off_out, self.off_flags = 0, xb, 2 * xb, 2 * xb + ab, 2 * xb + 2 * ab
        self.nbytes = self.off_flags + _align(L.GM_PEER_FLAG_BYTES)

    def phase_alloc_export(self, x):
        """Phase 1 (local, no collective): allocate the arena, export its IPC handle, move the points in."""
        if self.world > L.GM_MAX_PEERS:
            raise RuntimeError(f'peer update supports up to {L.GM_MAX_PEERS} ranks (one NVLink domain)')
        if self._need_even and x.shape[0] % self.world != 0:
            raise RuntimeError(f'{x.shape[0]} rows do not split evenly over {self.world} ranks')
        lib = L.lib()
        base = ctypes.c_void_p()
        with torch.cuda.device(self.dev):
            L.check(lib.gm_peer_alloc(self.nbytes, ctypes.byref(base)), 'gm_peer_alloc')
            self.base = base.value
            handle = ctypes.create_string_buffer(L.GM_PEER_HANDLE_BYTES)
            L.check(lib.gm_peer_export(ctypes.c_void_p(self.base), handle), 'gm_peer_export')
        self._handle = handle.raw
        raw = torch.as_tensor(_RawCuda(self.base, self.nbytes), device=self.dev)
        self._raw = raw
        nb = x.numel() * x.element_size()
        n_acc = self.n_acc
        self.x = raw[self.off_x:self.off_x + nb].view(x.dtype).view(x.shape)
        self.grad = raw[self.off_g:self.off_g + nb].view(x.dtype).view(x.shape)
        self.acc = raw[self.off_acc:self.off_acc + 8 * n_acc].view(torch.float64)
        self.acc_out = raw[self.off_out:self.off_out + 8 * n_acc].view(torch.float64)
        self.x.copy_(x)

    def exchange_handles(self):
        """The one collective between phase 1 and 2; only called once every rank has passed phase 1."""
        handles = [None] * self.world
        dist.all_gather_object(handles, (os.getpid(), self._handle), group=self.group)
        return handles

    def phase_open(self, handles):
        """Phase 2 (local): map every peer's arena through CUDA IPC and fill the table the kernel walks."""
        lib = L.lib()
        self.peer_base = []
        with torch.cuda.device(self.dev):
            for r, (pid, h) in enumerate(handles):
                if r == self.rank:
                    self.peer_base.append(self.base)
                    continue
                p = ctypes.c_void_p()
                hbuf = ctypes.create_string_buffer(L.GM_PEER_HANDLE_BYTES)
                hbuf.raw = h
                L.check(lib.gm_peer_open(hbuf, ctypes.byref(p)), f'gm_peer_open(rank {r})')
                self.peer_base.append(p.value)
                self._opened.append(p.value)
        self.table = L.Peers()
        self.table.world, self.table.rank, self.table.row_lo = self.world, self.rank, self.lo
        self.table.n_acc = self.n_acc
        self.table.acc_out = self.base + self.off_out
        # local workspace of the pipelined exchange (sum of every rank's gradient rows this rank owns); GM_PEER_PIPELINE=0
        # keeps the single fused kernel
        # Measured on B200 (bench.py, 2 M points): 2 GPUs 1.361 ms per step pipelined vs 1.452 fused; 4 GPUs 1.44-1.52
        # (2-8 chunks) vs 1.420 fused; 8 GPUs 1.62 vs 1.43 -- with more peers the fused kernel's thousands of blocks
        # already overlap pulls and pushes, and the extra launches cost more than they buy.  Default: pipelined for 2 ranks.
        self.gsum = None
        mode = os.environ.get('GM_PEER_PIPELINE', 'auto')
        if self._need_even and (mode == '1' or (mode == 'auto' and self.world == 2)):
            self.gsum = torch.empty_like(self.x[self.lo:self.hi])
            self.table.gsum = self.gsum.data_ptr()
        for r, b in enumerate(self.peer_base):
            self.table.x[r] = b + self.off_x
            self.table.grad[r] = b + self.off_g
            self.table.flags[r] = b + self.off_flags
            self.table.acc[r] = b + self.off_acc
        self.epoch = 0

    def phase_finish(self):
        """Phase 3: only after global success -- every arena is zero-filled and mapped before anyone raises a flag."""
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)

    def own(self, t):
        return t[self.lo:self.hi]

    def next_table(self):
        """The peer table for the next lock-step call (epoch advanced by one)."""
        self.epoch += 1
        self.table.epoch = self.epoch
        return self.table

    def close(self):
        lib = L.lib()
        for p in self._opened:
            lib.gm_peer_close(ctypes.c_void_p(p))
        self._opened = []
        if self.base:
            lib.gm_peer_free(ctypes.c_void_p(self.base))
            self.base = None


# ---------------------------------------------------------------------------------------------------------------------
# Row-sharded embeddings: every rank holds ONLY the rows it owns (cyclic ownership, include/gm_kernels.h
# gm_row_shards_t); the pair kernel gathers remote rows and reduces remote gradient rows over NVLink peer memory.
# ---------------------------------------------------------------------------------------------------------------------
def cyclic_shard(t, rank, world):
    """Rows rank, rank + world, rank + 2 world, ... of the (N, ...) tensor `t` (a contiguous copy)."""
    return t[rank::world].contiguous()


def cyclic_unshard(shards):
    """Inverse of cyclic_shard: shards[r] holds rows r, r + world, ...; the row counts may differ by one."""
    world = len(shards)
    n = sum(s.shape[0] for s in shards)
    full = shards[0].new_empty((n,) + tuple(shards[0].shape[1:]))
    for r, s in enumerate(shards):
        full[r::world] = s
    return full


def gather_cyclic(shard, group=None):
    """The full (N, ...) table from every rank's cyclic shard (equal row counts): an all-gather, then cyclic_unshard.
    Works on any backend (NCCL on GPUs, gloo in the CPU tests); for validation / snapshots, not on the step path."""
    world = dist.get_world_size(group)
    shards = [torch.empty_like(shard) for _ in range(world)]
    dist.all_gather(shards, shard.contiguous(), group=group)
    return cyclic_unshard(shards)


class ShardedArena(PeerArena):
    """Per-rank arena [point shard | gradient shard | step accumulator | accumulator sum | flag block] of a
    ROW-SHARDED embedding, mapped by every other rank through CUDA IPC.  `x` is this rank's shard (cyclic_shard of the
    full table; every rank must pass the same number of rows -- N divisible by the world size).  x_ptrs / grad_ptrs
    are the tables gm_pairs_loss_fused_sharded takes; next_table() the flag / accumulator table of gm_peer_barrier."""
    _need_even = False

    def _init_fields(self, x, n_acc, group):
        super()._init_fields(x, n_acc, group)
        if self.world & (self.world - 1):
            raise RuntimeError('row-sharded embeddings need a power-of-two number of ranks (cyclic ownership)')
        self.rows, self.lo, self.hi = x.shape[0], 0, x.shape[0]

    def phase_open(self, handles):
        super().phase_open(handles)
        self.table.row_lo = 0
        self.x_ptrs = [b + self.off_x for b in self.peer_base]
        self.grad_ptrs = [b + self.off_g for b in self.peer_base]


def run_phases(group, device, phases):
    """Runs `phases` = [(local_fn, collective_fn or None), ...] in lock step over the ranks of `group`: after every
    local_fn a MIN all-reduce of a status flag decides whether ALL ranks continue (then collective_fn, if any, runs on
    all of them) or ALL ranks stop.  A rank whose local_fn raised still takes part in that vote, so every rank
    executes the same sequence of collectives whether or not it failed.  Returns (ok, first local exception)."""
    err = None
    for local_fn, collective_fn in phases:
        ok = True
        try:
            local_fn()
        except Exception as e:  # noqa: BLE001 -- any failure on any rank => all ranks fall back together
            err, ok = e, False
        flag = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) != 1:
            return False, err
        if collective_fn is not None:
            collective_fn()
    return True, None


def try_peer_arena(x, n_acc, group, cls=None):
    """PeerArena (or `cls`, a subclass) if EVERY rank of `group` could allocate, export and map the arenas (NCCL backend, CUDA IPC between
    the processes, at most GM_MAX_PEERS ranks); otherwise None on every rank, and the caller keeps the NCCL
    reduce-scatter / all-gather owner update.  GM_PEER_UPDATE=0 disables the attempt."""
    if group is None or not x.is_cuda or dist.get_backend(group) != 'nccl':
        return None
    world = dist.get_world_size(group)
    if world < 2 or os.environ.get('GM_PEER_UPDATE', '1') == '0':
        return None
    cls = cls or PeerArena
    arena = cls.__new__(cls)
    arena._init_fields(x, n_acc, group)
    box = {}
    ok, err = run_phases(group, x.device, [
        (lambda: arena.phase_alloc_export(x), lambda: box.update(handles=arena.exchange_handles())),
        (lambda: arena.phase_open(box['handles']), None),
    ])
    if ok:
        arena.phase_finish()  # barrier only after global success
        return arena
    arena.close()
    if err is not None and dist.get_rank(group) == 0:
        import logging
        logging.getLogger(__name__).warning('peer-memory owner update unavailable (%s); using NCCL collectives', err)
    return None
