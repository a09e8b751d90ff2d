"""Sampled-pair training step: the hot path of BASELINE configs 3 and 5.

The reference only ever trains on all pairs of a node batch (train.py:198-228);
explicit (i, j, hop-count) pair batches are the capability BASELINE.json adds.
One step = zero the gradient, ONE fused kernel (gather, distance, loss term,
gradient, scatter-add), optional NCCL all-reduce of the dense gradient across
ranks, ONE fused optimizer kernel.  The semantics of a step are exactly those of

    loss = objective(targets(hops), softplus(scale) * man.dist(x[I], x[J], squared=True))
    optimizer.zero_grad(); loss.backward(); optimizer.step()

with the reference's manifold / objective / optimizer definitions.
"""
import torch
from torch.nn.functional import softplus

from . import _ops
from .parallel import allreduce_step_buffers


class PairTrainer:
    """Drives (I, J, hops) pair batches through a single-manifold embedding.

    embedding : ManifoldEmbedding with one factor, on a CUDA device
    optimizer : RiemannianAdam / RiemannianSGD over embedding.xs
    objective : QuotientLoss / StressLoss
    max_hops_sq : max over the graph of hop^2 (GraphDataset normalisation, data/dataset.py:11-12)
    process_group : optional torch.distributed group; gradients (and the loss) are summed over it, so that
        every rank applies the same update to its replica of the embedding (pair-sharded data parallelism).
    """

    def __init__(self, embedding, optimizer, objective, max_hops_sq, alpha=1.0, process_group=None):
        if embedding.n_components != 1:
            raise ValueError('PairTrainer drives a single-manifold embedding; use BatchedObjective for products')
        self.emb, self.opt, self.obj = embedding, optimizer, objective
        self.max_hops_sq, self.alpha, self.pg = float(max_hops_sq), alpha, process_group
        self.x = embedding.xs[0]
        self.man = embedding.manifolds[0]
        self.grad = torch.zeros_like(self.x, memory_format=torch.contiguous_format)
        self.x.grad = self.grad
        self.acc = torch.zeros(2, dtype=torch.float64, device=self.x.device)
        self._sp = float(softplus(embedding.scales[0].detach()))
        self._staging = None
        self._copy_stream = None

    # ---- device-resident inputs ---------------------------------------------------------------------------------
    def step(self, idx_i, idx_j, hops, epoch=1):
        """One training step on device tensors; returns the (device, float64) loss of the batch."""
        pairs = _ops.PairSet.from_lists(idx_i, idx_j, self.x.device)
        targets = _ops.TargetSpec.hops(hops, self.max_hops_sq)
        loss_spec = self.obj.loss_spec(epoch=epoch, alpha=self.alpha)
        self.grad.zero_()
        self.acc.zero_()
        _ops.pairs_loss_fused(self.man.spec, self.x.detach(), pairs, targets, loss_spec, self._sp, self.grad, self.acc)
        allreduce_step_buffers(self.grad, self.acc, self.pg)
        self.opt.step()
        return self.acc[0]

    # ---- host-resident inputs (what a data loader hands over) ----------------------------------------------------------
    def _ensure_staging(self, P, hop_dtype):
        if self._staging is None or self._staging[0][0].numel() < P:
            dev = self.x.device
            self._staging = [(torch.empty(P, dtype=torch.int32, device=dev), torch.empty(P, dtype=torch.int32, device=dev),
                              torch.empty(P, dtype=hop_dtype, device=dev)) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slot = 0
            self._pending = None

    def step_host(self, idx_i, idx_j, hops, epoch=1, next_batch=None):
        """One step from PINNED host tensors (int32 indices, uint8/int16 hop counts).  Copies the batch to the device,
        runs the step and returns the loss as a Python float (a device->host read).  If `next_batch` (a tuple of
        pinned host tensors) is given its upload is overlapped with this step's kernels on a second stream."""
        P = idx_i.numel()
        self._ensure_staging(P, hops.dtype)
        cur = torch.cuda.current_stream(self.x.device)
        if self._pending is not None and self._pending[0] is idx_i:
            slot, ev = self._pending[1], self._pending[2]
            cur.wait_event(ev)
        else:
            slot = self._slot
            di, dj, dh = self._staging[slot]
            di[:P].copy_(idx_i, non_blocking=True)
            dj[:P].copy_(idx_j, non_blocking=True)
            dh[:P].copy_(hops, non_blocking=True)
        di, dj, dh = self._staging[slot]
        self._pending = None
        if next_batch is not None:
            nslot = 1 - slot
            ni, nj, nh = self._staging[nslot]
            n = next_batch[0].numel()
            self._copy_stream.wait_stream(cur)  # the other slot was consumed by the previous step
            with torch.cuda.stream(self._copy_stream):
                ni[:n].copy_(next_batch[0], non_blocking=True)
                nj[:n].copy_(next_batch[1], non_blocking=True)
                nh[:n].copy_(next_batch[2], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._pending = (next_batch[0], nslot, ev)
        self._slot = 1 - slot
        loss = self.step(di[:P], dj[:P], dh[:P], epoch=epoch)
        return loss.item()
