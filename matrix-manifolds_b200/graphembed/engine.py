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
import os

import torch
from torch.nn.functional import softplus

from . import _lib as L
from . import _ops
from . import _torch_ops
from .modules import ManifoldParameter, _softplus_value
from .parallel import (RowShards, ShardedArena, allreduce_step_buffers, cyclic_shard, gather_cyclic,
                       try_peer_arena)


def pack_hops(idx_j, hops):
    """Second-endpoint index and hop count of every pair in ONE int32 word: (hops << 24) | idx_j.  Needs fewer than
    2^24 rows and hop counts < 255 (uint8).  Works on host or device tensors; this is the 4-byte-per-pair upload format
    of `PairTrainer.step*(..., hops=None)` (GM_TGT_HOPS_PACKED)."""
    if idx_j.dtype != torch.int32 or hops.dtype != torch.uint8:
        raise ValueError('pack_hops: int32 indices and uint8 hop counts')
    if idx_j.numel() and int(idx_j.max()) >= (1 << 24):
        raise ValueError('pack_hops: row ids must be below 2^24')
    return (idx_j | (hops.to(torch.int32) << 24)).contiguous()


def pack_hops3(idx_j, hops):
    """(j, hop count) in THREE bytes per pair: little-endian 24-bit words, bits 0-20 = j (fewer than 2^21 rows), bits
    21-23 = hops - 1 (1 <= hops <= 8).  Returns a uint8 tensor of 3*P bytes padded to a multiple of 4; the upload format
    of `PairTrainer.step_host_grouped(..., idx_j=<this>, hops=None)` -- 25 % fewer bytes over PCIe than pack_hops, which
    is what bounds the end-to-end step on one GPU.  Expanded on the device by gm_unpack_pairs3."""
    if idx_j.dtype != torch.int32 or hops.dtype != torch.uint8:
        raise ValueError('pack_hops3: int32 indices and uint8 hop counts')
    if idx_j.numel():
        if int(idx_j.max()) >= (1 << 21) or int(idx_j.min()) < 0:
            raise ValueError('pack_hops3: row ids must be below 2^21')
        if int(hops.min()) < 1 or int(hops.max()) > 8:
            raise ValueError('pack_hops3: hop counts must be in 1..8')
    w = idx_j.to(torch.int64) | ((hops.to(torch.int64) - 1) << 21)
    b = torch.stack([w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF], dim=1).to(torch.uint8).reshape(-1)
    pad = (-b.numel()) % 4
    if pad:
        b = torch.cat([b, torch.zeros(pad, dtype=torch.uint8, device=b.device)])
    return b.contiguous()


def pack_hops2(offsets, idx_j, hops):
    """Source-grouped batch -> the TWO-byte-per-pair upload format of gm_unpack_pairs2.  offsets int64 (G + 1,): group
    g is pairs offsets[g] <= k < offsets[g + 1]; idx_j int32 (P,), hops uint8 (P,).  Every group's targets are put in
    ascending row order (the returned permutation `order` says how: pair k of the packed batch is pair order[k] of the
    input) and stored as gaps: word k = (j_k - j_{k-1}) | (hops_k - 1) << 13, the first gap of a group taken from
    bases[g] = the group's smallest row.  Returns (words int16 (P,), bases int32 (G,), order int64 (P,)), or None when
    the batch does not fit (a gap above 8191 rows or a hop count outside 1..8): the caller then uploads 4-byte words."""
    if idx_j.dtype != torch.int32 or hops.dtype != torch.uint8 or offsets.dtype != torch.int64:
        raise ValueError('pack_hops2: int64 offsets, int32 indices, uint8 hop counts')
    P, G = idx_j.numel(), offsets.numel() - 1
    if P == 0:
        return torch.empty(0, dtype=torch.int16), torch.zeros(G, dtype=torch.int32), torch.empty(0, dtype=torch.int64)
    if int(hops.min()) < 1 or int(hops.max()) > 8 or int(idx_j.min()) < 0:
        return None
    counts = offsets[1:] - offsets[:-1]
    group = torch.repeat_interleave(torch.arange(G, dtype=torch.int64), counts)
    # sort by (group, row): one stable sort of the rows, then a stable sort of the groups
    o1 = torch.sort(idx_j.long(), stable=True).indices
    order = o1[torch.sort(group[o1], stable=True).indices]
    j = idx_j.long()[order]
    first = torch.zeros(P, dtype=torch.bool)
    nonempty = counts > 0
    first[offsets[:-1][nonempty]] = True
    gap = torch.zeros(P, dtype=torch.int64)
    gap[1:] = j[1:] - j[:-1]
    gap[first] = 0
    if int(gap.max()) > 0x1FFF:
        return None
    bases = torch.zeros(G, dtype=torch.int32)
    bases[nonempty] = j[offsets[:-1][nonempty]].to(torch.int32)
    words = gap | ((hops.long()[order] - 1) << 13)
    words = torch.where(words >= 0x8000, words - 0x10000, words).to(torch.int16)  # the 16 bits, as torch's signed type
    return words.contiguous(), bases.contiguous(), order


class PairTrainer:
    """Drives (I, J, hops) pair batches through a single-manifold embedding.

    embedding : ManifoldEmbedding with one factor, on a CUDA device
    optimizer : RiemannianAdam / RiemannianSGD over embedding.xs
    objective : QuotientLoss / StressLoss
    max_hops_sq : max over the graph of hop^2 (GraphDataset normalisation, data/dataset.py:11-12)
    process_group : optional torch.distributed group; gradients (and the loss) are summed over it (pair-sharded data
        parallelism).
    owner_update : with a process group, reduce-scatter the gradient and let every rank update (and keep optimizer
        state for) only the rows it owns, then all-gather the new points; otherwise all-reduce and update replicas.
        Both give the same trajectory up to summation order.  Default: owner update whenever the rows split evenly.
        On GPUs of one NVLink domain the owner update is a single kernel over peer memory (gm_optim_step_peer: pull
        and sum the owned gradient rows from every rank, update, push the new rows to every rank) instead of
        ncclReduceScatter + update + ncclAllGather; it is chosen automatically when CUDA IPC between the ranks works
        (`self.peer` is then the PeerArena; GM_PEER_UPDATE=0 forces the NCCL collectives).
    """

    def __init__(self, embedding, optimizer, objective, max_hops_sq, alpha=1.0, process_group=None, owner_update=None,
                 deterministic=False):
        if embedding.n_components != 1:
            raise ValueError('PairTrainer drives a single-manifold embedding; use BatchedObjective for products')
        self.deterministic = bool(deterministic)
        self.emb, self.opt, self.obj = embedding, optimizer, objective
        self.max_hops_sq, self.alpha, self.pg = float(max_hops_sq), alpha, process_group
        self.x = embedding.xs[0]
        self.n_points = self.x.shape[0]
        self.man = embedding.manifolds[0]
        self.grad = torch.zeros_like(self.x, memory_format=torch.contiguous_format)
        self.x.grad = self.grad
        self.acc = torch.zeros(2, dtype=torch.float64, device=self.x.device)
        self._staging = None
        self._copy_stream = None
        self.shards = None
        self.peer = None
        self._grad_is_clean = False
        self._fold_zero_grad = False
        if process_group is not None and torch.distributed.get_world_size(process_group) > 1:
            even = self.x.shape[0] % torch.distributed.get_world_size(process_group) == 0
            if owner_update is None:
                owner_update = even
            if owner_update:
                self.peer = try_peer_arena(self.x.data, 2, process_group)
                if self.peer is not None:
                    # points, gradient and step accumulator move into the IPC-mapped arena
                    self.x.data = self.peer.x
                    self.grad = self.peer.grad
                    self.x.grad = self.grad
                    self.acc = self.peer.acc
                    self.shards = self.peer  # same own()/lo/hi/rows interface as RowShards
                else:
                    self.shards = RowShards(self.x.shape[0], process_group)
                # the optimizer now drives a parameter that aliases the owned rows of x
                own = ManifoldParameter(self.shards.own(self.x.data), manifold=self.man)
                if self.peer is not None:
                    own.grad = self.peer.own(self.grad)  # placeholder: the kernel sums every rank's rows itself
                    own._gm_peer_arena = self.peer
                else:
                    own.grad = torch.zeros_like(own.data)
                self._own = own
                replaced = False
                for g in optimizer.param_groups:
                    for k, prm in enumerate(g['params']):
                        if prm is self.x:
                            g['params'][k] = own
                            replaced = True
                if not replaced:
                    raise ValueError('optimizer does not hold the embedding parameter')
        if self.shards is None and os.environ.get('GM_FOLD_ZERO_GRAD', '0') == '1':
            # the trainer owns the gradient buffer, so the optimizer kernel can hand every row back zeroed
            # (gm_optim_t.zero_grad) instead of a full-table memset at the top of the next step.  OFF by default:
            # measured on B200 (2 M SPD4 points) the step is 1.28 ms with it and 1.19 ms with the memset -- the
            # memset leaves the 128 MB table L2-resident for the pair kernel's reductions, the fused zeroing does not
            self.x._gm_zero_grad_after_step = True
            self._fold_zero_grad = True

    # ---- device-resident inputs ---------------------------------------------------------------------------------
    def step(self, idx_i, idx_j, hops, epoch=1, segments=0):
        """One training step on device tensors; returns the (device, float64) loss of the batch.  `segments`: the
        batch is in window_order(..., segments) (gm_pairs_t.segments: the kernel walks the windows one after another)."""
        pairs = _ops.PairSet.from_lists(idx_i, idx_j, self.x.device, segments=segments)
        if hops is None:  # hop counts packed into the top byte of idx_j (pack_hops)
            if self.n_points > (1 << 24) or pairs.idx64:
                raise ValueError('packed hop counts need int32 indices and fewer than 2^24 points')
            targets = _ops.TargetSpec.hops_packed(self.max_hops_sq)
        else:
            targets = _ops.TargetSpec.hops(hops, self.max_hops_sq)
        return self._step_pairs(pairs, targets, epoch)

    def step_sampled(self, sources, levels, per_src, seed, slots=None, epoch=1):
        """One training step whose pairs are DRAWN INSIDE the pair kernel (GM_PAIRS_SAMPLED): for every BFS source
        sources[g] (int32 device tensor) `per_src` targets j != i come from the counter hash of (seed, pair number) and
        their hop counts from row slots[g] (default g) of the resident uint8 (S, N) matrix `levels`.  Nothing per pair
        is uploaded, stored or read: the step's input is the list of sources.  The draw is stated in include/gm_kernels.h
        and can be reproduced on the host bit for bit, so `step(I, pack_hops(J, H), None)` on such lists is the same step."""
        pairs = _ops.PairSet.sampled(sources, levels, per_src, seed, slots=slots)
        if levels.shape[1] != self.n_points:
            raise ValueError('the level matrix must have one column per embedded point')
        return self._step_pairs(pairs, _ops.TargetSpec.hops_packed(self.max_hops_sq), epoch)

    def step_sampled_host(self, sources, levels, per_src, seed, slots=None, epoch=1, defer_loss=False):
        """step_sampled from PINNED host tensors: `sources` (and `slots`) int32 (G,) are the step's whole upload --
        8 bytes per source instead of 4-9 bytes per pair.  Returns the loss as a Python float (a device->host read), or
        with defer_loss=True the loss of the previous deferred step (see step_host_grouped)."""
        G = sources.numel()
        dev = self.x.device
        if getattr(self, '_src_staging', None) is None or self._src_staging[0][0].numel() < G:
            self._src_staging = [(torch.empty(G, dtype=torch.int32, device=dev),
                                  torch.empty(G, dtype=torch.int32, device=dev)) for _ in range(2)]
            self._src_slot = 0
        ds, dl = self._src_staging[self._src_slot]
        self._src_slot ^= 1  # the other buffer may still be read by the previous step's kernel
        ds[:G].copy_(sources, non_blocking=True)
        if slots is not None:
            dl[:G].copy_(slots, non_blocking=True)
        loss = self.step_sampled(ds[:G], levels, per_src, seed, slots=None if slots is None else dl[:G], epoch=epoch)
        if defer_loss:
            return self._queue_loss_read()
        return loss.item()

    def _pairs_deterministic(self, pairs, targets, loss_spec, sp):
        """The fused kernel's job with a fixed summation order: grad (zero on entry) and acc are filled."""
        if pairs.mode != L.GM_PAIRS_LIST:
            raise ValueError('deterministic accumulation takes explicit pair lists')
        x = self.x.detach()
        I = pairs.idx_i
        J = pairs.idx_j
        if targets.mode == L.GM_TGT_HOPS_PACKED:  # split the packed word: row id | hop count << 24
            hops = ((J >> 24) & 0xff).to(torch.uint8)
            J = (J & 0x00ffffff).contiguous()
            targets = _ops.TargetSpec.hops(hops, targets.max_sq)
        plain = _ops.PairSet.from_lists(I, J, x.device)
        d2 = _ops.pairs_dist2(self.man.spec, x, x, plain)
        acc, g = _ops.product_loss([d2], [sp], targets, loss_spec)
        self.acc.copy_(acc[:2])
        xi, xj = x.index_select(0, I.long()), x.index_select(0, J.long())
        gi, gj = torch.empty_like(xi), torch.empty_like(xj)
        _ops.pairs_grad(self.man.spec, xi, xj, _ops.PairSet.elementwise(plain.P), g, gi, gj, coef=sp)
        _ops.scatter_add_rows_deterministic(torch.cat([gi, gj]), torch.cat([I.long(), J.long()]), self.grad)

    def _step_pairs(self, pairs, targets, epoch):
        loss_spec = self.obj.loss_spec(epoch=epoch, alpha=self.alpha)
        if not self._grad_is_clean:
            self.grad.zero_()
        self.acc.zero_()
        if self.deterministic:
            self._pairs_deterministic(pairs, targets, loss_spec, _softplus_value(self.emb.scales[0]))
            if self.peer is not None:
                self.opt.step()
                return self.peer.acc_out[0]
            if self.shards is not None or (self.pg is not None and torch.distributed.get_world_size(self.pg) > 1):
                raise RuntimeError('deterministic accumulation: single GPU, or the peer-memory owner update (which '
                                   'sums the ranks in rank order)')
            self.opt.step()
            self._grad_is_clean = self._fold_zero_grad
            return self.acc[0]
        # softplus(scale) is re-read every step (cached on the parameter until somebody steps it or loads a snapshot);
        # the trainer itself does not train the scale: acc[1] (sum l' d2) is there for a caller who does
        sp = _softplus_value(self.emb.scales[0])
        if (pairs.mode == L.GM_PAIRS_LIST and self.man.spec.kind != L.GM_UNIVERSAL
                and targets.mode in (L.GM_TGT_HOPS_PACKED, L.GM_TGT_HOPS_U8, L.GM_TGT_HOPS_U16)):
            # explicit pair lists: through the registered custom op (torch.ops.graphembed_b200.pairs_loss_fused)
            torch.ops.graphembed_b200.pairs_loss_fused(
                self.x.detach(), pairs.idx_i, pairs.idx_j, targets.data, *_torch_ops.manifold_args(self.man.spec),
                loss_spec.kind, bool(loss_spec.inc_l1), bool(loss_spec.inc_l2), float(loss_spec.alpha),
                float(loss_spec.eps), float(targets.max_sq), float(sp), self.grad, self.acc, pairs.segments)
        else:
            _ops.pairs_loss_fused(self.man.spec, self.x.detach(), pairs, targets, loss_spec, sp, self.grad, self.acc)
        if self.peer is not None:
            self.opt.step()  # ONE kernel: cross-GPU barrier, pull+sum gradients, update, push points, sum the loss
            return self.peer.acc_out[0]
        if self.shards is None:
            allreduce_step_buffers(self.grad, self.acc, self.pg)
            self.opt.step()
            # single GPU / replicated update: the optimizer kernel zeroed every gradient row after reading it
            self._grad_is_clean = self._fold_zero_grad
        else:
            self.shards.reduce_scatter(self.grad, out=self._own.grad)
            torch.distributed.all_reduce(self.acc, group=self.pg)
            self.opt.step()
            self.shards.all_gather(self.x.data)
        return self.acc[0]

    # ---- host-resident inputs (what a data loader hands over) ----------------------------------------------------------
    def _ensure_staging(self, P, hop_dtype, G=0):
        if (self._staging is None or self._staging[0][0].numel() < P or self._staging[0][3].numel() < G
                or self._staging[0][2].dtype != hop_dtype):
            dev = self.x.device
            self._staging = [(torch.empty(P, dtype=torch.int32, device=dev), torch.empty(P, dtype=torch.int32, device=dev),
                              torch.empty(P, dtype=hop_dtype, device=dev), torch.empty(max(G, 1), dtype=torch.int32, device=dev),
                              torch.empty(max(G, 1) + 1, dtype=torch.int64, device=dev)) for _ in range(2)]
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
            di, dj, dh = self._staging[slot][:3]
            di[:P].copy_(idx_i, non_blocking=True)
            dj[:P].copy_(idx_j, non_blocking=True)
            dh[:P].copy_(hops, non_blocking=True)
        di, dj, dh = self._staging[slot][:3]
        self._pending = None
        if next_batch is not None:
            nslot = 1 - slot
            ni, nj, nh = self._staging[nslot][:3]
            n = next_batch[0].numel()
            if n > ni.numel() or next_batch[2].dtype != nh.dtype:
                raise ValueError('next_batch does not fit the staging buffers (more pairs, or another hop dtype, than '
                                 'the batch of this step)')
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

    # ---- pipelined loss read-back --------------------------------------------------------------------------------
    def _queue_loss_read(self):
        """Async device->host copy of this step's [loss, d loss/d scale] accumulator into pinned memory; returns the
        loss of the PREVIOUS queued step (None for the first).  The host then runs one step ahead of the GPU."""
        if getattr(self, '_loss_host', None) is None:
            self._loss_host = torch.empty(2, 2, dtype=torch.float64).pin_memory()
            self._loss_events = [torch.cuda.Event(), torch.cuda.Event()]
            self._loss_slot, self._loss_pending = 0, False
        prev = self.flush_loss()
        s = self._loss_slot
        self._loss_host[s].copy_(self.acc if self.peer is None else self.peer.acc_out, non_blocking=True)
        self._loss_events[s].record(torch.cuda.current_stream(self.x.device))
        self._loss_pending = True
        return prev

    def flush_loss(self):
        """Loss of the last step issued with defer_loss=True (waits for it), or None if nothing is pending."""
        if not getattr(self, '_loss_pending', False):
            return None
        s = self._loss_slot
        self._loss_events[s].synchronize()
        self._loss_pending = False
        self._loss_slot = 1 - s
        return float(self._loss_host[s][0])

    def step_host_grouped(self, sources, offsets, idx_j, hops, epoch=1, next_batch=None, defer_loss=False,
                          segments=0, bases=None):
        """One step from PINNED host tensors in source-grouped (CSR-like) form, the natural output of a sampler that
        draws targets per BFS source: pairs offsets[g] <= k < offsets[g+1] are (sources[g], idx_j[k]) with hop count
        hops[k].  sources int32 (G,), offsets int64 (G+1,), idx_j int32 (P,), hops uint8/int16 (P,).  Uploads 5 bytes
        per pair instead of 9 -- or 4 with hops=None and idx_j = pack_hops(j, hops) -- and the first-endpoint index
        vector is expanded on the device (gm_expand_groups).
        `next_batch` = the next step's (sources, offsets, idx_j, hops), uploaded on a second stream meanwhile.
        defer_loss=True returns the loss of the previous deferred step instead of blocking on this one (its own loss
        is copied to pinned host memory asynchronously; `flush_loss()` returns the last one), so that the host can
        enqueue step k+1 while the GPU runs step k.
        Two bytes per pair: idx_j = the int16 words of pack_hops2 (every group's targets sorted by row, stored as
        gaps), hops=None, bases = its int32 (G,) vector; `next_batch` then carries the bases as a fifth entry."""
        packed3 = hops is None and idx_j.dtype == torch.uint8  # pack_hops3: 3 bytes per pair over PCIe
        packed2 = hops is None and idx_j.dtype == torch.int16  # pack_hops2: 2 bytes per pair
        if packed2 and bases is None:
            raise ValueError('2-byte pair words need the bases vector of pack_hops2')
        P, G = (int(offsets[-1]) if packed3 else idx_j.numel()), sources.numel()
        packed = hops is None  # idx_j carries the hop counts (pack_hops / pack_hops3)
        self._ensure_staging(P, torch.uint8 if packed else hops.dtype, G)
        if packed3 and (getattr(self, '_staging3', None) is None or self._staging3[0].numel() < idx_j.numel()):
            self._staging3 = [torch.empty(idx_j.numel(), dtype=torch.uint8, device=self.x.device) for _ in range(2)]
        if packed2 and (getattr(self, '_staging2', None) is None or self._staging2[0][0].numel() < P
                        or self._staging2[0][1].numel() < G):
            self._staging2 = [(torch.empty(P, dtype=torch.int16, device=self.x.device),
                               torch.empty(max(G, 1), dtype=torch.int32, device=self.x.device)) for _ in range(2)]
        cur = torch.cuda.current_stream(self.x.device)

        def upload(slot, batch):
            di, dj, dh, ds, do = self._staging[slot]
            s_, o_, j_, h_ = batch[:4]
            if j_.dtype == torch.int16:  # 2-byte words + bases: staged raw, expanded on the device before the step
                w2, b2 = self._staging2[slot]
                if s_.numel() > ds.numel() or j_.numel() > w2.numel() or len(batch) < 5 or batch[4].numel() > b2.numel():
                    raise ValueError('batch does not fit the staging buffers of the first batch of this call')
                ds[:s_.numel()].copy_(s_, non_blocking=True)
                do[:o_.numel()].copy_(o_, non_blocking=True)
                w2[:j_.numel()].copy_(j_, non_blocking=True)
                b2[:batch[4].numel()].copy_(batch[4], non_blocking=True)
                return
            jcap = self._staging3[slot].numel() if j_.dtype == torch.uint8 else dj.numel()
            if s_.numel() > ds.numel() or j_.numel() > jcap or (h_ is not None and h_.dtype != dh.dtype):
                raise ValueError('batch does not fit the staging buffers (more sources / pairs, or another hop dtype, '
                                 'than the first batch of this call)')
            ds[:s_.numel()].copy_(s_, non_blocking=True)
            do[:o_.numel()].copy_(o_, non_blocking=True)
            if j_.dtype == torch.uint8:  # 3-byte words: staged raw, expanded on the device just before the step
                self._staging3[slot][:j_.numel()].copy_(j_, non_blocking=True)
            else:
                dj[:j_.numel()].copy_(j_, non_blocking=True)
            if h_ is not None:
                dh[:h_.numel()].copy_(h_, non_blocking=True)

        if self._pending is not None and self._pending[0] is idx_j:
            slot, ev = self._pending[1], self._pending[2]
            cur.wait_event(ev)
        else:
            slot = self._slot
            upload(slot, (sources, offsets, idx_j, hops, bases))
        di, dj, dh, ds, do = self._staging[slot]
        self._pending = None
        if next_batch is not None:
            nslot = 1 - slot
            self._copy_stream.wait_stream(cur)  # the other slot was consumed by the previous step
            with torch.cuda.stream(self._copy_stream):
                upload(nslot, next_batch)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._pending = (next_batch[2], nslot, ev)
        self._slot = 1 - slot
        if packed2:  # one kernel: gaps -> rows (segmented scan per group) and the group's first endpoint per pair
            w2, b2 = self._staging2[slot]
            _ops.unpack_pairs2(w2[:P], b2[:G], do[:G + 1], dj, group_rows=ds[:G], out_i=di)
        else:
            _ops.expand_groups(ds[:G], do[:G + 1], di[:P])
        if packed3:
            _ops.unpack_pairs3(self._staging3[slot], P, dj)
        loss = self.step(di[:P], dj[:P], None if packed else dh[:P], epoch=epoch, segments=segments)
        if defer_loss:  # read this step's loss back asynchronously, hand out the previous step's
            return self._queue_loss_read()
        return loss.item()


class ShardedPairTrainer(PairTrainer):
    """PairTrainer over ROW-SHARDED embeddings (SURVEY 8(e) second scheme, BASELINE config 5 "embeddings row-sharded over
    NVLink"): every rank keeps only the rows v with v % world == rank of the point table, their gradient rows and
    their optimizer state; the pair batch is sharded over the ranks as before (every rank passes ITS pairs, with
    GLOBAL row ids).  One step is

        fused pair kernel   rows gathered from whichever shard holds them (16-byte cp.async over NVLink for remote
                            ones), gradient rows added into the owning shard (red.global.add over NVLink)
        barrier 0           every rank's reductions into my shard are final; loss summed over the ranks
        optimizer kernel    local shard only (gradient rows handed back zeroed)
        barrier 1           every shard updated

    -- no replicated table, no reduce-scatter / all-gather phase, no NCCL call on the step path; the NVLink traffic
    (about 2 x 64 B per remote pair endpoint for SPD 4x4 fp32) rides inside the pair kernel instead of following it.
    Same trajectory as PairTrainer on one GPU up to floating-point summation order.  Needs CUDA IPC between the ranks
    (one NVLink domain), a power-of-two world size dividing the number of points, and an SPD manifold; raises
    otherwise (use PairTrainer's replicated owner update then).

    embedding : every rank passes the same full ManifoldEmbedding (same initial points); after construction the full
        table is released (`keep_full=True` keeps it, stale) and `gather()` rebuilds it from the shards.
    """

    def __init__(self, embedding, optimizer, objective, max_hops_sq, process_group, alpha=1.0, keep_full=False):
        if embedding.n_components != 1:
            raise ValueError('ShardedPairTrainer drives a single-manifold embedding')
        dist = torch.distributed
        self.emb, self.opt, self.obj = embedding, optimizer, objective
        self.max_hops_sq, self.alpha, self.pg = float(max_hops_sq), alpha, process_group
        self.man = embedding.manifolds[0]
        if self.man.spec.kind not in (L.GM_SPD_AI, L.GM_SPD_STEIN):
            raise ValueError('row-sharded training is built for the SPD pair kernels')
        full = embedding.xs[0]
        self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        self.n_points = full.shape[0]
        if self.n_points % self.world:
            raise ValueError(f'{self.n_points} points do not split evenly over {self.world} ranks')
        arena = try_peer_arena(cyclic_shard(full.data, self.rank, self.world), 2, process_group, cls=ShardedArena)
        if arena is None:
            raise RuntimeError('row-sharded training needs CUDA IPC peer memory between all ranks (one NVLink domain, '
                               'NCCL process group, GM_PEER_UPDATE != 0)')
        self.peer = arena
        self.x, self.grad, self.acc = arena.x, arena.grad, arena.acc
        own = ManifoldParameter(arena.x, manifold=self.man)
        own.grad = arena.grad
        own._gm_zero_grad_after_step = True  # the local optimizer kernel hands every gradient row back zeroed
        self._own = own
        replaced = False
        for g in optimizer.param_groups:
            for k, prm in enumerate(g['params']):
                if prm is full:
                    g['params'][k] = own
                    replaced = True
        if not replaced:
            raise ValueError('optimizer does not hold the embedding parameter')
        self._full = full
        if not keep_full:
            full.data = full.data.new_empty((0,) + tuple(full.shape[1:]))
        self._staging = None
        self._copy_stream = None
        self.shards = None
        self._grad_is_clean = True
        self._fold_zero_grad = True

    def _step_pairs(self, pairs, targets, epoch):
        loss_spec = self.obj.loss_spec(epoch=epoch, alpha=self.alpha)
        self.acc.zero_()
        sp = _softplus_value(self.emb.scales[0])
        a = self.peer
        _ops.pairs_loss_fused_sharded(self.man.spec, a.x_ptrs, a.grad_ptrs, self.x.dtype, self.x.device, pairs, targets,
                                      loss_spec, sp, self.acc)
        table = a.next_table()
        _ops.peer_barrier(table, 0, self.x.device)
        self.opt.step()
        _ops.peer_barrier(table, 1, self.x.device)
        return a.acc_out[0]

    def gather(self):
        """The full (N, ...) point table, rebuilt from every rank's shard (an all-gather; for validation / snapshots)."""
        full = gather_cyclic(self.x, self.pg)
        self._full.data = full
        return full


def window_order(idx_j, n_points, segments):
    """Permutation that puts a pair batch into (window of the target row, original position) order: window w holds the
    pairs whose target j lies in rows [w n / segments, (w + 1) n / segments).  Applied to a source-grouped batch it
    keeps every source's pairs consecutive inside each window.  Launch the reordered lists with `segments=segments`
    (gm_pairs_t.segments): the pair kernel then walks window after window with its whole grid, so the gradient rows it
    scatters into and the point rows it gathers stay inside one L2-sized window at a time (measured on B200: the fused
    SPD 4x4 kernel leaves its 3.3 GB-per-launch random DRAM traffic bound and becomes issue bound).  idx_j: int tensor
    (a packed hop count in the top byte is ignored)."""
    j = idx_j.long() & 0x00ffffff if idx_j.dtype == torch.int32 else idx_j.long()
    window = (j * int(segments)) // int(n_points)
    return torch.sort(window, stable=True).indices


def window_groups(sources, offsets, idx_j, n_points, segments):
    """A source-grouped batch (sources int32 (G,), offsets int64 (G + 1,), targets idx_j (P,)) in window order: returns
    (order, group_rows, group_offsets) with `order` = window_order(idx_j, n_points, segments) and the (window, source)
    groups of the reordered batch -- group w * G + g holds the pairs of source g whose target lies in window w (possibly
    none), group_rows[w * G + g] = sources[g].  Feed `idx_j[order]` etc. with these to step_host_grouped(segments=...)
    or pack_hops2."""
    G = sources.numel()
    counts = offsets[1:] - offsets[:-1]
    group = torch.repeat_interleave(torch.arange(G, dtype=torch.int64), counts)
    order = window_order(idx_j, n_points, segments)
    j = idx_j.long() & 0x00ffffff if idx_j.dtype == torch.int32 else idx_j.long()
    key = ((j[order] * int(segments)) // int(n_points)) * G + group[order]
    group_counts = torch.bincount(key, minlength=int(segments) * G)
    group_offsets = torch.cat([torch.zeros(1, dtype=torch.int64), group_counts.cumsum(0)])
    return order, sources.repeat(int(segments)).contiguous(), group_offsets


class ProductPairTrainer:
    """(I, J, hops) pair batches through a PRODUCT-manifold embedding -- BASELINE config 3 ("product SPD 3x3 x Lorentz 5,
    sampled pairs").  The distance of a pair is sum_f softplus(scale_f) * d_f^2 (modules.py:84-88); one step is

        gm_pairs_product_fused (ONE launch: every factor's d2, the loss, sum l' d2_f per factor, every factor's
        gradient scatter-add)  ->  the optimizer kernels

    for products of at most one SPD factor with up to three of Lorentz / Sphere / Euclidean, and otherwise

        F x gm_pairs_dist2  ->  gm_product_loss (loss term, dL/dm per pair, sum l' d2_f per factor)
        ->  F x gm_pairs_grad (scatter-add of softplus(scale_f) * dL/dm * d(d2_f)/dx)  ->  the optimizer kernels

    with the reference's semantics of `loss.backward(); optimizer.step()`.  `scale_optimizer` (optional: any optimizer
    over embedding.scales, e.g. the second RiemannianAdam group of experiments/run_grid.py:25-28) receives
    d loss / d scale_f = sigmoid(scale_f) * sum_k l'_k d2_f,k and is stepped after the points.  With a process group
    the pair batch is sharded over the ranks and gradients, loss and scale gradients are all-reduced (replicated
    update)."""

    def __init__(self, embedding, optimizer, objective, max_hops_sq, alpha=1.0, scale_optimizer=None,
                 process_group=None, fused=None):
        self.emb, self.opt, self.obj = embedding, optimizer, objective
        self.max_hops_sq, self.alpha, self.pg = float(max_hops_sq), alpha, process_group
        self.scale_opt = scale_optimizer
        self.xs = list(embedding.xs)
        self.mans = list(embedding.manifolds)
        self.scales = list(getattr(embedding, 'scales', ()))
        if any(hasattr(m, 'get_c') for m in self.mans):
            raise ValueError('Universal factors carry a curvature gradient: train them through BatchedObjective')
        self.grads = [torch.zeros_like(x, memory_format=torch.contiguous_format) for x in self.xs]
        for x, g in zip(self.xs, self.grads):
            x.grad = g
        # one launch for all factors where the fused product kernel takes the factor list (fused=False: the F + 1 + F
        # launches above, kept for the other products and as the A/B of the fused kernel)
        can_fuse = _ops.product_fusable([m.spec for m in self.mans], [x.dtype for x in self.xs])
        self.fused = can_fuse if fused is None else (bool(fused) and can_fuse)

    def step(self, idx_i, idx_j, hops, epoch=1):
        """One training step on device tensors (int32/int64 indices, uint8/int16 hop counts); returns the (device,
        float64) loss of the batch."""
        dev = self.xs[0].device
        pairs = _ops.PairSet.from_lists(idx_i, idx_j, dev)
        targets = _ops.TargetSpec.hops(hops, self.max_hops_sq)
        loss_spec = self.obj.loss_spec(epoch=epoch, alpha=self.alpha)
        sps = [_softplus_value(s) for s in self.scales] if self.scales else [1.0] * len(self.xs)
        if self.fused:
            for gx in self.grads:
                gx.zero_()
            acc = _ops.pairs_product_fused([m.spec for m in self.mans], [x.detach() for x in self.xs], pairs, targets,
                                           loss_spec, sps, self.grads)
        else:
            d2s = [_ops.pairs_dist2(m.spec, x.detach(), x.detach(), pairs) for m, x in zip(self.mans, self.xs)]
            acc, g = _ops.product_loss(d2s, sps, targets, loss_spec)
            for m, x, gx, sp in zip(self.mans, self.xs, self.grads, sps):
                gx.zero_()
                _ops.pairs_grad(m.spec, x.detach(), x.detach(), pairs, g, gx, gx, coef=sp)
        if self.pg is not None and torch.distributed.get_world_size(self.pg) > 1:
            for gx in self.grads:
                torch.distributed.all_reduce(gx, group=self.pg)
            torch.distributed.all_reduce(acc, group=self.pg)
        self.opt.step()
        if self.scale_opt is not None:
            for f, s in enumerate(self.scales):
                s.grad = (acc[1 + f] * torch.sigmoid(s.detach().double())).to(s.dtype).reshape(s.shape)
            self.scale_opt.step()
        return acc[0]
