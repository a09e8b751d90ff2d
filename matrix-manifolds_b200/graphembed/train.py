"""TrainingEngine -- the epoch loop around the hot path, with the constructor keywords, defaults, batching rule,
file names and call order of the reference's graphembed/train.py:22-352.

Per epoch (`_train`, train.py:198-228): `perm = torch.randperm(N)`; node batches of `batch_size` (whole graph if
None), tail batches shorter than `drop_last_n` dropped; per batch ONE fused kernel evaluates all B(B-1)/2 pair
distances, the loss and the gradient (BatchedObjective), then the fused optimizer kernels run.  Validation
(`_validate`, train.py:230-265) streams all N(N-1)/2 pairs through the distance + moments kernels
(metrics.validation_moments) instead of materialising them.

Multi-GPU: one process per GPU (torchrun).  With `process_group` set, every rank evaluates a contiguous slice of
each batch's pair triangle, the dense gradients and the loss are summed with one all-reduce per step, and all ranks
apply the same update.  The reference's nn.DataParallel path (train.py:107-109,203-204) drops pairs across device
chunks; this keeps every pair.  Plots (`add_figure`) are not produced."""
import logging
import math
import os
import tempfile

import torch

from . import metrics as metrics_mod
from .modules import BatchedObjective
from .utils import Timer, check_mkdir, latest_path_by_basename_numeric_order

logger = logging.getLogger(__name__)

# every attribute of the engine with its default (train.py:22-44); the first three are mandatory
default_attrs_ = dict(
    embedding=None, optimizer=None, objective_fn=None, alpha=None, n_epochs=2000, batch_size=None, drop_last_n=50,
    burnin_epochs=None, burnin_lower_lr=False, burnin_higher_lr=False, perturb_every_epochs=None,
    stabilize_every_epochs=None, lr_scheduler=None, min_lr=None, metrics=None, main_metric_idx=None,
    lazy_metrics=None, val_every_epochs=None, save_metrics_every_epochs=None, save_every_epochs=None, save_dir=None,
    snapshot_path=None)


class ScalarLog:
    """SummaryWriter stand-in that also keeps every scalar in memory: `history[tag] = [(step, value), ...]`.
    Forwards to a real tensorboard writer when one can be created.  Device tensors are accepted and only converted to
    Python floats when somebody looks (`history`, `flush()`, `close()`): logging the per-step loss must not force a
    device synchronisation per step, which is what dominates small node batches (BASELINE configs 2-3)."""

    def __init__(self, log_dir=None, tensorboard=True):
        self._history = {}
        self._pending = []  # (tag, step, 0-d device tensor) not yet converted
        self._tb = None
        if tensorboard and log_dir is not None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._tb = SummaryWriter(log_dir=log_dir)
            except Exception:  # tensorboard is optional
                self._tb = None

    def add_scalar(self, tag, value, step):
        if isinstance(value, torch.Tensor) and value.is_cuda:
            self._pending.append((tag, int(step), value.detach()))
            return
        self._store(tag, int(step), float(value))

    def _store(self, tag, step, v):
        self._history.setdefault(tag, []).append((step, v))
        if self._tb is not None:
            self._tb.add_scalar(tag, v, step)

    def flush(self):
        """Convert the queued device scalars (one stacked device->host copy) in the order they were logged."""
        if self._pending:
            pending, self._pending = self._pending, []
            values = torch.stack([v.double().reshape(()) for _, _, v in pending]).cpu().tolist()
            for (tag, step, _), v in zip(pending, values):
                self._store(tag, step, v)

    @property
    def history(self):
        self.flush()
        return self._history

    def add_figure(self, *args, **kwargs):
        pass

    def close(self):
        self.flush()
        if self._tb is not None:
            self._tb.close()


class _Snapshot:
    """What `best_struct['embedding']` needs to be: something with a state_dict() (train.py:324-341)."""

    def __init__(self, module):
        self._state = {k: v.detach().clone() for k, v in module.state_dict().items()}

    def state_dict(self):
        return self._state


class TrainingEngine:

    def __init__(self, **kwargs):
        self.__dict__.update(default_attrs_)
        self.process_group = None
        self.tensorboard = True
        self.__dict__.update(kwargs)
        if not self.embedding or not self.optimizer or not self.objective_fn:
            raise ValueError('`embedding`, `optimizer`, and `objective_fn` must be specified to construct a '
                             'TrainingEngine.')
        if self.burnin_lower_lr and self.burnin_higher_lr:
            raise ValueError('`burnin_lower_lr` and `burnin_higher_lr` are mutually exclusive.')
        if not isinstance(self.optimizer, (tuple, list)):
            self.optimizer = [self.optimizer]
        if self.lr_scheduler is not None and not isinstance(self.lr_scheduler, (tuple, list)):
            self.lr_scheduler = [self.lr_scheduler]
        if self.burnin_epochs is None:
            self.burnin_epochs = 0
        if self.perturb_every_epochs is None:
            self.perturb_every_epochs = self.n_epochs + 1
        if self.stabilize_every_epochs is None:
            self.stabilize_every_epochs = self.n_epochs + 1
        if self.metrics is None:
            self.metrics = ['pearsonr', 'average_distortion']
        if self.main_metric_idx is None:
            self.main_metric_idx = 0
        if self.val_every_epochs is None:
            self.val_every_epochs = self.n_epochs + 1
        if self.save_metrics_every_epochs is None:
            self.save_metrics_every_epochs = self.n_epochs // self.val_every_epochs * self.val_every_epochs
        if self.save_every_epochs is None:
            self.save_every_epochs = self.n_epochs
        if self.save_dir is None:
            self.save_dir = tempfile.gettempdir()
            check_mkdir(self.save_dir)
            logger.info('The save dir is (%s)', self.save_dir)
        if self.snapshot_path:
            self._load()

    # ---- multi-GPU helpers ---------------------------------------------------------------------------------------
    @property
    def _world(self):
        return 1 if self.process_group is None else torch.distributed.get_world_size(self.process_group)

    @property
    def _rank(self):
        return 0 if self.process_group is None else torch.distributed.get_rank(self.process_group)

    def __call__(self, graph_dataset, last_step=0):
        shard = None if self._world == 1 else (self._rank, self._world)
        self.batched_obj = BatchedObjective(self.objective_fn, graph_dataset, self.embedding, shard=shard)
        if self.burnin_epochs > 0 and self.alpha is not None:
            self._burnin_pre()
            self._burnin(graph_dataset)
            self._burnin_post()
        self.pending_metric_results = []
        self.pending_metric_idx = 0
        self.best_struct = dict(epoch=0, loss=1e8, embedding=_Snapshot(self.embedding))
        self.global_step = last_step
        self.writer = ScalarLog(self.save_dir if self._rank == 0 else None, self.tensorboard)
        try:
            for epoch in range(1, self.n_epochs + 1):
                self._run_epoch(graph_dataset, self.alpha, epoch)
                if self._check_early_break():
                    logger.warning('Early breaking (epoch=%d)', epoch)
                    break
        finally:
            if self._rank == 0:
                self._save_best()
            self.writer.close()
        self._consume_pending_metric_results(wait=True)

    # ---- burn-in (train.py:135-166) --------------------------------------------------------------------------------
    def _scale_lrs(self, factor):
        for optim in self.optimizer:
            for group in optim.param_groups:
                group['lr'] *= factor

    def _burnin_pre(self):
        self.embedding.burnin(True)
        self.global_step = 0
        self.writer = ScalarLog(os.path.join(self.save_dir, 'burnin') if self._rank == 0 else None, self.tensorboard)
        if self.burnin_lower_lr:
            self._scale_lrs(0.1)
        elif self.burnin_higher_lr:
            self._scale_lrs(10)

    def _burnin(self, graph_dataset):
        for i, alpha in enumerate([self.alpha / d for d in range(4, 0, -1)]):
            logger.info(f'Running burn-in epochs with alpha={alpha:.5f}')
            for epoch in range(i * self.burnin_epochs + 1, (i + 1) * self.burnin_epochs + 1):
                self._train(graph_dataset, alpha, epoch)
                if epoch % self.stabilize_every_epochs:  # (sic) stabilises when NOT a multiple, train.py:157
                    with Timer('stabilizing'), torch.no_grad():
                        self.embedding.stabilize()

    def _burnin_post(self):
        if self.burnin_lower_lr:
            self._scale_lrs(10)
        elif self.burnin_higher_lr:
            self._scale_lrs(0.1)
        self.embedding.burnin(False)
        self.writer.close()

    # ---- one epoch ------------------------------------------------------------------------------------------------
    def _run_epoch(self, graph_dataset, alpha, epoch):
        with Timer('training'):
            loss = self._train(graph_dataset, alpha, epoch)
        self._lr_scheduler_step(loss, epoch)
        with Timer('checking if better model'):
            self._check_best(loss, epoch)
        if epoch % self.perturb_every_epochs == 0:
            with Timer('perturbing'), torch.no_grad():
                self.embedding.perturb(1 / epoch)
        if epoch % self.stabilize_every_epochs == 0:
            with Timer('stabilizing'), torch.no_grad():
                self.embedding.stabilize()
        if epoch % self.val_every_epochs == 0:
            with Timer('validating'), torch.no_grad():
                self._validate(graph_dataset, epoch)
            self._consume_pending_metric_results()
        if epoch % self.save_every_epochs == 0 and self._rank == 0:
            self._save(epoch)

    def _train(self, graph_dataset, alpha, epoch):
        n_points = len(graph_dataset)
        bs = n_points if self.batch_size is None else min(n_points, self.batch_size)
        perm = torch.randperm(n_points)  # default device / default generator, as train.py:206
        if self._epoch_kernel_ready(graph_dataset):
            return self._train_epoch_kernel(graph_dataset, perm, bs, alpha, epoch)
        total_loss = None  # summed on the device: one host read per epoch instead of `loss.item()` per step
        for i in range(0, n_points, bs):
            indices = perm[i:(i + bs)]
            if len(indices) < self.drop_last_n:
                break
            if self._lean_ready(graph_dataset):
                with torch._C.DisableTorchFunction():  # see _ops._no_function_modes
                    loss = self._lean_step(graph_dataset, indices, alpha, epoch)
                    for optim in self.optimizer:
                        optim.step()
            else:
                loss = self.batched_obj(indices, alpha=alpha, epoch=epoch).sum()
                for optim in self.optimizer:
                    optim.zero_grad()
                loss.backward()
                if self._world > 1:
                    loss = self._combine_ranks(loss)
                for optim in self.optimizer:
                    optim.step()
            self.global_step += 1
            step_loss = loss.detach()
            self.writer.add_scalar(str(self.objective_fn), step_loss / len(indices), self.global_step)
            total_loss = step_loss.double() if total_loss is None else total_loss + step_loss.double()
        self.writer.flush()
        total_loss = 0 if total_loss is None else total_loss.item()
        logger.debug('epoch %d, train loss %.5f', epoch, total_loss / n_points)
        return total_loss

    # ---- lean step: the same arithmetic as forward/zero_grad/backward above without the autograd tape -------------------
    def _lean_ready(self, graph_dataset):
        """Single GPU, an objective with a fused form and GPU-resident targets: the step is `zero the gradient buffer ->
        pair kernels -> optimizer kernels` and nothing else -- gm_pairs_loss_fused for one (non-curved) manifold,
        gm_pairs_dist2 x F + gm_product_loss + gm_pairs_grad x F for products and Universal factors.  For node batches
        of a few hundred nodes (BASELINE configs 2-3) the autograd path spends ~10x the kernel time on the host."""
        st = getattr(self, '_lean', None)
        if st is not None and st['dataset'] is graph_dataset:
            return st['ok']
        emb = self.embedding
        pd = getattr(graph_dataset, 'pdists', None)
        ok = (self._world == 1 and getattr(emb, 'fused_pair_kernels', False) and hasattr(emb, 'manifolds')
              and 1 <= len(emb.xs) <= 8
              and getattr(self.objective_fn, 'loss_spec', None) is not None
              and self.objective_fn.loss_spec(epoch=1, alpha=1.0) is not None
              and pd is not None and pd.is_cuda and pd.device == emb.device and pd.dtype == emb.xs[0].dtype
              and all(x.is_contiguous() and x.dtype == emb.xs[0].dtype for x in emb.xs)
              and not torch.is_anomaly_enabled())
        st = dict(ok=ok, dataset=graph_dataset)
        if ok:
            from . import _ops
            dev = emb.device
            owned = {id(p) for optim in self.optimizer for g in optim.param_groups for p in g['params']}
            # every factor's gradient table, the step accumulator and the curvature-gradient slots share one
            # allocation: one memset per step
            sizes = [(x.numel() * x.element_size() + 15) // 16 * 16 for x in emb.xs]
            F = len(emb.xs)
            buf = torch.zeros(sum(sizes) + 16 + 8 * ((F + 1) // 2 * 2), dtype=torch.uint8, device=dev)
            grads, off = [], 0
            for x, sz in zip(emb.xs, sizes):
                grads.append(buf[off:off + x.numel() * x.element_size()].view(x.dtype).view(x.shape))
                off += sz
            acc = buf[off:off + 16].view(torch.float64)
            cgrads = buf[off + 16:off + 16 + 8 * F].view(torch.float64)
            scales = list(getattr(emb, 'scales', ()))
            st.update(buf=buf, grads=grads, acc=acc, cgrads=cgrads, targets=_ops.TargetSpec.dense(pd), scales=scales,
                      scale_trained=[id(s) in owned for s in scales],
                      curved=[hasattr(m, 'get_c') for m in emb.manifolds],
                      curv_trained=[hasattr(m, 'get_c') and id(m.c) in owned for m in emb.manifolds])
        self._lean = st
        return ok

    def _lean_step(self, graph_dataset, indices, alpha, epoch):
        from . import _ops
        from .modules import _softplus_value
        st, emb = self._lean, self.embedding
        xs, mans, scales = list(emb.xs), list(emb.manifolds), st['scales']
        F = len(xs)
        pairs = _ops.PairSet.triu(len(indices), indices, emb.device)
        loss_spec = self.objective_fn.loss_spec(epoch=epoch, alpha=alpha)
        grads = st['grads']
        st['buf'].zero_()
        sps = [_softplus_value(s) for s in scales] if scales else [1.0] * F
        if F == 1 and not st['curved'][0]:
            acc = st['acc']
            _ops.pairs_loss_fused(mans[0].spec, xs[0].detach(), pairs, st['targets'], loss_spec, sps[0], grads[0], acc)
            loss = acc[0].to(xs[0].dtype).clone()  # the accumulator is zeroed again next step
        else:
            cs = []  # curvature tensors get_c() of the Universal factors (autograd leaves behind them: man.c)
            for m, curved, trained in zip(mans, st['curved'], st['curv_trained']):
                if not curved:
                    cs.append(None)
                elif trained and m.c.requires_grad:
                    with torch.enable_grad():
                        cs.append(m.get_c())
                else:
                    cs.append(m.get_c().detach())
            cdet = [None if c is None else c.detach() for c in cs]
            d2s = [_ops.pairs_dist2(m.spec, x.detach(), x.detach(), pairs, c=c) for m, x, c in zip(mans, xs, cdet)]
            acc, g = _ops.product_loss(d2s, sps, st['targets'], loss_spec, pairs=pairs)
            for f, (m, x, c) in enumerate(zip(mans, xs, cdet)):
                want_c = cs[f] is not None and cs[f].requires_grad
                _ops.pairs_grad(m.spec, x.detach(), x.detach(), pairs, g, grads[f], grads[f], coef=sps[f], c=c,
                                c_grad=st['cgrads'][f:f + 1] if want_c else None)
                if curved := st['curved'][f]:
                    m.c.grad = None
                    if want_c:  # chain rule through get_c() (sign / softplus parametrisation, universal.py:28-32)
                        cs[f].backward(st['cgrads'][f:f + 1].to(cs[f].dtype).reshape(cs[f].shape))
            loss = acc[0].to(xs[0].dtype)
        for x, gx in zip(xs, grads):
            x.grad = gx  # what loss.backward() leaves behind (x[indices] backward: a dense (N, ...) gradient)
        for f, s in enumerate(scales):
            s.grad = None
            if st['scale_trained'][f] and s.requires_grad:  # d loss / d scale_f = sigmoid(scale_f) * sum_k l'_k d2_f,k
                s.grad = (acc[1 + f] * torch.sigmoid(s.detach().double())).to(s.dtype)
        return loss

    # ---- whole epoch in one native call (gm_train_epoch / gm_train_epoch_product) ----------------------------------------
    def _epoch_kernel_ready(self, graph_dataset):
        """The lean step's preconditions plus: no Universal factor, ONE Riemannian optimizer whose single parameter group
        holds exactly the embedding's point tensors (no trained scale, no AdamNc).  Then the slices of the epoch need no
        host work between them and the whole loop of train.py:207-226 runs inside one native call."""
        if not self._lean_ready(graph_dataset):
            return False
        st = self._lean
        if 'epoch_ok' not in st:
            from .optim import RiemannianAdam, RiemannianSGD
            emb = self.embedding
            ok = (not any(st['curved']) and not any(st['scale_trained'])
                  and len(self.optimizer) == 1 and isinstance(self.optimizer[0], (RiemannianAdam, RiemannianSGD))
                  and len(self.optimizer[0].param_groups) == 1
                  and len(self.optimizer[0].param_groups[0]['params']) == len(emb.xs)
                  and all(p is x for p, x in zip(self.optimizer[0].param_groups[0]['params'], emb.xs))
                  and all(x.shape[0] == emb.xs[0].shape[0] for x in emb.xs)
                  and not self.optimizer[0].param_groups[0].get('nc', False)
                  and os.environ.get('GM_EPOCH_KERNEL', '1') == '1')
            st['epoch_ok'] = ok
        return st['epoch_ok']

    def _train_epoch_kernel(self, graph_dataset, perm, bs, alpha, epoch):
        import ctypes
        from . import _lib as L
        from .modules import _softplus_value
        with torch._C.DisableTorchFunction():
            st, emb, opt = self._lean, self.embedding, self.optimizer[0]
            xs, mans = list(emb.xs), list(emb.manifolds)
            F, x0 = len(xs), xs[0]
            dev, dtype = x0.device, x0.dtype
            if perm.device != dev or perm.dtype not in (torch.int32, torch.int64) or not perm.is_contiguous():
                perm = perm.to(device=dev, dtype=torch.int64).contiguous()  # drawn on the default device
            n_points = perm.numel()
            max_steps = (n_points + bs - 1) // bs
            group = opt.param_groups[0]
            scales = st['scales']
            sps = [_softplus_value(s) for s in scales] if scales else [1.0] * F
            cfgs, b1s, b2s = [], [], []
            for x, man, gx in zip(xs, mans, st['grads']):
                x.grad = gx
                cfg, b1, b2 = opt._kernel_args(group, x)
                cfg.grassmann_retr_qr = int(getattr(man, 'retr_kind', 'svd') == 'qr')
                cfgs.append(cfg)
                b1s.append(b1)
                b2s.append(b2)
            t = st['targets'].c_struct()
            l = self.objective_fn.loss_spec(epoch=epoch, alpha=alpha).c_struct()
            n_steps = ctypes.c_int64(0)
            acc = torch.zeros(max_steps, 1 + F, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                if F == 1:
                    m = mans[0].spec.c_struct(dtype, dev)
                    rc = L.lib().gm_train_epoch(ctypes.byref(m), ctypes.byref(cfgs[0]), L.ptr(x0.data), L.ptr(x0.grad),
                                                L.ptr(b1s[0]), L.ptr(b2s[0]), x0.shape[0], L.ptr(perm),
                                                int(perm.dtype == torch.int64), n_points, bs, self.drop_last_n,
                                                ctypes.byref(t), ctypes.byref(l), sps[0], L.ptr(acc), max_steps,
                                                ctypes.byref(n_steps), L.stream_ptr(dev))
                    L.check(rc, 'gm_train_epoch')
                else:
                    bmax = min(bs, n_points)
                    ws = torch.empty(F + 1, bmax * (bmax - 1) // 2, dtype=dtype, device=dev)  # d2 per factor + dL/dm
                    vp = ctypes.c_void_p
                    arr = lambda ts: (vp * F)(*[None if t_ is None else t_.data_ptr() for t_ in ts])  # noqa: E731
                    m_arr = (L.Manifold * F)(*[man.spec.c_struct(dtype, dev) for man in mans])
                    o_arr = (L.Optim * F)(*cfgs)
                    rc = L.lib().gm_train_epoch_product(
                        F, m_arr, o_arr, arr([x.data for x in xs]), arr([x.grad for x in xs]), arr(b1s), arr(b2s),
                        x0.shape[0], L.ptr(perm), int(perm.dtype == torch.int64), n_points, bs, self.drop_last_n,
                        ctypes.byref(t), ctypes.byref(l), (ctypes.c_double * F)(*sps), arr([ws[f] for f in range(F)]),
                        L.ptr(ws[F]), L.ptr(acc), max_steps, ctypes.byref(n_steps), L.stream_ptr(dev))
                    L.check(rc, 'gm_train_epoch_product')
            k = int(n_steps.value)
            for x in xs:
                opt._advance(x, k)
            # per-step scalars exactly as the step loop logs them: loss / len(indices), in the embedding's dtype
            losses = acc[:k, 0].to(dtype)
            sizes = torch.tensor([min(bs, n_points - i * bs) for i in range(k)], dtype=dtype, device=dev)
            per_node = losses / sizes
            tag = str(self.objective_fn)
            for i in range(k):
                self.global_step += 1
                self.writer.add_scalar(tag, per_node[i], self.global_step)
            self.writer.flush()
            total_loss = float(losses.double().sum().item()) if k else 0
        logger.debug('epoch %d, train loss %.5f', epoch, total_loss / n_points)
        return total_loss

    def _combine_ranks(self, loss):
        """Sum the per-rank partial gradients and loss (each rank covered a slice of the batch's pairs)."""
        dist = torch.distributed
        for optim in self.optimizer:
            for group in optim.param_groups:
                for p in group['params']:
                    if p.grad is not None:
                        dist.all_reduce(p.grad, group=self.process_group)
        loss = loss.detach().clone()
        dist.all_reduce(loss, group=self.process_group)
        return loss

    def _validate(self, graph_dataset, epoch):
        n = len(graph_dataset)
        total = n * (n - 1) // 2
        streamed = all(m in metrics_mod.STREAMED_METRICS for m in self.metrics) and not self.lazy_metrics \
            and hasattr(self.embedding, 'manifolds')
        values = {}
        if streamed:
            base, extra = divmod(total, self._world)
            lo = self._rank * base + min(self._rank, extra)
            hi = lo + base + (1 if self._rank < extra else 0)
            acc = metrics_mod.validation_moments(self.embedding, graph_dataset, pair_range=(lo, hi))
            if self._world > 1:
                torch.distributed.all_reduce(acc, group=self.process_group)
            mom = metrics_mod.PairMoments(acc.cpu())
            values = {m: mom.metric(m) for m in self.metrics}
        else:
            gpdists = graph_dataset[None].sqrt()
            mpdists = self.embedding.compute_dists(None).sqrt_()
            if self.lazy_metrics:
                mp_np = None
                for name, f in self.lazy_metrics.items():
                    if getattr(f, 'takes_device_tensor', False):  # GPU FastPrecision: no host round trip
                        self.pending_metric_results.append((epoch, name, f(mpdists)))
                        continue
                    if mp_np is None:
                        mp_np = mpdists.cpu().numpy()
                    self.pending_metric_results.append((epoch, name, f(mp_np)))
            gpdists = gpdists.to(mpdists.device)
            values = {m: float(getattr(metrics_mod, m)(mpdists, gpdists)) for m in self.metrics}
        self.embedding.add_stats(self.writer, epoch)
        val_obj = None
        for i, m in enumerate(self.metrics):
            self.writer.add_scalar(m, values[m], epoch)
            if i == self.main_metric_idx:
                val_obj = values[m]
        logger.info('epoch %d, val obj %.5f', epoch, val_obj)
        return val_obj

    def _consume_pending_metric_results(self, wait=False):
        import numpy as np
        while self.pending_metric_idx < len(self.pending_metric_results):
            epoch, name, future = self.pending_metric_results[self.pending_metric_idx]
            if not wait and not future.done():
                break
            value = future.result(None if wait else 0)
            if isinstance(value, float):
                self.writer.add_scalar(name, value, epoch)
            else:  # (means, stds) of a per-layer metric
                means, stds = value
                self.writer.add_scalar('AUC_{}'.format(name), metrics_mod.area_under_curve(means)[0], epoch)
                if epoch % self.save_metrics_every_epochs == 0:
                    np.save(os.path.join(self.save_dir, f'mean_{name}_{epoch}'), means)
                    np.save(os.path.join(self.save_dir, f'std_{name}_{epoch}'), stds)
            self.pending_metric_idx += 1

    def _lr_scheduler_step(self, loss, epoch):
        if self.lr_scheduler:
            for lrs in self.lr_scheduler:
                if isinstance(lrs, torch.optim.lr_scheduler.ReduceLROnPlateau):
                    lrs.step(loss)
                else:
                    lrs.step(epoch)

    def _check_early_break(self):
        if self.min_lr is None:
            return False
        return all(group['lr'] <= self.min_lr + 1e-10 for optim in self.optimizer for group in optim.param_groups)

    def _check_best(self, loss, epoch):
        if loss < self.best_struct['loss']:
            self.best_struct = dict(epoch=epoch, loss=loss, embedding=_Snapshot(self.embedding))

    # ---- on-disk formats (train.py:331-352) ---------------------------------------------------------------------
    def _save(self, epoch):
        torch.save(self.embedding.state_dict(), os.path.join(self.save_dir, f'embedding_{epoch}.pth'))

    def _save_best(self):
        epoch, loss, embedding = (self.best_struct[k] for k in ('epoch', 'loss', 'embedding'))
        torch.save(embedding.state_dict(), os.path.join(self.save_dir, 'best_embedding.pth'))
        with open(os.path.join(self.save_dir, 'best_loss_{}'.format(epoch)), 'w') as f:
            f.write(f'{loss:.6f}')

    def _load(self):
        path = latest_path_by_basename_numeric_order(os.path.join(self.snapshot_path, 'embedding_*.pth'))
        self.embedding.load_state_dict(torch.load(path, map_location=self.embedding.device))


class SineLRScheduler(torch.optim.lr_scheduler._LRScheduler):
    """lr_i(t) = base_i + (eta_max_i - base_i) (1 - cos(pi t / T_max_i)) / 2  (train.py:355-374)."""

    def __init__(self, optimizer, T_max, eta_max=1, last_epoch=-1):
        groups = len(optimizer.param_groups)
        self.T_max = list(T_max) if isinstance(T_max, (list, tuple)) else [T_max] * groups
        self.eta_max = list(eta_max) if isinstance(eta_max, (list, tuple)) else [eta_max] * groups
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return [base + (self.eta_max[i] - base) * (1 - math.cos(math.pi * self.last_epoch / self.T_max[i])) / 2
                for i, base in enumerate(self.base_lrs)]
