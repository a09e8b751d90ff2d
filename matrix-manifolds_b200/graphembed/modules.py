"""Model glue: ManifoldParameter, ManifoldEmbedding, BatchedObjective -- the
interface of the reference's graphembed/modules.py:9-105.

What is different underneath: `compute_dists(indices)` never materialises
x[indices] (the gather is fused into the pair kernel), and `BatchedObjective`
runs distance + loss + gradient as ONE kernel per step for a single manifold
(gm_pairs_loss_fused) or distance / loss / gradient kernels per factor for a
product (gm_pairs_dist2 + gm_product_loss + gm_pairs_grad).
"""
import abc

import torch
from torch.nn.functional import softplus

from . import _ops


class ManifoldParameter(torch.nn.Parameter):

    def __new__(cls, data=None, manifold=None, requires_grad=True):
        if data is None:
            data = torch.Tensor()
        instance = torch.Tensor._make_subclass(cls, data, requires_grad)
        instance.manifold = manifold
        return instance

    def proj_(self):
        self.manifold.projx(self, inplace=True)

    def __repr__(self):
        return 'Parameter on {} containing:\n'.format(self.manifold) + torch.Tensor.__repr__(self)


class EmbeddingBase(torch.nn.Module):

    @property
    @abc.abstractmethod
    def device(self):
        pass

    @property
    @abc.abstractmethod
    def curvature_params(self):
        pass

    def burnin(self, value=True):
        for p in self.curvature_params:  # no curvature learning during burn-in
            p.requires_grad_(not value)


class ManifoldEmbedding(EmbeddingBase):
    """n points on each of the given manifolds; the product distance is
    sum_f softplus(scale_f) * dist_f^2 (modules.py:84-88)."""

    fused_pair_kernels = True  # BatchedObjective may run distance + loss + gradient as fused kernels

    def __init__(self, n, manifolds, device=None, dtype=None):
        super().__init__()
        self.n = n
        self.n_components = len(manifolds)
        self.manifolds = manifolds
        hint = None
        if device is not None or dtype is not None:
            hint = torch.empty(0, device=device, dtype=dtype or torch.get_default_dtype())
        self.xs = torch.nn.ParameterList([
            ManifoldParameter(data=(m.rand(n) if hint is None else m.rand(n, out=hint)).contiguous(), manifold=m)
            for m in manifolds
        ])
        # softplus(0.5) ~ 0.97
        self.scales = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.tensor(0.5, device=self.xs[0].device, dtype=self.xs[0].dtype))
             for _ in manifolds])

    @property
    def device(self):
        return self.xs[0].device

    @property
    def curvature_params(self):
        return self.scales

    @torch.no_grad()
    def perturb(self, norm):
        for x, man in zip(self.xs, self.manifolds):
            x.copy_(man.retr(x, man.randvec(x, norm)))

    @torch.no_grad()
    def stabilize(self):
        for x in self.xs:
            x.proj_()

    @torch.no_grad()
    def add_stats(self, writer, epoch):
        for i in range(self.n_components):
            writer.add_scalar(f'scale{i}', self.scales[i], epoch)

    def compute_dists(self, i=None):
        terms = []
        for x, s, man in zip(self.xs, self.scales, self.manifolds):
            d2 = man.pdist(x, squared=True) if i is None else man.batch_pdist2(x, i)
            terms.append(softplus(s) * d2)
        return sum(terms)

    def __len__(self):
        return self.n


def _softplus_value(scale):
    """float(softplus(scale)) without a device->host read per step: the value is cached on the parameter until it is
    modified in place (tensor._version changes: torch optimizers) or stepped by RiemannianAdam / RiemannianSGD, whose
    kernels write through the raw pointer and drop the cache themselves (optim/_common.py::fused_step)."""
    key = (scale._version, scale.data_ptr())  # in-place updates bump the version, `.data = ...` moves the storage
    hit = getattr(scale, '_gm_softplus', None)
    if hit is not None and hit[0] == key:
        return hit[1]
    value = float(softplus(scale.detach()))
    scale._gm_softplus = (key, value)
    return value


def _curvature_of(manifold):
    """The curvature tensor of a Universal factor (it receives a gradient), None for every other manifold."""
    get_c = getattr(manifold, 'get_c', None)
    return None if get_c is None else get_c()


class _FusedObjective(torch.autograd.Function):
    """loss(targets(pairs), sum_f sp_f * dist_f^2(pairs)), sp_f = softplus(s_f) (ManifoldEmbedding) or 1
    (products.Embedding), with the gradients w.r.t. every x_f, s_f and Universal curvature c_f produced in the
    forward pass; backward only rescales them.  params = (*xs, *scales [if has_scales], *curvatures-or-None)."""

    @staticmethod
    def forward(ctx, pairs, targets, loss_spec, manifolds, n_factors, has_scales, *params):
        F = n_factors
        xs = params[:F]
        scales = params[F:2 * F] if has_scales else ()
        cs = params[len(params) - F:]
        sps = [_softplus_value(s) for s in scales] if has_scales else [1.0] * F
        grads = [torch.zeros_like(x, memory_format=torch.contiguous_format) for x in xs]
        c_needs = ctx.needs_input_grad[len(ctx.needs_input_grad) - F:]
        cgrads = [torch.zeros(1, dtype=torch.float64, device=xs[0].device) if (c is not None and need) else None
                  for c, need in zip(cs, c_needs)]
        cvals = [None if c is None else c.detach() for c in cs]
        if F == 1:
            acc, _ = _ops.pairs_loss_fused(manifolds[0].spec, xs[0].detach(), pairs, targets, loss_spec, sps[0],
                                           grads[0], c=cvals[0], c_grad=cgrads[0])
        else:
            d2s = [_ops.pairs_dist2(m.spec, x.detach(), x.detach(), pairs, c=c) for m, x, c in zip(manifolds, xs, cvals)]
            acc, g = _ops.product_loss(d2s, sps, targets, loss_spec, pairs=pairs)
            for m, x, gx, sp, c, cg in zip(manifolds, xs, grads, sps, cvals, cgrads):
                _ops.pairs_grad(m.spec, x.detach(), x.detach(), pairs, g, gx, gx, coef=sp, c=c, c_grad=cg)
        dscale = [acc[1 + f] * torch.sigmoid(scales[f].detach().double()) for f in range(len(scales))]
        ctx.save_for_backward(*grads, *dscale, *[cg for cg in cgrads if cg is not None])
        ctx.n_factors, ctx.n_scales = F, len(scales)
        ctx.scale_dtypes = [s.dtype for s in scales]
        ctx.c_meta = [None if cg is None else (c.dtype, c.shape) for c, cg in zip(cs, cgrads)]
        return acc[0].to(xs[0].dtype)

    @staticmethod
    def backward(ctx, upstream):
        saved = ctx.saved_tensors
        F, S = ctx.n_factors, ctx.n_scales
        gx = [g * upstream for g in saved[:F]]
        gs = [(d * upstream).to(dt) for d, dt in zip(saved[F:F + S], ctx.scale_dtypes)]
        rest = list(saved[F + S:])
        gc = [None if meta is None else (rest.pop(0) * upstream).to(meta[0]).reshape(meta[1]) for meta in ctx.c_meta]
        return (None, None, None, None, None, None, *gx, *gs, *gc)


class BatchedObjective(torch.nn.Module):

    def __init__(self, objective_fn, dataset, embedding, shard=None):
        """shard = (rank, world): evaluate only this rank's contiguous slice of every batch's pair triangle; the
        caller sums losses and gradients over the ranks (TrainingEngine._combine_ranks)."""
        super().__init__()
        self.objective_fn = objective_fn
        self.dataset = dataset
        self.embedding = embedding
        self.shard = shard

    def forward(self, indices, *args, **kwargs):
        emb = self.embedding
        spec_fn = getattr(self.objective_fn, 'loss_spec', None)
        loss_spec = spec_fn(**kwargs) if (spec_fn is not None and not args) else None
        fusable = (loss_spec is not None and getattr(emb, 'fused_pair_kernels', False)
                   and getattr(self.dataset, 'pdists', None) is not None
                   and self.dataset.pdists.device == emb.device and self.dataset.pdists.dtype == emb.xs[0].dtype)
        if not fusable:
            if self.shard is not None:
                raise RuntimeError('pair-sharded training needs an objective with a fused loss_spec (QuotientLoss, '
                                   'StressLoss) and GPU-resident targets')
            return self.objective_fn(self.dataset[indices].to(emb.device), emb.compute_dists(indices), *args, **kwargs)
        if indices is None:
            pairs = _ops.PairSet.triu(emb.n)
        else:
            pairs = _ops.PairSet.triu(len(indices), indices, emb.device)
        lo = 0
        if self.shard is not None:
            full_k0 = pairs.k0
            pairs = pairs.slice(*self.shard)
            lo = pairs.k0 - full_k0
        n_factors = len(emb.xs)
        # the target gather pdists[idx][:, idx] -> triu (data/dataset.py:19-27) is fused into the pair kernel (one
        # factor) or into the product-loss kernel (several): both index the dense matrix with the pair's node ids
        targets = _ops.TargetSpec.dense(self.dataset.pdists)
        scales = tuple(getattr(emb, 'scales', ()))  # products.Embedding has none: plain sum of squared distances
        curvatures = [_curvature_of(m) for m in emb.manifolds]
        return _FusedObjective.apply(pairs, targets, loss_spec, list(emb.manifolds), n_factors, bool(scales), *emb.xs,
                                     *scales, *curvatures)
