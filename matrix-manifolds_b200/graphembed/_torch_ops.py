"""torch custom-op registration of the hot-path kernels: `torch.ops.graphembed_b200.*`.

north_star asks for "a thin C-ABI torch custom-op layer": the ops below are registered with torch.library (schema +
CUDA implementation only -- there is no CPU kernel, so calling them with CPU tensors raises NotImplementedError from the
dispatcher: no fallback).  The transport underneath stays the ctypes binding of libgm_b200.so (graphembed._lib); what
the registration adds is the dispatcher-visible, schema-checked, torch.ops-addressable surface that tools expecting
torch ops (profilers, torch.library.opcheck, export) see.

  graphembed_b200::pair_dist2      squared manifold distance of the pairs (x[I[k]], x[J[k]])        (gm_pairs_dist2)
  graphembed_b200::pairs_loss_fused distance + loss + gradient + scatter-add of one pair batch, in place (gm_pairs_loss_fused)
  graphembed_b200::optim_step      fused in-place Riemannian optimizer update                          (gm_optim_step)
  graphembed_b200::bfs_levels      multi-source BFS hop counts                                         (gm_bfs_multi_source)

Manifolds are passed as plain integers / floats (kind, n, p, flags, wmin, wmax = the fields of gm_manifold_t), so the
schemas contain only types the dispatcher knows.
"""
import torch

from . import _lib as L
from . import _ops

NS = 'graphembed_b200'
_lib = torch.library.Library(NS, 'DEF')

_lib.define('pair_dist2(Tensor x, Tensor idx_i, Tensor idx_j, int kind, int n, int p, int flags, float wmin, '
            'float wmax) -> Tensor')
_lib.define('pairs_loss_fused(Tensor x, Tensor idx_i, Tensor idx_j, Tensor? hops, int kind, int n, int p, int flags, '
            'float wmin, float wmax, int loss_kind, bool inc_l1, bool inc_l2, float alpha, float eps, float max_hops_sq, '
            'float scale_sp, Tensor(a!) grad, Tensor(b!) acc, int segments=0) -> ()')
_lib.define('optim_step(Tensor(a!) x, Tensor grad, Tensor(b!)? buf1, Tensor(c!)? buf2, int kind, int n, int p, int flags, '
            'float wmin, float wmax, int opt_kind, bool exact, bool has_clip, int step, bool has_momentum, '
            'bool first_step, bool retr_qr, bool zero_grad, float lr, float beta1, float beta2, float momentum, '
            'float dampening, float max_grad_norm) -> ()')
_lib.define('bfs_levels(Tensor rowptr, Tensor colidx, Tensor sources) -> Tensor')


def _spec(kind, n, p, flags, wmin, wmax, x):
    if kind == L.GM_UNIVERSAL:
        raise RuntimeError('the Universal manifold carries a curvature tensor: use the graphembed API, not the flat ops')
    if kind in (L.GM_SPD_AI, L.GM_SPD_STEIN):
        shape = (n, n)
    elif kind == L.GM_GRASSMANN:
        shape = (n, p)
    else:
        shape = (n,)
    return _ops.ManifoldSpec(kind, n, p=p, flags=flags, wmin=wmin, wmax=wmax, point_shape=shape)


def _pair_dist2(x, idx_i, idx_j, kind, n, p, flags, wmin, wmax):
    pairs = _ops.PairSet.from_lists(idx_i, idx_j, x.device)
    return _ops.pairs_dist2(_spec(kind, n, p, flags, wmin, wmax, x), x, x, pairs)


def _pairs_loss_fused(x, idx_i, idx_j, hops, kind, n, p, flags, wmin, wmax, loss_kind, inc_l1, inc_l2, alpha, eps,
                      max_hops_sq, scale_sp, grad, acc, segments=0):
    pairs = _ops.PairSet.from_lists(idx_i, idx_j, x.device, segments=segments)
    tg = _ops.TargetSpec.hops_packed(max_hops_sq) if hops is None else _ops.TargetSpec.hops(hops, max_hops_sq)
    loss = _ops.LossSpec(loss_kind, inc_l1, inc_l2, alpha=alpha, eps=eps)
    _ops.pairs_loss_fused(_spec(kind, n, p, flags, wmin, wmax, x), x, pairs, tg, loss, scale_sp, grad, acc)


def _optim_step(x, grad, buf1, buf2, kind, n, p, flags, wmin, wmax, opt_kind, exact, has_clip, step, has_momentum,
                first_step, retr_qr, zero_grad, lr, beta1, beta2, momentum, dampening, max_grad_norm):
    cfg = L.Optim(kind=opt_kind, exact=int(exact), has_clip=int(has_clip), step=step, has_momentum=int(has_momentum),
                  first_step=int(first_step), grassmann_retr_qr=int(retr_qr), zero_grad=int(zero_grad), lr=lr,
                  beta1=beta1, beta2=beta2,
                  momentum=momentum, dampening=dampening, max_grad_norm=max_grad_norm, eps=1e-8)
    _ops.optim_step(_spec(kind, n, p, flags, wmin, wmax, x), cfg, x, grad, buf1, buf2)


def _bfs_levels(rowptr, colidx, sources):
    from .data.graph import bfs_levels
    return bfs_levels(rowptr, colidx, sources=sources, device=rowptr.device)


_lib.impl('pair_dist2', _pair_dist2, 'CUDA')
_lib.impl('pairs_loss_fused', _pairs_loss_fused, 'CUDA')
_lib.impl('optim_step', _optim_step, 'CUDA')
_lib.impl('bfs_levels', _bfs_levels, 'CUDA')


def manifold_args(spec):
    """(kind, n, p, flags, wmin, wmax) of a ManifoldSpec, in the order the op schemas take them."""
    return (spec.kind, spec.n, spec.p, spec.flags, float(spec.wmin), float(spec.wmax))


def optim_step(spec, cfg, x, grad, buf1=None, buf2=None):
    """_ops.optim_step through the registered op (what RiemannianAdam / RiemannianSGD call for every parameter)."""
    torch.ops.graphembed_b200.optim_step(
        x, grad, buf1, buf2, *manifold_args(spec), cfg.kind, bool(cfg.exact), bool(cfg.has_clip), cfg.step,
        bool(cfg.has_momentum), bool(cfg.first_step), bool(cfg.grassmann_retr_qr), bool(cfg.zero_grad), cfg.lr,
        cfg.beta1, cfg.beta2, cfg.momentum, cfg.dampening, cfg.max_grad_norm)
