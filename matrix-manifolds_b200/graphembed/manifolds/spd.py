"""Symmetric positive definite matrices with the affine-invariant metric or the
symmetric Stein divergence.  API of the reference's graphembed/manifolds/spd.py
(ctor :23-30, dist/pdist :171-181, stein :183-194, rand :201-208, randvec
:210-221, to_vec/from_vec :66-80); arithmetic in csrc/gm_manifolds.cuh (pairs)
and csrc/gm_pointops.cuh (SpdPt)."""
import math

import torch

from .. import _lib as L
from .. import _ops
from .base import Manifold, _like


class SymmetricPositiveDefinite(Manifold):
    _dist_keep_axes = 1

    def __init__(self, n, *, fast_symeig=True, fast_chol=True, use_stein_div=False, wmin=1e-8, wmax=1e8):
        self.n, self.wmin, self.wmax = n, wmin, wmax
        self.use_stein_div = use_stein_div
        flags = 0
        if fast_symeig and n in (2, 3) and not use_stein_div:
            flags |= L.GM_FAST_EIG  # closed-form eigenvalues incl. the reference's eps terms
        if fast_chol and n == 2:
            flags |= L.GM_FAST_CHOL
        kind = L.GM_SPD_STEIN if use_stein_div else L.GM_SPD_AI
        super().__init__(_ops.ManifoldSpec(kind, n, 0, flags, wmin, wmax, point_shape=(n, n)))
        if use_stein_div:  # same aliases the reference installs
            self.stein_div, self.stein_pdiv = self.dist, self.pdist

    # Vec(.) of Pennec et al. Sec. 3.5: sqrt(2)-weighted upper triangle
    @staticmethod
    def to_vec(x):
        n = x.shape[-1]
        i, j = torch.triu_indices(n, n, device=x.device)
        w = torch.where(i == j, 1.0, math.sqrt(2)).to(x.dtype)
        return (x[..., i, j] * w).reshape(-1) if x.ndim == 2 else x[..., i, j] * w

    @staticmethod
    def from_vec(v):
        dimv = v.shape[-1]
        n = (math.isqrt(1 + 8 * dimv) - 1) // 2
        i, j = torch.triu_indices(n, n, device=v.device)
        vals = torch.where(i == j, v, v / math.sqrt(2))
        x = v.new_zeros(v.shape[:-1] + (n, n))
        x[..., i, j] = vals
        x[..., j, i] = vals
        return x

    @property
    def ndim(self):
        return 2

    @property
    def dim(self):
        return self.n * (self.n + 1) // 2

    def zero(self, *shape, out=None):
        return torch.eye(self.n, **_like(out)).repeat(*shape, 1, 1)

    def inner(self, x, u, v, keepdim=False):
        assert not (x.requires_grad or u.requires_grad or v.requires_grad)
        return super().inner(x, u, v, keepdim)

    def rand(self, *shape, out=None, ir=1e-1):
        eye = self.zero(*shape, out=out)
        u = torch.randn(*shape, self.dim, dtype=eye.dtype, device=eye.device)
        u.div_(u.norm(dim=-1, keepdim=True)).mul_(ir)
        return self.exp(eye, self.from_vec(u))

    def randvec(self, x, norm=1):
        u = torch.randn(x.shape[:-2] + (self.dim,), dtype=x.dtype, device=x.device)
        u.div_(u.norm(dim=-1, keepdim=True)).mul_(norm)
        root = _ops.point_op(self._spec, L.GM_OP_SPD_SQRTM, x)
        # x^{1/2} U x^{1/2}: parallel transport of U from the identity to x
        return root @ self.from_vec(u) @ root

    def seccurv(self, x, u, v):
        raise NotImplementedError('sectional-curvature sampling is analysis tooling, outside the training hot path')

    def __str__(self):
        return 'Manifold of {n}x{n} positive definite matrices'.format(n=self.n)
