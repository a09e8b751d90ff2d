"""Grassmannian Gr(n, p) of p-dimensional subspaces of R^n, points are n x p
matrices with orthonormal columns (graphembed/manifolds/grassmann.py).
Arithmetic: GrassmannCore (pairs) and GrassmannPt (point ops)."""
import torch

from .. import _lib as L
from .. import _ops
from .base import Manifold, _like


class Grassmann(Manifold):
    _dist_keep_axes = 1

    def __init__(self, n, p, retr='svd', requires_grad=True):
        if retr not in ('svd', 'qr'):
            raise ValueError('Unknown retraction type {}'.format(retr))
        self.n, self.p, self.requires_grad = n, p, requires_grad
        self.retr_kind = retr
        flags = L.GM_FAST_SVD if p == 2 else 0  # closed-form 2x2 singular values, as the reference (grassmann.py:27-30)
        super().__init__(_ops.ManifoldSpec(L.GM_GRASSMANN, n, p, flags, point_shape=(n, p)))

    @property
    def ndim(self):
        return 2

    @property
    def dim(self):
        return self.p * (self.n - self.p)

    def zero(self, *shape, out=None):
        return torch.eye(self.n, self.p, **_like(out)).repeat(*shape, 1, 1)

    def retr(self, x, u):
        op = L.GM_OP_RETR_QR if self.retr_kind == 'qr' else L.GM_OP_RETR
        return _ops.point_op(self._spec, op, x, u)

    def retr_qr_(self, x, u):
        return _ops.point_op(self._spec, L.GM_OP_RETR_QR, x, u)

    def retr_svd_(self, x, u):
        return _ops.point_op(self._spec, L.GM_OP_RETR, x, u)

    def rand(self, *shape, out=None, ir=1e-2):
        x = self.zero(*shape, out=out)
        return self.exp(x, self.randvec(x, norm=ir))

    def rand_uniform(self, *shape, out=None):
        return self.projx(torch.randn(*shape, self.n, self.p, **_like(out)), inplace=True)

    def randvec(self, x, norm):
        u = self.proju(x, torch.randn(x.shape, dtype=x.dtype, device=x.device))
        return u.div_(u.norm(dim=(-2, -1), keepdim=True)).mul_(norm)

    def __str__(self):
        return 'Grassmann manifold of {}x{} matrices'.format(self.n, self.p)
