"""kappa-stereographic ("Universal") manifold with a learnable curvature parameter: the Poincare ball for c > 0, the
stereographic projection of the sphere for c < 0 (reference: graphembed/manifolds/universal.py:11-97 over
graphembed/manifolds/impl/math.py).  Arithmetic: csrc/gm_manifolds.cuh (Kappa, VecMan<VEC_UNIVERSAL>: distance,
gradients w.r.t. both points AND the curvature) and csrc/gm_pointops.cuh (UniversalPt).

The curvature is handed to the kernels as a device scalar, so a curvature optimizer can update `c` between steps
without a host synchronisation.  Reference quirks kept on purpose:
  * `norm` evaluates the conformal factor with c = 1.0, because universal.py:44-48 calls math.norm without `c`;
  * `projx(x)` returns x itself -- projected only when `inplace=True` (universal.py:53-57);
  * `dist` clamps the value AFTER the optional squaring (universal.py:76-81).
"""
import torch
from torch.nn.functional import softplus

from .. import _lib as L
from .. import _ops
from ..utils import EPS
from .base import Manifold, _like


class Universal(Manifold, torch.nn.Module):

    def __init__(self, n, c_init=0.01, c_min=0.001, keep_sign_fixed=False, device=None, dtype=None):
        torch.nn.Module.__init__(self)
        Manifold.__init__(self, _ops.ManifoldSpec(L.GM_UNIVERSAL, n, point_shape=(n,), c_source=self.get_c))
        self.n = n
        self.c_min = c_min
        self.sign = None if not keep_sign_fixed else 1 if c_init > 0 else -1
        self.c = torch.nn.Parameter(torch.tensor([c_init], dtype=dtype or torch.get_default_dtype(), device=device))

    @property
    def ndim(self):
        return 1

    @property
    def dim(self):
        return self.n

    # ---- curvature (universal.py:28-38) --------------------------------------------------------------------------
    def get_c(self):
        if self.sign:
            return self.sign * (self.c_min + softplus(self.c))
        return self.c.sign() * self.c_min + self.c

    def get_K(self):
        return -self.get_c()

    def get_R(self):
        return 1.0 / torch.sqrt(torch.abs(self.get_c()))

    # ---- points -------------------------------------------------------------------------------------------------
    def zero(self, *shape, out=None):
        return torch.zeros(*shape, self.n, **self._hint(out))

    def zero_vec(self, *shape, out=None):
        return torch.zeros(*shape, self.n, **self._hint(out))

    def _hint(self, out):
        return _like(out) if out is not None else dict(dtype=self.c.dtype, device=self.c.device)

    def proju(self, x, u, inplace=False):
        return u

    def inner(self, x, u, v, keepdim=False):
        r = _ops.point_op(self._spec, L.GM_OP_INNER, x, u, v, scalar=True)  # lambda_x^2 <u, v> per point
        if keepdim:
            return r.unsqueeze(-1)
        # reference quirk (impl/math.py:225-228): the conformal factor is always computed with keepdim=True, so with
        # keepdim=False the (N, 1) factor broadcasts against the (N,) dot products into an (N, N) matrix
        e = torch.zeros_like(x)
        e[..., 0] = 1
        lam2 = _ops.point_op(self._spec, L.GM_OP_INNER, x, e, e, scalar=True)
        uv = _ops.point_op(_ops.ManifoldSpec(L.GM_EUCLIDEAN, self.n, point_shape=(self.n,)), L.GM_OP_INNER, x, u, v,
                           scalar=True)
        return lam2.unsqueeze(-1) * uv

    def norm(self, x, u, squared=False, keepdim=False):
        r = _ops.point_op(self._spec, L.GM_OP_NORM2, x, u, scalar=True)
        if not squared:
            r = r.sqrt()
        return r.unsqueeze(-1) if keepdim else r

    def projx(self, x, inplace=False):
        if inplace:
            x.copy_(_ops.point_op(self._spec, L.GM_OP_PROJX, x))
        return x

    def exp(self, x, u, project=True):
        if not project:
            raise NotImplementedError('Universal.exp(project=False) is not on the training path')
        return super().exp(x, u)

    # ---- distances: differentiable w.r.t. the points and the curvature ---------------------------------------------
    def dist(self, x, y, squared=False, keepdim=False):
        if squared:
            d = _ops.dist2_elementwise(self._spec, x, y, c=self.get_c())
        else:  # clamp d (not d^2) at EPS: value-only floor of EPS^2 on the squared distance, then the root
            d = _ops.dist2_elementwise(self._spec, x, y, c=self.get_c(), wmin=EPS[x.dtype]**2).sqrt()
        return d.unsqueeze(-1) if keepdim else d

    def pdist(self, x, squared=False):
        assert x.ndim == self.ndim + 1
        return self._pairs(x, _ops.PairSet.triu(x.shape[0]), squared)

    def pair_dist2(self, x, idx_i, idx_j):
        return self._pairs(x, _ops.PairSet.from_lists(idx_i, idx_j, x.device), True)

    def batch_pdist2(self, x, nodes):
        return self._pairs(x, _ops.PairSet.triu(len(nodes), nodes, x.device), True)

    def _pairs(self, x, pairs, squared):
        if squared:
            return _ops.dist2_indexed(self._spec, x, pairs, c=self.get_c())
        return _ops.dist2_indexed(self._spec, x, pairs, c=self.get_c(), wmin=EPS[x.dtype]**2).sqrt()

    # ---- sampling (universal.py:86-92) ------------------------------------------------------------------------------
    @torch.no_grad()
    def rand(self, *shape, out=None, ir=1e-2):
        x = torch.empty(*shape, self.n, **self._hint(out)).uniform_(-ir, ir)
        return self.projx(x, inplace=True)

    def randvec(self, x, norm=1):
        raise NotImplementedError

    def __str__(self):
        return f'Universal {self.n}-dimensional manifold'
