"""Abstract manifold interface -- same surface as the reference's
graphembed/manifolds/base.py:7-81 (dist/pdist/exp/log/retr/proju/projx/
egrad2rgrad/inner/norm/transp/rand/randvec/zero/zero_vec, properties ndim/dim),
but every numerical method dispatches to the sm_100a kernels in libgm_b200.so
through graphembed._ops; subclasses only describe themselves with a ManifoldSpec
and provide the random initialisers.
"""
import abc

import torch

from .. import _lib as L
from .. import _ops


def _like(out):
    """dtype/device from the reference-style `out=` argument (a tensor used only as a hint)."""
    if out is None:
        return {}
    return dict(dtype=out.dtype, device=out.device)


class Manifold(abc.ABC):
    #: number of trailing `1` axes `dist(..., keepdim=True)` adds (the reference is not uniform here)
    _dist_keep_axes = 0

    def __init__(self, spec):
        self._spec = spec

    # ---- structure ----------------------------------------------------------
    @property
    def spec(self):
        return self._spec

    @property
    @abc.abstractmethod
    def ndim(self):
        """Number of trailing axes that make up one point."""

    @property
    @abc.abstractmethod
    def dim(self):
        """Intrinsic dimension."""

    @abc.abstractmethod
    def zero(self, *shape, out=None):
        """The reference point ("origin") repeated over `shape`."""

    def zero_vec(self, *shape, out=None):
        return torch.zeros(*shape, *self._spec.point_shape, **_like(out))

    def _keep(self, t, x, axes=None):
        axes = self.ndim if axes is None else axes
        return t.reshape(*t.shape, *([1] * axes))

    # ---- metric ---------------------------------------------------------------
    def inner(self, x, u, v, keepdim=False):
        r = _ops.point_op(self._spec, L.GM_OP_INNER, self._anchor(x, u), u, v, scalar=True)
        return self._keep(r, u) if keepdim else r

    def norm(self, x, u, squared=False, keepdim=False):
        r = _ops.point_op(self._spec, L.GM_OP_NORM2, self._anchor(x, u), u, scalar=True)
        if not squared:
            r = r.sqrt()
        return self._keep(r, u) if keepdim else r

    def _anchor(self, x, u):
        # Sphere/Euclidean/Grassmann call inner()/norm() with x=None in the reference
        return u if x is None else x

    # ---- projections ------------------------------------------------------------
    def proju(self, x, u, inplace=False):
        r = _ops.point_op(self._spec, L.GM_OP_PROJU, x, u)
        if inplace:
            u.copy_(r)
            return u
        return r

    def projx(self, x, inplace=False):
        r = _ops.point_op(self._spec, L.GM_OP_PROJX, x)
        if inplace:
            x.copy_(r)
            return x
        return r

    def egrad2rgrad(self, x, u):
        return _ops.point_op(self._spec, L.GM_OP_EGRAD2RGRAD, x, u)

    # ---- geodesics ----------------------------------------------------------------
    def exp(self, x, u):
        return _ops.point_op(self._spec, L.GM_OP_EXP, x, u)

    def retr(self, x, u):
        return _ops.point_op(self._spec, L.GM_OP_RETR, x, u)

    def log(self, x, y):
        return _ops.point_op(self._spec, L.GM_OP_LOG, x, y)

    def transp(self, x, y, u):
        return _ops.point_op(self._spec, L.GM_OP_TRANSP, x, y, u)

    # ---- distances (differentiable) ---------------------------------------------------
    def dist(self, x, y, squared=False, keepdim=False):
        d2 = _ops.dist2_elementwise(self._spec, x, y)
        d = d2 if squared else d2.sqrt()
        return self._keep(d, x, self._dist_keep_axes) if keepdim else d

    def pdist(self, x, squared=False):
        """Condensed distances between all a<b rows of x (row-major upper triangle)."""
        assert x.ndim == self.ndim + 1
        d2 = _ops.dist2_indexed(self._spec, x, _ops.PairSet.triu(x.shape[0]))
        return d2 if squared else d2.sqrt()

    def pair_dist2(self, x, idx_i, idx_j):
        """Squared distances dist2(x[idx_i[k]], x[idx_j[k]]) with the gather and the gradient scatter-add fused."""
        return _ops.dist2_indexed(self._spec, x, _ops.PairSet.from_lists(idx_i, idx_j, x.device))

    def batch_pdist2(self, x, nodes):
        """pdist(x[nodes], squared=True) without materialising x[nodes]."""
        return _ops.dist2_indexed(self._spec, x, _ops.PairSet.triu(len(nodes), nodes, x.device))

    # ---- sampling -------------------------------------------------------------------------
    @abc.abstractmethod
    def rand(self, *shape, out=None):
        pass

    def rand_uniform(self, *shape, out=None):
        raise NotImplementedError

    @abc.abstractmethod
    def randvec(self, x, norm=1):
        pass

    @abc.abstractmethod
    def __str__(self):
        pass
