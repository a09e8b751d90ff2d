"""Flat space of tensors of a given shape (graphembed/manifolds/euclidean.py)."""
import numpy as np
import torch

from .. import _lib as L
from .. import _ops
from .base import Manifold, _like
from .sphere import _shape_name


class Euclidean(Manifold):

    def __init__(self, *shape):
        if len(shape) == 0:
            raise ValueError('Need shape parameters.')
        self.shape = shape
        self.dims = tuple(range(-len(shape), 0))
        self._name = _shape_name('Euclidean', shape)
        self._dist_keep_axes = len(shape)
        super().__init__(_ops.ManifoldSpec(L.GM_EUCLIDEAN, int(np.prod(shape)), point_shape=shape))

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dim(self):
        return int(np.prod(self.shape))

    def zero(self, *shape, out=None):
        return torch.zeros(*shape, *self.shape, **_like(out))

    def rand(self, *shape, out=None, ir=1e-2):
        return torch.empty(*shape, *self.shape, **_like(out)).uniform_(-ir, ir)

    def randvec(self, x, norm=1):
        u = torch.randn(x.shape, dtype=x.dtype, device=x.device)
        return u.div_(u.norm(dim=self.dims, keepdim=True)).mul_(norm)

    def __str__(self):
        return self._name
