"""Unit sphere of tensors of a given shape (graphembed/manifolds/sphere.py).
Arithmetic: VecMan<VEC_SPHERE> / SpherePt on the flattened point."""
import numpy as np
import torch

from .. import _lib as L
from .. import _ops
from .base import Manifold, _like


def _shape_name(kind, shape):
    if len(shape) == 1:
        return '{} manifold of {}-vectors'.format(kind, *shape)
    if len(shape) == 2:
        return '{} manifold of {}x{} matrices'.format(kind, *shape)
    return '{} manifold of shape {} tensors'.format(kind, shape)


class Sphere(Manifold):

    def __init__(self, *shape):
        if len(shape) == 0:
            raise ValueError('Need shape parameters.')
        self.shape = shape
        self.dims = tuple(range(-len(shape), 0))
        self._name = _shape_name('Sphere', shape)
        self._dist_keep_axes = len(shape)
        super().__init__(_ops.ManifoldSpec(L.GM_SPHERE, int(np.prod(shape)), point_shape=shape))

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dim(self):
        return int(np.prod(self.shape)) - 1

    def zero(self, *shape, out=None):
        x = torch.zeros(*shape, int(np.prod(self.shape)), **_like(out))
        x[..., 0] = -1
        return x.reshape(*shape, *self.shape)

    def rand(self, *shape, out=None, ir=1e-2):
        x = self.zero(*shape, out=out)
        return self.retr(x, self.randvec(x, norm=ir))

    def rand_uniform(self, *shape, out=None):
        return self.projx(torch.randn(*shape, *self.shape, **_like(out)), inplace=True)

    def rand_ball(self, *shape, out=None):
        xs = self.rand_uniform(*shape, out=out)
        rs = torch.rand(*shape, dtype=xs.dtype, device=xs.device).pow_(1 / (self.dim + 1))
        return xs.mul_(rs.reshape(*shape, *((1,) * len(self.shape))))

    def randvec(self, x, norm=1):
        u = self.proju(x, torch.randn(x.shape, dtype=x.dtype, device=x.device))
        return u.div_(u.norm(dim=self.dims, keepdim=True)).mul_(norm)

    def __str__(self):
        return self._name
