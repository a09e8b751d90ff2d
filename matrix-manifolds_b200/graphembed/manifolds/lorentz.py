"""Hyperboloid (Lorentz) model of hyperbolic space; `n` is the AMBIENT dimension
as in the reference (graphembed/manifolds/lorentz.py:11-26).  Arithmetic:
csrc/gm_manifolds.cuh (VecMan<VEC_LORENTZ>) and csrc/gm_pointops.cuh (LorentzPt)."""
import torch

from .. import _lib as L
from .. import _ops
from .base import Manifold, _like
from .sphere import Sphere


class Lorentz(Manifold):
    _dist_keep_axes = 0  # the reference's Lorentz.dist ignores keepdim (lorentz.py:72-77)

    def __init__(self, n):
        self.n = n
        self.sphere = Sphere(n - 1)
        super().__init__(_ops.ManifoldSpec(L.GM_LORENTZ, n, point_shape=(n,)))

    @staticmethod
    def to_poincare_ball(x):
        return x[..., 1:] / (x[..., :1] + 1)

    @property
    def ndim(self):
        return 1

    @property
    def dim(self):
        return self.n - 1

    def zero(self, *shape, out=None):
        x = torch.zeros(*shape, self.n, **_like(out))
        x[..., 0] = 1
        return x

    def egrad2rgrad(self, x, u, inplace=False):
        r = super().egrad2rgrad(x, u)
        if inplace:
            u.copy_(r)
            return u
        return r

    def rand(self, *shape, out=None, ir=1e-2):
        x = torch.empty(*shape, self.n, **_like(out)).uniform_(-ir, ir)
        return self.projx(x, inplace=True)

    def randvec(self, x, norm=1):
        shape = x.shape[:-1]
        dirs = self.sphere.rand_uniform(*shape, out=x)
        vs = torch.cat([torch.zeros(*shape, 1, dtype=x.dtype, device=x.device), dirs], dim=-1).mul_(norm)
        return self.transp(self.zero(*shape, out=x), x, vs)

    def __str__(self):
        return 'Lorentzian space of dimension {}'.format(self.n)
