"""Product of Universal (kappa-stereographic) factors with learnable curvatures -- the interface of the reference's
graphembed/products/embedding.py:8-61.  The squared product distance is the plain sum of the factors' squared
distances (no learnable scales, unlike modules.ManifoldEmbedding)."""
import torch

from ..manifolds import Universal
from ..modules import EmbeddingBase, ManifoldParameter


class Embedding(EmbeddingBase):

    fused_pair_kernels = True  # BatchedObjective: distance + loss + point/curvature gradients as fused kernels

    def __init__(self, n, ds, r_max=5.0, device=None, dtype=None, **kwargs):
        super().__init__()
        self.n = n
        self.ds = ds
        self.r_max = r_max
        self.manifolds = torch.nn.ModuleList([Universal(d, device=device, dtype=dtype, **kwargs) for d in self.ds])
        self.xs = torch.nn.ParameterList(
            [ManifoldParameter(data=man.rand(n).contiguous(), manifold=man) for man in self.manifolds])

    @property
    def device(self):
        return self.xs[0].device

    @property
    def curvature_params(self):
        for man in self.manifolds:
            yield man.c

    @torch.no_grad()
    def stabilize(self):
        for x in self.xs:
            # norm constraint |x| <= r_max (products/embedding.py:39-42), then back onto the manifold
            norm = x.norm(p=2, dim=-1, keepdim=True)
            norm.div_(self.r_max).clamp_(min=1)
            x.div_(norm)
            x.proj_()

    @torch.no_grad()
    def add_stats(self, writer, epoch):
        for i, man in enumerate(self.manifolds):
            writer.add_scalar(f'curv{i}', man.get_K(), epoch)

    def compute_dists(self, indices=None):
        return sum([
            man.pdist(x, squared=True) if indices is None else man.batch_pdist2(x, indices)
            for man, x in zip(self.manifolds, self.xs)
        ])

    def __len__(self):
        return self.n
