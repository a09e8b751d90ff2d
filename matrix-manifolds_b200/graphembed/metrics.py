"""Validation metrics of the reference's graphembed/metrics.py:13-56 -- `pearsonr`, `average_distortion`,
`average_pearsonr`, `spearmanr`, `area_under_curve` with the same call signatures -- plus the streamed form
`TrainingEngine._validate` uses here: `validation_moments(embedding, dataset)` walks the N(N-1)/2 pairs in
chunks (distance kernel -> gm_pairs_metrics) and never materialises the two distance vectors
(train.py:231-232 builds both, 228 M values each for condmat)."""
import math

import numpy as np
import torch

from . import _ops
from .utils import squareform1


class PairMoments:
    """[count, sum |m-g|/g, sum m, sum g, sum m^2, sum g^2, sum m g] over a set of pairs (float64, host)."""

    def __init__(self, acc):
        self.n, self.sum_rel, self.sm, self.sg, self.smm, self.sgg, self.smg = (float(v) for v in acc[:7])

    @property
    def average_distortion(self):
        return self.sum_rel / self.n

    @property
    def pearsonr(self):
        cov = self.smg - self.sm * self.sg / self.n
        vm = self.smm - self.sm * self.sm / self.n
        vg = self.sgg - self.sg * self.sg / self.n
        return cov / math.sqrt(vm * vg)

    def metric(self, name):
        if name not in ('pearsonr', 'average_distortion'):
            raise KeyError(name)
        return getattr(self, name)


STREAMED_METRICS = ('pearsonr', 'average_distortion')


def _vector_moments(mpdists, gpdists):
    if mpdists.shape != gpdists.shape or mpdists.ndim != 1:
        raise ValueError('expected two distance vectors of the same length')
    gp = gpdists.to(device=mpdists.device, dtype=mpdists.dtype)
    acc = _ops.pairs_metrics([mpdists], [1.0], _ops.PairSet.elementwise(mpdists.numel()), _ops.TargetSpec.vector(gp),
                             squared=False)
    return acc


def pearsonr(x, y):
    """Mimics scipy.stats.pearsonr (metrics.py:13-17); one fused moments kernel."""
    a = _vector_moments(x, y)
    n = a[0]
    cov = a[6] - a[2] * a[3] / n
    return (cov / torch.sqrt((a[4] - a[2] * a[2] / n) * (a[5] - a[3] * a[3] / n))).to(x.dtype)


def average_distortion(mpdists, gpdists):
    """mean(|m - g| / g) (metrics.py:46-56); one fused moments kernel."""
    a = _vector_moments(mpdists, gpdists)
    return (a[1] / a[0]).to(mpdists.dtype)


def average_pearsonr(mpdists, gpdists):
    """Per-node correlation, averaged (metrics.py:20-33).  O(N^2) dense work on square forms, off the hot path:
    plain tensor ops on whatever device the inputs live on."""
    m, g = squareform1(mpdists), squareform1(gpdists.to(mpdists.device))
    m = m - m.mean(dim=1)
    g = g - g.mean(dim=1)
    return ((m * g).sum(dim=1) / (m.norm(dim=1) * g.norm(dim=1))).mean()


def spearmanr(x, y):
    import scipy.stats
    return scipy.stats.spearmanr(x.cpu().numpy(), y.cpu().numpy()).correlation


def area_under_curve(vs, step=None):
    if step is None:
        step = len(vs)
    return [0.5 * np.mean(vs[(i + 1):(i + step)] + vs[i:(i + step - 1)])
            for i in range(0, len(vs) // step * step, step)]


@torch.no_grad()
def validation_moments(embedding, dataset, chunk_pairs=1 << 24, pair_range=None):
    """Moments of (sqrt of the product distance, sqrt of the normalised squared graph distance) over all pairs of the
    embedding -- what train.py:231-232 materialises -- streamed in chunks of `chunk_pairs`.  `pair_range` = (k0, k1)
    restricts the walk to a slice of the triangle (one rank of a pair-sharded validation); the returned accumulator
    (8 float64 on the device) can be summed across ranks before building `PairMoments`."""
    from torch.nn.functional import softplus
    n = embedding.n
    total = n * (n - 1) // 2
    k0, k1 = (0, total) if pair_range is None else pair_range
    dev = embedding.device
    pd = dataset.pdists
    if pd.device != dev or pd.dtype != embedding.xs[0].dtype:
        pd = pd.to(device=dev, dtype=embedding.xs[0].dtype)
    targets = _ops.TargetSpec.dense(pd)
    # products.Embedding has no scales: plain sum of the factors' squared distances (products/embedding.py:52-57)
    sps = [float(softplus(s.detach())) for s in embedding.scales] if hasattr(embedding, 'scales') \
        else [1.0] * len(embedding.xs)
    acc = torch.zeros(8, dtype=torch.float64, device=dev)
    for lo in range(k0, k1, chunk_pairs):
        pairs = _ops.PairSet.triu(n, k0=lo, P=min(chunk_pairs, k1 - lo))
        d2s = [_ops.pairs_dist2(m.spec, x.detach(), x.detach(), pairs) for m, x in zip(embedding.manifolds, embedding.xs)]
        _ops.pairs_metrics(d2s, sps, pairs, targets, squared=True, acc=acc)
    return acc
