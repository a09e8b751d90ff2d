#!/usr/bin/env python
"""DEV TOOL: one launch each of the secondary kernels, for `ncu --set full -k regex:rank_metrics|vec_pair_kernel`:
gm_rank_metrics on a 16384-node graph with random distinct distances, and the fused Universal (kappa-stereographic)
pair kernel on 2^22 sampled pairs of 1 M points.  Prints their CUDA-event times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
sys.path.insert(0, ROOT)

import networkx as nx  # noqa: E402
import torch  # noqa: E402

from bench import scale_free_edges  # noqa: E402
from graphembed import _lib as L, _ops  # noqa: E402
from graphembed.manifolds import Universal  # noqa: E402
from graphembed.pyx import FastPrecision  # noqa: E402


def timed(fn, reps=3):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device('cuda', 0)
    n = 16384
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(scale_free_edges(n, 3, 0).tolist())
    fp = FastPrecision(g, device=dev)
    P = n * (n - 1) // 2
    pd = ((torch.randperm(P, device=dev) + 1).double() / P).float()
    ms = timed(lambda: fp.layer_mean_f1_scores(pd), reps=2)
    print(f'gm_rank_metrics N={n} ({fp.n_layers} layers, {P} distances): {ms:.2f} ms -> '
          f'{n * n / ms / 1e6:.2f} G (root,node) ranks/s')
    N, Pp = 1_000_000, 1 << 22
    man = Universal(8, c_init=0.3, device=dev, dtype=torch.float32)
    x = man.rand(N, ir=0.5)
    I = torch.randint(N, (Pp,), device=dev, dtype=torch.int32)
    J = (I + 1 + torch.randint(N - 1, (Pp,), device=dev, dtype=torch.int32)) % N
    hops = torch.randint(1, 9, (Pp,), device=dev, dtype=torch.uint8)
    grad = torch.zeros_like(x)
    cg = torch.zeros(1, dtype=torch.float64, device=dev)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    pairs = _ops.PairSet.from_lists(I, J, dev)
    ms = timed(lambda: _ops.pairs_loss_fused(man.spec, x, pairs, _ops.TargetSpec.hops(hops, 64.0), spec, 1.0, grad, c_grad=cg))
    bpp = 4 * 8 * 4 + 4 + 8
    print(f'Universal(8) fused pair kernel, 2^22 pairs of 1M points: {ms:.3f} ms -> {Pp / ms / 1e6:.2f} G pairs/s, '
          f'{Pp * bpp / ms / 1e6:.0f} GB/s algorithmic ({bpp} B/pair)')


if __name__ == '__main__':
    main()
