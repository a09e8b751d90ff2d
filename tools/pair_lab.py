#!/usr/bin/env python
"""A/B harness for the fused SPD 4x4 pair kernel on the bench workload (BASELINE config 5: 2M points, 2^24 sampled
pairs per launch, source-grouped, hop counts packed into the top byte of j).

  python tools/pair_lab.py --make /tmp/wl.pt            build the workload once (BFS kernel + sampler)
  GM_PAIR_PARK=0 python tools/pair_lab.py --run /tmp/wl.pt [--iters 20]
                                                        time the kernel alone with CUDA events under the current
                                                        environment (the launcher reads its knobs once per process)
Prints one JSON line: kernel ms (mean / min), loss and a gradient checksum (to compare variants), launches.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
sys.path.insert(0, ROOT)


def make(path, nodes, log2_pairs, trained_steps):
    import bench
    from graphembed.engine import PairTrainer
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam
    dev = torch.device('cuda', 0)
    torch.manual_seed(42)
    emb = ManifoldEmbedding(nodes, [SymmetricPositiveDefinite(4)], device=dev, dtype=torch.float32)
    batches, max_sq = bench.make_pair_batches(nodes, log2_pairs, 2, dev, seed=1234)
    x0 = emb.xs[0].detach().clone()
    if trained_steps:  # also keep a copy of the points after a few optimizer steps (spread-out spectrum)
        opt = RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True)
        tr = PairTrainer(emb, opt, QuotientLoss(), max_hops_sq=max_sq)
        for k in range(trained_steps):
            tr.step(batches[k % 2][0].to(dev), batches[k % 2][5].to(dev), None, epoch=1)
    torch.save({'x0': x0.cpu(), 'x1': emb.xs[0].detach().cpu(), 'I': [b[0] for b in batches],
                'JP': [b[5] for b in batches], 'max_sq': max_sq}, path)
    print('workload written:', path)


def run(path, iters, which, sort_j=0):
    from graphembed import _lib as L, _ops
    from graphembed.manifolds import SymmetricPositiveDefinite
    dev = torch.device('cuda', 0)
    wl = torch.load(path)
    x = wl[which].to(dev).contiguous()
    man = SymmetricPositiveDefinite(4)
    Is = [t.to(dev) for t in wl['I']]
    Js = [t.to(dev) for t in wl['JP']]
    if sort_j:  # order every group of `sort_j` consecutive pairs (one source's pairs, or a slice of them) by target row
        def by_target(jp):
            g = jp.view(-1, sort_j)
            order = (g & 0x00ffffff).argsort(dim=1)
            return g.gather(1, order).reshape(-1).contiguous()
        Js = [by_target(j) for j in Js]
    tg = _ops.TargetSpec.hops_packed(wl['max_sq'])
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    grad = torch.zeros_like(x)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    times = []
    for k in range(iters + 3):
        pairs = _ops.PairSet.from_lists(Is[k % 2], Js[k % 2], dev)
        grad.zero_()
        acc.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 0.97, grad, acc)
        e1.record()
        torch.cuda.synchronize()
        if k >= 3:
            times.append(e0.elapsed_time(e1))
    env = {k: v for k, v in os.environ.items() if k.startswith('GM_')}
    print(json.dumps({'env': env, 'points': which, 'sort_j': sort_j, 'ms_mean': sum(times) / len(times), 'ms_min': min(times),
                      'loss': acc[0].item(), 'sum_ld2': acc[1].item(), 'grad_abs_sum': grad.double().abs().sum().item(),
                      'grad_sq': grad.double().pow(2).sum().item(), 'finite': bool(torch.isfinite(grad).all())}))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--make')
    ap.add_argument('--run')
    ap.add_argument('--nodes', type=int, default=2_000_000)
    ap.add_argument('--pairs-log2', type=int, default=24)
    ap.add_argument('--trained-steps', type=int, default=30)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--points', default='x0', choices=['x0', 'x1'])
    ap.add_argument('--sort-j', type=int, default=0, help='sort the targets inside every group of this many pairs')
    a = ap.parse_args()
    if a.make:
        make(a.make, a.nodes, a.pairs_log2, a.trained_steps)
    else:
        run(a.run, a.iters, a.points, a.sort_j)
