#!/usr/bin/env python
"""Times the fused pair kernel (distance + loss + gradient) of a vector manifold on 2^22 source-grouped sampled pairs
over a 2 M-point table -- the measurement VERDICT r1 #9 asks for (Lorentz(11) as BASELINE config 2a embeds, and the
16-byte-row sizes around it).

  python tools/vec_lab.py [--pairs-log2 22] [--nodes 2000000]
Prints one JSON line per (manifold, n): kernel ms, pairs/s, algorithmic GB/s (two row reads + the second endpoint's
gradient read-modify-write + 4 B of indices, 1 B hop count; the first endpoint's row and gradient are amortised over
its 4096-pair group) and its fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nodes', type=int, default=2_000_000)
    ap.add_argument('--pairs-log2', type=int, default=22)
    ap.add_argument('--per-src', type=int, default=4096)
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    from graphembed import _lib as L, _ops
    from graphembed import manifolds as M
    import bench
    dev = torch.device('cuda', 0)
    peak = bench.hbm_peak()[0] if hasattr(bench, 'hbm_peak') else 6454.6
    P = 1 << a.pairs_log2
    g = torch.Generator(device='cpu').manual_seed(7)
    n_src = P // a.per_src
    src = torch.randperm(a.nodes, generator=g)[:n_src]
    I = src.repeat_interleave(a.per_src).to(torch.int32)
    J = torch.randint(0, a.nodes, (P,), generator=g, dtype=torch.int32)
    J = torch.where(J == I, (J + 1) % a.nodes, J)
    hops = torch.randint(1, 9, (P,), generator=g, dtype=torch.int32)
    JP = (J | (hops << 24)).to(dev)
    I = I.to(dev)
    tg = _ops.TargetSpec.hops_packed(64)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    cases = [('lorentz', 11), ('lorentz', 12), ('lorentz', 8), ('lorentz', 16), ('sphere', 12), ('euclidean', 12)]
    for fam, n in cases:
        man = {'lorentz': M.Lorentz, 'sphere': M.Sphere, 'euclidean': M.Euclidean}[fam](n)
        torch.manual_seed(3)
        x = man.rand(a.nodes, out=torch.empty(0, device=dev, dtype=torch.float32), ir=0.5).contiguous()
        grad = torch.zeros_like(x)
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        pairs = _ops.PairSet.from_lists(I, JP, dev)
        times = []
        for k in range(a.iters + 3):
            grad.zero_(); acc.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 0.97, grad, acc)
            e1.record()
            torch.cuda.synchronize()
            if k >= 3:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        row = x[0].numel() * 4
        bpp = 2 * row + 2 * row + 4 + (2 * row + 4) / a.per_src  # y read, gy RMW, packed j; x row + gx + i per group
        gbs = bpp * P / ms / 1e6
        print(json.dumps({'manifold': f'{fam}({n})', 'pairs': P, 'kernel_ms': round(ms, 4), 'kernel_ms_min': round(min(times), 4),
                          'pairs_per_s': P / ms * 1e3, 'bytes_per_pair': round(bpp, 1), 'algorithmic_GBps': round(gbs, 1),
                          'hbm_peak_GBps': peak, 'frac': round(gbs / peak, 4), 'loss': acc[0].item(),
                          'finite': bool(torch.isfinite(grad).all())}))


if __name__ == '__main__':
    main()
