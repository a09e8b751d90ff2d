#!/usr/bin/env python
"""A/B of the in-kernel pair draw (GM_PAIRS_SAMPLED) against explicit pair lists on the bench workload; optional
cudaLimitMaxL2FetchGranularity (GM_L2_FETCH=32|64|128)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
sys.path.insert(0, ROOT)


def main():
    import bench
    from graphembed import _lib as L, _ops
    from graphembed.manifolds import SymmetricPositiveDefinite
    dev = torch.device('cuda', 0)
    torch.zeros(1, device=dev)
    gran = os.environ.get('GM_L2_FETCH')
    if gran:
        rt = ctypes.CDLL('libcudart.so.12')
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(gran)))  # cudaLimitMaxL2FetchGranularity
        val = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        print('set L2 fetch granularity rc', rc, 'now', val.value)
    N = 2_000_000
    torch.manual_seed(42)
    man = SymmetricPositiveDefinite(4)
    x = man.rand(N, out=torch.empty(0, device=dev, dtype=torch.float32)).contiguous()
    batches, max_sq, levels, per_src, graph = bench.make_pair_batches(N, 24, 2, dev, seed=1234, keep_levels=True)
    tg = _ops.TargetSpec.hops_packed(max_sq)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    grad = torch.zeros_like(x)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    res = {}
    for mode in ('lists', 'sampled'):
        times = []
        for k in range(13):
            b = batches[k % 2]
            if mode == 'lists':
                pairs = _ops.PairSet.from_lists(b[0].to(dev), b[5].to(dev), dev)
            else:
                pairs = _ops.PairSet.sampled(b[3].to(dev), levels[k % 2], per_src, seed=1000 + k)
            grad.zero_(); acc.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 0.97, grad, acc)
            e1.record()
            torch.cuda.synchronize()
            if k >= 3:
                times.append(e0.elapsed_time(e1))
        res[mode] = sum(times) / len(times)
    print(json.dumps({'l2_fetch': gran, **res}))


if __name__ == '__main__':
    main()
