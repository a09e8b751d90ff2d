#!/usr/bin/env python
"""One multi-source BFS (1024 sources) of the 2 M-node bench graph: for ncu launch lists / A/B (GM_BFS_V1=1)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from graphembed.data import bfs_levels, edges_to_csr  # noqa: E402

dev = torch.device('cuda', 0)
N = int(os.environ.get('BFS_N', 2_000_000))
S = int(os.environ.get('BFS_S', 1024))
rowptr, colidx = edges_to_csr(N, bench.scale_free_edges(N, 4, 1234))
rp, ci = torch.as_tensor(rowptr, device=dev), torch.as_tensor(colidx, device=dev)
g = torch.Generator().manual_seed(1)
for k in range(2):
    src = torch.randperm(N, generator=g)[:S].int().to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lv = bfs_levels(rp, ci, sources=src, device=dev, level_bytes=1)
    e1.record()
    torch.cuda.synchronize()
    print('bfs ms', e0.elapsed_time(e1), 'depth', int(lv.max()))
    del lv
