#!/usr/bin/env python
"""Secondary measurement (NOT the driver's bench line -- that is bench.py on BASELINE config 5): epoch time and pair
throughput of BASELINE.json configs 1-4 on one B200, through the public training API (TrainingEngine / BatchedObjective /
RiemannianSGD|Adam), on synthetic graphs of the named sizes (the reference's edge lists do not travel to the GPU box).

    python tools/bench_configs.py [--epochs 3] [--only 1,2a,...] > profiles/rNN_configs.json
    torchrun --nproc-per-node G tools/bench_configs.py --only 4      # pair-sharded over G GPUs

One JSON line per config: pairs per epoch, steps per epoch, median epoch ms (CUDA-synchronised wall clock around
TrainingEngine._train, validation off), pairs/s, the pair kernel's share measured with CUDA events, and the algorithmic
HBM rate (SURVEY 8d bytes per pair) for the sampled/batched configs.  Targets come from the GPU BFS for N <= 5000 and,
for config 4, from hop counts of the same BFS kept as a dense fp32 matrix resident in HBM (1.8 GB).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'matrix-manifolds_b200'))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def graph_targets(n, dev, dtype, seed=0):
    """Dense (n, n) normalised squared hop-count targets of a connected scale-free graph, built on the GPU."""
    from bench import scale_free_edges
    from graphembed.data import bfs_levels, edges_to_csr
    from graphembed import _lib as L
    rowptr, colidx = edges_to_csr(n, scale_free_edges(n, 3, seed))
    levels = bfs_levels(rowptr, colidx, device=dev, level_bytes=1)
    max_sq = float(levels.max().item())**2
    dense = torch.empty(n, n, dtype=dtype, device=dev)
    rc = L.lib().gm_levels_to_dense_targets(1, L.ptr(levels), n, max_sq, L.dtype_code(dtype), L.ptr(dense),
                                            L.stream_ptr(dev))
    L.check(rc, 'gm_levels_to_dense_targets')
    return dense


class DenseDataset:
    """GraphDataset (data/dataset.py:9-27) over an already-normalised dense target matrix."""

    def __init__(self, dense):
        self.pdists = dense

    @property
    def device(self):
        return self.pdists.device

    def __len__(self):
        return len(self.pdists)

    def __getitem__(self, idx=None):
        sub = self.pdists if idx is None else self.pdists[idx.to(self.device)][:, idx.to(self.device)]
        i, j = torch.triu_indices(len(sub), len(sub), 1, device=self.device)
        return sub[i, j]


def bytes_per_pair(E_list, s):
    return sum(4 * E * s for E in E_list) + s + 8


def run_config(name, n, mk_manifolds, dtype, opt_name, batch_nodes, epochs, dev, E_list, pg=None):
    from graphembed import _ops
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam, RiemannianSGD
    from graphembed.train import TrainingEngine
    os.makedirs(f"/tmp/bench_configs/r{0 if pg is None else torch.distributed.get_rank(pg)}", exist_ok=True)
    torch.manual_seed(42)
    ds = DenseDataset(graph_targets(n, dev, dtype))
    emb = ManifoldEmbedding(n, mk_manifolds(), device=dev, dtype=dtype)
    if opt_name == 'rsgd':
        opt = RiemannianSGD(emb.xs, lr=1e-3, max_grad_norm=20, exact=True)
    else:
        opt = RiemannianAdam(emb.xs, lr=1e-3, max_grad_norm=100, exact=True)
    world = 1 if pg is None else torch.distributed.get_world_size(pg)
    rank = 0 if pg is None else torch.distributed.get_rank(pg)
    eng = TrainingEngine(embedding=emb, optimizer=opt, objective_fn=QuotientLoss(), n_epochs=epochs, alpha=1.0,
                         batch_size=batch_nodes, tensorboard=False, save_dir=f'/tmp/bench_configs/r{rank}',
                         process_group=pg)
    # time the pair kernels inside the epochs
    kernel_events = []

    def timed(fn):
        def wrapper(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **kw)
            e1.record()
            kernel_events.append((e0, e1))
            return r
        return wrapper

    saved = (_ops.pairs_loss_fused, _ops.pairs_dist2, _ops.pairs_grad)
    _ops.pairs_loss_fused, _ops.pairs_dist2, _ops.pairs_grad = (timed(f) for f in saved)
    epoch_ms = []
    orig_train = eng._train

    def timed_train(*a, **kw):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        r = orig_train(*a, **kw)
        torch.cuda.synchronize(dev)
        epoch_ms.append((time.perf_counter() - t0) * 1e3)
        return r

    eng._train = timed_train
    try:
        eng(ds)
    finally:
        _ops.pairs_loss_fused, _ops.pairs_dist2, _ops.pairs_grad = saved
    bs = n if batch_nodes is None else min(n, batch_nodes)
    steps = sum(1 for i in range(0, n, bs) if min(bs, n - i) >= 50)
    pairs = sum(b * (b - 1) // 2 for b in (min(bs, n - i) for i in range(0, n, bs)) if b >= 50)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_events) / max(len(epoch_ms), 1)
    med = float(np.median(epoch_ms[1:] if len(epoch_ms) > 1 else epoch_ms))
    if pg is not None:  # every rank evaluates its slice of each batch's pair triangle; report the slowest rank
        t = torch.tensor([med], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=pg)
        med = float(t.item())
        if rank != 0:
            return
    s = 4 if dtype == torch.float32 else 8
    line = dict(config=name, n_gpus=world, nodes=n, dtype='f32' if s == 4 else 'f64', optimizer=opt_name,
                batch_nodes=batch_nodes,
                steps_per_epoch=steps, pairs_per_epoch=pairs, epoch_ms=med, pairs_per_s=pairs / (med * 1e-3),
                pair_kernels_ms_per_epoch_rank0=kernel_ms, bytes_per_pair=bytes_per_pair(E_list, s),
                final_loss=float(eng.writer.history['quotient_loss'][-1][1]))
    print(json.dumps(line), flush=True)
    del eng, emb, ds
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=4)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    from graphembed.manifolds import Grassmann, Lorentz, SymmetricPositiveDefinite as SPD
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    pg = None
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:  # torchrun: every batch's pair triangle is sharded over the ranks
        torch.distributed.init_process_group('nccl', device_id=dev)
        pg = torch.distributed.group.WORLD
    torch.set_default_device(dev)  # as run.py does: the per-epoch randperm and its slices stay on the GPU
    f32, f64 = torch.float32, torch.float64
    configs = [
        # BASELINE configs[0]: tree1000-sized graph, SPD 3x3, full-pair loss, RSGD, fp64
        ('1: N=1000 SPD3 f64 full batch RSGD', 1000, lambda: [SPD(3)], f64, 'rsgd', None, [9]),
        # configs[1]: power-grid-sized graph (4941 nodes), node batches of 512, RAdam, fp32
        ('2a: N=4941 Lorentz(11) f32 batch 512 RAdam', 4941, lambda: [Lorentz(11)], f32, 'radam', 512, [11]),
        ('2b: N=4941 SPD4-Stein f32 batch 512 RAdam', 4941, lambda: [SPD(4, use_stein_div=True)], f32, 'radam', 512, [16]),
        # configs[2]: facebook-sized graph (4039 nodes)
        # fp64: the reference's default Grassmann init is non-finite in fp32 (sigma == 1 exactly)
        ('3a: N=4039 Gr(2,6) f64 batch 512 RAdam', 4039, lambda: [Grassmann(6, 2)], f64, 'radam', 512, [12]),
        ('3b: N=4039 SPD3 x Lorentz(5) f32 batch 512 RAdam', 4039, lambda: [SPD(3), Lorentz(5)], f32, 'radam', 512, [9, 5]),
        # configs[3]: condmat-sized graph (21363 nodes), SPD 6x6, all 228 M pairs per step
        ('4: N=21363 SPD6 f32 full batch RAdam', 21363, lambda: [SPD(6)], f32, 'radam', None, [36]),
    ]
    only = {t.strip() for t in args.only.split(',') if t.strip()}
    for name, n, mk, dtype, opt, bn, E in configs:
        if only and name.split(':')[0] not in only:
            continue
        run_config(name, n, mk, dtype, opt, bn, args.epochs, dev, E, pg)
    if pg is not None:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
