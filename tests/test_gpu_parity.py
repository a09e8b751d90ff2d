"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C-ABI of libgm_b200.so,
against (a) golden vectors produced by the real reference and (b) the pinned oracle on fresh seeded inputs.

Tolerances are the ones BASELINE.json's north_star states: 1e-10 relative (fp64), 1e-5 relative (fp32) on
distances, gradients, losses; bit-exact for BFS / indexing.  fp32 results are held to an error budget against the
reference's fp64 answer on the same fp32 inputs (helpers.assert_parity): within 1e-5 of it, or -- where the
reference's own fp32 arithmetic is further away than that -- at most twice as far as the reference itself.
"""
import numpy as np
import pytest
import torch

import manifolds_oracle as O
from helpers import (CASES, DTYPES, assert_parity, assert_parity_scalar, is_spd, load_golden, load_truth, make_oracle,
                     make_product, parity_errors, rel_err, sym, tol)

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _fix(name):
    return sym if is_spd(name) else (lambda t: t)


def _truth(name, tag):
    return load_truth(name) if tag == 'f32' else None


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_dist_elementwise_vs_golden(name, tag):
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_product(name)
    x, y = g['x'].to(DEV).requires_grad_(), g['y'].to(DEV).requires_grad_()
    d2 = man.dist(x, y, squared=True)
    (d2 * g['w'].to(DEV)).sum().backward()
    fix = _fix(name)
    assert_parity(d2.detach(), g, 'dist2', tag, T)
    assert_parity(x.grad, g, 'gx', tag, T, fix)
    assert_parity(y.grad, g, 'gy', tag, T, fix)
    # non-squared distance and its gradient flow through torch's sqrt like the reference's
    d = man.dist(g['x'].to(DEV), g['y'].to(DEV))
    assert_parity(d * d, g, 'dist2', tag, T)


LOSSES = (('quot', dict(epoch=3, alpha=1.7)), ('quot_l1', dict(epoch=3, alpha=1.7)), ('stress', dict()))


def _loss_fn(lname):
    from graphembed.objectives import QuotientLoss, StressLoss
    return {'quot': QuotientLoss(), 'quot_l1': QuotientLoss(inc_l2=False), 'stress': StressLoss()}[lname]


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_pdist_and_losses_vs_golden(name, tag):
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_product(name)
    fix = _fix(name)
    targets = g['targets'].to(DEV)
    for lname, kw in LOSSES:
        x = g['x'].to(DEV).requires_grad_()
        pd2 = man.pdist(x, squared=True)
        loss = _loss_fn(lname)(targets, 0.9 * pd2, **kw)
        loss.backward()
        assert_parity(pd2.detach(), g, 'pdist2', tag, T)
        assert_parity_scalar(loss.item(), g, f'loss_{lname}', tag, T)
        assert_parity(x.grad, g, f'grad_{lname}', tag, T, fix)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_fused_kernel_vs_golden(name, tag):
    """gm_pairs_loss_fused (distance + loss + gradient in one launch) against the reference's autograd."""
    from graphembed import _ops, _lib as L
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_product(name)
    fix = _fix(name)
    x = g['x'].to(DEV).contiguous()
    n = x.shape[0]
    for lname, spec in (('quot', _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.7, eps=1 / 4)),
                        ('quot_l1', _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, False, alpha=1.7, eps=1 / 4)),
                        ('stress', _ops.LossSpec(L.GM_LOSS_STRESS))):
        grad = torch.zeros_like(x)
        acc, d2 = _ops.pairs_loss_fused(man.spec, x, _ops.PairSet.triu(n), _ops.TargetSpec.vector(g['targets'].to(DEV)),
                                        spec, 0.9, grad, want_d2=True)
        assert_parity(d2, g, 'pdist2', tag, T)
        assert_parity_scalar(acc[0].item(), g, f'loss_{lname}', tag, T)
        assert_parity(grad, g, f'grad_{lname}', tag, T, fix)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_point_ops_vs_golden(name, tag):
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_product(name)
    x, y, u, v, eg = (g[k].to(DEV) for k in ('x', 'y', 'u', 'v', 'eg'))
    assert_parity(man.exp(x, u), g, 'exp', tag, T)
    assert_parity(man.retr(x, u), g, 'retr', tag, T)
    assert_parity(man.log(x, y), g, 'log', tag, T)
    assert_parity(man.proju(x, eg), g, 'proju', tag, T)
    assert_parity(man.egrad2rgrad(x, eg), g, 'egrad2rgrad', tag, T)
    assert_parity(man.transp(x, y, u), g, 'transp', tag, T)
    assert_parity(man.inner(x, u, v), g, 'inner', tag, T)
    assert_parity(man.norm(x, u, squared=True).reshape(g['norm2'].shape), g, 'norm2', tag, T)
    assert man.norm(x, u, keepdim=True).shape == (x.shape[0],) + (1,) * man.ndim


OPTS = {
    'radam_clip': ('radam', dict(lr=0.05, max_grad_norm=1.5)),
    'radam_exact': ('radam', dict(lr=0.05, exact=True)),
    'rsgd_exact_clip': ('rsgd', dict(lr=0.05, max_grad_norm=0.5, exact=True)),
    'rsgd_momentum': ('rsgd', dict(lr=0.05, momentum=0.9, dampening=0.1)),
}


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('oname', sorted(OPTS))
@pytest.mark.parametrize('name', sorted(CASES))
def test_optimizer_trajectories_vs_golden(name, oname, tag):
    from graphembed.modules import ManifoldParameter
    from graphembed.optim import RiemannianAdam, RiemannianSGD
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_product(name)
    kind, kw = OPTS[oname]
    p = ManifoldParameter(g['x'].to(DEV).contiguous(), manifold=man)
    opt = (RiemannianAdam if kind == 'radam' else RiemannianSGD)([p], **kw)
    for k in range(3):
        p.grad = g['opt_grads'][k].to(DEV)
        opt.step()
        assert_parity(p.data, g, f'{oname}_x', tag, T, index=k, what=f'{oname} step {k}')
    st = opt.state[p]
    for key in ('exp_avg', 'exp_avg_sq', 'momentum_buffer'):
        if f'{oname}_{key}' in g:
            assert_parity(st[key], g, f'{oname}_{key}', tag, T)


@pytest.mark.parametrize('tag', ['spd3_rsgd', 'prod_radam'])
def test_training_run_vs_golden(tag):
    """BASELINE config 1 in miniature (tree -> SPD 3x3, fp64, QuotientLoss, RSGD exact clip 20) and a product
    SPD3 x Lorentz5 with RAdam, free-running for 4 steps through the drop-in modules; targets from the BFS kernel."""
    from graphembed.data import GraphDataset, bfs_levels, edges_to_csr
    from graphembed.data.graph import levels_to_condensed
    from graphembed.manifolds import SymmetricPositiveDefinite, Lorentz
    from graphembed.modules import ManifoldEmbedding, BatchedObjective
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam, RiemannianSGD
    with np.load(f'{__import__("helpers").GOLDEN}/training_run_f64.npz') as z:
        g = {k: torch.from_numpy(z[k]) for k in z.files}
    n = 31
    rowptr, colidx = edges_to_csr(n, g['edges'].numpy())
    levels = bfs_levels(rowptr, colidx)
    cond = levels_to_condensed(levels, torch.float64)
    assert torch.equal(cond.cpu(), g['hops_condensed'])  # BFS targets bit-exact
    ds = GraphDataset(cond)
    mans = [SymmetricPositiveDefinite(3)] if tag == 'spd3_rsgd' else [SymmetricPositiveDefinite(3), Lorentz(5)]
    emb = ManifoldEmbedding(n, mans, device=DEV, dtype=torch.float64)
    with torch.no_grad():
        for i, x in enumerate(emb.xs):
            x.copy_(g[f'{tag}_x0_{i}'].to(DEV))
    opt = (RiemannianSGD(emb.xs, lr=0.01, max_grad_norm=20, exact=True) if tag == 'spd3_rsgd' else
           RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True))
    bobj = BatchedObjective(QuotientLoss(), ds, emb)
    perm = g[f'{tag}_perm'].to(DEV)
    losses = []
    for step in range(4):
        idx = perm if step % 2 == 0 else perm[:20]
        loss = bobj(idx, alpha=1.0, epoch=step + 1).sum()
        opt.zero_grad()
        loss.backward()
        if step == 0:
            for i, x in enumerate(emb.xs):
                fix = sym if i == 0 else (lambda t: t)
                assert rel_err(fix(x.grad), fix(g[f'{tag}_grad0_{i}'])) < 1e-10
        opt.step()
        losses.append(loss.item())
    ref = g[f'{tag}_losses']
    assert np.allclose(np.array(losses), ref.numpy(), rtol=1e-9, atol=0), (losses, ref)
    for i, x in enumerate(emb.xs):
        assert rel_err(x.data, g[f'{tag}_xT_{i}']) < 1e-9


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('name', ['spd3', 'spd4', 'spd6', 'stein4', 'lorentz11', 'sphere5', 'grassmann6_2', 'euclidean7'])
def test_pair_list_and_batch_gather_vs_oracle(name, dtype):
    """LIST pairs (random I, J incl. repeats) and TRIU-with-node-gather against the oracle on seeded inputs.
    fp32: error budget against the oracle's fp64 answer on the same fp32 inputs (see helpers.assert_parity)."""
    gen = torch.Generator().manual_seed(5)
    g = load_golden(name, 'f64')
    man, orc = make_product(name), make_oracle(name)
    xs = torch.cat([g['x'], g['y']]).to(dtype)  # 24 valid points
    n = xs.shape[0]
    P = 500
    I = torch.randint(n, (P,), generator=gen)
    J = (I + 1 + torch.randint(n - 1, (P,), generator=gen)) % n
    w = torch.rand(P, generator=gen, dtype=dtype) + 0.5
    fix = _fix(name)

    def oracle_pairs(dt):
        xo = xs.detach().clone().to(dt).requires_grad_()
        d2o = orc.dist2(xo[I], xo[J])
        (d2o * w.to(dt)).sum().backward()
        return d2o.detach(), xo.grad

    def oracle_batch(dt, nodes):
        xo = xs.detach().clone().to(dt).requires_grad_()
        pdo = orc.pdist2(xo[nodes])
        pdo.sum().backward()
        return pdo.detach(), xo.grad

    def check(got, same, truth, fx=None):
        fx = fx or (lambda t: t)
        if dtype == torch.float64:
            assert rel_err(fx(got), fx(same)) < 1e-10
        else:
            e_got, e_ref = parity_errors(fx(got), fx(same), fx(truth))
            assert e_got <= max(1e-5, 2 * e_ref), (e_got, e_ref)

    d2o, gxo = oracle_pairs(dtype)
    d2t, gxt = oracle_pairs(torch.float64)
    xg = xs.to(DEV).requires_grad_()
    d2 = man.pair_dist2(xg, I.to(DEV), J.to(DEV))
    (d2 * w.to(DEV)).sum().backward()
    check(d2.detach(), d2o, d2t)
    check(xg.grad, gxo, gxt, fix)
    # int32 indices take the same path
    d2b = man.pair_dist2(xs.to(DEV), I.int().to(DEV), J.int().to(DEV))
    assert torch.equal(d2b, d2.detach())
    # pdist(x[nodes]) with the gather fused
    nodes = torch.randperm(n, generator=gen)[:17]
    pdo, gpo = oracle_batch(dtype, nodes)
    pdt, gpt = oracle_batch(torch.float64, nodes)
    xg = xs.to(DEV).requires_grad_()
    pd = man.batch_pdist2(xg, nodes.to(DEV))
    pd.sum().backward()
    check(pd.detach(), pdo, pdt)
    check(xg.grad, gpo, gpt, fix)


def test_bfs_bit_exact_random_graphs():
    """Multi-source BFS against scipy's BFS-based shortest_path on random connected graphs, all level widths,
    partial source sets, and a path graph deeper than 254 hops (forces the uint16 retry)."""
    import networkx as nx
    from scipy.sparse.csgraph import shortest_path
    from graphembed.data import bfs_levels, edges_to_csr
    for seed, (n, m) in enumerate([(50, 2), (333, 1), (1000, 3)]):
        g = nx.barabasi_albert_graph(n, m, seed=seed)
        ref = shortest_path(nx.to_scipy_sparse_array(g), unweighted=True).astype(np.int64)
        rowptr, colidx = edges_to_csr(n, np.array(g.edges()))
        lv = bfs_levels(rowptr, colidx)
        assert lv.dtype == torch.uint8 and np.array_equal(lv.cpu().numpy().astype(np.int64), ref)
        src = np.random.RandomState(seed).choice(n, size=min(70, n - 3), replace=False)
        lv = bfs_levels(rowptr, colidx, sources=src, level_bytes=4)
        assert np.array_equal(lv.cpu().numpy().astype(np.int64), ref[src])
    g = nx.path_graph(400)
    rowptr, colidx = edges_to_csr(400, np.array(g.edges()))
    lv = bfs_levels(rowptr, colidx)
    assert lv.dtype == torch.int16
    ref = np.abs(np.arange(400)[:, None] - np.arange(400)[None, :])
    assert np.array_equal(lv.cpu().numpy().astype(np.int64), ref)
    # disconnected: unreachable marker
    rowptr, colidx = edges_to_csr(4, np.array([[0, 1], [2, 3]]))
    lv = bfs_levels(rowptr, colidx).cpu().numpy()
    assert lv[0, 1] == 1 and lv[0, 2] == 255 and lv[2, 3] == 1


def test_dataset_targets_match_oracle():
    from graphembed.data import GraphDataset
    gen = torch.Generator().manual_seed(0)
    hops = torch.randint(1, 12, (45,), generator=gen).double()
    ds = GraphDataset(hops.to(DEV))
    ref = O.dataset_targets(hops, torch.float64)
    assert torch.equal(ds[None].cpu(), ref)
    idx = torch.tensor([7, 2, 9, 0])
    dense = torch.zeros(10, 10, dtype=torch.float64)
    i, j = torch.triu_indices(10, 10, 1)
    dense[i, j] = ref
    dense = dense + dense.T
    assert torch.equal(ds[idx.to(DEV)].cpu(), O.batch_targets(dense, idx))


@pytest.mark.parametrize('dtype,N,P', [(torch.float32, 1 << 16, 1 << 20), (torch.float64, 1 << 16, 1 << 20),
                                       (torch.float32, 2_000_000, 1 << 24)])  # the last: BASELINE config 5's full size
def test_full_size_properties_spd4(dtype, N, P):
    """BASELINE-sized invariants that need no oracle: symmetry, affine invariance, fused == unfused, the sum of
    all gradient rows equals the gradient computed pair-by-pair."""
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import SymmetricPositiveDefinite
    torch.manual_seed(0)
    man = SymmetricPositiveDefinite(4)
    x = man.rand(N, out=torch.empty(0, device=DEV, dtype=dtype), ir=1.0)
    I = torch.randint(N, (P,), device=DEV, dtype=torch.int32)
    J = (I + 1 + torch.randint(N - 1, (P,), device=DEV, dtype=torch.int32)) % N
    d_ij = man.pair_dist2(x, I, J)
    d_ji = man.pair_dist2(x, J, I)
    rt = 5e-5 if dtype == torch.float32 else 1e-10
    assert rel_err(d_ij, d_ji) < rt
    a = torch.randn(4, 4, device=DEV, dtype=dtype) + 3 * torch.eye(4, device=DEV, dtype=dtype)
    xa = a @ x @ a.T
    xa = 0.5 * (xa + xa.transpose(-2, -1))
    assert rel_err(man.pair_dist2(xa, I, J), d_ij) < rt * 20
    hops = torch.randint(1, 9, (P,), device=DEV, dtype=torch.uint8)
    tg = _ops.TargetSpec.hops(hops, 64.0)
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.0, eps=0.5)
    pairs = _ops.PairSet.from_lists(I, J, DEV)
    grad = torch.zeros_like(x)
    acc, d2 = _ops.pairs_loss_fused(man.spec, x, pairs, tg, spec, 0.97, grad, want_d2=True)
    assert torch.equal(d2, d_ij)
    t = (hops.to(dtype) ** 2) / 64.0
    acc2, g = _ops.product_loss([d_ij], [0.97], _ops.TargetSpec.vector(t), spec)
    # fp32: the fused training kernel evaluates the loss quotients with MUFU reciprocals (~2 ulp per term),
    # the unfused product_loss with IEEE division
    assert abs(acc[0].item() - acc2[0].item()) <= (1e-6 if dtype == torch.float32 else 1e-9) * abs(acc2[0].item())
    grad2 = torch.zeros_like(x)
    _ops.pairs_grad(man.spec, x, x, pairs, g, grad2, grad2, coef=0.97)
    assert rel_err(grad, grad2) < (1e-4 if dtype == torch.float32 else 1e-11)
    assert torch.isfinite(grad).all()


def test_grouped_host_step_matches_list_step():
    """step_host_grouped (sources, offsets, j, hops) == step_host (i, j, hops) on the same pairs: same loss and the
    same updated points, bit for bit; expand_groups handles empty groups and ragged tails."""
    from graphembed import _ops
    from graphembed.engine import PairTrainer
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam
    # expand_groups alone, ragged with empty groups
    rows = torch.tensor([7, 3, 9, 1, 4], dtype=torch.int32, device=DEV)
    offs = torch.tensor([0, 5, 5, 18, 18, 21], dtype=torch.int64, device=DEV)
    out = torch.empty(21, dtype=torch.int32, device=DEV)
    _ops.expand_groups(rows, offs, out)
    want = torch.repeat_interleave(rows.cpu(), (offs[1:] - offs[:-1]).cpu())
    assert torch.equal(out.cpu(), want)
    results = []
    for grouped in (False, True):
        torch.manual_seed(3)
        n, G, per = 300, 37, 129
        emb = ManifoldEmbedding(n, [SymmetricPositiveDefinite(4)], device=torch.device(DEV), dtype=torch.float32)
        opt = RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True)
        tr = PairTrainer(emb, opt, QuotientLoss(), max_hops_sq=81.0)
        g = torch.Generator().manual_seed(5)
        src = torch.randperm(n, generator=g)[:G].int()
        offsets = torch.arange(G + 1, dtype=torch.int64) * per
        I = src.repeat_interleave(per).contiguous()
        J = torch.randint(n - 1, (G * per,), generator=g, dtype=torch.int32)
        J = torch.where(J >= I, J + 1, J).contiguous()
        hops = torch.randint(1, 10, (G * per,), generator=g, dtype=torch.uint8)
        pin = lambda t: t.pin_memory()  # noqa: E731
        losses = []
        for step in range(3):
            if grouped:
                losses.append(tr.step_host_grouped(pin(src), pin(offsets), pin(J), pin(hops), epoch=1))
            else:
                losses.append(tr.step_host(pin(I), pin(J), pin(hops), epoch=1))
        results.append((losses, emb.xs[0].detach().cpu().clone()))
    # the gradient scatter uses floating-point atomics, so two runs agree to rounding, not bit for bit
    assert max(abs(a - b) / abs(b) for a, b in zip(results[0][0], results[1][0])) < 1e-5
    assert rel_err(results[0][1], results[1][1]) < 1e-5


def _owner_update_worker(rank, world, port, out):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from graphembed.engine import PairTrainer
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam
    from graphembed.parallel import shard_range
    n, P = 512, 1 << 14
    g = torch.Generator().manual_seed(11)
    I = torch.randint(n, (P,), generator=g, dtype=torch.int32)
    J = ((I + 1 + torch.randint(n - 1, (P,), generator=g, dtype=torch.int32)) % n).int()
    hops = torch.randint(1, 9, (P,), generator=g, dtype=torch.uint8)
    res = {}
    for mode in ('single', 'owner', 'replica'):
        torch.manual_seed(0)
        emb = ManifoldEmbedding(n, [SymmetricPositiveDefinite(4)], device=dev, dtype=torch.float64)
        opt = RiemannianAdam(emb.xs, lr=0.01, max_grad_norm=100, exact=True)
        if mode == 'single':
            tr = PairTrainer(emb, opt, QuotientLoss(), max_hops_sq=64.0)
            lo, hi = 0, P
        else:
            tr = PairTrainer(emb, opt, QuotientLoss(), max_hops_sq=64.0, process_group=dist.group.WORLD,
                             owner_update=(mode == 'owner'))
            lo, hi = shard_range(P, rank, world)
        for _ in range(3):
            loss = tr.step(I[lo:hi].to(dev), J[lo:hi].to(dev), hops[lo:hi].to(dev), epoch=1)
        res[mode] = (float(loss.item()), emb.xs[0].detach().cpu())
    if rank == 0:
        out.put({k: (v[0], v[1].numpy()) for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_owner_update_matches_single_gpu():
    """2 ranks over NCCL: pair-sharded batch + (a) reduce-scatter / owner update / all-gather and (b) all-reduce /
    replicated update both reproduce the single-GPU trajectory (fp64, 1e-10)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_owner_update_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    l0, x0 = res['single']
    for mode in ('owner', 'replica'):
        l, x = res[mode]
        assert abs(l - l0) <= 1e-10 * abs(l0), mode
        assert np.abs(x - x0).max() <= 1e-10 * np.abs(x0).max(), mode


@pytest.mark.parametrize('kind,n,dtype', [('lorentz', 11, torch.float32), ('lorentz', 11, torch.float64),
                                          ('sphere', 8, torch.float32), ('euclidean', 5, torch.float64),
                                          ('universal', 6, torch.float64)])
@pytest.mark.parametrize('B', [512, 1500])
def test_node_batch_enumeration_matches_explicit_pair_list(kind, n, dtype, B):
    """Fused step of the vector manifolds over a node batch (TRIU enumeration, targets gathered from the dense matrix
    by node id) against the same pairs given as an explicit (i, j) LIST: same d2, loss, gradient (and curvature
    gradient)."""
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import Euclidean, Lorentz, Sphere, Universal
    torch.manual_seed(1)
    N = 5000
    hint = torch.empty(0, device=DEV, dtype=dtype)
    if kind == 'lorentz':
        man = Lorentz(n)
        x = man.rand(N, out=hint, ir=0.5)
    elif kind == 'sphere':
        man = Sphere(n)
        x = man.rand_uniform(N, out=hint)
    elif kind == 'euclidean':
        man = Euclidean(n)
        x = man.rand(N, out=hint, ir=1.0)
    else:
        man = Universal(n, c_init=0.3, device=DEV, dtype=dtype)
        x = man.rand(N, ir=0.5)
    x = x.contiguous()
    nodes = torch.randperm(N, device=DEV)[:B]
    dense = torch.rand(N, N, device=DEV, dtype=dtype) * 0.9 + 0.1
    iu = torch.triu_indices(B, B, 1, device=DEV)
    I, J = nodes[iu[0]], nodes[iu[1]]
    spec = _ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.3, eps=0.25)
    outs = []
    for pairs in (_ops.PairSet.triu(B, nodes, DEV), _ops.PairSet.from_lists(I, J, DEV)):
        grad = torch.zeros_like(x)
        cg = torch.zeros(1, dtype=torch.float64, device=DEV) if kind == 'universal' else None
        acc, d2 = _ops.pairs_loss_fused(man.spec, x, pairs, _ops.TargetSpec.dense(dense), spec, 0.9, grad, want_d2=True,
                                        c_grad=cg)
        outs.append((acc.clone(), d2, grad, cg))
    t = 1e-11 if dtype == torch.float64 else 2e-5
    assert torch.equal(outs[0][1], outs[1][1])
    assert rel_err(outs[0][0], outs[1][0]) < (1e-12 if dtype == torch.float64 else 1e-6)
    assert rel_err(outs[0][2], outs[1][2]) < t
    if kind == 'universal':
        assert rel_err(outs[0][3], outs[1][3]) < 1e-10


@pytest.mark.parametrize('kind,n', [('lorentz', 8), ('sphere', 4), ('euclidean', 12)])
def test_vector_reduction_rows_match_oracle(kind, n):
    """fp32 rows of 16-byte multiples (n % 4 == 0) accumulate their gradients with 128-bit reductions; the sampled-pair
    gradient (incl. many pairs hitting the same row, and a warp whose lanes all share the first endpoint) against the
    oracle's autograd."""
    from graphembed.manifolds import Euclidean, Lorentz, Sphere
    torch.manual_seed(4)
    N, P = 300, 6000
    g = torch.Generator().manual_seed(4)
    if kind == 'lorentz':
        man, orc = Lorentz(n), O.LorentzOracle(n)
        x = orc.rand(N, ir=0.7, dtype=torch.float32, generator=g)
    elif kind == 'sphere':
        man, orc = Sphere(n), O.SphereOracle(n)
        x = orc.rand_uniform(N, dtype=torch.float32, generator=g)
    else:
        man, orc = Euclidean(n), O.EuclideanOracle(n)
        x = torch.randn(N, n, generator=g)
    I = torch.randint(N, (P,), generator=g)
    I[:64] = 7  # two warps whose lanes share the `i` row: the warp-aggregated branch
    J = (I + 1 + torch.randint(N - 1, (P,), generator=g)) % N
    w = torch.rand(P, generator=g) + 0.5
    xr = x.clone().requires_grad_()
    (orc.dist2(xr[I], xr[J]) * w).sum().backward()
    xt = x.double().requires_grad_()
    (orc.dist2(xt[I], xt[J]) * w.double()).sum().backward()
    xd = x.to(DEV).requires_grad_()
    d2 = man.pair_dist2(xd, I.to(DEV), J.to(DEV))
    (d2 * w.to(DEV)).sum().backward()
    e_got, e_ref = parity_errors(xd.grad, xr.grad, xt.grad)
    assert e_got <= max(1e-5, 2 * e_ref), (e_got, e_ref)
