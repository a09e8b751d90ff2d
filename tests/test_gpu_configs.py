"""GPU parity at BASELINE-config scale against the REAL reference (fixtures: tests/golden/make_golden_r2.py).

* BFS hop counts of the four shipped graphs (tree1000, power, facebook, condmat -- edge lists as the reference's
  loader numbers them, tests/golden/graphs/*.npz) bit-exact against scipy's BFS: sha256 of the full hop matrix, row
  sums, histogram, sample rows.
* BASELINE config 1 at full size: data/tree1000.edges.gz -> SPD 3x3, fp64, all 499 500 pairs per step, QuotientLoss,
  RiemannianSGD(lr .01, exact, clip 20), 5 free-running epochs through TrainingEngine against the reference's own
  TrainingEngine (train.py:198-265): per-epoch loss, distortion, pearsonr, final points at 1e-10 -- also with the
  scale parameter trained by a second RiemannianSGD group as experiments/run_grid.py:30-33 does.
* One teacher-forced RiemannianAdam step of configs 2a / 2b / 3a / 3b / 4 on a 512-node batch of the shipped graph
  (run_grid.py:25-28,131): loss, gradient rows, scale gradient, updated points and Adam state; fp64 at 1e-10, fp32
  through the error budget of helpers.assert_parity against the reference's fp64 step on the same inputs.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, assert_parity, assert_parity_scalar, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def load_graph(name):
    with np.load(os.path.join(GOLDEN, 'graphs', f'{name}.npz')) as z:
        return {k: z[k] for k in z.files}


_levels_cache = {}


def graph_levels(name):
    """(N, N) uint8 hop matrix from the BFS kernel (cached per process: condmat's is 456 MB)."""
    from graphembed.data import bfs_levels, edges_to_csr
    if name not in _levels_cache:
        _levels_cache.clear()
        g = load_graph(name)
        rowptr, colidx = edges_to_csr(int(g['n']), g['edges'])
        _levels_cache[name] = bfs_levels(rowptr, colidx, device=DEV)
    return _levels_cache[name]


@pytest.mark.parametrize('name', ['tree1000', 'power', 'facebook', 'condmat'])
def test_bfs_bit_exact_on_shipped_graphs(name):
    g = load_graph(name)
    lv = graph_levels(name)
    n = int(g['n'])
    assert lv.dtype == torch.uint8 and lv.shape == (n, n)
    assert int(lv.max()) == int(g['max_hops'])
    assert np.array_equal(lv.sum(dim=1, dtype=torch.int64).cpu().numpy(), g['row_sums'])
    assert np.array_equal(torch.bincount(lv.reshape(-1).long(), minlength=256).cpu().numpy(), g['hist'])
    rows = torch.from_numpy(g['sample_rows']).to(DEV)
    assert np.array_equal(lv[rows].cpu().numpy(), g['sample_levels'])
    host = lv.cpu().numpy()
    assert hashlib.sha256(host.tobytes()).hexdigest() == str(g['sha256'])
    assert np.array_equal(host, host.T)


def _dataset(name, dtype):
    """GraphDataset (dense squared / max-normalised targets, data/dataset.py:9-13) from the BFS kernel's levels."""
    from graphembed.data import GraphDataset
    from graphembed.data.graph import levels_to_condensed
    cond = levels_to_condensed(graph_levels(name), dtype)
    return GraphDataset(cond)


@pytest.mark.parametrize('tag', ['xs', 'curv'])
def test_config1_tree1000_five_epochs_vs_reference_engine(tag, tmp_path):
    from graphembed.manifolds import SymmetricPositiveDefinite
    from graphembed.modules import ManifoldEmbedding
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianSGD
    from graphembed.train import TrainingEngine
    with np.load(os.path.join(GOLDEN, 'config1_tree1000_f64.npz')) as z:
        g = {k: z[k] for k in z.files}
    n = 1000
    ds = _dataset('tree1000', torch.float64)
    emb = ManifoldEmbedding(n, [SymmetricPositiveDefinite(3)], device=DEV, dtype=torch.float64)
    with torch.no_grad():
        emb.xs[0].copy_(torch.from_numpy(g[f'{tag}_x0']).to(DEV))
    groups = [dict(params=emb.xs, lr=0.01, exact=True, max_grad_norm=20)]
    if tag == 'curv':
        groups.append(dict(params=emb.curvature_params, lr=1e-4, max_grad_norm=500))
    opt = RiemannianSGD(groups, lr=0.01)
    obj = QuotientLoss()
    eng = TrainingEngine(embedding=emb, optimizer=opt, objective_fn=obj, alpha=1.0, n_epochs=5, val_every_epochs=1,
                         save_dir=str(tmp_path), tensorboard=False)
    torch.manual_seed(1234)
    eng(ds)
    hist = eng.writer.history
    got = np.array([v for _, v in hist[str(obj)]])
    assert got.shape == g[f'{tag}_step_loss'].shape
    assert np.allclose(got, g[f'{tag}_step_loss'], rtol=1e-10, atol=0), (got, g[f'{tag}_step_loss'])
    for m in ('pearsonr', 'average_distortion'):
        got = np.array([v for _, v in hist[m]])
        assert np.allclose(got, g[f'{tag}_{m}'], rtol=1e-10, atol=0), (m, got, g[f'{tag}_{m}'])
    assert rel_err(emb.xs[0].data, torch.from_numpy(g[f'{tag}_xT'])) < 1e-10
    assert abs(float(emb.scales[0].detach()) - float(g[f'{tag}_scaleT'])) <= 1e-10 * abs(float(g[f'{tag}_scaleT']))
    if tag == 'curv':
        assert abs(float(emb.scales[0].detach()) - 0.5) > 1e-3  # the scale really was trained (and tracked by the kernels)


STEP_CONFIGS = {
    # tag: (graph, factors, dtype tag)
    '2a': ('power', [('lorentz', dict(n=11))], 'f32'),
    '2b': ('power', [('spd', dict(n=4, use_stein_div=True))], 'f32'),
    '3a': ('facebook', [('grassmann', dict(n=6, p=2))], 'f64'),
    '3b': ('facebook', [('spd', dict(n=3)), ('lorentz', dict(n=5))], 'f32'),
    '4': ('condmat', [('spd', dict(n=6))], 'f32'),
}


def _factor(fam, kw):
    from graphembed import manifolds as M
    if fam == 'spd':
        return M.SymmetricPositiveDefinite(**kw)
    if fam == 'lorentz':
        return M.Lorentz(kw['n'])
    return M.Grassmann(kw['n'], kw['p'])


@pytest.mark.parametrize('with_curv', [False, True])
@pytest.mark.parametrize('cfg', sorted(STEP_CONFIGS))
def test_config_step_vs_reference(cfg, with_curv):
    from graphembed.modules import BatchedObjective, ManifoldEmbedding
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam
    gname, factors, tag = STEP_CONFIGS[cfg]
    dtype = torch.float32 if tag == 'f32' else torch.float64
    with np.load(os.path.join(GOLDEN, f'config{cfg}_step_{tag}.npz')) as z:
        raw = {k: torch.from_numpy(z[k]) for k in z.files}
    pre = 'curv_' if with_curv else ''
    g = {k[len(pre):]: v for k, v in raw.items() if k.startswith(pre) and not k.startswith(pre + 'truth_')
         and (with_curv or not k.startswith('curv_'))}
    T = {k[len(pre + 'truth_'):]: v for k, v in raw.items() if k.startswith(pre + 'truth_')} if tag == 'f32' else None
    idx = raw['idx'].to(DEV)
    ds = _dataset(gname, dtype)
    n = len(ds)
    assert int(raw['max_hops']) == int(load_graph(gname)['max_hops'])
    torch.manual_seed(0)
    mans = [_factor(f, kw) for f, kw in factors]
    emb = ManifoldEmbedding(n, mans, device=DEV, dtype=dtype)
    with torch.no_grad():
        for f, x in enumerate(emb.xs):
            x[idx] = raw[f'x0_{f}'].to(DEV)
    groups = [dict(params=emb.xs, lr=0.01, exact=True, max_grad_norm=100)]
    if with_curv:
        groups.append(dict(params=emb.curvature_params, lr=0.01))
    opt = RiemannianAdam(groups)
    bobj = BatchedObjective(QuotientLoss(), ds, emb)
    loss = bobj(idx, alpha=1.0, epoch=1).sum()
    opt.zero_grad()
    loss.backward()
    assert_parity_scalar(loss.item(), g, 'loss', tag, T)
    sym = lambda t: 0.5 * (t + t.transpose(-2, -1))  # noqa: E731
    for f, x in enumerate(emb.xs):
        fix = sym if factors[f][0] == 'spd' else None
        assert_parity(x.grad[idx], g, f'grad_{f}', tag, T, fix)
        rest = x.grad.clone()
        rest[idx] = 0
        assert not rest.any()
    for f, s in enumerate(emb.scales):
        assert_parity(s.grad.reshape(1), {'k': g[f'scale_grad_{f}'].reshape(1)}, 'k', tag,
                      None if T is None else {'k': T[f'scale_grad_{f}'].reshape(1)}, what=f'scale_grad_{f}')
    opt.step()
    for f, x in enumerate(emb.xs):
        assert_parity(x.data[idx], g, f'x1_{f}', tag, T)
        st = opt.state[x]
        assert_parity(st['exp_avg'][idx], g, f'exp_avg_{f}', tag, T)
        assert_parity(st['exp_avg_sq'][idx], g, f'exp_avg_sq_{f}', tag, T)
    for f, s in enumerate(emb.scales):
        assert_parity(s.data.reshape(1), {'k': g[f'scale1_{f}'].reshape(1)}, 'k', tag,
                      None if T is None else {'k': T[f'scale1_{f}'].reshape(1)}, what=f'scale1_{f}')
