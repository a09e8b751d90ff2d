"""CPU (`-m "not gpu"`): the oracle restatement against the BASELINE-config-scale fixtures produced by the real
reference (tests/golden/make_golden_r2.py): config 1 at full size through the epoch-loop oracle, and the
teacher-forced 512-node steps of configs 2a / 3a / 3b on the shipped graphs.  (The same fixtures drive the CUDA path
in tests/test_gpu_configs.py.)"""
import os

import numpy as np
import pytest
import torch

import manifolds_oracle as O
from helpers import GOLDEN, rel_err


def _graph_hops(name, rows=None):
    """Hop counts from the C BFS oracle on the shipped edge list (all sources, or the given rows)."""
    import bfs_oracle
    from scipy.sparse import coo_matrix
    with np.load(os.path.join(GOLDEN, 'graphs', f'{name}.npz')) as z:
        n, edges = int(z['n']), z['edges'].astype(np.int64)
    a = coo_matrix((np.ones(2 * len(edges)), (np.r_[edges[:, 0], edges[:, 1]], np.r_[edges[:, 1], edges[:, 0]])),
                   shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.sort_indices()
    rowptr, colidx = a.indptr.astype(np.int32), a.indices.astype(np.int32)
    if rows is None:
        return bfs_oracle.all_pairs_hops(rowptr, colidx)
    return bfs_oracle.hops_from(rowptr, colidx, np.asarray(rows, dtype=np.int32))


def test_bfs_oracle_on_shipped_graphs():
    for name in ('tree1000', 'power'):
        with np.load(os.path.join(GOLDEN, 'graphs', f'{name}.npz')) as z:
            g = {k: z[k] for k in z.files}
        hops = _graph_hops(name)
        assert int(hops.max()) == int(g['max_hops'])
        assert np.array_equal(hops.sum(axis=1), g['row_sums'])
        assert np.array_equal(hops[g['sample_rows']], g['sample_levels'])


def test_config1_full_size_epoch_oracle_vs_reference_engine():
    import engine_oracle as E
    with np.load(os.path.join(GOLDEN, 'config1_tree1000_f64.npz')) as z:
        g = {k: torch.from_numpy(z[k]) for k in z.files}
    hops = _graph_hops('tree1000')
    cond = torch.from_numpy(hops[np.triu_indices(1000, 1)].astype(np.float64))
    step_fn = lambda f, o, x, grad, st: O.rsgd_step(o, x, grad, st, lr=0.01, max_grad_norm=20, exact=True)  # noqa: E731
    out = E.run_engine([O.SpdOracle(3)], [g['xs_x0'].clone()], [torch.tensor(0.5, dtype=torch.float64)], cond,
                       lambda t, m, alpha, epoch: O.quotient_loss(t, m, alpha, epoch), step_fn, 5, 1.0, seed=1234)
    assert np.allclose(out['step_loss'], g['xs_step_loss'].numpy(), rtol=1e-10)
    assert np.allclose(out['average_distortion'], g['xs_average_distortion'].numpy(), rtol=1e-10)
    assert np.allclose(out['pearsonr'], g['xs_pearsonr'].numpy(), rtol=1e-9)
    assert rel_err(out['xs'][0], g['xs_xT']) < 1e-10


STEPS = {
    '2a': ('power', lambda: [O.LorentzOracle(11)], 'f32'),
    '3a': ('facebook', lambda: [O.GrassmannOracle(6, 2)], 'f64'),
    '3b': ('facebook', lambda: [O.SpdOracle(3), O.LorentzOracle(5)], 'f32'),
}


@pytest.mark.parametrize('cfg', sorted(STEPS))
def test_config_step_oracle_vs_reference(cfg):
    """fp64 restatement of the step on the fixture's inputs against the reference's fp64 result (the `truth_` keys of
    the fp32 fixtures, the fixture itself for fp64 configs)."""
    gname, mk, tag = STEPS[cfg]
    with np.load(os.path.join(GOLDEN, f'config{cfg}_step_{tag}.npz')) as z:
        raw = {k: torch.from_numpy(z[k]) for k in z.files}
    pre = 'truth_' if tag == 'f32' else ''
    idx = raw['idx']
    hops = _graph_hops(gname, idx.numpy())[:, idx.numpy()].astype(np.float64)
    tsq = torch.from_numpy(hops).pow(2) / float(int(raw['max_hops']) ** 2)
    iu = torch.triu_indices(len(idx), len(idx), 1)
    targets = tsq[iu[0], iu[1]]
    oracles = mk()
    xs = [raw[f'x0_{f}'].double().requires_grad_() for f in range(len(oracles))]
    scales = [torch.tensor(0.5, dtype=torch.float64) for _ in oracles]
    m = O.product_dist2(oracles, xs, scales, lambda o, x: o.pdist2(x))
    loss = O.quotient_loss(targets, m, 1.0, 1)
    loss.backward()
    ref = float(raw[pre + 'loss'])
    assert abs(loss.item() - ref) <= 1e-10 * abs(ref)
    for f, (o, x) in enumerate(zip(oracles, xs)):
        sym = (lambda t: 0.5 * (t + t.transpose(-2, -1))) if isinstance(o, O.SpdOracle) else (lambda t: t)
        assert rel_err(sym(x.grad), sym(raw[pre + f'grad_{f}'])) < 1e-10
        x1 = O.radam_step(o, x.detach(), x.grad, {}, lr=0.01, max_grad_norm=100, exact=True)
        assert rel_err(x1, raw[pre + f'x1_{f}']) < 1e-10
