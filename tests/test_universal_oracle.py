"""Pins oracle/manifolds_oracle.py::UniversalOracle (kappa-stereographic manifold) to golden vectors
produced by the real reference (graphembed/manifolds/universal.py + impl/math.py).  CPU only."""
import numpy as np
import pytest
import torch

import manifolds_oracle as O
from helpers import load_golden, rel_err
from helpers_universal import OPTS, UNIVERSAL_CASES, oracle_for

TOL = {'f64': 1e-12, 'f32': 2e-5}


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_dist_grads_and_curvature_grad(name, tag):
    g = load_golden(name, tag)
    man, c_param = oracle_for(g)
    assert rel_err(man.c.detach(), g['c']) < 1e-7
    x, y = g['x'].clone().requires_grad_(), g['y'].clone().requires_grad_()
    d2 = man.dist2(x, y)
    (d2 * g['w']).sum().backward()
    assert rel_err(d2.detach(), g['dist2']) < TOL[tag]
    assert rel_err(x.grad, g['gx']) < TOL[tag] * 10
    assert rel_err(y.grad, g['gy']) < TOL[tag] * 10
    assert rel_err(c_param.grad, g['gc']) < TOL[tag] * 10
    assert rel_err(man.dist(g['x'], g['y']).detach(), g['dist']) < TOL[tag]


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_pdist_losses(name, tag):
    g = load_golden(name, tag)
    for lname, fn in (('quot', lambda t, m: O.quotient_loss(t, m, 1.7, 3)),
                      ('quot_l1', lambda t, m: O.quotient_loss(t, m, 1.7, 3, inc_l2=False)),
                      ('stress', O.stress_loss)):
        man, c_param = oracle_for(g)
        x = g['x'].clone().requires_grad_()
        pd2 = man.pdist2(x)
        loss = fn(g['targets'], 0.9 * pd2)
        loss.backward()
        assert rel_err(pd2.detach(), g['pdist2']) < TOL[tag]
        assert abs(loss.item() - g[f'loss_{lname}'].item()) <= TOL[tag] * 10 * abs(g[f'loss_{lname}'].item())
        assert rel_err(x.grad, g[f'grad_{lname}']) < TOL[tag] * 50
        assert rel_err(c_param.grad, g[f'gradc_{lname}']) < TOL[tag] * 50


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_point_ops(name, tag):
    g = load_golden(name, tag)
    man, _ = oracle_for(g)
    x, y, u, v, eg = g['x'], g['y'], g['u'], g['v'], g['eg']
    t = TOL[tag] * 20
    with torch.no_grad():
        assert rel_err(man.exp(x, u), g['exp']) < t
        assert rel_err(man.retr(x, u), g['retr']) < t
        assert rel_err(man.log(x, y), g['log']) < t * 50
        assert rel_err(man.proju(x, eg), g['proju']) < t
        assert rel_err(man.egrad2rgrad(x, eg), g['egrad2rgrad']) < t
        assert rel_err(man.transp(x, y, u), g['transp']) < t
        assert rel_err(man.inner(x, u, v), g['inner']) < t
        assert rel_err(man.norm(x, u).pow(2).reshape(-1), g['norm2'].reshape(-1)) < t
        assert rel_err(man.projx(g['far']), g['projx']) < t


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('oname', sorted(OPTS))
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_optimizer_trajectories(name, oname, tag):
    g = load_golden(name, tag)
    man, _ = oracle_for(g)
    kind, kw = OPTS[oname]
    x, state = g['x'].clone(), {}
    with torch.no_grad():
        for k in range(3):
            step = O.radam_step if kind == 'radam' else O.rsgd_step
            x = step(man, x, g['opt_grads'][k], state, **kw)
            assert rel_err(x, g[f'{oname}_x'][k]) < TOL[tag] * 50
    for key in ('exp_avg', 'exp_avg_sq', 'momentum_buffer'):
        if f'{oname}_{key}' in g:
            assert rel_err(state[key], g[f'{oname}_{key}']) < TOL[tag] * 50


def universal_training_oracle(g):
    """Restatement of tests/golden/make_golden.py::make_universal_training_run with the oracle pieces (the reference
    loop is products/embedding.py:24-57 + modules.py:102-105 + optim/radam.py:43-98 + torch.optim.SGD on c)."""
    ns = [g['x0_0'].shape[1], g['x0_1'].shape[1]]
    xs = [g['x0_0'].clone(), g['x0_1'].clone()]
    c_params = [torch.tensor([float(c)], dtype=torch.float64) for c in g['c0']]
    hops = g['hops_condensed']
    n = xs[0].shape[0]
    dense = torch.zeros(n, n, dtype=torch.float64)
    iu = torch.triu_indices(n, n, 1)
    t = O.dataset_targets(hops, torch.float64)
    dense[iu[0], iu[1]] = t
    dense = dense + dense.T
    perm = g['perm'].long()
    states = [{}, {}]
    losses, cs, cgrads, grad0 = [], [], [], None
    for step in range(4):
        idx = perm if step % 2 == 0 else perm[:20]
        cps = [c.clone().requires_grad_() for c in c_params]
        mans = [O.UniversalOracle(d, O.universal_get_c(cp)) for d, cp in zip(ns, cps)]
        xr = [x.clone().requires_grad_() for x in xs]
        m = sum(man.pdist2(x[idx]) for man, x in zip(mans, xr))
        loss = O.quotient_loss(O.batch_targets(dense, idx), m, 1.0, step + 1)
        loss.backward()
        if step == 0:
            grad0 = [x.grad.clone() for x in xr]
        cgrads.append([cp.grad.item() for cp in cps])
        with torch.no_grad():
            mans_ng = [O.UniversalOracle(d, O.universal_get_c(cp.detach())) for d, cp in zip(ns, c_params)]
            xs = [O.radam_step(man, x, xg.grad, st, lr=0.02, max_grad_norm=100, exact=True)
                  for man, x, xg, st in zip(mans_ng, xs, xr, states)]
            c_params = [cp - 1e-4 * cg.grad for cp, cg in zip(c_params, cps)]  # torch.optim.SGD(lr=1e-4)
            # products/embedding.py:36-46 stabilize(): norm constraint r_max = 5, then projx with the NEW curvature
            mans_new = [O.UniversalOracle(d, O.universal_get_c(cp)) for d, cp in zip(ns, c_params)]
            xs = [man.projx(x / (x.norm(p=2, dim=-1, keepdim=True) / 5.0).clamp(min=1)) for man, x in zip(mans_new, xs)]
        losses.append(loss.item())
        cs.append([cp.item() for cp in c_params])
    return dict(losses=np.array(losses), cs=np.array(cs), cgrads=np.array(cgrads), grad0=grad0, xs=xs)


def test_products_embedding_training_run():
    g = load_golden('universal_training_run', 'f64')
    out = universal_training_oracle(g)
    assert np.allclose(out['losses'], g['losses'].numpy(), rtol=1e-10)
    assert np.allclose(out['cgrads'], g['cgrads'].numpy(), rtol=1e-9)
    assert np.allclose(out['cs'], g['cs'].numpy(), rtol=1e-10)
    for i in range(2):
        assert rel_err(out['grad0'][i], g[f'grad0_{i}']) < 1e-10
        assert rel_err(out['xs'][i], g[f'xT_{i}']) < 1e-10


def products_engine_oracle(g, n_epochs=3, batch=40, drop_last_n=5):
    """Restatement of the reference's products.TrainingEngine run recorded in tests/golden/products_engine_run_f64.npz
    (make_golden.py::make_products_engine_run): train.py:198-228 step loop, stabilize after every epoch
    (products/train.py:13-14, products/embedding.py:36-46), validation metrics (train.py:230-265) and the curvature
    scalars `curv{i}` = get_K() (products/embedding.py:48-50)."""
    ns = [g['x0_0'].shape[1], g['x0_1'].shape[1]]
    xs = [g['x0_0'].clone(), g['x0_1'].clone()]
    c_params = [torch.tensor([float(c)], dtype=torch.float64) for c in g['c0']]
    n = xs[0].shape[0]
    t_cond = O.dataset_targets(g['hops_condensed'], torch.float64)
    dense = torch.zeros(n, n, dtype=torch.float64)
    iu = torch.triu_indices(n, n, 1)
    dense[iu[0], iu[1]] = t_cond
    dense = dense + dense.T
    states = [{}, {}]
    out = dict(step_loss=[], pearsonr=[], average_distortion=[], curv0=[], curv1=[])
    torch.manual_seed(1234)
    for epoch in range(1, n_epochs + 1):
        perm = torch.randperm(n)
        for i in range(0, n, batch):
            idx = perm[i:i + batch]
            if len(idx) < drop_last_n:
                break
            cps = [c.clone().requires_grad_() for c in c_params]
            mans = [O.UniversalOracle(d, O.universal_get_c(cp)) for d, cp in zip(ns, cps)]
            xr = [x.clone().requires_grad_() for x in xs]
            m = sum(man.pdist2(x[idx]) for man, x in zip(mans, xr))
            loss = O.quotient_loss(O.batch_targets(dense, idx), m, 1.0, epoch)
            loss.backward()
            with torch.no_grad():
                plain = [O.UniversalOracle(d, O.universal_get_c(cp.detach())) for d, cp in zip(ns, c_params)]
                xs = [O.radam_step(man, x, xg.grad, st, lr=0.02, max_grad_norm=100, exact=True)
                      for man, x, xg, st in zip(plain, xs, xr, states)]
                c_params = [cp - 1e-4 * cg.grad for cp, cg in zip(c_params, cps)]
            out['step_loss'].append(loss.item() / len(idx))
        with torch.no_grad():
            mans = [O.UniversalOracle(d, O.universal_get_c(cp)) for d, cp in zip(ns, c_params)]
            xs = [man.projx(x / (x.norm(p=2, dim=-1, keepdim=True) / 5.0).clamp(min=1)) for man, x in zip(mans, xs)]
            md = sum(man.pdist2(x) for man, x in zip(mans, xs)).sqrt()
            gd = t_cond.sqrt()
            out['pearsonr'].append(O.pearsonr(md, gd).item())
            out['average_distortion'].append(O.average_distortion(md, gd).item())
            out['curv0'].append(-O.universal_get_c(c_params[0]).item())
            out['curv1'].append(-O.universal_get_c(c_params[1]).item())
    out.update(xs=xs, cT=[cp.item() for cp in c_params])
    return out


def test_products_training_engine_run():
    g = load_golden('products_engine_run', 'f64')
    out = products_engine_oracle(g)
    for key in ('step_loss', 'pearsonr', 'average_distortion', 'curv0', 'curv1'):
        assert np.allclose(out[key], g[key].numpy(), rtol=1e-9), key
    assert np.allclose(out['cT'], g['cT'].numpy(), rtol=1e-10)
    for i in range(2):
        assert rel_err(out['xs'][i], g[f'xT_{i}']) < 1e-10
