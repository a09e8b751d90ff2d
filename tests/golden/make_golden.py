"""Generates tests/golden/*.npz by running the REAL reference (dalab/matrix-manifolds
`graphembed`, imported from /root/reference via oracle/ref_import.py) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
The fixtures are small (a dozen points per case) and are committed; the GPU box has
no /root/reference, so the parity tests read these files instead.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
import ref_import  # noqa: E402

ref_import.load()
from graphembed.manifolds import (SymmetricPositiveDefinite, Lorentz, Sphere, Grassmann, Euclidean)  # noqa: E402
from graphembed.modules import ManifoldParameter, ManifoldEmbedding, BatchedObjective  # noqa: E402
from graphembed.objectives import QuotientLoss, StressLoss  # noqa: E402
from graphembed.optim import RiemannianAdam, RiemannianSGD  # noqa: E402
from graphembed.data.dataset import GraphDataset  # noqa: E402

N = 12
CASES = {
    # name: (constructor, ctor kwargs, init fn name, init kwargs)
    'spd2': (SymmetricPositiveDefinite, dict(n=2), 'rand', dict(ir=1.0)),
    'spd2_exact': (SymmetricPositiveDefinite, dict(n=2, fast_symeig=False, fast_chol=False), 'rand', dict(ir=1.0)),
    'spd3': (SymmetricPositiveDefinite, dict(n=3), 'rand', dict(ir=1.0)),
    'spd3_default_init': (SymmetricPositiveDefinite, dict(n=3), 'rand', dict()),
    'spd4': (SymmetricPositiveDefinite, dict(n=4), 'rand', dict(ir=1.0)),
    'spd4_default_init': (SymmetricPositiveDefinite, dict(n=4), 'rand', dict()),
    'spd6': (SymmetricPositiveDefinite, dict(n=6), 'rand', dict(ir=1.0)),
    'stein2': (SymmetricPositiveDefinite, dict(n=2, use_stein_div=True), 'rand', dict(ir=1.0)),
    'stein4': (SymmetricPositiveDefinite, dict(n=4, use_stein_div=True), 'rand', dict(ir=1.0)),
    'lorentz11': (Lorentz, dict(n=11), 'rand', dict(ir=1.0)),
    'lorentz5_default_init': (Lorentz, dict(n=5), 'rand', dict()),
    'sphere5': (Sphere, (5,), 'rand_uniform', dict()),
    'euclidean7': (Euclidean, (7,), 'rand', dict(ir=1.0)),
    'grassmann6_2': (Grassmann, dict(n=6, p=2), 'rand_uniform', dict()),
    'grassmann7_3': (Grassmann, dict(n=7, p=3), 'rand_uniform', dict()),
}


def build(ctor, kw):
    return ctor(*kw) if isinstance(kw, tuple) else ctor(**kw)


def make_case(name, dtype, seed):
    ctor, kw, init, ikw = CASES[name]
    torch.set_default_dtype(dtype)
    torch.manual_seed(seed)
    man = build(ctor, kw)
    x = getattr(man, init)(N, **ikw).contiguous()
    y = getattr(man, init)(N, **ikw).contiguous()
    out = dict(x=x.numpy(), y=y.numpy())
    # elementwise dist^2 + gradients of a weighted sum
    w = torch.linspace(0.5, 1.5, N, dtype=dtype)
    xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
    d2 = man.dist(xr, yr, squared=True)
    (d2 * w).sum().backward()
    out.update(dist2=d2.detach().numpy(), w=w.numpy(), gx=xr.grad.numpy(), gy=yr.grad.numpy())
    # pdist^2, QuotientLoss (both terms) and Stress against synthetic targets, gradients
    P = N * (N - 1) // 2
    g = torch.rand(P, dtype=dtype) * 0.9 + 0.1
    out['targets'] = g.numpy()
    for lname, fn, kwargs in (('quot', QuotientLoss(), dict(epoch=3, alpha=1.7)),
                              ('quot_l1', QuotientLoss(inc_l2=False), dict(epoch=3, alpha=1.7)),
                              ('stress', StressLoss(), dict())):
        xr = x.clone().requires_grad_()
        pd2 = man.pdist(xr, squared=True)
        loss = fn(g, 0.9 * pd2, **kwargs)
        loss.backward()
        out[f'pdist2'] = pd2.detach().numpy()
        out[f'loss_{lname}'] = np.array(loss.item())
        out[f'grad_{lname}'] = xr.grad.numpy()
    # point ops
    u = man.randvec(x, 0.7).contiguous()
    v = man.randvec(x, 0.3).contiguous()
    eg = torch.randn_like(x)
    out.update(u=u.numpy(), v=v.numpy(), eg=eg.numpy(), exp=man.exp(x, u).numpy(), retr=man.retr(x, u).numpy(),
               log=man.log(x, y).numpy(), proju=man.proju(x, eg.clone()).numpy(),
               egrad2rgrad=man.egrad2rgrad(x, eg.clone()).numpy(), transp=man.transp(x, y, u).numpy(),
               inner=man.inner(x, u, v).numpy(), norm2=man.norm(x, u, squared=True).numpy())
    # optimizer trajectories: 3 steps with fixed Euclidean gradients (the 2nd one zero)
    grads = [torch.randn_like(x), torch.zeros_like(x), torch.randn_like(x)]
    out['opt_grads'] = np.stack([t.numpy() for t in grads])
    for oname, mk in (('radam_clip', lambda ps: RiemannianAdam(ps, lr=0.05, max_grad_norm=1.5)),
                      ('radam_exact', lambda ps: RiemannianAdam(ps, lr=0.05, exact=True)),
                      ('rsgd_exact_clip', lambda ps: RiemannianSGD(ps, lr=0.05, max_grad_norm=0.5, exact=True)),
                      ('rsgd_momentum', lambda ps: RiemannianSGD(ps, lr=0.05, momentum=0.9, dampening=0.1))):
        p = ManifoldParameter(x.clone(), manifold=man)
        opt = mk([p])
        traj = []
        for gk in grads:
            p.grad = gk.clone()
            opt.step()
            traj.append(p.data.clone().numpy())
        out[f'{oname}_x'] = np.stack(traj)
        st = opt.state[p]
        for key in ('exp_avg', 'exp_avg_sq', 'momentum_buffer'):
            if key in st:
                out[f'{oname}_{key}'] = st[key].numpy()
    return out


def make_training_run():
    """BASELINE config-1-shaped run on a tiny tree: SPD 3x3, fp64, full-batch QuotientLoss, RSGD(exact, clip 20),
    through the reference's own ManifoldEmbedding / BatchedObjective, 4 steps; plus a product
    SPD3 x Lorentz(5) run with RAdam."""
    import networkx as nx
    from scipy.sparse.csgraph import shortest_path
    torch.set_default_dtype(torch.float64)
    g = nx.balanced_tree(2, 4)  # 31 nodes
    n = g.number_of_nodes()
    hops = shortest_path(nx.to_scipy_sparse_array(g), unweighted=True)
    iu = np.triu_indices(n, 1)
    cond = torch.tensor(hops[iu])
    out = dict(edges=np.array(g.edges()), hops_condensed=cond.numpy())
    for tag, mans, mkopt in (
            ('spd3_rsgd', lambda: [SymmetricPositiveDefinite(3)],
             lambda ps: RiemannianSGD(ps, lr=0.01, max_grad_norm=20, exact=True)),
            ('prod_radam', lambda: [SymmetricPositiveDefinite(3), Lorentz(5)],
             lambda ps: RiemannianAdam(ps, lr=0.01, max_grad_norm=100, exact=True))):
        torch.manual_seed(42)
        ds = GraphDataset(cond.clone())
        emb = ManifoldEmbedding(n, mans())
        for i, x in enumerate(emb.xs):
            out[f'{tag}_x0_{i}'] = x.data.clone().numpy()
        opt = mkopt(emb.xs)
        bobj = BatchedObjective(QuotientLoss(), ds, emb)
        losses = []
        perm = torch.randperm(n)
        out[f'{tag}_perm'] = perm.numpy()
        for step in range(4):
            idx = perm if step % 2 == 0 else perm[:20]  # full batch and a node mini-batch
            loss = bobj(idx, alpha=1.0, epoch=step + 1).sum()
            opt.zero_grad()
            loss.backward()
            if step == 0:
                for i, x in enumerate(emb.xs):
                    out[f'{tag}_grad0_{i}'] = x.grad.clone().numpy()
            opt.step()
            losses.append(loss.item())
        out[f'{tag}_losses'] = np.array(losses)
        for i, x in enumerate(emb.xs):
            out[f'{tag}_xT_{i}'] = x.data.clone().numpy()
    return out


class _Recorder:
    """Stands in for tensorboard's SummaryWriter inside the reference TrainingEngine."""
    last = None

    def __init__(self, log_dir=None):
        self.scalars = {}
        _Recorder.last = self

    def add_scalar(self, tag, value, step):
        self.scalars.setdefault(tag, []).append((int(step), float(value)))

    def add_figure(self, *a, **k):
        pass

    add_histogram = add_figure


def make_objectives(dtype):
    """KL divergence with the stochastic-neighbour model (objectives.py:48-76), both directions: value and gradient."""
    from graphembed.objectives import KLDiveregenceLoss
    torch.set_default_dtype(dtype)
    torch.manual_seed(11)
    n = 13
    P = n * (n - 1) // 2
    g = torch.rand(P, dtype=dtype) * 0.9 + 0.1
    m0 = torch.rand(P, dtype=dtype) * 2.0 + 0.05
    out = dict(g=g.numpy(), m=m0.numpy(), alpha=np.array(3.5))
    for inc in (True, False):
        m = m0.clone().requires_grad_()
        loss = KLDiveregenceLoss('sne', inclusive=inc)(g, m, alpha=3.5)
        loss.backward()
        out[f'kl_{int(inc)}_loss'] = np.array(loss.item())
        out[f'kl_{int(inc)}_grad'] = m.grad.numpy()
    # validation metrics on plain vectors (metrics.py:13-17,46-56)
    import graphembed.metrics as M
    out['pearsonr'] = np.array(M.pearsonr(m0, g).item())
    out['average_distortion'] = np.array(M.average_distortion(m0, g).item())
    return out


def make_engine_runs():
    """The reference's own TrainingEngine (train.py) for 3 epochs with validation every epoch on a 63-node tree:
    (a) BASELINE config 1 in miniature: SPD 3x3, QuotientLoss, RSGD(exact, clip 20), full batch;
    (b) the same with node mini-batches of 52 (tail batch of 11 < drop_last_n is dropped);
    (c) example_config.yaml in miniature: SPD 2x2, KL/SNE loss, alpha 10, RAdam(clip 100), stabilize every 2 epochs;
    (d) product SPD3 x Lorentz5, QuotientLoss, RAdam exact."""
    import tempfile
    import types
    import networkx as nx
    from scipy.sparse.csgraph import shortest_path
    import graphembed.train as T
    from graphembed.objectives import KLDiveregenceLoss
    plt = sys.modules['matplotlib.pyplot']
    plt.scatter = plt.gcf = plt.close = lambda *a, **k: None
    T.SummaryWriter = _Recorder
    torch.set_default_dtype(torch.float64)
    g = nx.balanced_tree(2, 5)  # 63 nodes
    n = g.number_of_nodes()
    hops = shortest_path(nx.to_scipy_sparse_array(g), unweighted=True)
    cond = torch.tensor(hops[np.triu_indices(n, 1)])
    out = dict(edges=np.array(g.edges()), hops_condensed=cond.numpy())
    runs = {
        'spd3_full': (lambda: [SymmetricPositiveDefinite(3)], QuotientLoss,
                      lambda ps: RiemannianSGD(ps, lr=0.01, max_grad_norm=20, exact=True), dict(alpha=1.0)),
        'spd3_batched': (lambda: [SymmetricPositiveDefinite(3)], QuotientLoss,
                         lambda ps: RiemannianSGD(ps, lr=0.01, max_grad_norm=20, exact=True),
                         dict(alpha=1.0, batch_size=52)),
        'spd2_kl': (lambda: [SymmetricPositiveDefinite(2)], lambda: KLDiveregenceLoss('sne', inclusive=True),
                    lambda ps: RiemannianAdam(ps, lr=0.01, max_grad_norm=100, exact=False),
                    dict(alpha=10.0, stabilize_every_epochs=2)),
        'prod_radam': (lambda: [SymmetricPositiveDefinite(3), Lorentz(5)], QuotientLoss,
                       lambda ps: RiemannianAdam(ps, lr=0.01, max_grad_norm=100, exact=True), dict(alpha=1.0)),
    }
    for tag, (mans, mkobj, mkopt, extra) in runs.items():
        torch.manual_seed(42)
        emb = ManifoldEmbedding(n, mans())
        for i, x in enumerate(emb.xs):
            out[f'{tag}_x0_{i}'] = x.data.clone().numpy()
        obj = mkobj()
        with tempfile.TemporaryDirectory() as tmp:
            eng = T.TrainingEngine(embedding=emb, optimizer=mkopt(emb.xs), objective_fn=obj, n_epochs=3,
                                   val_every_epochs=1, save_dir=tmp, **extra)
            torch.manual_seed(1234)  # the engine draws one randperm per epoch from the global CPU generator
            eng(GraphDataset(cond.clone()))
            out[f'{tag}_files'] = np.array(sorted(os.listdir(tmp)))
            best = [f for f in os.listdir(tmp) if f.startswith('best_loss_')][0]
            out[f'{tag}_best'] = np.array([float(best.split('_')[-1]), float(open(os.path.join(tmp, best)).read())])
            sd = torch.load(os.path.join(tmp, 'best_embedding.pth'))
            out[f'{tag}_state_keys'] = np.array(sorted(sd.keys()))
        rec = _Recorder.last.scalars
        out[f'{tag}_step_loss'] = np.array([v for _, v in rec[str(obj)]])
        out[f'{tag}_pearsonr'] = np.array([v for _, v in rec['pearsonr']])
        out[f'{tag}_average_distortion'] = np.array([v for _, v in rec['average_distortion']])
        for i, x in enumerate(emb.xs):
            out[f'{tag}_xT_{i}'] = x.data.clone().numpy()
    return out


UNIVERSAL_CASES = {
    # name: (n, ctor kwargs, init radius): hyperbolic (c > 0), spherical (c < 0), sign-fixed softplus parametrisation
    'universal5_pos': (5, dict(c_init=0.5), 0.5),
    'universal5_neg': (5, dict(c_init=-0.7), 0.5),
    'universal3_fixed_sign': (3, dict(c_init=-0.3, keep_sign_fixed=True), 0.4),
    'universal4_default_init': (4, dict(), 1e-2),
}


def make_universal(name, dtype, seed):
    """Universal (kappa-stereographic) manifold, graphembed/manifolds/universal.py: distances and gradients w.r.t.
    the points AND the curvature parameter, point ops, optimizer trajectories, and a products.Embedding training
    run with a curvature optimizer."""
    from graphembed.manifolds import Universal
    n, kw, ir = UNIVERSAL_CASES[name]
    torch.set_default_dtype(dtype)
    torch.manual_seed(seed)
    man = Universal(n, **kw)
    with torch.no_grad():  # rand() -> projx(inplace) set_()s a tensor that depends on c (not differentiable today)
        x = man.rand(N, ir=ir).contiguous()
        y = man.rand(N, ir=ir).contiguous()
    out = dict(x=x.numpy(), y=y.numpy(), c_param=man.c.detach().numpy(), c=man.get_c().detach().numpy(),
               c_min=np.array(man.c_min), sign=np.array(0 if man.sign is None else man.sign))
    w = torch.linspace(0.5, 1.5, N, dtype=dtype)
    xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
    d2 = man.dist(xr, yr, squared=True)
    (d2 * w).sum().backward()
    out.update(dist2=d2.detach().numpy(), w=w.numpy(), gx=xr.grad.numpy(), gy=yr.grad.numpy(),
               gc=man.c.grad.clone().numpy())
    man.c.grad = None
    with torch.no_grad():
        out['dist'] = man.dist(x, y).numpy()
    P = N * (N - 1) // 2
    g = torch.rand(P, dtype=dtype) * 0.9 + 0.1
    out['targets'] = g.numpy()
    for lname, fn, kwargs in (('quot', QuotientLoss(), dict(epoch=3, alpha=1.7)),
                              ('quot_l1', QuotientLoss(inc_l2=False), dict(epoch=3, alpha=1.7)),
                              ('stress', StressLoss(), dict())):
        xr = x.clone().requires_grad_()
        pd2 = man.pdist(xr, squared=True)
        loss = fn(g, 0.9 * pd2, **kwargs)
        loss.backward()
        out['pdist2'] = pd2.detach().numpy()
        out[f'loss_{lname}'] = np.array(loss.item())
        out[f'grad_{lname}'] = xr.grad.numpy()
        out[f'gradc_{lname}'] = man.c.grad.clone().numpy()
        man.c.grad = None
    with torch.no_grad():
        u = (torch.randn_like(x) * 0.3).contiguous()
        v = (torch.randn_like(x) * 0.2).contiguous()
        eg = torch.randn_like(x)
        far = (torch.randn(N, n) * 3.0).contiguous()  # outside the ball for c > 0
        far_proj = far.clone()
        man.projx(far_proj, inplace=True)
        out.update(u=u.numpy(), v=v.numpy(), eg=eg.numpy(), far=far.numpy(), projx=far_proj.numpy(),
                   exp=man.exp(x, u).numpy(), retr=man.retr(x, u).numpy(), log=man.log(x, y).numpy(),
                   proju=man.proju(x, eg.clone()).numpy(), egrad2rgrad=man.egrad2rgrad(x, eg.clone()).numpy(),
                   transp=man.transp(x, y, u).numpy(), inner=man.inner(x, u, v).numpy(),
                   norm2=man.norm(x, u, squared=True).numpy())
        grads = [torch.randn_like(x), torch.zeros_like(x), torch.randn_like(x)]
        out['opt_grads'] = np.stack([t.numpy() for t in grads])
        for oname, mk in (('radam_clip', lambda ps: RiemannianAdam(ps, lr=0.05, max_grad_norm=1.5)),
                          ('radam_exact', lambda ps: RiemannianAdam(ps, lr=0.05, exact=True)),
                          ('rsgd_exact_clip', lambda ps: RiemannianSGD(ps, lr=0.05, max_grad_norm=0.5, exact=True)),
                          ('rsgd_momentum', lambda ps: RiemannianSGD(ps, lr=0.05, momentum=0.9, dampening=0.1))):
            p = ManifoldParameter(x.clone(), manifold=man)
            opt = mk([p])
            traj = []
            for gk in grads:
                p.grad = gk.clone()
                opt.step()
                traj.append(p.data.clone().numpy())
            out[f'{oname}_x'] = np.stack(traj)
            st = opt.state[p]
            for key in ('exp_avg', 'exp_avg_sq', 'momentum_buffer'):
                if key in st:
                    out[f'{oname}_{key}'] = st[key].detach().numpy()
    return out


def make_universal_training_run():
    """products.Embedding (two Universal factors, one hyperbolic one spherical) on a 31-node tree, fp64:
    QuotientLoss through the reference's BatchedObjective, RAdam on the points + SGD on the curvatures, stabilize()
    after every step; 4 steps (full batch / node mini-batch alternating)."""
    import networkx as nx
    from scipy.sparse.csgraph import shortest_path
    from graphembed.products import Embedding
    torch.set_default_dtype(torch.float64)
    g = nx.balanced_tree(2, 4)
    n = g.number_of_nodes()
    hops = shortest_path(nx.to_scipy_sparse_array(g), unweighted=True)
    cond = torch.tensor(hops[np.triu_indices(n, 1)])
    out = dict(edges=np.array(g.edges()), hops_condensed=cond.numpy())
    torch.manual_seed(42)
    ds = GraphDataset(cond.clone())
    with torch.no_grad():
        emb = Embedding(n, [3, 2], c_init=0.4)
        emb.manifolds[1].c.fill_(-0.6)
        for x in emb.xs:  # spread the points (the default ir=1e-2 start is degenerate for 4 steps)
            x.mul_(30.0)
            x.proj_()
    out['c0'] = np.array([m.c.item() for m in emb.manifolds])
    for i, x in enumerate(emb.xs):
        out[f'x0_{i}'] = x.data.clone().numpy()
    opt = RiemannianAdam(emb.xs, lr=0.02, max_grad_norm=100, exact=True)
    copt = torch.optim.SGD(list(emb.curvature_params), lr=1e-4)
    bobj = BatchedObjective(QuotientLoss(), ds, emb)
    perm = torch.randperm(n)
    out['perm'] = perm.numpy()
    losses, cs, cgrads = [], [], []
    for step in range(4):
        idx = perm if step % 2 == 0 else perm[:20]
        loss = bobj(idx, alpha=1.0, epoch=step + 1).sum()
        opt.zero_grad()
        copt.zero_grad()
        loss.backward()
        if step == 0:
            for i, x in enumerate(emb.xs):
                out[f'grad0_{i}'] = x.grad.clone().numpy()
        cgrads.append([m.c.grad.item() for m in emb.manifolds])
        with torch.no_grad():
            opt.step()
        copt.step()
        emb.stabilize()
        losses.append(loss.item())
        cs.append([m.c.item() for m in emb.manifolds])
    out.update(losses=np.array(losses), cs=np.array(cs), cgrads=np.array(cgrads))
    for i, x in enumerate(emb.xs):
        out[f'xT_{i}'] = x.data.clone().numpy()
    return out


def make_precision():
    """Ranking metrics: the reference's own pure-Python mean average precision (graphembed/metrics.py:61-96, the
    regression anchor its tests hold for FastPrecision, tests/test_metrics.py:14-23) on Erdos-Renyi graphs with random
    fp32 distances -- exactly the reference test's inputs -- plus the graph distances for the F1 == 1 known answer."""
    import networkx as nx
    from scipy.sparse.csgraph import shortest_path
    from scipy.spatial.distance import squareform
    from graphembed.metrics import py_mean_average_precision
    out = {}
    rng = np.random.RandomState(5)
    for tag, (n, p) in dict(a=(50, 0.1), b=(100, 0.1), c=(100, 0.5), d=(300, 0.02)).items():
        g = nx.erdos_renyi_graph(n, p, seed=int(rng.randint(1 << 30)))
        comp = max(nx.connected_components(g), key=len)  # tests/conftest.py rand_graph keeps the largest component
        g = nx.convert_node_labels_to_integers(g.subgraph(comp).copy())
        n = g.number_of_nodes()
        pd = rng.rand(n * (n - 1) // 2).astype(np.float32)
        out[f'{tag}_edges'] = np.array(g.edges())
        out[f'{tag}_n'] = np.array(n)
        out[f'{tag}_pdists'] = pd
        out[f'{tag}_map'] = np.array(py_mean_average_precision(squareform(pd), g))
        hops = shortest_path(nx.to_scipy_sparse_array(g), unweighted=True)
        out[f'{tag}_hops'] = hops[np.triu_indices(n, 1)]
        out[f'{tag}_map_hops'] = np.array(py_mean_average_precision(hops + 1e-9 * rng.rand(n, n), g))
    return out


def make_products_engine_run():
    """The reference's products.TrainingEngine (graphembed/products/train.py) for 3 epochs with validation every epoch:
    products.Embedding of a hyperbolic and a spherical Universal factor on the 63-node tree, QuotientLoss, node
    mini-batches of 40, RAdam(exact) on the points + SGD on the curvatures, stabilize() after every epoch."""
    import tempfile
    import networkx as nx
    from scipy.sparse.csgraph import shortest_path
    import graphembed.train as T
    from graphembed.products import Embedding, TrainingEngine
    class _Anything:  # the per-manifold monitors draw matplotlib figures (monitor.py:53-90): swallow all of it
        def __call__(self, *a, **k):
            return self

        def __getattr__(self, name):
            return self

    plt = sys.modules['matplotlib.pyplot']
    for name in ('figure', 'scatter', 'gcf', 'close', 'subplots', 'plot'):
        setattr(plt, name, _Anything())
    T.SummaryWriter = _Recorder
    torch.set_default_dtype(torch.float64)
    g = nx.balanced_tree(2, 5)
    n = g.number_of_nodes()
    hops = shortest_path(nx.to_scipy_sparse_array(g), unweighted=True)
    cond = torch.tensor(hops[np.triu_indices(n, 1)])
    out = dict(edges=np.array(g.edges()), hops_condensed=cond.numpy())
    torch.manual_seed(42)
    with torch.no_grad():
        emb = Embedding(n, [3, 2], c_init=0.4)
        emb.manifolds[1].c.fill_(-0.6)
        for x in emb.xs:
            x.mul_(30.0)
            x.proj_()
    out['c0'] = np.array([m.c.item() for m in emb.manifolds])
    for i, x in enumerate(emb.xs):
        out[f'x0_{i}'] = x.data.clone().numpy()
    opt = RiemannianAdam(emb.xs, lr=0.02, max_grad_norm=100, exact=True)
    copt = torch.optim.SGD(list(emb.curvature_params), lr=1e-4)
    # the reference optimizers update parameters in place outside no_grad(); wrap their step like torch >= 1.x needs
    for o in (opt,):
        step = o.step
        o.step = (lambda st: (lambda *a, **k: _no_grad_call(st, *a, **k)))(step)
    obj = QuotientLoss()
    with tempfile.TemporaryDirectory() as tmp:
        eng = TrainingEngine(embedding=emb, optimizer=[opt, copt], objective_fn=obj, n_epochs=3, val_every_epochs=1,
                             alpha=1.0, batch_size=40, drop_last_n=5, save_dir=tmp)
        torch.manual_seed(1234)
        eng(GraphDataset(cond.clone()))
        out['files'] = np.array(sorted(os.listdir(tmp)))
        sd = torch.load(os.path.join(tmp, 'best_embedding.pth'))
        out['state_keys'] = np.array(sorted(sd.keys()))
    rec = _Recorder.last.scalars
    out['step_loss'] = np.array([v for _, v in rec[str(obj)]])
    out['pearsonr'] = np.array([v for _, v in rec['pearsonr']])
    out['average_distortion'] = np.array([v for _, v in rec['average_distortion']])
    out['curv0'] = np.array([v for _, v in rec['curv0']])
    out['curv1'] = np.array([v for _, v in rec['curv1']])
    out['cT'] = np.array([m.c.item() for m in emb.manifolds])
    for i, x in enumerate(emb.xs):
        out[f'xT_{i}'] = x.data.clone().numpy()
    return out


def _no_grad_call(fn, *a, **k):
    with torch.no_grad():
        return fn(*a, **k)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'products_engine':
        np.savez_compressed(os.path.join(HERE, 'products_engine_run_f64.npz'), **make_products_engine_run())
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'precision':  # SURVEY 8f-4 fixture only
        np.savez_compressed(os.path.join(HERE, 'precision_map.npz'), **make_precision())
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'universal':  # SURVEY 8f-3 fixtures only
        for name in UNIVERSAL_CASES:
            for dtype, tag in ((torch.float64, 'f64'), (torch.float32, 'f32')):
                np.savez_compressed(os.path.join(HERE, f'{name}_{tag}.npz'), **make_universal(name, dtype, seed=7))
        np.savez_compressed(os.path.join(HERE, 'universal_training_run_f64.npz'), **make_universal_training_run())
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'new':  # only the fixtures added after the first batch
        for dtype, tag in ((torch.float64, 'f64'), (torch.float32, 'f32')):
            np.savez_compressed(os.path.join(HERE, f'objectives_{tag}.npz'), **make_objectives(dtype))
        np.savez_compressed(os.path.join(HERE, 'engine_runs_f64.npz'), **make_engine_runs())
        return
    for name in CASES:
        for dtype, tag in ((torch.float64, 'f64'), (torch.float32, 'f32')):
            np.savez_compressed(os.path.join(HERE, f'{name}_{tag}.npz'), **make_case(name, dtype, seed=7))
    np.savez_compressed(os.path.join(HERE, 'training_run_f64.npz'), **make_training_run())
    for dtype, tag in ((torch.float64, 'f64'), (torch.float32, 'f32')):
        np.savez_compressed(os.path.join(HERE, f'objectives_{tag}.npz'), **make_objectives(dtype))
    np.savez_compressed(os.path.join(HERE, 'engine_runs_f64.npz'), **make_engine_runs())
    for name in UNIVERSAL_CASES:
        for dtype, tag in ((torch.float64, 'f64'), (torch.float32, 'f32')):
            np.savez_compressed(os.path.join(HERE, f'{name}_{tag}.npz'), **make_universal(name, dtype, seed=7))
    np.savez_compressed(os.path.join(HERE, 'universal_training_run_f64.npz'), **make_universal_training_run())
    np.savez_compressed(os.path.join(HERE, 'precision_map.npz'), **make_precision())
    np.savez_compressed(os.path.join(HERE, 'products_engine_run_f64.npz'), **make_products_engine_run())
    print('wrote', len(os.listdir(HERE)) - 1, 'fixtures to', HERE)


if __name__ == '__main__':
    main()
