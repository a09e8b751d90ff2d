"""Round-2 fixtures, all produced by the REAL reference (dalab/matrix-manifolds `graphembed`, imported from
/root/reference through oracle/ref_import.py) on CPU.  Run in the build container only:

    python tests/golden/make_golden_r2.py [graphs] [truth] [config1] [steps] [precision]

  graphs  : tests/golden/graphs/<name>.npz -- the integer-labelled edge lists of the graphs BASELINE.json's configs
            1-4 name (tree1000, power, facebook, condmat), exactly as the reference's loader numbers the nodes
            (data/graph.py:27-41: nx.read_edgelist + convert_node_labels_to_integers), plus exact BFS statistics of
            each (sha256 of the full uint8 hop matrix, per-source row sums, a few full rows) from
            scipy.sparse.csgraph.shortest_path(unweighted=True) for the bit-exact BFS tests.
  truth   : <case>_f32truth.npz -- the reference evaluated in fp64 ON THE fp32 FIXTURE'S INPUTS (upcast), i.e. the
            exact answer the fp32 kernels and the fp32 reference both approximate.  The fp32 parity tests use it as
            an error budget: err(kernel_fp32) <= max(1e-5, 2 * err(reference_fp32)).
  config1 : config1_tree1000_f64.npz -- BASELINE config 1 at full size through the reference's own TrainingEngine:
            data/tree1000.edges.gz -> SPD 3x3, fp64, all 499 500 pairs per step, QuotientLoss, RiemannianSGD(lr .01,
            exact, clip 20) (experiments/run_grid.py:30-33), 5 epochs, validation every epoch; and the same with the
            scale ("curvature") parameter in a second RiemannianSGD group as run_grid.py:30-33 builds it.
  precision: precision_f1.npz -- per-layer F1 statistics of the REFERENCE's own native FastPrecision
            (graphembed/pyx/impl/precision.cpp compiled unmodified into oracle/_ref/, see oracle/Makefile) on the four
            random graphs of precision_map.npz with random fp32 / fp64 distances: LayerMeanF1Scores (all degrees and a
            degree window), LayerMeanAverageF1Scores, MeanAveragePrecision, NodesPerLayer.
  steps   : config<k>_step_<dtype>.npz -- one teacher-forced training step (512-node batch = 130 816 pairs, the
            reference's canonical batch, run_grid.py:131) of configs 2a / 2b / 3a / 3b / 4 on the shipped graphs:
            loss, gradient rows, points / optimizer state after one RiemannianAdam step (run_grid.py:25-28), in the
            config's dtype and (for fp32) the fp64 truth of the same step.
"""
import gzip
import hashlib
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

ref_import.load()
import make_golden as G1  # noqa: E402  (re-uses CASES / UNIVERSAL_CASES / _Recorder; imports the reference too)
from graphembed.manifolds import SymmetricPositiveDefinite, Lorentz, Grassmann  # noqa: E402
from graphembed.modules import ManifoldParameter, ManifoldEmbedding, BatchedObjective  # noqa: E402
from graphembed.objectives import QuotientLoss, StressLoss  # noqa: E402
from graphembed.optim import RiemannianAdam, RiemannianSGD  # noqa: E402
from graphembed.data.dataset import GraphDataset  # noqa: E402

DATA = '/root/reference/data'
GRAPHS = {'tree1000': 'tree1000.edges.gz', 'power': 'power.edges.gz', 'facebook': 'facebook.edges.gz',
          'condmat': 'condmat.edges.gz'}


# ---------------------------------------------------------------------------------------------------------------
# graphs
# ---------------------------------------------------------------------------------------------------------------
def load_nx(name):
    import networkx as nx
    g = nx.read_edgelist(os.path.join(DATA, GRAPHS[name]), create_using=nx.Graph)  # data/graph.py:34-39
    g = nx.convert_node_labels_to_integers(g)
    assert nx.number_connected_components(g) == 1
    return g


def hop_matrix(g):
    import networkx as nx
    from scipy.sparse.csgraph import shortest_path
    n = g.number_of_nodes()
    a = nx.to_scipy_sparse_array(g, nodelist=range(n), format='csr')
    out = np.empty((n, n), dtype=np.uint8)
    step = 2048
    for lo in range(0, n, step):  # chunks of sources keep the float64 temporaries small
        idx = np.arange(lo, min(n, lo + step))
        d = shortest_path(a, unweighted=True, indices=idx)
        assert np.isfinite(d).all() and d.max() < 255
        out[idx] = d.astype(np.uint8)
    return out


def make_graph(name):
    g = load_nx(name)
    n = g.number_of_nodes()
    edges = np.array(g.edges(), dtype=np.int32)
    hops = hop_matrix(g)
    rows = np.unique(np.concatenate([[0, 1, n // 2, n - 1], np.random.RandomState(0).randint(0, n, 4)]))
    return dict(n=np.array(n), edges=edges, max_hops=np.array(int(hops.max())),
                sha256=np.array(hashlib.sha256(hops.tobytes()).hexdigest()),
                row_sums=hops.sum(axis=1, dtype=np.int64), hist=np.bincount(hops.reshape(-1), minlength=256),
                sample_rows=rows.astype(np.int64), sample_levels=hops[rows]), hops


# ---------------------------------------------------------------------------------------------------------------
# fp64 truth on the fp32 fixtures' inputs
# ---------------------------------------------------------------------------------------------------------------
def _t(a):
    return torch.from_numpy(np.asarray(a)).double()


def truth_case(name):
    """Everything make_golden.make_case stores, recomputed by the reference in fp64 from the fp32 fixture's inputs."""
    ctor, kw, _, _ = G1.CASES[name]
    torch.set_default_dtype(torch.float64)
    man = G1.build(ctor, kw)
    with np.load(os.path.join(HERE, f'{name}_f32.npz')) as z:
        f = {k: z[k] for k in z.files}
    x, y, w, g = _t(f['x']), _t(f['y']), _t(f['w']), _t(f['targets'])
    out = {}
    xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
    d2 = man.dist(xr, yr, squared=True)
    (d2 * w).sum().backward()
    out.update(dist2=d2.detach().numpy(), gx=xr.grad.numpy(), gy=yr.grad.numpy())
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
    u, v, eg = _t(f['u']), _t(f['v']), _t(f['eg'])
    out.update(exp=man.exp(x, u).numpy(), retr=man.retr(x, u).numpy(), log=man.log(x, y).numpy(),
               proju=man.proju(x, eg.clone()).numpy(), egrad2rgrad=man.egrad2rgrad(x, eg.clone()).numpy(),
               transp=man.transp(x, y, u).numpy(), inner=man.inner(x, u, v).numpy(),
               norm2=man.norm(x, u, squared=True).numpy())
    grads = [_t(gk) for gk in f['opt_grads']]
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


def truth_universal(name):
    from graphembed.manifolds import Universal
    n, kw, _ = G1.UNIVERSAL_CASES[name]
    torch.set_default_dtype(torch.float64)
    with np.load(os.path.join(HERE, f'{name}_f32.npz')) as z:
        f = {k: z[k] for k in z.files}
    man = Universal(n, **kw)
    with torch.no_grad():
        man.c.copy_(_t(f['c_param']).reshape(man.c.shape))
    x, y, w, g = _t(f['x']), _t(f['y']), _t(f['w']), _t(f['targets'])
    out = dict(c=man.get_c().detach().numpy())
    xr, yr = x.clone().requires_grad_(), y.clone().requires_grad_()
    d2 = man.dist(xr, yr, squared=True)
    (d2 * w).sum().backward()
    out.update(dist2=d2.detach().numpy(), gx=xr.grad.numpy(), gy=yr.grad.numpy(), gc=man.c.grad.clone().numpy())
    man.c.grad = None
    with torch.no_grad():
        out['dist'] = man.dist(x, y).numpy()
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
        u, v, eg, far = _t(f['u']), _t(f['v']), _t(f['eg']), _t(f['far'])
        far_proj = far.clone()
        man.projx(far_proj, inplace=True)
        out.update(projx=far_proj.numpy(), exp=man.exp(x, u).numpy(), retr=man.retr(x, u).numpy(),
                   log=man.log(x, y).numpy(), proju=man.proju(x, eg.clone()).numpy(),
                   egrad2rgrad=man.egrad2rgrad(x, eg.clone()).numpy(), transp=man.transp(x, y, u).numpy(),
                   inner=man.inner(x, u, v).numpy(), norm2=man.norm(x, u, squared=True).numpy())
        grads = [_t(gk) for gk in f['opt_grads']]
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


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 1 at full size
# ---------------------------------------------------------------------------------------------------------------
def make_config1(hops):
    import graphembed.train as T
    plt = sys.modules['matplotlib.pyplot']
    plt.scatter = plt.gcf = plt.close = lambda *a, **k: None
    T.SummaryWriter = G1._Recorder
    torch.set_default_dtype(torch.float64)
    n = hops.shape[0]
    cond = torch.tensor(hops[np.triu_indices(n, 1)].astype(np.float64))
    out = {}
    for tag, mkopt in (
            ('xs', lambda emb: RiemannianSGD([dict(params=emb.xs, lr=0.01, exact=True, max_grad_norm=20)], lr=0.01)),
            ('curv', lambda emb: RiemannianSGD([dict(params=emb.xs, lr=0.01, exact=True, max_grad_norm=20),
                                                dict(params=emb.curvature_params, lr=1e-4, max_grad_norm=500)],
                                               lr=0.01))):
        torch.manual_seed(42)
        emb = ManifoldEmbedding(n, [SymmetricPositiveDefinite(3)])
        out[f'{tag}_x0'] = emb.xs[0].data.clone().numpy()
        obj = QuotientLoss()
        with tempfile.TemporaryDirectory() as tmp:
            eng = T.TrainingEngine(embedding=emb, optimizer=mkopt(emb), objective_fn=obj, alpha=1.0, n_epochs=5,
                                   val_every_epochs=1, save_dir=tmp)
            torch.manual_seed(1234)  # one randperm per epoch from the global CPU generator (train.py:206)
            eng(GraphDataset(cond.clone()))
        rec = G1._Recorder.last.scalars
        out[f'{tag}_step_loss'] = np.array([v for _, v in rec[str(obj)]])
        out[f'{tag}_pearsonr'] = np.array([v for _, v in rec['pearsonr']])
        out[f'{tag}_average_distortion'] = np.array([v for _, v in rec['average_distortion']])
        out[f'{tag}_xT'] = emb.xs[0].data.clone().numpy()
        out[f'{tag}_scaleT'] = emb.scales[0].data.clone().numpy()
        if 'scale0' in rec:
            out[f'{tag}_scale_log'] = np.array([v for _, v in rec['scale0']])
        print(tag, out[f'{tag}_step_loss'], out[f'{tag}_average_distortion'], out[f'{tag}_scaleT'])
    return out


# ---------------------------------------------------------------------------------------------------------------
# teacher-forced steps of configs 2-4
# ---------------------------------------------------------------------------------------------------------------
STEP_CONFIGS = {
    # tag: (graph, factories of the factors, dtype)
    '2a': ('power', lambda: [Lorentz(11)], torch.float32),
    '2b': ('power', lambda: [SymmetricPositiveDefinite(4, use_stein_div=True)], torch.float32),
    '3a': ('facebook', lambda: [Grassmann(6, 2)], torch.float64),
    '3b': ('facebook', lambda: [SymmetricPositiveDefinite(3), Lorentz(5)], torch.float32),
    '4': ('condmat', lambda: [SymmetricPositiveDefinite(6)], torch.float32),
}
BATCH = 512


def _one_step(mans, dtype, x0_rows, idx, ds, n, with_curv):
    """One step of the reference: BatchedObjective(QuotientLoss) on the node batch `idx`, backward, RiemannianAdam as
    run_grid.py:25-28 builds it (points lr .01 exact clip 100; scales lr .01 in a second group if with_curv)."""
    torch.set_default_dtype(dtype)
    emb = ManifoldEmbedding(n, mans())
    with torch.no_grad():
        for x, r in zip(emb.xs, x0_rows):
            x[idx] = torch.from_numpy(r).to(dtype)
    groups = [dict(params=emb.xs, lr=0.01, exact=True, max_grad_norm=100)]
    if with_curv:
        groups.append(dict(params=emb.curvature_params, lr=0.01))
    opt = RiemannianAdam(groups)
    bobj = BatchedObjective(QuotientLoss(), ds, emb)
    loss = bobj(idx, alpha=1.0, epoch=1).sum()
    opt.zero_grad()
    loss.backward()
    res = dict(loss=np.array(loss.item()))
    for f, x in enumerate(emb.xs):
        res[f'grad_{f}'] = x.grad[idx].clone().numpy()
        rest = x.grad.clone()
        rest[idx] = 0
        assert not rest.any()  # nothing outside the batch
    for f, s in enumerate(emb.scales):
        res[f'scale_grad_{f}'] = s.grad.clone().numpy()
    with torch.no_grad():
        opt.step()
    for f, x in enumerate(emb.xs):
        res[f'x1_{f}'] = x.data[idx].clone().numpy()
        st = opt.state[x]
        res[f'exp_avg_{f}'] = st['exp_avg'][idx].clone().numpy()
        res[f'exp_avg_sq_{f}'] = st['exp_avg_sq'][idx].clone().numpy()
    for f, s in enumerate(emb.scales):
        res[f'scale1_{f}'] = s.data.clone().numpy()
    return res


def make_step(tag, hops):
    gname, mans, dtype = STEP_CONFIGS[tag]
    n = hops.shape[0]
    torch.set_default_dtype(dtype)
    torch.manual_seed(42)
    emb0 = ManifoldEmbedding(n, mans())  # the reference's initialiser on CPU, seed 42 (run.py:107)
    idx = torch.randperm(n)[:BATCH]
    x0_rows = [x.data[idx].clone().numpy() for x in emb0.xs]
    # the dataset only has to serve this batch: a (512, 512) block of squared, max-normalised hop counts laid out as
    # GraphDataset does for the full graph (data/dataset.py:9-27), driven through the reference's own class on the
    # condensed vector of the batch's induced hop matrix, with the global maximum appended so that max() matches
    sub = hops[idx.numpy()][:, idx.numpy()].astype(np.float64)
    out = dict(idx=idx.numpy().astype(np.int64), max_hops=np.array(int(hops.max())))
    for f, r in enumerate(x0_rows):
        out[f'x0_{f}'] = r

    class _BatchDataset:  # GraphDataset.__getitem__ (dataset.py:19-27) restricted to the rows the step touches
        def __init__(self, dt):
            full = torch.from_numpy(sub).to(dt)
            self.block = full.pow(2).div_(float(int(hops.max()) ** 2))
            self.pos = {int(v): k for k, v in enumerate(idx.tolist())}

        def __getitem__(self, indices):
            p = torch.tensor([self.pos[int(v)] for v in indices.tolist()])
            blk = self.block[p][:, p]
            mask = torch.triu(torch.ones(len(p), len(p)), diagonal=1).bool()
            return blk.masked_select(mask)

    # check the restricted dataset against the reference's own GraphDataset where the full one is affordable
    if n < 6000:
        cond = torch.tensor(hops[np.triu_indices(n, 1)].astype(np.float64)).to(dtype)
        full = GraphDataset(cond)
        assert torch.equal(full[idx], _BatchDataset(dtype)[idx])
    for with_curv in (False, True):
        key = 'curv_' if with_curv else ''
        res = _one_step(mans, dtype, x0_rows, idx, _BatchDataset(dtype), n, with_curv)
        out.update({key + k: v for k, v in res.items()})
        if dtype == torch.float32:
            x0_64 = [r.astype(np.float64) for r in x0_rows]
            res = _one_step(mans, torch.float64, x0_64, idx, _BatchDataset(torch.float64), n, with_curv)
            out.update({key + 'truth_' + k: v for k, v in res.items()})
    print(tag, gname, 'loss', out['loss'], out.get('truth_loss'))
    return out


def make_precision_f1():
    sys.path.insert(0, os.path.join(HERE, '..'))
    import precision_ref as R
    from helpers_precision import TAGS, csr_of, load_precision_golden
    assert R.available(), 'run `make -C oracle` first (builds oracle/_ref/libprecision_ref.so from the reference)'
    g = load_precision_golden()
    out = {}
    rng = np.random.RandomState(11)
    for tag in TAGS:
        n = int(g[f'{tag}_n'])
        ref = R.FastPrecisionRef(*csr_of(n, g[f'{tag}_edges']))
        pd32 = g[f'{tag}_pdists'].astype(np.float32)
        pd64 = rng.rand(n * (n - 1) // 2)
        out[f'{tag}_pd64'] = pd64
        out[f'{tag}_npl'] = ref.nodes_per_layer()
        for name, pd in (('f32', pd32), ('f64', pd64)):
            m, s_ = ref.layer_mean_f1_scores(pd)
            out[f'{tag}_{name}_f1_mean'], out[f'{tag}_{name}_f1_std'] = m, s_
            m, s_ = ref.layer_mean_f1_scores(pd, 3, 8)
            out[f'{tag}_{name}_f1w_mean'], out[f'{tag}_{name}_f1w_std'] = m, s_
            m, s_ = ref.layer_mean_average_f1_scores(pd.astype(np.float64))
            out[f'{tag}_{name}_af_mean'], out[f'{tag}_{name}_af_std'] = m, s_
            out[f'{tag}_{name}_map'] = np.array(ref.mean_average_precision(pd))
    return out


def main():
    what = set(sys.argv[1:]) or {'graphs', 'truth', 'config1', 'steps', 'precision'}
    if 'precision' in what:
        np.savez_compressed(os.path.join(HERE, 'precision_f1.npz'), **make_precision_f1())
    os.makedirs(os.path.join(HERE, 'graphs'), exist_ok=True)
    hops = {}

    def get_hops(name):
        if name not in hops:
            meta, h = make_graph(name)
            hops[name] = h
            if 'graphs' in what:
                np.savez_compressed(os.path.join(HERE, 'graphs', f'{name}.npz'), **meta)
                print('graph', name, int(meta['n']), 'nodes', len(meta['edges']), 'edges, max hops', int(meta['max_hops']))
        return hops[name]

    if 'graphs' in what:
        for name in GRAPHS:
            get_hops(name)
    if 'truth' in what:
        for name in G1.CASES:
            np.savez_compressed(os.path.join(HERE, f'{name}_f32truth.npz'), **truth_case(name))
        for name in G1.UNIVERSAL_CASES:
            np.savez_compressed(os.path.join(HERE, f'{name}_f32truth.npz'), **truth_universal(name))
    if 'config1' in what:
        np.savez_compressed(os.path.join(HERE, 'config1_tree1000_f64.npz'), **make_config1(get_hops('tree1000')))
    if 'steps' in what:
        for tag, (gname, _, dtype) in STEP_CONFIGS.items():
            dt = 'f32' if dtype == torch.float32 else 'f64'
            np.savez_compressed(os.path.join(HERE, f'config{tag}_step_{dt}.npz'), **make_step(tag, get_hops(gname)))


if __name__ == '__main__':
    main()
