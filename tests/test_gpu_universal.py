"""GPU parity tests (-m gpu) for the Universal (kappa-stereographic) manifold and products.Embedding -- SURVEY 8f-3:
the CUDA path through the C-ABI against golden vectors from the real reference (graphembed/manifolds/universal.py,
graphembed/products/embedding.py).  1e-10 relative in fp64; fp32 through the error budget of helpers.assert_parity (1e-5, or twice the reference's own
fp32 error against its fp64 answer on the same inputs)."""
import numpy as np
import pytest
import torch

from helpers import assert_parity, assert_parity_scalar, load_golden, load_truth, rel_err
from helpers_universal import OPTS, UNIVERSAL_CASES, check_curvature_grad

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def make_manifold(name, g, dtype):
    from graphembed.manifolds import Universal
    n, kw = UNIVERSAL_CASES[name]
    man = Universal(n, device=DEV, dtype=dtype, **kw)
    assert rel_err(man.get_c().detach(), g['c']) < 1e-6
    return man


def _truth(name, tag):
    return load_truth(name) if tag == 'f32' else None


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_dist_point_and_curvature_gradients(name, tag):
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_manifold(name, g, g['x'].dtype)
    x, y = g['x'].to(DEV).requires_grad_(), g['y'].to(DEV).requires_grad_()
    d2 = man.dist(x, y, squared=True)
    (d2 * g['w'].to(DEV)).sum().backward()
    assert_parity(d2.detach(), g, 'dist2', tag, T)
    assert_parity(x.grad, g, 'gx', tag, T)
    assert_parity(y.grad, g, 'gy', tag, T)
    check_curvature_grad(man.c.grad, g, name, tag, 'gc', T)  # d/dc through get_c(): sign / softplus forms
    with torch.no_grad():
        assert_parity(man.dist(g['x'].to(DEV), g['y'].to(DEV)), g, 'dist', tag, T)
        assert man.dist(g['x'].to(DEV), g['y'].to(DEV), keepdim=True).shape == (g['x'].shape[0], 1)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_pdist_losses_and_fused_kernel(name, tag):
    from graphembed import _ops, _lib as L
    from graphembed.objectives import QuotientLoss, StressLoss
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_manifold(name, g, g['x'].dtype)
    targets = g['targets'].to(DEV)
    specs = dict(quot=_ops.LossSpec(L.GM_LOSS_QUOTIENT, True, True, alpha=1.7, eps=1 / 4),
                 quot_l1=_ops.LossSpec(L.GM_LOSS_QUOTIENT, True, False, alpha=1.7, eps=1 / 4),
                 stress=_ops.LossSpec(L.GM_LOSS_STRESS))
    for lname, fn, kw in (('quot', QuotientLoss(), dict(epoch=3, alpha=1.7)),
                          ('quot_l1', QuotientLoss(inc_l2=False), dict(epoch=3, alpha=1.7)),
                          ('stress', StressLoss(), dict())):
        # (a) autograd path: pdist kernel forward, gradient kernel backward (points + curvature)
        man.c.grad = None
        x = g['x'].to(DEV).requires_grad_()
        pd2 = man.pdist(x, squared=True)
        loss = fn(targets, 0.9 * pd2, **kw)
        loss.backward()
        assert_parity(pd2.detach(), g, 'pdist2', tag, T)
        assert_parity_scalar(loss.item(), g, f'loss_{lname}', tag, T)
        assert_parity(x.grad, g, f'grad_{lname}', tag, T)
        check_curvature_grad(man.c.grad, g, name, tag, f'gradc_{lname}', T)
        # (b) one fused launch: distance + loss + point gradient + d(loss)/dc
        xd = g['x'].to(DEV).contiguous()
        grad = torch.zeros_like(xd)
        cg = torch.zeros(1, dtype=torch.float64, device=DEV)
        acc, d2 = _ops.pairs_loss_fused(man.spec, xd, _ops.PairSet.triu(xd.shape[0]), _ops.TargetSpec.vector(targets),
                                        specs[lname], 0.9, grad, want_d2=True, c_grad=cg)
        assert_parity(d2, g, 'pdist2', tag, T)
        assert_parity_scalar(acc[0].item(), g, f'loss_{lname}', tag, T)
        assert_parity(grad, g, f'grad_{lname}', tag, T)
        # chain rule through get_c(): d get_c / d c_param is 1 (free sign) or sign * sigmoid(c_param)
        sign = int(g['sign'])
        chain = 1.0 if not sign else sign * torch.sigmoid(g['c_param'].double()).item()
        check_curvature_grad(cg.cpu() * chain, g, name, tag, f'gradc_{lname}', T)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_point_ops(name, tag):
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_manifold(name, g, g['x'].dtype)
    x, y, u, v, eg, far = (g[k].to(DEV) for k in ('x', 'y', 'u', 'v', 'eg', 'far'))
    with torch.no_grad():
        assert_parity(man.exp(x, u), g, 'exp', tag, T)
        assert_parity(man.retr(x, u), g, 'retr', tag, T)
        assert_parity(man.log(x, y), g, 'log', tag, T)
        assert man.proju(x, eg) is eg
        assert_parity(man.egrad2rgrad(x, eg), g, 'egrad2rgrad', tag, T)
        assert_parity(man.transp(x, y, u), g, 'transp', tag, T)
        assert_parity(man.inner(x, u, v), g, 'inner', tag, T)  # (N, N): the reference's broadcast, kept
        assert_parity(man.norm(x, u, squared=True).reshape(g['norm2'].shape), g, 'norm2', tag, T)
        same = man.projx(far)  # not in place: returns its argument untouched (universal.py:53-57)
        assert same is far and torch.equal(far.cpu(), g['far'])
        assert_parity(man.projx(far.clone(), inplace=True), g, 'projx', tag, T)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('oname', sorted(OPTS))
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_optimizer_trajectories(name, oname, tag):
    from graphembed.modules import ManifoldParameter
    from graphembed.optim import RiemannianAdam, RiemannianSGD
    g, T = load_golden(name, tag), _truth(name, tag)
    man = make_manifold(name, g, g['x'].dtype)
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


@pytest.mark.parametrize('fused', [True, False])
def test_products_embedding_training_run(fused):
    """The reference's products.Embedding run (RAdam on the points, SGD on the two curvatures, stabilize every step):
    losses, curvature gradients, curvature trajectory, first-step point gradients and final points.  fused=True takes
    BatchedObjective's kernel path (gm_pairs_dist2 + gm_product_loss + gm_pairs_grad with c_grad), fused=False the
    generic autograd path through Embedding.compute_dists."""
    from graphembed.data import GraphDataset
    from graphembed.modules import BatchedObjective
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam
    from graphembed.products import Embedding
    g = load_golden('universal_training_run', 'f64')
    n = g['x0_0'].shape[0]
    emb = Embedding(n, [3, 2], c_init=0.4, device=DEV, dtype=torch.float64)
    with torch.no_grad():
        emb.manifolds[1].c.fill_(-0.6)
        for i, x in enumerate(emb.xs):
            x.copy_(g[f'x0_{i}'].to(DEV))
    ds = GraphDataset(g['hops_condensed'].clone())
    if fused:
        ds.pdists = ds.pdists.to(device=DEV, dtype=torch.float64)
    emb.fused_pair_kernels = fused
    opt = RiemannianAdam(emb.xs, lr=0.02, max_grad_norm=100, exact=True)
    copt = torch.optim.SGD(list(emb.curvature_params), lr=1e-4)
    bobj = BatchedObjective(QuotientLoss(), ds, emb)
    perm = g['perm'].long().to(DEV)
    losses, cs, cgrads = [], [], []
    for step in range(4):
        idx = perm if step % 2 == 0 else perm[:20]
        loss = bobj(idx, alpha=1.0, epoch=step + 1).sum()
        opt.zero_grad()
        copt.zero_grad()
        loss.backward()
        if step == 0:
            for i, x in enumerate(emb.xs):
                assert rel_err(x.grad, g[f'grad0_{i}']) < 1e-10
        cgrads.append([m.c.grad.item() for m in emb.manifolds])
        opt.step()
        copt.step()
        emb.stabilize()
        losses.append(loss.item())
        cs.append([m.c.item() for m in emb.manifolds])
    assert np.allclose(losses, g['losses'].numpy(), rtol=1e-10)
    assert np.allclose(cgrads, g['cgrads'].numpy(), rtol=1e-9)
    assert np.allclose(cs, g['cs'].numpy(), rtol=1e-10)
    for i, x in enumerate(emb.xs):
        assert rel_err(x.data, g[f'xT_{i}']) < 1e-9


def test_universal_at_scale_properties():
    """Size-independent properties at 2^20 sampled pairs of 100k points (fp32): symmetry d(x,y) == d(y,x), the fused
    kernel's loss equals the sum over the per-pair path, and the curvature gradient agrees with a central difference of
    the loss in c."""
    from graphembed import _ops, _lib as L
    from graphembed.manifolds import Universal
    torch.manual_seed(5)
    N, P = 100_000, 1 << 20
    for c0 in (0.3, -0.4):
        man = Universal(6, c_init=c0, device=DEV, dtype=torch.float64)
        x = man.rand(N, ir=0.6)
        I = torch.randint(N, (P,), device=DEV, dtype=torch.int32)
        J = (I + 1 + torch.randint(N - 1, (P,), device=DEV, dtype=torch.int32)) % N
        with torch.no_grad():
            dij = man.pair_dist2(x, I, J)
            dji = man.pair_dist2(x, J, I)
        assert rel_err(dij, dji) < 1e-12
        tg = torch.rand(P, device=DEV, dtype=torch.float64) + 0.5
        spec = _ops.LossSpec(L.GM_LOSS_STRESS)
        pairs = _ops.PairSet.from_lists(I, J, DEV)

        def loss_at(cval):
            acc, _ = _ops.pairs_loss_fused(man.spec, x, pairs, _ops.TargetSpec.vector(tg), spec, 1.0,
                                           torch.zeros_like(x), c=torch.tensor([cval], device=DEV, dtype=torch.float64))
            return acc[0].item()

        cg = torch.zeros(1, dtype=torch.float64, device=DEV)
        grad = torch.zeros_like(x)
        acc, d2 = _ops.pairs_loss_fused(man.spec, x, pairs, _ops.TargetSpec.vector(tg), spec, 1.0, grad, want_d2=True,
                                        c_grad=cg)
        assert abs(acc[0].item() - ((d2 - tg)**2).sum().item()) < 1e-9 * acc[0].item()
        c = man.get_c().item()
        h = 1e-6
        fd = (loss_at(c + h) - loss_at(c - h)) / (2 * h)
        assert abs(cg.item() - fd) < 1e-6 * abs(fd)


def test_products_training_engine_vs_reference_engine(tmp_path):
    """graphembed.products.TrainingEngine (3 epochs, node batches of 40, RAdam on the points + SGD on the two
    curvatures, stabilize every epoch, validation every epoch) against the reference's own products.TrainingEngine
    (tests/golden/products_engine_run_f64.npz): step losses, pearsonr / average_distortion, the logged curvatures,
    final curvature parameters and points, files written."""
    import os
    from graphembed.data import GraphDataset
    from graphembed.objectives import QuotientLoss
    from graphembed.optim import RiemannianAdam
    from graphembed.products import Embedding, TrainingEngine
    g = load_golden('products_engine_run', 'f64')
    n = g['x0_0'].shape[0]
    emb = Embedding(n, [3, 2], c_init=0.4, device=DEV, dtype=torch.float64)
    with torch.no_grad():
        emb.manifolds[1].c.fill_(-0.6)
        for i, x in enumerate(emb.xs):
            x.copy_(g[f'x0_{i}'].to(DEV))
    opt = RiemannianAdam(emb.xs, lr=0.02, max_grad_norm=100, exact=True)
    copt = torch.optim.SGD(list(emb.curvature_params), lr=1e-4)
    obj = QuotientLoss()
    eng = TrainingEngine(embedding=emb, optimizer=[opt, copt], objective_fn=obj, n_epochs=3, val_every_epochs=1,
                         alpha=1.0, batch_size=40, drop_last_n=5, save_dir=str(tmp_path), tensorboard=False)
    ds = GraphDataset(g['hops_condensed'].to(device=DEV, dtype=torch.float64))
    torch.manual_seed(1234)  # one CPU randperm per epoch, as in the fixture
    eng(ds)
    assert eng._lean['ok'] and not eng._lean['epoch_ok']  # trained curvatures: lean step per slice, not the epoch kernel
    h = eng.writer.history
    for key, tag in (('step_loss', str(obj)), ('pearsonr', 'pearsonr'), ('average_distortion', 'average_distortion'),
                     ('curv0', 'curv0'), ('curv1', 'curv1')):
        got = np.array([v for _, v in h[tag]])
        assert np.allclose(got, g[key].numpy(), rtol=1e-8), (key, got, g[key])
    assert np.allclose([m.c.item() for m in emb.manifolds], g['cT'].numpy(), rtol=1e-9)
    for i, x in enumerate(emb.xs):
        assert rel_err(x.data, g[f'xT_{i}']) < 1e-9
    assert sorted(os.listdir(tmp_path)) == list(g['files'])
    assert sorted(torch.load(tmp_path / 'best_embedding.pth').keys()) == list(g['state_keys'])
