"""The reference's OWN known-answer tests for this path, run against the CUDA kernels (-m gpu), with the reference's
tolerance (atol 1e-4, tests/utils.py) and its input generators (tests/conftest.py: seeds 0-4, rand_spd = U U^T + I,
rand_sym).  Sources, all under /root/reference/graphembed/tests/:
  test_spd.py:15-23   unit distance       d(I, exp_I(u)) == ||u|| == 1
  test_spd.py:26-33   exp / log           log_x(exp_x(u)) == u,  ||u||_x == d(x, exp_x(u))
  test_spd.py:36-44   no NaN distances
  test_spd.py:47-59   distance formula    d(x, y) == || log eig(y^-1 x) ||
  test_spd.py:62-67   inner / norm        sqrt(<u, u>_x) == ||u||_x
  test_spd.py:70-78   gradient            rgrad(d^2 / 2) == -log_x(y),  n = 2..9
  test_spd.py:81-88   Stein pdiv vs div
  test_ortho.py:12-17, 28-36  Grassmann   d(x, y) == ||log_x(y)||,  rgrad(d^2 / 2) == -log_x(y)
  test_isometry.py:51-107     SPD(2) with fixed determinant is isometric to sqrt(2) * H^2
  test_sphere.py:9-15         antipodal points are pi apart
  test_optim.py:13-41         RSGD / RAdam on the sphere find the dominant eigenvector
"""
import math
from itertools import product

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
ATOL = 1e-4


@pytest.fixture(params=range(5))
def seed(request):
    torch.manual_seed(request.param)
    np.random.seed(request.param)
    return request.param


def rand_sym(n, d, dtype=torch.float64):
    x = torch.rand(n, d, d, dtype=dtype, device=DEV)
    return 0.5 * (x + x.transpose(1, 2))


def rand_spd(n, d, dtype=torch.float64):
    x = torch.rand(n, d, d, dtype=dtype, device=DEV)
    return x @ x.transpose(1, 2) + torch.eye(d, dtype=dtype, device=DEV)


def close(a, b, atol=ATOL):
    a, b = torch.as_tensor(a, dtype=torch.float64).cpu(), torch.as_tensor(b, dtype=torch.float64).cpu()
    return bool(((a - b).abs() <= atol).all())


def SPD(*a, **k):
    from graphembed.manifolds import SymmetricPositiveDefinite
    return SymmetricPositiveDefinite(*a, **k)


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('d', range(2, 10))
def test_spd_unit_distance(d, seed, dtype):
    spd = SPD(d)
    u_vec = torch.randn(spd.dim, dtype=dtype, device=DEV)
    u = spd.from_vec(u_vec / u_vec.norm())
    x = torch.eye(d, dtype=dtype, device=DEV)
    assert close(1.0, spd.norm(x.unsqueeze(0), u.unsqueeze(0)))
    y = spd.exp(x.unsqueeze(0), u.unsqueeze(0))
    assert close(1.0, spd.dist(x.unsqueeze(0), y))


@pytest.mark.parametrize('d', range(2, 10))
def test_spd_exp_log(seed, d):
    spd = SPD(d)
    x, u = rand_spd(10, d), rand_sym(10, d)
    y = spd.exp(x, u)
    assert close(u, spd.log(x, y))
    assert close(spd.norm(x, u), spd.dist(x, y))


@pytest.mark.parametrize('d,n', [(2, 10000), (3, 10000), (4, 3000)])
def test_spd_no_nan_dists(seed, d, n):
    spd = SPD(d)
    x = rand_spd(n, d, dtype=torch.float32)
    assert not torch.isnan(spd.pdist(x)).any()


@pytest.mark.parametrize('d', range(1, 10))
def test_spd_distance_formulas(seed, d):
    spd = SPD(d)
    x, y = rand_spd(2, d)
    ref = spd.dist(x.unsqueeze(0), y.unsqueeze(0))
    for a, b in ((y, x), (x, y)):  # eigenvalues of b^-1 a (not symmetric: general eigensolver, on the host)
        ev = torch.linalg.eigvals(torch.linalg.solve(b.cpu(), a.cpu())).real
        assert close(ref, ev.log().pow(2).sum().sqrt())


@pytest.mark.parametrize('d', range(2, 10))
def test_spd_inner_norm(seed, d):
    spd = SPD(d)
    xs = spd.rand(100, ir=1.0, out=torch.empty(0, dtype=torch.float64, device=DEV))
    us = spd.randvec(xs)
    assert close(spd.inner(xs, us, us)**0.5, spd.norm(xs, us))


@pytest.mark.parametrize('d', range(2, 10))
def test_spd_gradient_is_minus_log(seed, d):
    spd = SPD(d)
    x, y = spd.rand(2, ir=1.0, out=torch.empty(0, dtype=torch.float64, device=DEV))
    x = x.unsqueeze(0).clone().requires_grad_()
    y = y.unsqueeze(0)
    dist = 0.5 * spd.dist(x, y, squared=True)
    grad_e = torch.autograd.grad(dist.sum(), x)[0]
    grad = spd.egrad2rgrad(x.detach(), grad_e)
    assert close(grad, -spd.log(x.detach(), y))


@pytest.mark.parametrize('d', range(2, 10))
def test_spd_stein_pdiv_equals_div(seed, d):
    spd = SPD(d, use_stein_div=True)
    xs = spd.rand(10, ir=1.0, out=torch.empty(0, dtype=torch.float64, device=DEV))
    m = torch.triu_indices(10, 10, 1, device=DEV)
    assert close(spd.dist(xs[m[0]], xs[m[1]], squared=True), spd.pdist(xs, squared=True))


@pytest.mark.parametrize('n,p', list(product(range(5, 10), [2, 3, 4])))
def test_grassmann_log_and_gradient(seed, n, p):
    from graphembed.manifolds import Grassmann
    gras = Grassmann(n, p)
    hint = torch.empty(0, dtype=torch.float64, device=DEV)
    x, y = gras.rand_uniform(10, out=hint), gras.rand_uniform(10, out=hint)
    assert close(gras.dist(x, y), gras.norm(x, gras.log(x, y)))
    xr = x.clone().requires_grad_()
    dist = 0.5 * gras.dist(xr, y, squared=True)
    grad_e = torch.autograd.grad(dist.sum(), xr)[0]
    assert close(gras.egrad2rgrad(x, grad_e), -gras.log(x, y))


def _sspd2_to_h2(x):
    a, b, c = x[..., 0, 0], x[..., 1, 1], x[..., 0, 1]
    y = torch.stack([0.5 * (a + b), 0.5 * (a - b), c], dim=-1)
    ldot = -y[..., :1]**2 + (y[..., 1:]**2).sum(-1, keepdim=True)
    return y / torch.sqrt(-ldot)


@pytest.mark.parametrize('det', [0.5, 1.0, 2.0])
def test_spd2_is_isometric_to_scaled_h2(seed, det):
    from graphembed.manifolds import Lorentz
    spd, lorentz = SPD(2), Lorentz(3)
    x = spd.rand(100, ir=1.0, out=torch.empty(0, dtype=torch.float64, device=DEV))
    x = x / torch.linalg.det(x).sqrt().reshape(-1, 1, 1) * det
    assert close(x, spd.projx(x))
    y = _sspd2_to_h2(x).contiguous()
    assert close(spd.pdist(x), math.sqrt(2) * lorentz.pdist(y))


def test_sphere_antipodal_distance():
    from graphembed.manifolds import Sphere
    man = Sphere(5)
    x = torch.zeros(1, 5, device=DEV)
    y = torch.zeros(1, 5, device=DEV)
    x[0, 0], y[0, 0] = 1.0, -1.0
    assert close(man.dist(x, y), math.pi, atol=1e-3)  # fp32 acos near -1


@pytest.mark.parametrize('n,which', list(product([3, 4, 5], ['rsgd', 'radam'])))
def test_optimizers_find_the_dominant_eigenvector(seed, n, which):
    from graphembed.manifolds import Sphere
    from graphembed.modules import ManifoldParameter
    from graphembed.optim import RiemannianAdam, RiemannianSGD
    man = Sphere(n)
    A = rand_sym(1, n)[0]
    x = ManifoldParameter(man.rand(1, out=torch.empty(0, dtype=torch.float64, device=DEV)), manifold=man)
    optim = RiemannianSGD([x], lr=1e-1) if which == 'rsgd' else RiemannianAdam([x], lr=1e-1)
    for _ in range(200):
        optim.zero_grad()
        loss = -torch.einsum('i,ij,j', x[0], A, x[0])
        loss.backward()
        optim.step()
    assert close(1.0, x.detach().norm())
    w, v = torch.linalg.eigh(A)
    x_opt = x[0].detach()
    assert close((x_opt / v[:, -1]).abs(), torch.ones(n))
    assert close((A @ x_opt).norm(), w[-1])
