"""Shared test helpers: golden fixtures, oracle / product-manifold factories, tolerances."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# case name -> (family, kwargs).  Mirrors tests/golden/make_golden.py::CASES.
CASES = {
    'spd2': ('spd', dict(n=2)),
    'spd2_exact': ('spd', dict(n=2, fast_symeig=False, fast_chol=False)),
    'spd3': ('spd', dict(n=3)),
    'spd3_default_init': ('spd', dict(n=3)),
    'spd4': ('spd', dict(n=4)),
    'spd4_default_init': ('spd', dict(n=4)),
    'spd6': ('spd', dict(n=6)),
    'stein2': ('spd', dict(n=2, use_stein_div=True)),
    'stein4': ('spd', dict(n=4, use_stein_div=True)),
    'lorentz11': ('lorentz', dict(n=11)),
    'lorentz5_default_init': ('lorentz', dict(n=5)),
    'sphere5': ('sphere', dict(n=5)),
    'euclidean7': ('euclidean', dict(n=7)),
    'grassmann6_2': ('grassmann', dict(n=6, p=2)),
    'grassmann7_3': ('grassmann', dict(n=7, p=3)),
}
DTYPES = {'f64': torch.float64, 'f32': torch.float32}


def load_golden(name, tag):
    with np.load(os.path.join(GOLDEN, f'{name}_{tag}.npz')) as z:
        return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in 'fiub' else z[k]) for k in z.files}


def make_oracle(name):
    import manifolds_oracle as O
    fam, kw = CASES[name]
    if fam == 'spd':
        return O.SpdOracle(kw['n'], fast_symeig=kw.get('fast_symeig', True), fast_chol=kw.get('fast_chol', True),
                           stein=kw.get('use_stein_div', False))
    if fam == 'lorentz':
        return O.LorentzOracle(kw['n'])
    if fam == 'sphere':
        return O.SphereOracle(kw['n'])
    if fam == 'euclidean':
        return O.EuclideanOracle(kw['n'])
    return O.GrassmannOracle(kw['n'], kw['p'])


def make_product(name):
    """The drop-in (CUDA-backed) manifold object of a case."""
    from graphembed import manifolds as M
    fam, kw = CASES[name]
    if fam == 'spd':
        return M.SymmetricPositiveDefinite(**kw)
    if fam == 'lorentz':
        return M.Lorentz(kw['n'])
    if fam == 'sphere':
        return M.Sphere(kw['n'])
    if fam == 'euclidean':
        return M.Euclidean(kw['n'])
    return M.Grassmann(kw['n'], kw['p'])


def is_spd(name):
    return CASES[name][0] == 'spd'


def sym(t):
    return 0.5 * (t + t.transpose(-2, -1))


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    scale = b.abs().max().clamp(min=1e-300)
    return ((a - b).abs().max() / scale).item()


def tol(dtype, name=''):
    """Relative tolerance (of the max-norm) per BASELINE.json north_star: 1e-10 in fp64, 1e-5 in fp32.
    fp32 cases whose points start within ~0.1 of each other are ill-conditioned in the reference itself
    (log-eigenvalue / acosh cancellation), they get 2e-4."""
    if dtype == torch.float64:
        return 1e-10
    if 'default_init' in name or name.startswith('stein') or name.startswith('grassmann'):
        return 2e-4
    return 1e-5
