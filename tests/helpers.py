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
    """Relative tolerance (of the max-norm) per BASELINE.json north_star: 1e-10 in fp64, 1e-5 in fp32 -- no
    exceptions.  fp32 comparisons against the reference's fp32 golden vectors go through `assert_parity`, which
    measures both sides against the reference's fp64 answer on the same fp32 inputs (the *_f32truth.npz fixtures)."""
    return 1e-10 if dtype == torch.float64 else 1e-5


def load_truth(name):
    """The reference evaluated in fp64 on the fp32 fixture's inputs (tests/golden/make_golden_r2.py::truth_case)."""
    with np.load(os.path.join(GOLDEN, f'{name}_f32truth.npz')) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


F32_FLOOR = 1e-5  # north_star: 1e-5 relative in fp32
F32_FACTOR = 2.0  # ... or at most twice the reference's own fp32 error, where that is larger


def parity_errors(got, ref, truth):
    """(error of `got`, error of the reference's fp32 result `ref`), both relative to the max-norm of the fp64
    `truth` computed by the reference from the same fp32 inputs."""
    return rel_err(got, truth), rel_err(ref, truth)


def assert_parity(got, g, key, tag, truth=None, fix=None, index=None, what=''):
    """The parity bar of north_star for one tensor.
    fp64: |got - reference| <= 1e-10 (relative to the max-norm of the reference).
    fp32: error budget -- with T the reference's fp64 result on the same (fp32) inputs,
          err(got, T) <= max(1e-5, 2 * err(reference_fp32, T)):
          the kernel is within the contract's 1e-5 of the exact answer, or, where the reference's own fp32 arithmetic
          is further away than that (ill-conditioned inputs), no more than twice as far as the reference itself."""
    fix = fix or (lambda t: t)
    ref = g[key] if index is None else g[key][index]
    if tag == 'f64':
        e = rel_err(fix(got), fix(ref))
        assert e < 1e-10, f'{what or key}: fp64 rel err {e:.2e}'
        return e
    tr = truth[key] if index is None else truth[key][index]
    e_got, e_ref = parity_errors(fix(got), fix(ref), fix(tr))
    bound = max(F32_FLOOR, F32_FACTOR * e_ref)
    assert e_got <= bound, f'{what or key}: fp32 err vs fp64 truth {e_got:.2e} > max(1e-5, 2 x reference fp32 err ' \
                           f'{e_ref:.2e})'
    return e_got


def assert_parity_scalar(got, g, key, tag, truth=None):
    ref = float(g[key])
    if tag == 'f64':
        assert abs(got - ref) <= 1e-10 * abs(ref), (key, got, ref)
        return
    tr = float(truth[key])
    e_got, e_ref = abs(got - tr) / abs(tr), abs(ref - tr) / abs(tr)
    assert e_got <= max(F32_FLOOR, F32_FACTOR * e_ref), f'{key}: fp32 err {e_got:.2e} vs reference fp32 err {e_ref:.2e}'
