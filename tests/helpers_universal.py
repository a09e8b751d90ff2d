"""Fixtures / factories for the Universal (kappa-stereographic) manifold tests.  The golden files come
from the real reference: tests/golden/make_golden.py::make_universal, make_universal_training_run."""
from helpers import load_golden

UNIVERSAL_CASES = {
    # name: (n, ctor kwargs) -- mirrors tests/golden/make_golden.py::UNIVERSAL_CASES
    'universal5_pos': (5, dict(c_init=0.5)),
    'universal5_neg': (5, dict(c_init=-0.7)),
    'universal3_fixed_sign': (3, dict(c_init=-0.3, keep_sign_fixed=True)),
    'universal4_default_init': (4, dict()),
}
OPTS = {
    'radam_clip': ('radam', dict(lr=0.05, max_grad_norm=1.5)),
    'radam_exact': ('radam', dict(lr=0.05, exact=True)),
    'rsgd_exact_clip': ('rsgd', dict(lr=0.05, max_grad_norm=0.5, exact=True)),
    'rsgd_momentum': ('rsgd', dict(lr=0.05, momentum=0.9, dampening=0.1)),
}


def oracle_for(g):
    """(oracle, c_param leaf) rebuilt from a fixture: the curvature goes through get_c() so that d/dc_param is
    comparable with the reference's `man.c.grad`."""
    import manifolds_oracle as O
    c_param = g['c_param'].clone().requires_grad_()
    sign = int(g['sign'])
    c = O.universal_get_c(c_param, float(g['c_min']), sign if sign else None)
    return O.UniversalOracle(g['x'].shape[-1], c), c_param


def tol_u(tag, name=''):
    """1e-10 (fp64) / 1e-5 (fp32), no exceptions: fp32 results are compared through helpers.assert_parity's error
    budget against the reference's fp64 answer on the same fp32 inputs."""
    return 1e-10 if tag == 'f64' else 1e-5


def check_curvature_grad(got, g, name, tag, key, truth=None):
    """d(loss)/dc_param against the reference: 1e-10 in fp64; in fp32 within max(1e-5, 2 x the reference's own fp32
    error) of the reference's fp64 value on the same inputs.  (For the default init the quantity is cancellation
    noise in fp32 -- d^2 ~ 4|x - y|^2 barely depends on c and the reference's fp32 value is 1.5-3x off its fp64 one;
    the budget then only asks the kernel to be no further from the truth than twice that.)"""
    from helpers import assert_parity
    assert_parity(got.reshape(-1), {key: g[key].reshape(-1)}, key, tag,
                  None if truth is None else {key: truth[key].reshape(-1)})
