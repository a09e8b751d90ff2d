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
    """1e-10 (fp64) / 1e-5 (fp32); the default init (points within 1e-2 of each other and of the origin, d^2 ~ 1e-3)
    is cancellation-limited in fp32 in the reference itself -> 2e-4, as for the other default-init cases."""
    if tag == 'f64':
        return 1e-10
    return 2e-4 if 'default_init' in name else 1e-5


def check_curvature_grad(got, g, name, tag, key, t):
    """d(loss)/dc_param against the reference.  For the default init in fp32 the quantity is cancellation noise in the
    reference itself (d^2 ~ 4|x - y|^2 barely depends on c: the reference's fp32 values are 1.5x-3x off its own fp64
    values, see tests/golden/universal4_default_init_f{32,64}.npz), so there only sign and magnitude (within a factor
    2 of the reference's fp32 number) are required; the same inputs in fp64 are checked to 1e-10."""
    from helpers import rel_err
    if tag == 'f32' and 'default_init' in name:
        ratio = (got.double().cpu().reshape(-1) / g[key].double().reshape(-1)).item()
        assert 0.5 < ratio < 2.0
        return
    assert rel_err(got, g[key]) < t
