"""CPU-only check of the *kernel arithmetic*: csrc/gm_manifolds.cuh and csrc/gm_pointops.cuh are __host__ __device__,
tests/hostcheck compiles them for x86 and this test compares them with the pinned oracle on seeded inputs.
(The product never runs this host build; it exists so that kernel math regressions are caught without a GPU.)"""
import ctypes
import os
import subprocess

import pytest
import torch

from helpers import CASES, assert_parity, load_golden, load_truth, make_oracle, is_spd, rel_err, sym

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'hostcheck', 'libhostcheck.so')
KIND = {'spd': 0, 'lorentz': 2, 'sphere': 3, 'grassmann': 4, 'euclidean': 5}


@pytest.fixture(scope='module')
def hc():
    if not os.path.isfile(SO):
        try:
            subprocess.check_call(['make', '-C', os.path.join(HERE, 'hostcheck')])
        except (OSError, subprocess.CalledProcessError):
            pytest.skip('nvcc not available to build the host math checker')
    return ctypes.CDLL(SO)


def _desc(name):
    fam, kw = CASES[name]
    kind, n, p, flags = KIND[fam], kw['n'], kw.get('p', 0), 0
    if fam == 'spd':
        if kw.get('use_stein_div'):
            kind = 1
        elif kw.get('fast_symeig', True) and n in (2, 3):
            flags |= 1
        if kw.get('fast_chol', True) and n == 2:
            flags |= 2
    if fam == 'grassmann' and p == 2:
        flags |= 4
    return kind, n, p, flags


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_pair_math_matches_oracle(hc, name, tag):
    g = load_golden(name, tag)
    kind, n, p, flags = _desc(name)
    x, y = g['x'].contiguous(), g['y'].contiguous()
    P = x.shape[0]
    d2 = torch.empty(P, dtype=x.dtype)
    gx, gy = torch.empty(x.shape, dtype=x.dtype), torch.empty(y.shape, dtype=y.dtype)
    rc = hc.hc_pairs(kind, 0 if x.dtype == torch.float32 else 1, n, p, flags, ctypes.c_double(1e-8),
                     ctypes.c_double(1e8), ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()),
                     ctypes.c_long(P), ctypes.c_void_p(d2.data_ptr()), ctypes.c_void_p(gx.data_ptr()),
                     ctypes.c_void_p(gy.data_ptr()))
    assert rc == 0
    w = g['w'].view(-1, *([1] * (x.ndim - 1)))
    fix = sym if is_spd(name) else (lambda t: t)
    T = load_truth(name) if tag == 'f32' else None
    assert_parity(d2, g, 'dist2', tag, T)
    assert_parity(gx * w, g, 'gx', tag, T, fix)
    assert_parity(gy * w, g, 'gy', tag, T, fix)


# ---- Universal (kappa-stereographic) manifold, SURVEY 8f-3: kernel arithmetic vs golden vectors of the reference --------
from helpers_universal import OPTS as U_OPTS, UNIVERSAL_CASES, check_curvature_grad  # noqa: E402

_vp = ctypes.c_void_p


def _p(t):
    return None if t is None else _vp(t.data_ptr())


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_universal_pair_math(hc, name, tag):
    g = load_golden(name, tag)
    x, y = g['x'].contiguous(), g['y'].contiguous()
    P, n = x.shape
    d2, gc = torch.empty(P, dtype=x.dtype), torch.empty(P, dtype=x.dtype)
    gx, gy = torch.empty_like(x), torch.empty_like(y)
    c = float(g['c'].item())
    hc.hc_universal_pairs(0 if x.dtype == torch.float32 else 1, n, ctypes.c_double(c), ctypes.c_double(1e-8), _p(x),
                          _p(y), ctypes.c_long(P), _p(d2), _p(gx), _p(gy), _p(gc))
    T = load_truth(name) if tag == 'f32' else None
    w = g['w']
    assert_parity(d2, g, 'dist2', tag, T)
    assert_parity(gx * w[:, None], g, 'gx', tag, T)
    assert_parity(gy * w[:, None], g, 'gy', tag, T)
    sign = int(g['sign'])
    chain = 1.0 if not sign else sign * torch.sigmoid(g['c_param'].double()).item()  # d get_c / d c_param
    check_curvature_grad((gc.double() * w.double()).sum().reshape(1) * chain, g, name, tag, 'gc', T)
    # non-squared distance: value floor EPS on d  <=>  EPS^2 on d^2
    hc.hc_universal_pairs(0 if x.dtype == torch.float32 else 1, n, ctypes.c_double(c), ctypes.c_double(1e-16), _p(x),
                          _p(y), ctypes.c_long(P), _p(d2), None, None, None)
    assert_parity(d2.sqrt(), g, 'dist', tag, T)


def _hc_point(hc, g, op, x, u=None, v=None, scalar=False, opt=None, b1=None, b2=None):
    from graphembed import _lib as L
    n = x.shape[-1]
    out = torch.empty(x.shape[0] if scalar else x.shape, dtype=x.dtype)
    hc.hc_set_universal_c(ctypes.c_double(float(g['c'].item())))
    rc = hc.hc_point(L.GM_UNIVERSAL, 0 if x.dtype == torch.float32 else 1, n, 0, 0, ctypes.c_double(1e-8),
                     ctypes.c_double(1e8), 0, op, None if opt is None else ctypes.byref(opt), _p(x), _p(u), _p(v),
                     _p(out), _p(b1), _p(b2), ctypes.c_long(x.shape[0]))
    assert rc == 0
    return out


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_universal_point_ops(hc, name, tag):
    from graphembed import _lib as L
    g = load_golden(name, tag)
    x, y, u, v, eg, far = (g[k].contiguous() for k in ('x', 'y', 'u', 'v', 'eg', 'far'))
    T = load_truth(name) if tag == 'f32' else None
    assert_parity(_hc_point(hc, g, L.GM_OP_EXP, x, u), g, 'exp', tag, T)
    assert_parity(_hc_point(hc, g, L.GM_OP_RETR, x, u), g, 'retr', tag, T)
    assert_parity(_hc_point(hc, g, L.GM_OP_LOG, x, y), g, 'log', tag, T)
    assert_parity(_hc_point(hc, g, L.GM_OP_EGRAD2RGRAD, x, eg), g, 'egrad2rgrad', tag, T)
    assert_parity(_hc_point(hc, g, L.GM_OP_TRANSP, x, y, u), g, 'transp', tag, T)
    assert_parity(_hc_point(hc, g, L.GM_OP_NORM2, x, u, scalar=True).reshape(g['norm2'].shape), g, 'norm2', tag, T)
    diag = lambda d: None if d is None else {'inner': d['inner'].diagonal()}  # noqa: E731
    assert_parity(_hc_point(hc, g, L.GM_OP_INNER, x, u, v, scalar=True), diag(g), 'inner', tag, diag(T))
    assert_parity(_hc_point(hc, g, L.GM_OP_PROJX, far), g, 'projx', tag, T)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('oname', sorted(U_OPTS))
@pytest.mark.parametrize('name', sorted(UNIVERSAL_CASES))
def test_universal_optimizer_update(hc, name, oname, tag):
    from graphembed import _lib as L
    g = load_golden(name, tag)
    kind, kw = U_OPTS[oname]
    x = g['x'].clone().contiguous()
    radam = kind == 'radam'
    b1 = torch.zeros_like(x) if (radam or kw.get('momentum', 0) > 0) else None
    b2 = torch.zeros_like(x) if radam else None
    T = load_truth(name) if tag == 'f32' else None
    for k in range(3):
        opt = L.Optim(kind=L.GM_OPT_RADAM if radam else L.GM_OPT_RSGD, exact=int(kw.get('exact', False)),
                      has_clip=int('max_grad_norm' in kw), step=k + 1, has_momentum=int(kw.get('momentum', 0) > 0),
                      first_step=int(k == 0), grassmann_retr_qr=0, zero_grad=0, lr=kw['lr'], beta1=0.9, beta2=0.999,
                      momentum=kw.get('momentum', 0.0), dampening=kw.get('dampening', 0.0),
                      max_grad_norm=kw.get('max_grad_norm', 0.0), eps=1e-8)
        _hc_point(hc, g, -1, x, g['opt_grads'][k].contiguous(), opt=opt, b1=b1, b2=b2)
        assert_parity(x, g, f'{oname}_x', tag, T, index=k)
    for key, buf in (('exp_avg', b1 if radam else None), ('exp_avg_sq', b2), ('momentum_buffer', None if radam else b1)):
        if f'{oname}_{key}' in g and buf is not None:
            assert_parity(buf, g, f'{oname}_{key}', tag, T)


# ---- every other manifold: point ops and optimizer trajectories of the kernel arithmetic, same bar as the GPU tests --------
def _hc_point_any(hc, name, op, x, u=None, v=None, scalar=False, opt=None, b1=None, b2=None):
    kind, n, p, flags = _desc(name)
    out = torch.empty(x.shape[0] if scalar else x.shape, dtype=x.dtype)
    rc = hc.hc_point(kind, 0 if x.dtype == torch.float32 else 1, n, p, flags, ctypes.c_double(1e-8),
                     ctypes.c_double(1e8), 0, op, None if opt is None else ctypes.byref(opt), _p(x), _p(u), _p(v),
                     _p(out), _p(b1), _p(b2), ctypes.c_long(x.shape[0]))
    assert rc == 0
    return out


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_point_ops_all_manifolds(hc, name, tag):
    from graphembed import _lib as L
    g = load_golden(name, tag)
    T = load_truth(name) if tag == 'f32' else None
    x, y, u, v, eg = (g[k].contiguous() for k in ('x', 'y', 'u', 'v', 'eg'))
    assert_parity(_hc_point_any(hc, name, L.GM_OP_EXP, x, u), g, 'exp', tag, T)
    assert_parity(_hc_point_any(hc, name, L.GM_OP_RETR, x, u), g, 'retr', tag, T)
    assert_parity(_hc_point_any(hc, name, L.GM_OP_LOG, x, y), g, 'log', tag, T)
    assert_parity(_hc_point_any(hc, name, L.GM_OP_PROJU, x, eg), g, 'proju', tag, T)
    assert_parity(_hc_point_any(hc, name, L.GM_OP_EGRAD2RGRAD, x, eg), g, 'egrad2rgrad', tag, T)
    assert_parity(_hc_point_any(hc, name, L.GM_OP_TRANSP, x, y, u), g, 'transp', tag, T)
    assert_parity(_hc_point_any(hc, name, L.GM_OP_NORM2, x, u, scalar=True).reshape(g['norm2'].shape), g, 'norm2',
                  tag, T)
    assert_parity(_hc_point_any(hc, name, L.GM_OP_INNER, x, u, v, scalar=True).reshape(g['inner'].shape), g, 'inner',
                  tag, T)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
@pytest.mark.parametrize('oname', sorted(U_OPTS))
@pytest.mark.parametrize('name', sorted(CASES))
def test_optimizer_update_all_manifolds(hc, name, oname, tag):
    from graphembed import _lib as L
    g = load_golden(name, tag)
    T = load_truth(name) if tag == 'f32' else None
    kind, kw = U_OPTS[oname]
    x = g['x'].clone().contiguous()
    radam = kind == 'radam'
    b1 = torch.zeros_like(x) if (radam or kw.get('momentum', 0) > 0) else None
    b2 = torch.zeros_like(x) if radam else None
    for k in range(3):
        opt = L.Optim(kind=L.GM_OPT_RADAM if radam else L.GM_OPT_RSGD, exact=int(kw.get('exact', False)),
                      has_clip=int('max_grad_norm' in kw), step=k + 1, has_momentum=int(kw.get('momentum', 0) > 0),
                      first_step=int(k == 0), grassmann_retr_qr=0, zero_grad=0, lr=kw['lr'], beta1=0.9, beta2=0.999,
                      momentum=kw.get('momentum', 0.0), dampening=kw.get('dampening', 0.0),
                      max_grad_norm=kw.get('max_grad_norm', 0.0), eps=1e-8)
        _hc_point_any(hc, name, -1, x, g['opt_grads'][k].contiguous(), opt=opt, b1=b1, b2=b2)
        assert_parity(x, g, f'{oname}_x', tag, T, index=k, what=f'{oname} step {k}')
    for key, buf in (('exp_avg', b1 if radam else None), ('exp_avg_sq', b2), ('momentum_buffer', None if radam else b1)):
        if f'{oname}_{key}' in g and buf is not None:
            assert_parity(buf, g, f'{oname}_{key}', tag, T)
