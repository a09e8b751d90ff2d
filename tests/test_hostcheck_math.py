"""CPU-only check of the *kernel arithmetic*: csrc/gm_manifolds.cuh and csrc/gm_pointops.cuh are __host__ __device__,
tests/hostcheck compiles them for x86 and this test compares them with the pinned oracle on seeded inputs.
(The product never runs this host build; it exists so that kernel math regressions are caught without a GPU.)"""
import ctypes
import os
import subprocess

import pytest
import torch

from helpers import CASES, load_golden, make_oracle, is_spd, rel_err, sym

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
    t = 1e-10 if tag == 'f64' else (2e-4 if ('default_init' in name or 'stein' in name or 'grass' in name) else 2e-5)
    assert rel_err(d2, g['dist2']) < t
    assert rel_err(fix(gx * w), fix(g['gx'])) < t * 10
    assert rel_err(fix(gy * w), fix(g['gy'])) < t * 10
