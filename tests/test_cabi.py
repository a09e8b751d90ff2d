"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/gm_kernels.h
declares, rejects bad arguments with error codes (no launch), and the Python layer fails loudly without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'gm_kernels.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gm_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from graphembed import _lib
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in gm_kernels.h but not exported'
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    assert b'sm_100a' in lib.gm_version()


def test_argument_validation_without_gpu():
    from graphembed import _lib as L
    lib = L.lib()
    man = L.Manifold(kind=L.GM_SPD_AI, dtype=L.GM_F32, n=4, p=0, flags=0, reserved=0, wmin=1e-8, wmax=1e8)
    assert lib.gm_supported(ctypes.byref(man)) == 1
    bad = L.Manifold(kind=L.GM_SPD_AI, dtype=L.GM_F32, n=11, p=0, flags=0, reserved=0, wmin=1e-8, wmax=1e8)
    assert lib.gm_supported(ctypes.byref(bad)) == 0
    bad = L.Manifold(kind=L.GM_SPD_AI, dtype=L.GM_F32, n=4, p=0, flags=L.GM_FAST_CHOL, reserved=0, wmin=0, wmax=1)
    assert lib.gm_supported(ctypes.byref(bad)) == 0
    pairs = L.Pairs(mode=L.GM_PAIRS_TRIU, idx64=0, P=7, idx_i=None, idx_j=None, B=5, nodes=None)  # 5*4/2 != 7
    assert lib.gm_pairs_dist2(ctypes.byref(man), None, None, ctypes.byref(pairs), None, None) == -1
    pairs = L.Pairs(mode=L.GM_PAIRS_TRIU, idx64=0, P=10, idx_i=None, idx_j=None, B=5, nodes=None)
    assert lib.gm_pairs_dist2(ctypes.byref(man), None, None, ctypes.byref(pairs), None, None) == -3  # NULL data
    empty = L.Pairs(mode=L.GM_PAIRS_LIST, idx64=0, P=0, idx_i=None, idx_j=None, B=0, nodes=None)
    assert lib.gm_pairs_dist2(ctypes.byref(man), None, None, ctypes.byref(empty), None, None) == 0  # empty input
    assert lib.gm_bfs_workspace_bytes(1000, 64) >= 3 * 1000 * 8
    assert lib.gm_launch_count() == 0 or torch.cuda.is_available()


def test_no_cpu_fallback():
    from graphembed.manifolds import SymmetricPositiveDefinite, Lorentz
    from graphembed.modules import ManifoldParameter
    from graphembed.optim import RiemannianAdam
    man = SymmetricPositiveDefinite(3)
    x = torch.eye(3).repeat(4, 1, 1)
    with pytest.raises(RuntimeError, match='CUDA'):
        man.pdist(x)
    with pytest.raises(RuntimeError, match='CUDA'):
        man.exp(x, torch.zeros_like(x))
    with pytest.raises(RuntimeError, match='CUDA'):
        Lorentz(4).rand(3)
    p = ManifoldParameter(x.clone(), manifold=man)
    p.grad = torch.zeros_like(x)
    with pytest.raises(RuntimeError, match='CUDA'):
        RiemannianAdam([p]).step()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'matrix-manifolds_b200', 'graphembed')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no CPU', ''), f'{f} mentions the oracle'
                assert 'manifolds_oracle' not in src and 'ref_import' not in src


def test_reference_api_surface():
    """Same class names / ctor signatures / method names as the reference."""
    import inspect
    from graphembed import manifolds, optim, modules, objectives
    for cls in ('Manifold', 'Euclidean', 'Grassmann', 'Lorentz', 'SymmetricPositiveDefinite', 'Sphere'):
        assert hasattr(manifolds, cls)
    for meth in ('ndim', 'dim', 'zero', 'zero_vec', 'inner', 'norm', 'proju', 'projx', 'egrad2rgrad', 'exp', 'retr',
                 'log', 'dist', 'pdist', 'transp', 'rand', 'rand_uniform', 'randvec'):
        assert hasattr(manifolds.Manifold, meth), meth
    sig = inspect.signature(manifolds.SymmetricPositiveDefinite.__init__)
    assert list(sig.parameters)[1:] == ['n', 'fast_symeig', 'fast_chol', 'use_stein_div', 'wmin', 'wmax']
    assert list(inspect.signature(optim.RiemannianAdam.__init__).parameters)[1:] == \
        ['params', 'lr', 'betas', 'nc', 'max_grad_norm', 'exact']
    assert list(inspect.signature(optim.RiemannianSGD.__init__).parameters)[1:] == \
        ['params', 'lr', 'momentum', 'dampening', 'max_grad_norm', 'exact']
    assert {'xs', 'scales'} <= set(dir(modules.ManifoldEmbedding)) | {'xs', 'scales'}
    for name in ('QuotientLoss', 'StressLoss', 'Sum', 'ObjectiveFunction'):
        assert hasattr(objectives, name)
    with pytest.raises(ValueError):
        objectives.QuotientLoss(inc_l1=False, inc_l2=False)
    with pytest.raises(ValueError):
        manifolds.Grassmann(5, 2, retr='nope')
    assert manifolds.SymmetricPositiveDefinite(4).dim == 10 and manifolds.Lorentz(11).dim == 10
    assert manifolds.Grassmann(6, 2).dim == 8 and str(manifolds.Lorentz(5)) == 'Lorentzian space of dimension 5'
