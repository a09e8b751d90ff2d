"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference (dalab/matrix-manifolds
`graphembed`) from /root/reference so that golden fixtures can be generated and the
oracle restatement (oracle/manifolds_oracle.py) can be pinned against it.

/root/reference exists only in the build container, never on the GPU box, so nothing
on a `-m gpu` test path, in `smoke()` or in `bench.py` may import this module.

The reference targets a 2019 PyTorch nightly; three shims are needed on torch 2.11
(SURVEY.md section 8c):
  * matplotlib stubs (graphembed/monitor.py:1, graphembed/train.py:8)
  * torch.symeig  -> torch.linalg.eigh  (graphembed/linalg/torch_batch.py:127-135)
  * torch.solve   -> torch.linalg.solve (graphembed/manifolds/grassmann.py:85)
"""
import os
import sys
import types
from collections import namedtuple

import torch

REFERENCE_ROOT = '/root/reference/graphembed'


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'graphembed'))


def load():
    """Returns the reference `graphembed` package (imported under its own name)."""
    if not available():
        raise RuntimeError('reference tree not present (expected on the GPU box)')
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.cm',
                 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    # torch 2.11 still has the names but they raise: always override.
    _Out = namedtuple('symeig', ['eigenvalues', 'eigenvectors'])

    def symeig(x, eigenvectors=False, upper=True):
        w, v = torch.linalg.eigh(x, UPLO='U' if upper else 'L')
        return _Out(w, v)

    torch.symeig = symeig
    _Sol = namedtuple('solve', ['solution', 'LU'])
    torch.solve = lambda b, a: _Sol(torch.linalg.solve(a, b), None)
    import warnings
    warnings.filterwarnings('ignore', message='.*deprecated.*')
    # our drop-in package is also called `graphembed`; make sure the reference wins here
    for k in [k for k in sys.modules if k == 'graphembed' or k.startswith('graphembed.')]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import graphembed  # noqa: F401
        import graphembed.manifolds  # noqa: F401
        import graphembed.optim  # noqa: F401
        import graphembed.objectives  # noqa: F401
        import graphembed.modules  # noqa: F401
        import graphembed.data.dataset  # noqa: F401
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return sys.modules['graphembed']
