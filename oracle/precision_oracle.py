"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/precision_oracle.c (the plain-C restatement of the
reference's FastPrecision, graphembed/pyx/impl/precision.cpp).  See the header of the C file for the pinning status."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libprecision_oracle.so')
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            subprocess.check_call(['make', '-C', _HERE])
        _lib = ctypes.CDLL(_SO)
        _lib.fp_mean_average_precision.restype = ctypes.c_double
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class FastPrecisionOracle:
    """Same methods as the reference's PyFastPrecision (pyx/precision.pyx:48-127), on a CSR adjacency."""

    def __init__(self, rowptr, colidx):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.n = self.rowptr.size - 1

    def mean_average_precision(self, mpdists):
        mp = np.ascontiguousarray(mpdists, dtype=np.float64)
        return _load().fp_mean_average_precision(self.n, _ptr(self.rowptr), _ptr(self.colidx), _ptr(mp))

    def _f1(self, mpdists, mode, min_degree=1, max_degree=99999):
        mp = np.ascontiguousarray(mpdists, dtype=np.float64)
        means, stds = np.empty(self.n), np.empty(self.n)
        k = _load().fp_layer_f1(self.n, _ptr(self.rowptr), _ptr(self.colidx), _ptr(mp), int(min_degree), int(max_degree),
                                mode, _ptr(means), _ptr(stds))
        return means[:k], stds[:k]

    def layer_mean_f1_scores(self, mpdists, min_degree=1, max_degree=99999):
        return self._f1(mpdists, 0, min_degree, max_degree)

    def layer_mean_average_f1_scores(self, mpdists):
        return self._f1(mpdists, 1)

    def nodes_per_layer(self):
        out = np.empty(self.n, dtype=np.int32)
        k = _load().fp_nodes_per_layer(self.n, _ptr(self.rowptr), _ptr(self.colidx), _ptr(out))
        return out[:k]
