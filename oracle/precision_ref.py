"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libprecision_ref.so: the REFERENCE's own FastPrecision
(graphembed/pyx/impl/precision.cpp, compiled unmodified from /root/reference by oracle/Makefile with a stand-in for the
one Boost container it uses).  It exists to pin oracle/precision_oracle.c (the plain-C restatement) and to generate
tests/golden/precision_f1.npz; the product never loads it."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_ref', 'libprecision_ref.so')
_lib = None


def available():
    return os.path.isfile(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
        _lib.fpref_create.restype = ctypes.c_void_p
        _lib.fpref_map_f64.restype = ctypes.c_double
        _lib.fpref_map_f32.restype = ctypes.c_double
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class FastPrecisionRef:
    """The reference's FastPrecision on a CSR adjacency (same method names as pyx/precision.pyx:48-127)."""

    def __init__(self, rowptr, colidx):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        self.n = self.rowptr.size - 1
        self.h = ctypes.c_void_p(_load().fpref_create(self.n, _ptr(self.rowptr), _ptr(self.colidx)))

    def __del__(self):
        if getattr(self, 'h', None):
            _load().fpref_destroy(self.h)
            self.h = None

    def nodes_per_layer(self):
        out = np.empty(self.n + 1, dtype=np.int32)
        k = _load().fpref_nodes_per_layer(self.h, _ptr(out), out.size)
        return out[:k]

    def mean_average_precision(self, mpdists):
        mp = np.ascontiguousarray(mpdists)
        if mp.dtype == np.float32:
            return _load().fpref_map_f32(self.h, _ptr(mp))
        mp = np.ascontiguousarray(mp, dtype=np.float64)
        return _load().fpref_map_f64(self.h, _ptr(mp))

    def layer_mean_f1_scores(self, mpdists, min_degree=1, max_degree=(1 << 62)):
        mp = np.ascontiguousarray(mpdists)
        means, stds = np.empty(self.n), np.empty(self.n)
        fn = _load().fpref_layer_f1_f32 if mp.dtype == np.float32 else _load().fpref_layer_f1_f64
        if mp.dtype != np.float32:
            mp = np.ascontiguousarray(mp, dtype=np.float64)
        k = fn(self.h, _ptr(mp), ctypes.c_long(int(min_degree)), ctypes.c_long(int(max_degree)), _ptr(means), _ptr(stds),
               self.n)
        return means[:k], stds[:k]

    def layer_mean_average_f1_scores(self, mpdists):
        mp = np.ascontiguousarray(mpdists, dtype=np.float64)
        means, stds = np.empty(self.n), np.empty(self.n)
        k = _load().fpref_layer_avg_f1_f64(self.h, _ptr(mp), _ptr(means), _ptr(stds), self.n)
        return means[:k], stds[:k]
