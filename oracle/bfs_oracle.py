"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes wrapper of oracle/bfs_oracle.c (built into oracle/_build/)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libbfs_oracle.so')
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            import subprocess
            subprocess.check_call(['make', '-C', _HERE])
        _lib = ctypes.CDLL(_SO)
    return _lib


def hops_from(rowptr, colidx, sources):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    sources = np.ascontiguousarray(sources, dtype=np.int32)
    n = rowptr.size - 1
    out = np.empty((sources.size, n), dtype=np.int32)
    rc = _load().bfs_many(rowptr.ctypes.data_as(ctypes.c_void_p), colidx.ctypes.data_as(ctypes.c_void_p), n,
                          sources.ctypes.data_as(ctypes.c_void_p), sources.size, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def all_pairs_hops(rowptr, colidx):
    return hops_from(rowptr, colidx, np.arange(len(rowptr) - 1, dtype=np.int32))
