"""ctypes wrapper of ``oracle/csn.c`` (TEST INFRASTRUCTURE ONLY, see ``oracle/__init__.py``): the chordal
triangular solve with dense right-hand sides on compiled code and all host cores, for the timed CPU baseline of
``bench.py``.  ``oracle/supernodal.py:trsm`` stays the specification; ``tests/test_oracle_drivers.py`` pins this on it."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "csn.so")
_lib = None


def available():
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(_SO)
        p64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
        pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        lib.csn_trsm.restype = C.c_int
        lib.csn_trsm.argtypes = [C.c_int64, p64, p64, p64, p64, pd, pd, C.c_int64, C.c_int]
        _lib = lib
    return _lib


def trsm(symb, L, Bm, trans="N"):
    """Bm (n x k, C-contiguous, rows in the internal order of ``symb``) <- L^-1 Bm ('N') or L^-T Bm ('T'), in place.
    Returns the number of threads used."""
    lib = _load()
    assert Bm.flags["C_CONTIGUOUS"] and Bm.dtype == np.float64 and Bm.ndim == 2
    nn = np.ascontiguousarray(symb.nn, dtype=np.int64)
    rowptr = np.ascontiguousarray(symb.rowptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(symb.rowidx, dtype=np.int64)
    blkptr = np.ascontiguousarray(symb.blkptr, dtype=np.int64)
    Lf = np.ascontiguousarray(np.ravel(L), dtype=np.float64)
    return int(lib.csn_trsm(int(symb.nsn), nn, rowptr, rowidx, blkptr, Lf, Bm, int(Bm.shape[1]), 0 if trans == "N" else 1))
