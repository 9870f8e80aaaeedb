"""TEST INFRASTRUCTURE -- builds oracle/_build/liboracle.so (gcc) and exposes `paint` through ctypes."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build():
    src = os.path.join(HERE, "paint_c.c")
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, src, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_paint.restype = None
        _lib.oracle_paint.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def paint(pos, mass, Nmesh, BoxSize, resampler="cic", shift=0.0, out=None):
    """same contract as pmesh_oracle.paint (full mesh, f8 accumulation)"""
    sup = {"nnb": 1, "nearest": 1, "cic": 2, "tsc": 3, "pcs": 4}[resampler]
    N = (np.asarray(Nmesh, dtype="i8") * np.ones(3, dtype="i8")).copy()
    L = (np.asarray(BoxSize, dtype="f8") * np.ones(3)).copy()
    pos = np.ascontiguousarray(pos)
    if pos.dtype not in (np.float32, np.float64):
        pos = pos.astype("f8")
    if out is None:
        out = np.zeros(tuple(int(v) for v in N), dtype="f8")
    assert out.dtype == np.float64 and out.flags.c_contiguous
    m = None if mass is None else np.ascontiguousarray(mass, dtype="f8")
    lib().oracle_paint(pos.ctypes.data, int(pos.dtype == np.float32), len(pos),
                       m.ctypes.data if m is not None else None, sup, float(shift),
                       L.ctypes.data, N.ctypes.data, out.ctypes.data)
    return out
