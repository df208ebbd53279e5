"""ctypes access to oracle/libm_exact.so (platform libm array wrappers) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libm_exact.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "csrc", "libm_exact.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", _SO, "-lm"])
    return _SO


def _get():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.nsem_or_pow.argtypes = [dp, ctypes.c_double, dp, ctypes.c_size_t]
        for n in ("nsem_or_sqrt", "nsem_or_exp", "nsem_or_cos", "nsem_or_sin", "nsem_or_acos"):
            getattr(_lib, n).argtypes = [dp, dp, ctypes.c_size_t]
        _lib.nsem_or_atan2.argtypes = [dp, dp, dp, ctypes.c_size_t]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def pow_(x, e: float):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _get().nsem_or_pow(_p(x), float(e), _p(out), x.size)
    return out


def _unary(name, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    getattr(_get(), name)(_p(x), _p(out), x.size)
    return out


def sqrt_(x):
    return _unary("nsem_or_sqrt", x)


def exp_(x):
    return _unary("nsem_or_exp", x)


def cos_(x):
    return _unary("nsem_or_cos", x)


def sin_(x):
    return _unary("nsem_or_sin", x)


def acos_(x):
    return _unary("nsem_or_acos", x)


def atan2_(y, x):
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _get().nsem_or_atan2(_p(y), _p(x), _p(out), x.size)
    return out
