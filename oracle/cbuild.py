"""TEST INFRASTRUCTURE: gcc build + ctypes wrapper of the oracle's plain-C pieces (oracle/*.c).

Only tests/, __graft_entry__.build()/smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRCS = [os.path.join(_HERE, "gamma_der.c")]
_SO = os.path.join(_HERE, "_build", "libspmf_oracle_c.so")


def build():
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if os.path.exists(_SO) and all(os.path.getmtime(s) <= os.path.getmtime(_SO) for s in _SRCS):
        return _SO
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO] + _SRCS + ["-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        _lib.spmf_oracle_gamma_der.argtypes = [d, d, d, C.c_longlong]
        _lib.spmf_oracle_gamma_der.restype = None
    return _lib


def gamma_sample_der_alpha_c(alpha, g):
    """dg/dalpha, float64 numpy, scalar C loops (oracle/gamma_der.c) -- same expansions as
    spmf_oracle.gamma_sample_der_alpha."""
    a, x = np.broadcast_arrays(np.asarray(alpha, np.float64), np.asarray(g, np.float64))
    shape = a.shape
    a = np.ascontiguousarray(a).reshape(-1)
    x = np.ascontiguousarray(x).reshape(-1)
    out = np.empty_like(x)
    lib().spmf_oracle_gamma_der(a, x, out, x.size)
    return out.reshape(shape)
