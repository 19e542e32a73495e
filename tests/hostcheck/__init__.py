"""TEST INFRASTRUCTURE: g++ build + ctypes wrapper of tests/hostcheck/hostcheck.cpp, which runs the
__host__ __device__ bodies of the CUDA kernels (spmf_b200/csrc/spmf_model.cuh) on the CPU."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "hostcheck.cpp")
_SO = os.path.join(_HERE, "_build", "libhostcheck.so")
_DEPS = [_SRC] + [os.path.join(_HERE, "..", "..", "spmf_b200", "csrc", f)
                  for f in ("spmf_model.cuh", "spmf_math.cuh")]


def _build():
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if os.path.exists(_SO) and all(os.path.getmtime(d) <= os.path.getmtime(_SO) for d in _DEPS):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", _SRC,
                           "-o", _SO, "-lm"])


_build()
lib = C.CDLL(_SO)
_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ll = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
lib.hc_gamma_grad.argtypes = [_f, _f, _f, C.c_int]
lib.hc_digamma.argtypes = [_f, _f, C.c_int]
lib.hc_gamma_grad4.argtypes = [_f, _f, _f, C.c_int]
lib.hc_softplus.argtypes = [_f, _f, _f, C.c_int]
lib.hc_normals.argtypes = [_f, C.c_longlong, C.c_uint, C.c_uint, C.c_ulonglong]
lib.hc_gammas.argtypes = [_f, C.c_longlong, C.c_float, C.c_uint, C.c_ulonglong]
lib.hc_philox.argtypes = [C.c_uint] * 6 + [_u]
lib.hc_layout.argtypes = [C.c_int, C.c_int, C.c_int, _ll, _ll]
lib.hc_draw_operands.argtypes = [_f, _f, _f, C.c_int, C.c_int, C.c_int, _f, _f, _f]
lib.hc_backward_params.argtypes = [_f, _f, _f, C.c_int, C.c_int, C.c_int, _f, _f, _f, C.c_float,
                                   C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _f, _d]


def gamma_grad(a, x):
    a = np.ascontiguousarray(a, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(a)
    lib.hc_gamma_grad(a, x, out, a.size)
    return out


def gamma_grad4(a, x):
    a = np.ascontiguousarray(a, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(a)
    lib.hc_gamma_grad4(a, x, out, a.size)
    return out


def digamma(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib.hc_digamma(x, out, x.size)
    return out


def normals(n, stream=0, step=0, seed=0):
    out = np.empty(n, np.float32)
    lib.hc_normals(out, n, stream, step, seed)
    return out


def gammas(n, alpha, stream=0, seed=0):
    out = np.empty(n, np.float32)
    lib.hc_gammas(out, n, alpha, stream, seed)
    return out


def philox(ctr, key):
    out = np.empty(4, np.uint32)
    lib.hc_philox(*[int(c) for c in ctr], int(key[0]), int(key[1]), out)
    return out


def layout(D, K, S):
    t = np.empty(25, np.int64); n = np.empty(13, np.int64)
    lib.hc_layout(D, K, S, t, n)
    return t, n


INTERNAL_ORDER = ['v', 'w', 'u', 's', 'u_eta', 'u_tau', 's_eta', 's_tau',
                  'u_eta_a', 'u_tau_a', 's_eta_a', 's_tau_a']
NORMAL = ('v', 'w', 'u', 's')


def pack_params(params, D, K, S):
    """oracle param dict (reference shapes, float64 torch) -> flat fp32 buffer (v transposed)."""
    toff, _ = layout(D, K, S)
    flat = np.zeros(toff[-1], np.float32)
    for vi, name in enumerate(INTERNAL_ORDER):
        a, b = ((name + '/loc', name + '/scale_raw') if name in NORMAL
                else (name + '/conc_raw', name + '/scale_raw'))
        for w, key in enumerate((a, b)):
            t = params[key].detach().numpy()
            if name == 'v':
                t = t.T
            t = np.ascontiguousarray(t, np.float32).ravel()
            flat[toff[2 * vi + w]: toff[2 * vi + w] + t.size] = t
    return flat


def unpack_grads(flat, D, K, S, shapes):
    toff, _ = layout(D, K, S)
    out = {}
    for vi, name in enumerate(INTERNAL_ORDER):
        a, b = ((name + '/loc', name + '/scale_raw') if name in NORMAL
                else (name + '/conc_raw', name + '/scale_raw'))
        shp = shapes[name]
        n = shp[0] * shp[1]
        for w, key in enumerate((a, b)):
            t = flat[toff[2 * vi + w]: toff[2 * vi + w] + n]
            out[key] = t.reshape(shp[1], shp[0]).T.copy() if name == 'v' else t.reshape(shp).copy()
    return out


def pack_noise(noise, D, K, S):
    _, noff = layout(D, K, S)
    flat = np.zeros(noff[-1], np.float32)
    for vi, name in enumerate(INTERNAL_ORDER):
        t = noise[name].detach().numpy()
        if name == 'v':
            t = np.swapaxes(t, -1, -2)
        t = np.ascontiguousarray(t, np.float32).ravel()
        flat[noff[vi]: noff[vi] + t.size] = t
    return flat


def _eta2(eta, D):
    """[2][D] scale table of the kernels: decoder scale | encoder divisor (both eta_i for the linear link)."""
    e = np.ascontiguousarray(eta, np.float32).reshape(-1)
    return np.ascontiguousarray(np.concatenate([e, e]) if e.size == D else e, np.float32)


def draw_operands(P, N, eta, D, K, S):
    Ap = np.zeros((S, D, K), np.float32); EV = np.zeros((S, D, K), np.float32); PH = np.zeros((S, D), np.float32)
    lib.hc_draw_operands(P, N, _eta2(eta, D), D, K, S, Ap, EV, PH)
    return Ap, EV, PH


def backward_params(P, N, eta, D, K, S, GAp, GEV, Gphinz, batch_rows, u_tau_scale, s_tau_scale, decay,
                    w_entropy=1.0, w_prior=1.0, world=1):
    grads = np.zeros_like(P)
    parts = np.zeros((S, 16), np.float64)
    lib.hc_backward_params(P, N, _eta2(eta, D), D, K, S,
                           np.ascontiguousarray(GAp, np.float32), np.ascontiguousarray(GEV, np.float32),
                           np.ascontiguousarray(Gphinz, np.float32), batch_rows, u_tau_scale,
                           s_tau_scale, decay, w_entropy, w_prior, world, grads, parts)
    return grads, parts
