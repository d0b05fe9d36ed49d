"""TEST INFRASTRUCTURE ONLY - ctypes wrapper of oracle/truth_ld.c (long-double exact-GP posterior of a sphere
series kernel).  Used by tests/ to attribute float64 errors of the CUDA path and of the float64 oracle
(oracle/restated.py) to one side or the other.  The product never imports this module."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "truth_ld.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libmgbtruth.so")


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -shared; the .so is git-ignored and travels to the GPU box with the gpurun snapshot."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB + ".tmp.%d" % os.getpid()
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", tmp, SRC, "-lm"])
    os.replace(tmp, LIB)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.mgb_truth_series_posterior.restype = C.c_long
        lib.mgb_truth_series_posterior.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_int,
                                                   C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                                   C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        lib.mgb_truth_long_double_digits.restype = C.c_int
        _lib = lib
    return _lib


def gegenbauer_power_table(terms: int, alpha: float) -> np.ndarray:
    """G[n][p] = coefficient of z^p in C_n^alpha(z), with the float64 numbers of the reference's explicit power sum
    (math_utils/gegenbauer_polynomials.py:31-33): (-1)^i 2^(n-2i) Gamma(n-i+alpha) / (Gamma(alpha) i! (n-2i)!)."""
    g = np.zeros((terms, terms))
    ga = math.gamma(alpha)
    for n in range(terms):
        for i in range(n // 2 + 1):
            c = math.gamma(n - i + alpha) / (ga * math.factorial(i) * math.factorial(n - 2 * i))
            g[n, n - 2 * i] = math.pow(-1, i) * c * 2.0 ** (n - 2 * i)
    return g


def sphere_matern_weights(dim: int, nu: float, lengthscale: float, terms: int = 10) -> np.ndarray:
    """Normalised series weights of kernels_sphere.py:126-139 in float64 (the restated oracle's numbers)."""
    import torch

    from . import restated as R

    cst, g1 = R._sphere_tables(dim, terms, torch.float64)
    nu_t, ls = R._as_param(nu), R._as_param(lengthscale)
    e0 = torch.pow(2 * nu_t / ls ** 2, -(nu_t + dim / 2))
    a = [torch.pow(2 * nu_t / ls ** 2 + n * (n + dim - 1), -(nu_t + dim / 2)) / e0 * cst[n] for n in range(terms)]
    norm = sum(a[n] * g1[n] for n in range(terms))
    return np.array([float((a[n] / norm).reshape(())) for n in range(terms)])


def sphere_matern_posterior(xt, y, xc, dim, nu, lengthscale, outputscale=1.0, noise=1e-2, mean_const=0.0, terms=10,
                            threads=0):
    """Long-double posterior mean / variance (numpy float64 arrays, rounded once at the end)."""
    lib = load()
    if lib.mgb_truth_long_double_digits() < 64:
        raise RuntimeError("long double is not wider than double on this machine: no truth leg")
    xt = np.ascontiguousarray(xt, dtype=np.float64)
    xc = np.ascontiguousarray(xc, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.ascontiguousarray(sphere_matern_weights(dim, nu, lengthscale, terms))
    g = np.ascontiguousarray(gegenbauer_power_table(terms, (dim - 1.0) / 2.0))
    n, d = xt.shape
    m = xc.shape[0]
    mean, var = np.empty(m), np.empty(m)
    rc = lib.mgb_truth_series_posterior(xt.ctypes.data, n, y.ctypes.data, xc.ctypes.data, m, d, w.ctypes.data,
                                        g.ctypes.data, terms, -1.0 + 1e-15, 1.0 - 1e-15, outputscale, noise, mean_const,
                                        mean.ctypes.data, var.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError("mgb_truth_series_posterior failed: %d" % rc)
    return mean, var
