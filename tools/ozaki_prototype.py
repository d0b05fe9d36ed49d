"""CPU prototype (numpy) of the error-free INT8 slicing proposed in DESIGN.md section 11 for the posterior's N^2
contraction: how many 7-bit slices per operand does the S^10 variance need for the 1e-10 parity bar?

V = K* L^-T (m x n),  var_i = kss - sum_j V_ij^2.  Each operand row is scaled by a power of two (its largest
magnitude) and cut into `s` signed slices of `bits` magnitude bits; the slice products are integer matrix products
(exact in int32 for n * 2^(2 bits) < 2^31 - emulated here by float64 BLAS on integer-valued matrices, which is exact
below 2^53), recombined in float64.  Only slice pairs with p + q <= s + 1 are formed (the rest is below the truncation
error of the slicing itself).

usage: python tools/ozaki_prototype.py [n_train] [n_cand] [noise]     (CPU only; imports the oracle - test infrastructure)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
import torch

torch.set_default_dtype(torch.float64)
from oracle import restated as R  # noqa: E402


ROUND = True      # round-to-nearest slices (|q| <= 2^(bits-1), zero-mean remainders) instead of truncation


def slices(a, s, bits):
    """a (rows x k) -> list of s integer-valued float64 matrices and the per-row scale 2^e with
    a = 2^e * sum_p slices[p] * 2^(-bits (p + 1)) + O(2^(e - bits s))."""
    e = np.floor(np.log2(np.abs(a).max(axis=1, keepdims=True) + 1e-300)) + (2.0 if ROUND else 1.0)
    r = a / np.exp2(e)                       # |r| < 1 (truncation) or <= 1/2 (rounding)
    out = []
    for _ in range(s):
        r = r * (1 << bits)
        q = np.rint(r) if ROUND else np.trunc(r)
        out.append(q)
        r = r - q
    return out, np.exp2(e)


def sliced_product(a, b, s, bits):
    """a @ b.T from slice products; returns (product, number of integer GEMMs, max |int32 accumulator|)."""
    sa, ea = slices(a, s, bits)
    sb, eb = slices(b, s, bits)
    acc = np.zeros((a.shape[0], b.shape[0]))
    gemms, peak = 0, 0.0
    for d in range(2, s + 2):                # d = (p + 1) + (q + 1): one int32 accumulator per anti-diagonal
        diag = np.zeros_like(acc)
        for p in range(s):
            q = d - 2 - p
            if 0 <= q < s:
                diag += sa[p] @ sb[q].T
                gemms += 1
        peak = max(peak, float(np.abs(diag).max()))
        acc += diag * 2.0 ** (-bits * d)
    return acc * ea * eb.T, gemms, peak


def main(n=1000, m=512, noise=1e-2):
    gen = torch.Generator().manual_seed(0)
    xt = torch.nn.functional.normalize(torch.randn(n, 11, generator=gen), dim=-1)
    xc = torch.nn.functional.normalize(torch.randn(m, 11, generator=gen), dim=-1)
    xc[: m // 4] = torch.nn.functional.normalize(xt[: m // 4] + 1e-3 * torch.randn(m // 4, 11, generator=gen), dim=-1)
    k = R.sphere_matern(xt, xt, 10, 2.5, 0.5).numpy()
    ks = R.sphere_matern(xc, xt, 10, 2.5, 0.5).numpy()
    L = np.linalg.cholesky(k + noise * np.eye(n))
    rinv = np.linalg.solve(L, np.eye(n))     # lower triangular
    # reference in extended precision
    v_ref = ks.astype(np.longdouble) @ rinv.T.astype(np.longdouble)
    var_ref = (1.0 - (v_ref * v_ref).sum(1)).astype(np.float64)
    v64 = ks @ rinv.T
    var64 = 1.0 - (v64 * v64).sum(1)
    scale = np.abs(var_ref).max()
    print("n = %d training points, m = %d candidates (a quarter of them 1e-3 from a training point), noise %.1e; min var %.3e, max |L^-1| %.2e"
          % (n, m, noise, var_ref.min(), np.abs(rinv).max()))
    print("float64 BLAS:                      max |var - ref| / max|var| = %.2e" % (np.abs(var64 - var_ref).max() / scale))
    global ROUND
    for ROUND in (False, True):
      print("round-to-nearest slices" if ROUND else "truncated slices")
      for bits in (7,):
        for s in (6, 7, 8):
            v, gemms, peak = sliced_product(ks, rinv, s, bits)
            var = 1.0 - (v * v).sum(1)
            print("  " +"bits %d, %d slices: %2d integer GEMMs, max |accumulator| 2^%.1f, max |var - ref| / max|var| = %.2e"
                  % (bits, s, gemms, np.log2(max(peak, 1.0)), np.abs(var - var_ref).max() / scale))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]), *(float(a) for a in sys.argv[3:4]))
