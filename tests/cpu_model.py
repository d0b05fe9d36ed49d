"""TEST INFRASTRUCTURE: numpy model of the C-ABI ops' *definition* (include/mgb.h), used on CPU to check the
host-side coefficient / node-weight code against the golden vectors without a GPU.  It is not an
implementation of the reference algorithm (that is oracle/restated.py) and is never imported by the product."""
import numpy as np


def series_gram(desc, x1, x2, coef):
    """desc: dict(n_factors, factor_width, basis, scale, shift, lo, hi); coef (F, T)."""
    f_, w_ = desc["n_factors"], desc["factor_width"]
    coef = np.asarray(coef, dtype=np.float64).reshape(f_, -1)
    out = np.ones((x1.shape[0], x2.shape[0]))
    for f in range(f_):
        a = x1[:, f * w_:(f + 1) * w_]
        b = x2[:, f * w_:(f + 1) * w_]
        u = np.clip(desc["scale"] * (a @ b.T) + desc["shift"], desc["lo"], desc["hi"])
        if desc["basis"] == 0:
            s = np.polynomial.polynomial.polyval(u, coef[f])
        else:
            s = np.polynomial.chebyshev.chebval(u, coef[f])
        out = out * s
    return out


def spec_to_desc(spec):
    return dict(n_factors=spec.n_factors, factor_width=spec.factor_width, basis=spec.basis, scale=spec.scale,
                shift=spec.shift, lo=spec.lo, hi=spec.hi)
