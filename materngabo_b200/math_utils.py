"""Host helpers with the reference's math_utils API (BoManifolds/math_utils/gegenbauer_polynomials.py:12-43,
jacobi_theta_functions.py:16-20).  They are used for construction-time constants and small host-side
evaluations only; the device kernels evaluate the same polynomials in registers."""
from __future__ import annotations

import math

import torch


def gegenbauer_polynomial(n, alpha, z):
    """C_n^alpha(z) as the explicit power sum sum_i (-1)^i (2z)^(n-2i) Gamma(n-i+alpha) / (Gamma(alpha) i! (n-2i)!).
    Raises ValueError for alpha = 0 (math.gamma(0)), like the reference, i.e. S^1 needs the circle kernels."""
    gamma_alpha = math.gamma(alpha)
    total = 0.
    for i in range(n // 2 + 1):
        weight = math.gamma(n - i + alpha) / (gamma_alpha * math.factorial(i) * math.factorial(n - 2 * i))
        total = total + math.pow(-1, i) * torch.pow(2 * z, n - 2 * i) * weight
    return total


def jacobi_theta_function3(z, q, serie_nb_terms=200):
    """theta_3(z, q) = 1 + 2 sum_{n>=1} q^(n^2) cos(2 n z), truncated after serie_nb_terms terms."""
    n = torch.arange(1, serie_nb_terms + 1, dtype=z.dtype, device=z.device)
    qn = torch.pow(q.reshape(-1)[0].to(z.dtype), n ** 2)
    return 1. + 2 * (torch.cos(2 * z.unsqueeze(-1) * n) * qn).sum(-1)
