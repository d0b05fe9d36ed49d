"""CPU: the long-double truth leg (oracle/truth_ld.c) against the float64 oracle at sizes where float64 is
comfortably accurate - they must agree to ~1e-13, which pins the truth leg's equations, coefficient tables and clamp."""
import numpy as np
import pytest
import torch

from oracle import restated as R, truth


@pytest.mark.parametrize("dim,n,noise", [(10, 400, 1e-2), (3, 250, 1e-4), (2, 60, 1e-2)])
def test_truth_leg_agrees_with_float64_oracle(dim, n, noise):
    gen = torch.Generator().manual_seed(dim + n)
    xt = torch.nn.functional.normalize(torch.randn(n, dim + 1, generator=gen), dim=-1)
    xc = torch.nn.functional.normalize(torch.randn(64, dim + 1, generator=gen), dim=-1)
    xc[:8] = xt[:8]                                   # clamp active: <x, x> rounds above 1 - 1e-15
    y = torch.sin(3 * xt[:, 0])
    mt, vt = truth.sphere_matern_posterior(xt.numpy(), y.numpy(), xc.numpy(), dim, 2.5, 0.5, 1.3, noise, 0.2)
    kfn = lambda a, b: R.sphere_matern(a, b, dim, 2.5, 0.5)  # noqa: E731
    mr, vr = R.gp_posterior(kfn, xt, y, xc, 1.3, noise, 0.2)
    assert np.abs(mr.numpy() - mt).max() < 1e-12 * np.abs(mt).max()
    assert np.abs(vr.numpy() - vt).max() < 1e-12 * vt.max()
    assert truth.load().mgb_truth_long_double_digits() >= 64


def test_gegenbauer_power_table_matches_the_oracle_polynomials():
    z = torch.linspace(-1, 1, 41, dtype=torch.float64)
    g = truth.gegenbauer_power_table(10, 4.5)
    for n in range(10):
        ref = R.gegenbauer(n, 4.5, z)
        val = sum(g[n, p] * z ** p for p in range(10))
        assert float((val - ref).abs().max()) < 1e-12 * max(1.0, float(torch.as_tensor(ref).abs().max()))
