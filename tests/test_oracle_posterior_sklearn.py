"""not gpu: the posterior oracle (oracle/restated.py: gp_fit / gp_predict - the textbook exact-GP equations that
gpytorch's ExactGP prediction and botorch's analytic EI implement) against an INDEPENDENT third-party implementation of
the same algorithm: scikit-learn's GaussianProcessRegressor (Rasmussen & Williams, Algorithm 2.1; Cholesky of
K + alpha I, mean = K* alpha_vec, var = k** - ||L^-1 K*^T||^2) driven with the oracle's own manifold kernel.  gpytorch /
botorch are not installable in this image, so this is the closest available pin of the a19 arithmetic; the EI formula
is pinned against scipy's normal distribution."""
import math

import numpy as np
import pytest
import torch

from oracle import restated as R

sk = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import Kernel as SkKernel  # noqa: E402


class _OracleKernel(SkKernel):
    """scikit-learn kernel object evaluating outputscale * k_oracle(x, y) (no hyper-parameters of its own)."""

    def __init__(self, dim=3, outputscale=1.0):
        self.dim = dim
        self.outputscale = outputscale

    def __call__(self, X, Y=None, eval_gradient=False):
        a = torch.as_tensor(np.asarray(X), dtype=torch.float64)
        b = a if Y is None else torch.as_tensor(np.asarray(Y), dtype=torch.float64)
        k = (self.outputscale * R.sphere_matern(a, b, self.dim, 2.5, 0.5)).numpy()
        if eval_gradient:
            return k, np.empty((k.shape[0], k.shape[1], 0))
        return k

    def diag(self, X):
        a = torch.as_tensor(np.asarray(X), dtype=torch.float64)
        return (self.outputscale * R.sphere_matern(a.unsqueeze(-2), a.unsqueeze(-2), self.dim, 2.5, 0.5)).reshape(-1).numpy()

    def is_stationary(self):
        return False


@pytest.mark.parametrize("dim,n,noise,mean_const", [(3, 120, 1e-2, 0.0), (10, 300, 1e-3, 0.4), (2, 40, 5e-2, -0.2)])
def test_posterior_oracle_matches_scikit_learn(dim, n, noise, mean_const):
    gen = torch.Generator().manual_seed(dim * 100 + n)
    xt = torch.nn.functional.normalize(torch.randn(n, dim + 1, generator=gen), dim=-1)
    xc = torch.nn.functional.normalize(torch.randn(77, dim + 1, generator=gen), dim=-1)
    xc[:10] = xt[:10]
    y = torch.sin(3 * xt[:, 0]) + 0.1 * torch.randn(n, generator=gen)
    s = 1.3
    gpr = sk.GaussianProcessRegressor(kernel=_OracleKernel(dim, s), alpha=noise, optimizer=None, normalize_y=False)
    gpr.fit(xt.numpy(), (y - mean_const).numpy())
    mu, std = gpr.predict(xc.numpy(), return_std=True)
    kfn = lambda a, b: R.sphere_matern(a, b, dim, 2.5, 0.5)  # noqa: E731
    mo, vo = R.gp_posterior(kfn, xt, y, xc, outputscale=s, noise=noise, mean_const=mean_const)
    assert np.abs(mo.numpy() - (mu + mean_const)).max() < 1e-10 * max(1.0, np.abs(mu).max())
    # scikit-learn clips negative variances to 0 before the square root; compare where it did not clip
    var_sk = std ** 2
    ok = var_sk > 0
    assert ok.sum() >= 60
    assert np.abs(vo.numpy()[ok] - var_sk[ok]).max() < 1e-9 * s


def test_expected_improvement_matches_scipy_normal():
    from scipy.stats import norm

    gen = torch.Generator().manual_seed(1)
    mean = torch.randn(200, generator=gen, dtype=torch.float64)
    var = torch.rand(200, generator=gen, dtype=torch.float64) * 0.5
    var[:5] = 1e-12                                     # below botorch's clamp_min(1e-9)
    best = 0.1
    for maximize in (True, False):
        ei = R.expected_improvement(mean, var, best, maximize=maximize).numpy()
        sigma = np.sqrt(np.maximum(var.numpy(), 1e-9))
        u = (mean.numpy() - best) / sigma * (1.0 if maximize else -1.0)
        ref = sigma * (norm.pdf(u) + u * norm.cdf(u))
        assert np.abs(ei - ref).max() < 1e-13
