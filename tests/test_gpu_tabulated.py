"""-m gpu: the opt-in tabulated radial profile (materngabo_b200/tabulated.py) against the per-entry quadrature kernels
and the oracle."""
import pytest
import torch

from tests.conftest import TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _lor(cnt, dim, gen, spread=1.0):
    z = spread * torch.randn(cnt, dim, generator=gen)
    return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_hyperbolic_matern_table_matches_quadrature_and_oracle(dim):
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200.tabulated import TabulatedRadialKernel
    from oracle import restated as R

    gen = torch.Generator().manual_seed(dim)
    x1, x2 = _lor(700, dim, gen, 1.5), _lor(130, dim, gen, 1.5)
    x1[:130] = x2                                        # exact self pairs (distance floor 1e-4)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=2.5, nb_points_integral=64).to(DEV)
    k.lengthscale = 1.0
    tab = TabulatedRadialKernel.for_inputs(k, x1.to(DEV), x2.to(DEV))
    assert tab.max_abs_error <= 1e-12 * tab.scale
    out = tab(x1.to(DEV), x2.to(DEV))
    with torch.no_grad():
        direct = k.forward(x1.to(DEV), x2.to(DEV))
    assert rel_err(out, direct) < 1e-11
    ref = R.hyperbolic_integrated_matern(x1[:90], x2[:70], dim, 2.5, 1.0, 64)
    assert rel_err(out[:90, :70], ref) < TOL


def test_heat_table():
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200.tabulated import TabulatedRadialKernel

    gen = torch.Generator().manual_seed(7)
    x1, x2 = _lor(257, 2, gen).to(DEV), _lor(65, 2, gen).to(DEV)
    k = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=2, nb_points_integral=300).to(DEV)
    k.lengthscale = 1.3
    tab = TabulatedRadialKernel.for_inputs(k, x1, x2)
    with torch.no_grad():
        assert rel_err(tab(x1, x2), k.forward(x1, x2)) < 1e-11


def test_euclidean_table():
    from materngabo_b200 import kernels_euclidean as ke
    from materngabo_b200.tabulated import TabulatedRadialKernel

    gen = torch.Generator().manual_seed(8)
    e1, e2 = torch.randn(300, 3, generator=gen).to(DEV), torch.randn(77, 3, generator=gen).to(DEV)
    k = ke.EuclideanIntegratedMaternKernel(3, nu=2.5).to(DEV)
    k.lengthscale = 0.7
    tab = TabulatedRadialKernel.for_inputs(k, e1, e2)
    with torch.no_grad():
        assert rel_err(tab(e1, e2), k.forward(e1, e2)) < 1e-11
    assert tab(e1[:0], e2).shape == (0, 77)


def test_posterior_with_tabulated_candidate_rows():
    """Acquisition scoring of a large candidate set on H^2: tabulated candidate rows against the per-entry quadrature
    posterior and the oracle."""
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200.posterior import ExactGPPosterior
    from oracle import restated as R

    gen = torch.Generator().manual_seed(31)
    xt, xc = _lor(60, 3, gen), _lor(5000, 3, gen)
    y = torch.tanh(xt[:, 1]) - 0.2 * xt[:, 3]
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=3, nu=2.5, nb_points_integral=32).to(DEV)
    k.lengthscale = 1.0
    direct = ExactGPPosterior(k, xt.to(DEV), y.to(DEV), outputscale=1.2, noise=1e-2, mean_const=0.1).fit()
    tab = ExactGPPosterior(k, xt.to(DEV), y.to(DEV), outputscale=1.2, noise=1e-2, mean_const=0.1, tabulate=True).fit()
    m0, v0 = direct.posterior(xc.to(DEV))
    m1, v1 = tab.posterior(xc.to(DEV))
    assert rel_err(m1, m0) < 1e-9 and float((v1 - v0).abs().max()) < 1e-9
    kfn = lambda a, b: R.hyperbolic_integrated_matern(a, b, 3, 2.5, 1.0, 32)  # noqa: E731
    mr, vr = R.gp_posterior(kfn, xt, y, xc[:300], outputscale=1.2, noise=1e-2, mean_const=0.1)
    assert rel_err(m1[:300], mr) < 1e-9 and float((v1[:300].cpu() - vr).abs().max()) < 1e-8
