"""not gpu: the host-side coefficient / node-weight code (materngabo_b200/_weights.py + the kernel classes' series())
evaluated through a numpy model of the C-ABI op definition (tests/cpu_model.py) against the golden vectors.
This pins everything on the product path except the CUDA kernels themselves, without a GPU."""
import numpy as np
import pytest
import torch

from materngabo_b200 import _lib, _weights
from materngabo_b200 import kernels_so as kso
from materngabo_b200 import kernels_sphere as ks
from materngabo_b200 import kernels_torus as kt
from tests.conftest import load_golden, rel_err
from tests.cpu_model import series_gram, spec_to_desc

TOL = 1e-12


def _eval(kernel, x1, x2):
    spec, coef = kernel.series()
    return series_gram(spec_to_desc(spec), x1, x2, coef.detach().numpy())


@pytest.mark.parametrize("name", ["sphere_matern_S2", "sphere_matern_S3", "sphere_matern_S10", "sphere_matern_S2_self",
                                  "sphere_matern_S3_example"])
def test_sphere_matern(name):
    g = load_golden(name)
    k = ks.SphereRiemannianMaternKernel(dim=int(g["dim"]), nu=float(g["nu"]))
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(_eval(k, g["x1"], g["x2"]), g["K"]) < TOL


@pytest.mark.parametrize("name", ["sphere_gaussian_S2", "sphere_gaussian_S3", "sphere_gaussian_S10"])
def test_sphere_gaussian(name):
    g = load_golden(name)
    k = ks.SphereRiemannianGaussianKernel(dim=int(g["dim"]))
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(_eval(k, g["x1"], g["x2"]), g["K"]) < TOL


def test_integrated_and_circle():
    g = load_golden("sphere_intmatern_S3")
    k = ks.SphereRiemannianIntegratedMaternKernel(dim=3, nu=2.5)
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(_eval(k, g["x1"], g["x2"]), g["K"]) < TOL
    g = load_golden("circle_gaussian")
    k = ks.CircleRiemannianGaussianKernel()
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(_eval(k, g["x1"], g["x2"]), g["K"]) < TOL
    g = load_golden("circle_intmatern")
    k = ks.CircleRiemannianIntegratedMaternKernel(nu=2.5)
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(_eval(k, g["x1"], g["x2"]), g["K"]) < TOL


def test_so3():
    for name, cls, kw in (("so3_matern", kso.SORiemannianMaternKernel, {"nu": 2.5}),
                          ("so3_gaussian", kso.SORiemannianGaussianKernel, {}),
                          ("so3_matern_self", kso.SORiemannianMaternKernel, {"nu": 2.5})):
        g = load_golden(name)
        k = cls(dim=3, **kw)
        k.lengthscale = float(g["lengthscale"])
        assert rel_err(_eval(k, g["x1"], g["x2"]), g["K"]) < TOL


@pytest.mark.parametrize("dim", [3, 10])
def test_torus(dim):
    g = load_golden("torus_matern_T%d" % dim)
    k = kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=dim, nu=2.5)
    for c, kk in enumerate(k.torus_kernel.kernels):
        kk.lengthscale = float(g["lengthscales"][c])
    x1 = k._to_circles(torch.tensor(g["x1"])).numpy()
    x2 = k._to_circles(torch.tensor(g["x2"])).numpy()
    assert rel_err(_eval(k, x1, x2), g["K"]) < TOL
    spec, coef = k.series()
    assert spec.n_factors == dim and spec.factor_width == 2 and coef.shape[0] == dim
    assert coef.shape[1] <= _lib.MAX_TERMS


def test_hyper_gradient_flows_through_coefficients():
    """dL/d raw_lengthscale = (dL/dcoef from the device) . (dcoef/d raw_lengthscale from torch autograd)."""
    g = load_golden("sphere_matern_S3")
    k = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    k.lengthscale = float(g["lengthscale"])
    spec, coef = k.series()
    # dL/dcoef_k = sum_ij G_ij u_ij^k in the monomial basis
    u = np.clip(g["x1"] @ g["x2"].T, spec.lo, spec.hi)
    gcoef = np.array([(g["G"] * u ** kk).sum() for kk in range(coef.shape[1])])
    (grad,) = torch.autograd.grad((coef.reshape(-1) * torch.tensor(gcoef)).sum(), [k.raw_lengthscale])
    assert rel_err(grad.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-10


def test_basis_tables():
    m = _weights.chebyshev_to_monomial(6)
    assert m[5].tolist() == [0, 5, 0, -20, 0, 16]
    x = np.linspace(-1, 1, 9)
    gm, gc = _weights.gegenbauer_to_monomial(8, 1.5), _weights.gegenbauer_to_chebyshev(8, 3)
    for n in range(8):
        a = np.polynomial.polynomial.polyval(x, gm[n])
        b = np.polynomial.chebyshev.chebval(x, gc[n])
        assert np.allclose(a, b, rtol=0, atol=1e-12 * max(1.0, np.abs(a).max()))
    w = _weights.trapz_weights(torch.tensor([0.0, 1.0, 3.0, 4.0]))
    f = torch.tensor([1.0, 2.0, 0.5, 3.0])
    assert abs(float((w * f).sum()) - float(torch.trapz(f, torch.tensor([0.0, 1.0, 3.0, 4.0])))) < 1e-15


def test_large_term_count_switches_to_chebyshev():
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5, serie_nb_terms=20)
    k.lengthscale = 0.3
    spec, coef = k.series()
    assert spec.basis == _lib.BASIS_CHEBYSHEV and coef.shape == (1, 20)
    from oracle import restated as R
    x = torch.nn.functional.normalize(torch.randn(12, 3, generator=torch.Generator().manual_seed(0)), dim=-1)
    ref = R.sphere_matern(x, x, 2, 2.5, 0.3, terms=20)
    assert rel_err(series_gram(spec_to_desc(spec), x.numpy(), x.numpy(), coef.detach().numpy()), ref) < 1e-12


def test_float32_construction_regime_is_mirrored():
    """Constants are built in the default dtype at construction, as in the reference."""
    torch.set_default_dtype(torch.float32)
    try:
        k32 = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    finally:
        torch.set_default_dtype(torch.float64)
    assert k32.cst_nd[3].dtype == torch.float32
    k64 = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    assert k64.cst_nd[3].dtype == torch.float64
    assert abs(float(k32.cst_nd[3]) - float(k64.cst_nd[3])) > 0


def test_reference_error_conventions():
    with pytest.raises(ValueError):
        ks.SphereRiemannianMaternKernel(dim=1, nu=2.5)          # math.gamma(0)
    from materngabo_b200 import kernels_spd
    with pytest.raises(NotImplementedError):
        kernels_spd.SpdRiemannianGaussianKernel(3)
    from materngabo_b200._lib import MgbError
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    x = torch.nn.functional.normalize(torch.randn(4, 3), dim=-1)
    with pytest.raises(MgbError):                                # CPU tensors: loud failure, no fallback
        k.forward(x, x)


def test_circle_series_truncation_and_basis_choice():
    """Tail-sum truncation of the theta_3 series and the conditioning rule for the monomial re-expansion."""
    import math

    full = _weights.THETA3_TERMS + 1
    for ls, expect_basis in ((0.5, _lib.BASIS_MONOMIAL), (1.0, _lib.BASIS_MONOMIAL), (0.15, _lib.BASIS_CHEBYSHEV)):
        c = _weights.circle_integrated_matern_coef(torch.tensor(ls), torch.tensor(2.5), 50)
        assert 2 <= c.numel() < full
        # what was dropped is below 1e-16 of the sum: recompute the untruncated series
        saved = _weights._TRUNC
        try:
            _weights._TRUNC = 0.0
            c_all = _weights.circle_integrated_matern_coef(torch.tensor(ls), torch.tensor(2.5), 50)
        finally:
            _weights._TRUNC = saved
        assert float(c_all[c.numel():].abs().sum()) <= 1e-16 * float(c_all.abs().sum())
        coef, basis = _weights.chebyshev_or_monomial(c.reshape(1, -1))
        assert basis == expect_basis
        u = torch.cos(torch.linspace(0, math.pi, 301))
        cheb = sum(c[k] * torch.cos(k * torch.acos(u.clamp(-1, 1))) for k in range(c.numel()))
        if basis == _lib.BASIS_MONOMIAL:
            mono = torch.zeros_like(u)
            for k in range(coef.shape[-1] - 1, -1, -1):
                mono = mono * u + coef[0, k]
            assert float((mono - cheb).abs().max()) < 1e-14
        else:
            assert torch.equal(coef.reshape(-1), c)


def test_spd2_lattice_matches_the_actual_grids():
    """b_b / (2 t_a) = g_min tau^(sb b + sa (Pt - 1 - a)) for the reference's logspace grids (kernels_spd.py:262-265)."""
    from materngabo_b200 import kernels_spd as ksp

    for pb, pt, ls in ((64, 64, 1.0), (50, 50, 0.7), (40, 40, 2.0)):
        k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=pb, nb_points_integral_t=pt)
        k.lengthscale = ls
        tcoef, rscale, bscale, b, bw, lattice = k.nodes("cpu", with_lattice=True)
        assert lattice is not None
        sb, sa, g_min, log_tau = lattice
        bi = torch.arange(pb, dtype=torch.float64).unsqueeze(1)
        ai = torch.arange(pt, dtype=torch.float64).unsqueeze(0)
        model = g_min * torch.exp((sb * bi + sa * (pt - 1 - ai)) * log_tau)
        true = b.unsqueeze(1) * bscale.unsqueeze(0)
        assert float(((model - true) / true).abs().max()) < 1e-13
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=64, nb_points_integral_t=32)
    assert k.nodes("cpu", with_lattice=True)[-1] is None          # incommensurate enough: falls back to the node-by-node kernel
