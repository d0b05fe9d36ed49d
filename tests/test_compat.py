"""not gpu: the gpytorch stand-in (parameters, constraints, priors, active_dims, Scale/Product kernels) and the
parameter layout of the drop-in classes."""
import math

import torch

from materngabo_b200 import gpytorch_compat as gc
from materngabo_b200 import kernels_hyperbolic, kernels_so, kernels_spd, kernels_sphere, kernels_torus


def test_positive_constraint_roundtrip():
    c = gc.Positive()
    v = torch.tensor([0.1, 0.5, 3.0, 40.0])
    assert torch.allclose(c.transform(c.inverse_transform(v)), v, rtol=1e-12)
    g = gc.GreaterThan(0.6)
    assert torch.allclose(g.transform(g.inverse_transform(torch.tensor([0.7, 2.0]))), torch.tensor([0.7, 2.0]))
    assert float(c.transform(torch.zeros(()))) == math.log(2.0)


def test_parameter_names_shapes_and_setters():
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=3)
    names = dict(k.named_parameters())
    assert set(names) == {"raw_lengthscale", "raw_nu"}
    assert names["raw_lengthscale"].shape == (1, 1) and names["raw_nu"].shape == (1, 1)
    assert names["raw_nu"].requires_grad                        # nu=None -> learnable (kernels_sphere.py:73-75)
    k.lengthscale = 0.5
    k.nu = 1.5
    assert abs(float(k.lengthscale) - 0.5) < 1e-12 and abs(float(k.nu) - 1.5) < 1e-12
    k2 = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    assert not k2.raw_nu.requires_grad and abs(float(k2.nu) - 2.5) < 1e-12
    assert abs(float(k2.lengthscale) - math.log(2.0)) < 1e-12   # softplus(0), the reference's initial value
    k2.initialize(raw_lengthscale=torch.tensor(0.3))
    assert abs(float(k2.raw_lengthscale) - 0.3) < 1e-15
    sd = k2.state_dict()
    assert "raw_lengthscale" in sd and "raw_nu" in sd


def test_priors_are_registered():
    marker = object()
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=2, nu_prior=marker, lengthscale_prior=marker)
    got = {name for name, *_ in k.named_priors()}
    assert got == {"nu_prior", "lengthscale_prior"}


def test_torus_parameter_layout_matches_reference():
    k = kernels_torus.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=[1.5, 2.5, 0.5])
    names = set(dict(k.named_parameters()))
    assert "raw_lengthscale" in names                            # unused outer lengthscale, as in the reference
    for i in range(3):
        assert "torus_kernel.kernels.%d.raw_lengthscale" % i in names
        assert "torus_kernel.kernels.%d.raw_nu" % i in names
        assert k.torus_kernel.kernels[i].active_dims.tolist() == [2 * i, 2 * i + 1]
    assert [round(float(kk.nu), 6) for kk in k.torus_kernel.kernels] == [1.5, 2.5, 0.5]
    ang = torch.tensor([[0.0, math.pi / 2, math.pi]])
    circ = k._to_circles(ang)
    assert torch.allclose(circ, torch.tensor([[1.0, 0.0, 0.0, 1.0, -1.0, 0.0]]), atol=1e-15)


def test_constructor_signatures():
    kernels_sphere.SphereRiemannianIntegratedMaternKernel(3, nu=2.5, nu_prior=None, serie_nb_terms=8, nb_points_integral=40)
    kernels_sphere.CircleRiemannianIntegratedMaternKernel(nu=2.5, serie_nb_terms=20, nb_points_integral=50,
                                                          active_dims=torch.tensor([0, 1]))
    kernels_so.SORiemannianMaternKernel(3, nu=2.5, sigma_squared=1, summands_number_bound=10, lambda_coeff_bound=1e-10)
    kernels_so.SORiemannianGaussianKernel(3, 1, summands_number_bound=5)
    kernels_hyperbolic.HyperbolicRiemannianMillsonIntegratedMaternKernel(2, nu=2.5, nb_points_integral=64)
    kernels_hyperbolic.HyperbolicRiemannianMillsonGaussianKernel(2, nb_points_integral=100)
    kernels_spd.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=10, nb_points_integral_t=12)
    kernels_sphere.SphereApproximatedGaussianKernel(3)
    kernels_sphere.SphereRiemannianLaplaceKernel()
    assert float(kernels_sphere.SphereApproximatedGaussianKernel(3).beta) > 2.0  # beta_min = 2 for S^3


def test_scale_and_product_kernels_and_active_dims():
    class Dot(gc.Kernel):
        def forward(self, x1, x2, diag=False, **params):
            return (x1 * x2).sum(-1) if diag else x1 @ x2.transpose(-1, -2)

    x = torch.arange(12.0).reshape(3, 4)
    k = Dot(active_dims=[1, 2])
    assert torch.equal(k(x), x[:, 1:3] @ x[:, 1:3].t())
    s = gc.ScaleKernel(Dot())
    s.outputscale = 2.5
    assert torch.allclose(s(x, x), 2.5 * (x @ x.t()))
    assert torch.allclose(s(x, x, diag=True), 2.5 * (x * x).sum(-1))
    p = gc.ProductKernel(Dot(active_dims=[0]), Dot(active_dims=[3]))
    assert torch.equal(p(x, x), (x[:, :1] @ x[:, :1].t()) * (x[:, 3:] @ x[:, 3:].t()))
    d = Dot().covar_dist(x, x[:2], square_dist=True)
    assert d.shape == (3, 2) and float(d[0, 0]) == 0.0
