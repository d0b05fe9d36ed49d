"""not gpu, dev container only: the restated oracle against the UNMODIFIED reference imported from /root/reference
through oracle/reference_loader.py (skipped where the reference tree does not exist, e.g. on the GPU box)."""
import warnings

import pytest
import torch

from oracle import reference_loader as rl
from oracle import restated as R
from tests.conftest import rel_err

pytestmark = pytest.mark.skipif(not rl.available(), reason="/root/reference not mounted")


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def test_sphere_family_bit_level():
    ks = rl.load("BoManifolds.kernel_utils.kernels_sphere")
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(42)
    for dim in (2, 4, 6):
        x, y = _unit(torch.randn(31, dim + 1, generator=gen)), _unit(torch.randn(17, dim + 1, generator=gen))
        k = ks.SphereRiemannianMaternKernel(dim=dim, nu=1.5)
        k.lengthscale = 0.8
        assert rel_err(R.sphere_matern(x, y, dim, 1.5, 0.8), k.forward(x, y)) < 1e-14
        k = ks.SphereRiemannianGaussianKernel(dim=dim)
        k.lengthscale = 0.3
        assert rel_err(R.sphere_gaussian(x, y, dim, 0.3), k.forward(x, y)) < 1e-14


def test_float32_constant_regime():
    """constructed under a float32 default the reference's constants are float32-rounded."""
    ks = rl.load("BoManifolds.kernel_utils.kernels_sphere")
    torch.set_default_dtype(torch.float32)
    try:
        k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
        k.lengthscale = 0.5
    finally:
        torch.set_default_dtype(torch.float64)
    k = k.double()
    x = _unit(torch.randn(20, 3, generator=torch.Generator().manual_seed(1)))
    ref = k.forward(x, x)
    assert rel_err(R.sphere_matern(x, x, 2, 2.5, 0.5, const_dtype=torch.float32), ref) < 1e-7
    assert rel_err(R.sphere_matern(x, x, 2, 2.5, 0.5), ref) > 1e-9   # and it differs from the float64 regime


def test_so3_torus_hyperbolic_spd():
    warnings.simplefilter("ignore")
    gen = torch.Generator().manual_seed(7)
    kso = rl.load("BoManifolds.kernel_utils.kernels_so")
    with rl.quiet():
        k3 = kso.SORiemannianMaternKernel(dim=3, nu=1.5, summands_number_bound=7)
    k3.lengthscale = 0.9
    q, _ = torch.linalg.qr(torch.randn(15, 3, 3, generator=gen))
    q[:, :, 0] *= torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    q = q.reshape(15, 9)
    assert rel_err(R.so3_matern(q, q[:6], 1.5, 0.9, summands=7), k3.forward(q, q[:6])) < 1e-14
    kt = rl.load("BoManifolds.kernel_utils.kernels_torus")
    k_t = kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=2, nu=[1.5, 2.5])
    for kk, ls in zip(k_t.torus_kernel.kernels, (0.2, 0.6)):
        kk.lengthscale = ls
    a = torch.rand(9, 2, generator=gen) * 6.28
    assert rel_err(R.torus_matern(a, a[:4], 2, [1.5, 2.5], [0.2, 0.6]), k_t.forward(a, a[:4])) < 1e-14
    kh = rl.load("BoManifolds.kernel_utils.kernels_hyperbolic")
    z = torch.randn(10, 2, generator=gen)
    h = torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)
    km = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=1.5, nb_points_integral=30)
    km.lengthscale = 0.7
    assert rel_err(R.hyperbolic_integrated_matern(h, h[:5], 2, 1.5, 0.7, 30), km.forward(h, h[:5])) < 1e-14
    ksp = rl.load("BoManifolds.kernel_utils.kernels_spd")
    a = torch.randn(6, 2, 2, generator=gen)
    s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
    v = torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)
    kp = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=1.5, nb_points_integral_b=20, nb_points_integral_t=25)
    kp.lengthscale = 1.2
    assert rel_err(R.spd2_integrated_matern(v, v[:3], 1.5, 1.2, 20, 25), kp.forward(v, v[:3])) < 1e-13
