"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the dev container (where /root/reference is mounted):  ``python -m oracle.make_golden``

The reference has no golden vectors of its own, so these fixtures are outputs of
the reference's ``Kernel.forward`` (and torch autograd through it) on seeded inputs, float64 regime
(``torch.set_default_dtype(torch.float64)`` before construction).  Inputs are the deterministic
inputs of the reference's example scripts where they exist
(examples/kernels_manifolds_examples/kernels_sphere.py:66-96, kernels_so.py:59-79) and seeded random
points otherwise.  Each .npz stores inputs, hyper-parameters, K, and gradients of sum(K * G) with
respect to raw_lengthscale / x1 / x2 for a fixed seeded G.
"""
from __future__ import annotations

import math
import os
import warnings

import numpy as np
import torch

from oracle import reference_loader as rl

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def _rotations(n, gen):
    q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    q[:, :, 0] = q[:, :, 0] * torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    return q.reshape(n, 9)


def _lorentz(n, dim, gen):
    z = torch.randn(n, dim, generator=gen)
    return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)


def _spd2(n, gen):
    a = torch.randn(n, 2, 2, generator=gen)
    s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
    return torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)


def _sphere_geodesic():
    """100 points from (1,0,0,0) toward normalised (-0.9,0.1,0,0): example kernels_sphere.py:66-84."""
    base = torch.tensor([1.0, 0.0, 0.0, 0.0])
    end = _unit(torch.tensor([-0.9, 0.1, 0.0, 0.0]))
    ang = torch.acos((base * end).sum().clamp(-1, 1))
    u = _unit(end - (base * end).sum() * base)
    t = torch.linspace(0, 1, 100).unsqueeze(-1) * ang
    return torch.cos(t) * base + torch.sin(t) * u


def _run(kernel, x1, x2, gen, grads=True, params=("raw_lengthscale",)):
    x1 = x1.clone().requires_grad_(grads)
    x2 = x2.clone().requires_grad_(grads)
    k = kernel.forward(x1, x2)
    out = {"x1": x1.detach().numpy(), "x2": x2.detach().numpy(), "K": k.detach().numpy()}
    if grads:
        g = torch.randn(k.shape, generator=gen)
        plist = [getattr(kernel, p) for p in params]
        for p in plist:
            p.requires_grad_(True)
        res = torch.autograd.grad((k * g).sum(), [x1, x2] + plist, allow_unused=True)
        out["G"] = g.numpy()
        out["gx1"] = res[0].numpy()
        out["gx2"] = res[1].numpy()
        for name, val in zip(params, res[2:]):
            out["g_" + name] = val.numpy()
    return out


def main():
    warnings.simplefilter("ignore")
    torch.set_default_dtype(torch.float64)
    os.makedirs(OUT, exist_ok=True)
    gen = torch.Generator().manual_seed(20211108)
    ks = rl.load("BoManifolds.kernel_utils.kernels_sphere")
    kso = rl.load("BoManifolds.kernel_utils.kernels_so")
    kt = rl.load("BoManifolds.kernel_utils.kernels_torus")
    kh = rl.load("BoManifolds.kernel_utils.kernels_hyperbolic")
    ksp = rl.load("BoManifolds.kernel_utils.kernels_spd")
    ke = rl.load("BoManifolds.kernel_utils.kernels_euclidean")
    torch.set_default_dtype(torch.float64)
    cases = {}

    # ---- sphere ---------------------------------------------------------------------------
    for dim, m, n in ((2, 64, 48), (3, 40, 33), (10, 32, 40)):
        x1, x2 = _unit(torch.randn(m, dim + 1, generator=gen)), _unit(torch.randn(n, dim + 1, generator=gen))
        k = ks.SphereRiemannianMaternKernel(dim=dim, nu=2.5)
        k.lengthscale = 0.5
        cases["sphere_matern_S%d" % dim] = dict(_run(k, x1, x2, gen), dim=dim, nu=2.5, lengthscale=0.5)
        k = ks.SphereRiemannianGaussianKernel(dim=dim)
        k.lengthscale = 0.4
        cases["sphere_gaussian_S%d" % dim] = dict(_run(k, x1, x2, gen), dim=dim, lengthscale=0.4)
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=None)  # learnable nu: gradient to raw_nu
    k.lengthscale = 0.7
    k.nu = 1.5
    x1, x2 = _unit(torch.randn(20, 3, generator=gen)), _unit(torch.randn(16, 3, generator=gen))
    cases["sphere_matern_S2_nugrad"] = dict(_run(k, x1, x2, gen, params=("raw_lengthscale", "raw_nu")),
                                            dim=2, nu=1.5, lengthscale=0.7)
    geo = _sphere_geodesic()
    k = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    k.lengthscale = 0.5
    cases["sphere_matern_S3_example"] = dict(_run(k, geo, geo[:1], gen, grads=False), dim=3, nu=2.5, lengthscale=0.5)
    x1s = torch.cat([x1[:6], x1[:3]])  # includes exactly repeated points (clamp active)
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    k.lengthscale = 0.5
    cases["sphere_matern_S2_self"] = dict(_run(k, x1s, x1s, gen, grads=False), dim=2, nu=2.5, lengthscale=0.5)
    k = ks.SphereRiemannianIntegratedMaternKernel(dim=3, nu=2.5)
    k.lengthscale = 0.6
    x1, x2 = _unit(torch.randn(14, 4, generator=gen)), _unit(torch.randn(11, 4, generator=gen))
    cases["sphere_intmatern_S3"] = dict(_run(k, x1, x2, gen), dim=3, nu=2.5, lengthscale=0.6)

    # ---- circle / torus -------------------------------------------------------------------
    c1, c2 = _unit(torch.randn(18, 2, generator=gen)), _unit(torch.randn(13, 2, generator=gen))
    k = ks.CircleRiemannianGaussianKernel()
    k.lengthscale = 0.3
    cases["circle_gaussian"] = dict(_run(k, c1, c2, gen), lengthscale=0.3)
    k = ks.CircleRiemannianIntegratedMaternKernel(nu=2.5)
    k.lengthscale = 0.5
    cases["circle_intmatern"] = dict(_run(k, c1, c2, gen), nu=2.5, lengthscale=0.5)
    for dim, m, n in ((3, 16, 12), (10, 8, 6)):
        a1, a2 = torch.rand(m, dim, generator=gen) * 2 * math.pi, torch.rand(n, dim, generator=gen) * 2 * math.pi
        ls = [0.3 + 0.05 * c for c in range(dim)]
        k = kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=dim, nu=2.5)
        for c, kk in enumerate(k.torus_kernel.kernels):
            kk.lengthscale = ls[c]
        out = _run(k, a1, a2, gen, grads=False)
        cases["torus_matern_T%d" % dim] = dict(out, dim=dim, nu=2.5, lengthscales=np.array(ls))
        k = kt.TorusProductOfManifoldsRiemannianGaussianKernel(dim=dim)
        for c, kk in enumerate(k.torus_kernel.kernels):
            kk.lengthscale = ls[c]
        cases["torus_gaussian_T%d" % dim] = dict(_run(k, a1, a2, gen, grads=False), dim=dim, lengthscales=np.array(ls))

    # ---- SO(3) ----------------------------------------------------------------------------
    with rl.quiet():
        km = kso.SORiemannianMaternKernel(dim=3, nu=2.5)
        kg = kso.SORiemannianGaussianKernel(dim=3)
    km.lengthscale = 0.5
    kg.lengthscale = 0.5
    r1, r2 = _rotations(30, gen), _rotations(22, gen)
    cases["so3_matern"] = dict(_run(km, r1, r2, gen), nu=2.5, lengthscale=0.5)
    cases["so3_gaussian"] = dict(_run(kg, r1, r2, gen), lengthscale=0.5)
    cases["so3_matern_self"] = dict(_run(km, r1[:8], r1[:8], gen, grads=False), nu=2.5, lengthscale=0.5)

    # ---- hyperbolic -----------------------------------------------------------------------
    for dim in (1, 2, 3):
        h1, h2 = _lorentz(20, dim, gen), _lorentz(14, dim, gen)
        k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=2.5, nb_points_integral=64)
        k.lengthscale = 1.0
        cases["hyperbolic_matern_H%d" % dim] = dict(_run(k, h1, h2, gen), dim=dim, nu=2.5, lengthscale=1.0, P=64)
        k = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=dim)
        k.lengthscale = 1.3
        cases["hyperbolic_heat_H%d" % dim] = dict(_run(k, h1, h2, gen), dim=dim, lengthscale=1.3, P=1000)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=64)
    k.lengthscale = 1.0
    h1 = _lorentz(9, 2, gen)
    cases["hyperbolic_matern_H2_self"] = dict(_run(k, h1, h1, gen, grads=False), dim=2, nu=2.5,
                                              lengthscale=1.0, P=64)

    # ---- SPD(2) ---------------------------------------------------------------------------
    s1, s2 = _spd2(12, gen), _spd2(9, gen)
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=64, nb_points_integral_t=64)
    k.lengthscale = 1.0
    cases["spd2_matern"] = dict(_run(k, s1, s2, gen, grads=False), nu=2.5, lengthscale=1.0, Pb=64, Pt=64)
    cases["spd2_matern_self"] = dict(_run(k, s1[:6], s1[:6], gen, grads=False), nu=2.5, lengthscale=1.0, Pb=64, Pt=64)
    k = ksp.SpdRiemannianGaussianKernel(2, nb_points_integral=50)
    k.lengthscale = 1.4
    cases["spd2_heat"] = dict(_run(k, s1, s2, gen, grads=False), lengthscale=1.4, Pb=50)

    # ---- Euclidean factor -----------------------------------------------------------------
    e1, e2 = torch.randn(15, 3, generator=gen), torch.randn(10, 3, generator=gen)
    k = ke.EuclideanIntegratedMaternKernel(3, nu=2.5)
    k.lengthscale = 0.7
    cases["euclidean_intmatern"] = dict(_run(k, e1, e2, gen), dim=3, nu=2.5, lengthscale=0.7)

    # ---- SPD product kernels on pre-decomposed inputs (eigenvalues | rotation or sphere point) -----------
    def _eig(n_):
        return 0.5 + 2.0 * torch.rand(n_, 3, generator=gen)

    def _quat(n_):
        return _unit(torch.randn(n_, 4, generator=gen))

    p1 = torch.cat([_eig(14), _rotations(14, gen)], -1)
    p2 = torch.cat([_eig(10), _rotations(10, gen)], -1)
    with rl.quiet():
        k = ksp.SpdProductOfRSOManifoldsRiemannianMaternKernel(dim=3, nu=2.5, spd_input=False)
    k.Euclidean_kernel.lengthscale = 1.2
    k.so_kernel.lengthscale = 0.6
    cases["spd3_product_rso_matern"] = dict(_run(k, p1, p2, gen, grads=False), nu=2.5, ls_euclidean=1.2, ls_manifold=0.6)
    q1 = torch.cat([_eig(13), _quat(13)], -1)
    q2 = torch.cat([_eig(9), _quat(9)], -1)
    k = ksp.SpdProductOfRSManifoldsRiemannianMaternKernel(dim=3, nu=2.5, spd_input=False)
    k.Euclidean_kernel.lengthscale = 0.9
    k.sphere_kernel.lengthscale = 0.5
    cases["spd3_product_rs_matern"] = dict(_run(k, q1, q2, gen, grads=False), nu=2.5, ls_euclidean=0.9, ls_manifold=0.5)
    c1 = torch.cat([_eig(11)[:, :2], _unit(torch.randn(11, 2, generator=gen))], -1)
    c2 = torch.cat([_eig(8)[:, :2], _unit(torch.randn(8, 2, generator=gen))], -1)
    k = ksp.SpdProductOfRSManifoldsRiemannianMaternKernel(dim=2, nu=1.5, spd_input=False)
    k.Euclidean_kernel.lengthscale = 1.1
    k.sphere_kernel.lengthscale = 0.4
    cases["spd2_product_rs_matern"] = dict(_run(k, c1, c2, gen, grads=False), nu=1.5, ls_euclidean=1.1, ls_manifold=0.4)

    # ---- SPD(2) heat kernel: gradient with respect to the lengthscale only ---------------------------------
    k = ksp.SpdRiemannianGaussianKernel(2, nb_points_integral=50)
    k.lengthscale = 1.4
    k.raw_lengthscale.requires_grad_(True)
    kk = k.forward(s1, s2)
    gg = torch.randn(kk.shape, generator=gen)
    (gl,) = torch.autograd.grad((kk * gg).sum(), [k.raw_lengthscale])
    cases["spd2_heat_lsgrad"] = dict(x1=s1.numpy(), x2=s2.numpy(), K=kk.detach().numpy(), G=gg.numpy(),
                                     g_raw_lengthscale=gl.numpy(), lengthscale=1.4, Pb=50)

    # ---- input gradients and Hessian-vector products of the quadrature kernels (what the trust-region inner loop
    # differentiates: manifold_optimize.py:184-217, pymanopt_addons/tools/autodiff/_pytorch.py:89-116) ---------
    def _hvp_case(kernel, a, b):
        a = a.clone().requires_grad_(True)
        kk_ = kernel.forward(a, b)
        g_ = torch.randn(kk_.shape, generator=gen)
        v_ = torch.randn(a.shape, generator=gen)
        (ga,) = torch.autograd.grad((kk_ * g_).sum(), [a], create_graph=True)
        (hv,) = torch.autograd.grad((ga * v_).sum(), [a])
        return dict(x1=a.detach().numpy(), x2=b.numpy(), K=kk_.detach().numpy(), G=g_.numpy(), V=v_.numpy(),
                    gx1=ga.detach().numpy(), hvp_x1=hv.numpy())

    for dim in (1, 2, 3):
        h1, h2 = _lorentz(6, dim, gen), _lorentz(11, dim, gen)
        k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=2.5, nb_points_integral=32)
        k.lengthscale = 0.9
        cases["hyperbolic_matern_H%d_hvp" % dim] = dict(_hvp_case(k, h1, h2), dim=dim, nu=2.5, lengthscale=0.9, P=32)
        k = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=dim, nb_points_integral=200)
        k.lengthscale = 1.1
        cases["hyperbolic_heat_H%d_hvp" % dim] = dict(_hvp_case(k, h1, h2), dim=dim, lengthscale=1.1, P=200)
    s1, s2 = _spd2(7, gen), _spd2(9, gen)
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=40, nb_points_integral_t=36)
    k.lengthscale = 1.2
    cases["spd2_matern_xgrad"] = dict(_run(k, s1, s2, gen, params=()), nu=2.5, lengthscale=1.2, Pb=40, Pt=36)
    k = ksp.SpdRiemannianGaussianKernel(2, nb_points_integral=50)
    k.lengthscale = 1.4
    cases["spd2_heat_xgrad"] = dict(_run(k, s1, s2, gen, params=()), lengthscale=1.4, Pb=50)
    e1, e2 = torch.randn(5, 3, generator=gen), torch.randn(8, 3, generator=gen)
    k = ke.EuclideanIntegratedMaternKernel(3, nu=1.5)
    k.lengthscale = 0.8
    cases["euclidean_intmatern_hvp"] = dict(_hvp_case(k, e1, e2), dim=3, nu=1.5, lengthscale=0.8)

    for name, payload in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k_: np.asarray(v) for k_, v in payload.items()})
        print("%-32s K%s max|K|=%.3f" % (name, payload["K"].shape, np.abs(payload["K"]).max()))


if __name__ == "__main__":
    main()
