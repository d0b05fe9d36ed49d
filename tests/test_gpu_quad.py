"""-m gpu: quadrature kernels (hyperbolic H^1/H^2/H^3, SPD(2)) against golden vectors from the unmodified reference."""
import numpy as np
import pytest
import torch

from tests.conftest import TOL, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device=DEV)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_hyperbolic_matern(dim):
    from materngabo_b200 import kernels_hyperbolic as kh

    g = load_golden("hyperbolic_matern_H%d" % dim)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=2.5, nb_points_integral=int(g["P"]))
    k.lengthscale = float(g["lengthscale"])
    out = k.forward(_t(g["x1"]), _t(g["x2"]))
    assert rel_err(out, g["K"]) < TOL
    # hyper-parameter gradient (node weights are differentiable, the grid is detached as in the reference)
    k.raw_lengthscale.requires_grad_(True)
    out = k.forward(_t(g["x1"]), _t(g["x2"]))
    (gl,) = torch.autograd.grad((out * _t(g["G"])).sum(), [k.raw_lengthscale])
    assert rel_err(gl.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-9


def test_hyperbolic_matern_self_distance():
    from materngabo_b200 import kernels_hyperbolic as kh

    g = load_golden("hyperbolic_matern_H2_self")
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=64)
    k.lengthscale = 1.0
    assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_hyperbolic_heat(dim):
    from materngabo_b200 import kernels_hyperbolic as kh

    g = load_golden("hyperbolic_heat_H%d" % dim)
    k = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=dim, nb_points_integral=int(g["P"]))
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL
    # lengthscale gradient: t = kappa^2 / 2 is differentiable, the device returns dL/dt (incl. the normaliser)
    k.raw_lengthscale.requires_grad_(True)
    out = k.forward(_t(g["x1"]), _t(g["x2"]))
    (gl,) = torch.autograd.grad((out * _t(g["G"])).sum(), [k.raw_lengthscale])
    assert rel_err(gl.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-9


def test_hyperbolic_unsupported_dim():
    from materngabo_b200 import kernels_hyperbolic as kh

    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=4, nu=2.5)
    with pytest.raises(NotImplementedError):
        k.forward(torch.zeros(2, 5, device=DEV), torch.zeros(2, 5, device=DEV))


@pytest.mark.parametrize("name", ["spd2_matern", "spd2_matern_self"])
def test_spd2_matern(name):
    from materngabo_b200 import kernels_spd as ksp

    g = load_golden(name)
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=int(g["Pb"]), nb_points_integral_t=int(g["Pt"]))
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL


def test_spd2_heat():
    from materngabo_b200 import kernels_spd as ksp

    g = load_golden("spd2_heat")
    k = ksp.SpdRiemannianGaussianKernel(2, nb_points_integral=int(g["Pb"]))
    k.lengthscale = float(g["lengthscale"])
    assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL


def test_spd2_heat_lengthscale_gradient():
    from materngabo_b200 import kernels_spd as ksp

    g = load_golden("spd2_heat_lsgrad")
    k = ksp.SpdRiemannianGaussianKernel(2, nb_points_integral=int(g["Pb"]))
    k.lengthscale = float(g["lengthscale"])
    k.raw_lengthscale.requires_grad_(True)
    out = k.forward(_t(g["x1"]), _t(g["x2"]))
    assert rel_err(out, g["K"]) < TOL
    (gl,) = torch.autograd.grad((out * _t(g["G"])).sum(), [k.raw_lengthscale])
    assert rel_err(gl.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-9


def test_spd2_matern_hyper_gradient_matches_oracle_autograd():
    from materngabo_b200 import kernels_spd as ksp
    from oracle import restated as R

    g = load_golden("spd2_matern")
    G = torch.randn(g["K"].shape, generator=torch.Generator().manual_seed(1))
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=None, nb_points_integral_b=32, nb_points_integral_t=32)
    k.lengthscale = 1.1
    k.nu = 1.5
    out = k.forward(_t(g["x1"]), _t(g["x2"]))
    res = torch.autograd.grad((out * G.to(DEV)).sum(), [k.raw_lengthscale, k.raw_nu])
    raw_ls = k.raw_lengthscale.detach().clone().requires_grad_(True)
    raw_nu = k.raw_nu.detach().clone().requires_grad_(True)
    sp = torch.nn.functional.softplus
    ref = R.spd2_integrated_matern(torch.tensor(g["x1"]), torch.tensor(g["x2"]), sp(raw_nu), sp(raw_ls), 32, 32)
    rres = torch.autograd.grad((ref * G).sum(), [raw_ls, raw_nu])
    assert rel_err(out, ref) < TOL
    assert rel_err(res[0], rres[0]) < 1e-9 and rel_err(res[1], rres[1]) < 1e-9


def test_spd_dim3_raises_like_reference():
    from materngabo_b200 import kernels_spd as ksp

    with pytest.raises(NotImplementedError):
        ksp.SpdRiemannianIntegratedMaternKernel(3, nu=2.5)


def test_h2_config4_block_vs_oracle():
    """BASELINE config 4 geometry (H^2, 64 x 64 nodes) on a block the oracle finishes in seconds."""
    from materngabo_b200 import kernels_hyperbolic as kh
    from oracle import restated as R

    torch.manual_seed(0)
    z = torch.randn(2000, 2)
    x = torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=64)
    k.lengthscale = 1.0
    full = k.forward(x.to(DEV), x.to(DEV))                       # 2k x 2k, 4096 nodes per entry
    ref = R.hyperbolic_integrated_matern(x[:96], x[:80], 2, 2.5, 1.0, 64)
    assert rel_err(full[:96, :80], ref) < TOL
    assert rel_err(full, full.t()) < 1e-12
    assert torch.isfinite(full).all()


def test_spd2_config4_block_vs_oracle():
    from materngabo_b200 import kernels_spd as ksp
    from oracle import restated as R

    torch.manual_seed(0)
    a = torch.randn(512, 2, 2)
    s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
    v = torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=64, nb_points_integral_t=64)
    k.lengthscale = 1.0
    full = k.forward(v.to(DEV), v.to(DEV))
    ref = R.spd2_integrated_matern(v[:48], v[:40], 2.5, 1.0, 64, 64)
    assert rel_err(full[:48, :40], ref) < TOL
    assert rel_err(full, full.t()) < 1e-10


def test_euclidean_integrated_matern():
    from materngabo_b200 import kernels_euclidean as ke

    g = load_golden("euclidean_intmatern")
    k = ke.EuclideanIntegratedMaternKernel(3, nu=2.5)
    k.lengthscale = float(g["lengthscale"])
    k.raw_lengthscale.requires_grad_(True)
    out = k.forward(_t(g["x1"]), _t(g["x2"]))
    assert rel_err(out, g["K"]) < TOL
    (gl,) = torch.autograd.grad((out * _t(g["G"])).sum(), [k.raw_lengthscale])
    assert rel_err(gl.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-9


@pytest.mark.parametrize("name,cls,dim,nu", [("spd3_product_rso_matern", "SpdProductOfRSOManifoldsRiemannianMaternKernel", 3, 2.5),
                                             ("spd3_product_rs_matern", "SpdProductOfRSManifoldsRiemannianMaternKernel", 3, 2.5),
                                             ("spd2_product_rs_matern", "SpdProductOfRSManifoldsRiemannianMaternKernel", 2, 1.5)])
def test_spd_product_kernels(name, cls, dim, nu):
    """SPD(3) as R^3 x SO(3) / R^3 x S^3 and SPD(2) as R^2 x S^1 on pre-decomposed inputs."""
    from materngabo_b200 import kernels_spd as ksp

    g = load_golden(name)
    k = getattr(ksp, cls)(dim=dim, nu=nu, spd_input=False)
    k.Euclidean_kernel.lengthscale = float(g["ls_euclidean"])
    manifold_kernel = k.so_kernel if "RSO" in cls else k.sphere_kernel
    manifold_kernel.lengthscale = float(g["ls_manifold"])
    assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL


def _random_spd_mandel(cnt, dim, gen, spread=1.0):
    a = torch.randn(cnt, dim, dim, generator=gen)
    sm = a @ a.transpose(-1, -2) * spread + 0.3 * torch.eye(dim)
    if dim == 2:
        return torch.stack([sm[:, 0, 0], sm[:, 1, 1], sm[:, 0, 1] * 2 ** 0.5], -1)
    return torch.stack([sm[:, 0, 0], sm[:, 1, 1], sm[:, 2, 2], sm[:, 0, 1] * 2 ** 0.5, sm[:, 1, 2] * 2 ** 0.5,
                        sm[:, 0, 2] * 2 ** 0.5], -1)


@pytest.mark.parametrize("dim", [2, 3])
def test_spd_eigen_front_end_conventions_and_reconstruction(dim):
    """csrc/spd_eigh.cu: Q diag(ev) Q^T reconstructs S to 1e-13, Q is orthogonal, eigenvalues
    ascending, every eigenvector's largest component positive before the reference's Q = det(V) V, and the result
    agrees with the oracle's torch.linalg.eigh under the same convention - incl. nearly isotropic matrices, a diagonal
    matrix and widely spread eigenvalues."""
    from materngabo_b200 import _ops
    from oracle import restated as R

    gen = torch.Generator().manual_seed(40 + dim)
    x = _random_spd_mandel(500, dim, gen)
    x[0] = 0.0
    x[0, :dim] = torch.arange(1, dim + 1, dtype=torch.float64)            # already diagonal
    x[1] = x[0] + 1e-9 * torch.randn(x.shape[1], generator=gen)           # tiny off-diagonal part
    x[2:20] = _random_spd_mandel(18, dim, gen, spread=1e6)                # condition ~1e6
    out = _ops.spd_decompose(x.to(DEV), dim, 0).cpu()
    ev, q = out[:, :dim], out[:, dim:].reshape(-1, dim, dim)
    sm = R.mandel_to_sym(x)
    rec = q @ torch.diag_embed(ev) @ q.transpose(-1, -2)
    scale = sm.abs().amax(dim=(-1, -2), keepdim=True)
    assert float(((rec - sm).abs() / scale).max()) < 1e-13
    assert float((q.transpose(-1, -2) @ q - torch.eye(dim)).abs().max()) < 1e-14
    assert bool((ev[:, 1:] >= ev[:, :-1]).all())
    assert float((ev - torch.linalg.eigvalsh(sm)).abs().max() / float(scale.max())) < 1e-14
    if dim == 3:
        assert float((torch.linalg.det(q) - 1.0).abs().max()) < 1e-13     # det(V) V lands in SO(3)
    ref = R.spd_decompose(x, dim, 0)
    generic = slice(20, None)                                              # well separated eigenvalues
    assert float((out[generic] - ref[generic]).abs().max()) < 1e-10
    sph = _ops.spd_decompose(x.to(DEV), dim, 1).cpu()
    assert sph.shape == (500, dim + (2 if dim == 2 else 4))
    assert float((sph[generic] - R.spd_decompose(x, dim, 1)[generic]).abs().max()) < 1e-10
    assert float((sph[:, dim:].norm(dim=-1) - 1.0).abs().max()) < 1e-13
    assert _ops.spd_decompose(x[:0].to(DEV), dim, 0).shape == (0, dim + dim * dim)


@pytest.mark.parametrize("cls,dim,manifold", [("SpdProductOfRSOManifoldsRiemannianMaternKernel", 3, "so"),
                                              ("SpdProductOfRSManifoldsRiemannianMaternKernel", 3, "s"),
                                              ("SpdProductOfRSManifoldsRiemannianMaternKernel", 2, "s")])
def test_spd_product_kernels_with_spd_input(cls, dim, manifold):
    """spd_input=True end to end: Mandel vectors -> device eigen-decomposition -> product kernel, against the oracle's
    product kernel fed the SAME decomposition (the reference's own, torch.symeig, no longer exists)."""
    from materngabo_b200 import _ops, kernels_spd as ksp
    from oracle import restated as R

    gen = torch.Generator().manual_seed(50 + dim)
    x1, x2 = _random_spd_mandel(23, dim, gen), _random_spd_mandel(17, dim, gen)
    k = getattr(ksp, cls)(dim=dim, nu=2.5, spd_input=True)
    k.Euclidean_kernel.lengthscale = 1.1
    (k.so_kernel if manifold == "so" else k.sphere_kernel).lengthscale = 0.6
    out = k.forward(x1.to(DEV), x2.to(DEV))
    target = 0 if manifold == "so" else 1
    d1, d2 = _ops.spd_decompose(x1.to(DEV), dim, target).cpu(), _ops.spd_decompose(x2.to(DEV), dim, target).cpu()
    ref = R.spd_product_matern(d1, d2, dim, 2.5, 1.1, 0.6, manifold)
    assert out.shape == (23, 17) and rel_err(out, ref) < TOL
    pre = getattr(ksp, cls)(dim=dim, nu=2.5, spd_input=False)
    pre.Euclidean_kernel.lengthscale = 1.1
    (pre.so_kernel if manifold == "so" else pre.sphere_kernel).lengthscale = 0.6
    assert rel_err(pre.forward(d1.to(DEV), d2.to(DEV)), out) < 1e-14
    with pytest.raises(NotImplementedError):
        k.forward(x1.to(DEV).requires_grad_(True), x2.to(DEV))


# ---- input gradients / Hessian-vector products (trust-region inner loop) -------------------
def _hvp_check(kernel, g, tol_g=1e-9, tol_h=1e-8):
    a = _t(g["x1"]).requires_grad_(True)
    k = kernel.forward(a, _t(g["x2"]))
    assert rel_err(k, g["K"]) < TOL
    (ga,) = torch.autograd.grad((k * _t(g["G"])).sum(), [a], create_graph=True)
    assert rel_err(ga, g["gx1"]) < tol_g
    (hv,) = torch.autograd.grad((ga * _t(g["V"])).sum(), [a])
    assert rel_err(hv, g["hvp_x1"]) < tol_h


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_hyperbolic_matern_input_gradient_and_hvp(dim):
    from materngabo_b200 import kernels_hyperbolic as kh

    g = load_golden("hyperbolic_matern_H%d_hvp" % dim)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=2.5, nb_points_integral=int(g["P"]))
    k.lengthscale = float(g["lengthscale"])
    _hvp_check(k, g)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_hyperbolic_heat_input_gradient_and_hvp(dim):
    from materngabo_b200 import kernels_hyperbolic as kh

    g = load_golden("hyperbolic_heat_H%d_hvp" % dim)
    k = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=dim, nb_points_integral=int(g["P"]))
    k.lengthscale = float(g["lengthscale"])
    for p in k.parameters():
        p.requires_grad_(False)
    _hvp_check(k, g)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_hyperbolic_gradients_to_both_inputs_and_lengthscale(dim):
    """The golden cases of the forward tests also carry dL/dx1, dL/dx2 from the reference's autograd; asking for the
    input and the lengthscale gradients in one backward exercises the coefficient path of the profile op."""
    from materngabo_b200 import kernels_hyperbolic as kh

    g = load_golden("hyperbolic_matern_H%d" % dim)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=2.5, nb_points_integral=int(g["P"]))
    k.lengthscale = float(g["lengthscale"])
    a, b = _t(g["x1"]).requires_grad_(True), _t(g["x2"]).requires_grad_(True)
    out = k.forward(a, b)
    assert rel_err(out, g["K"]) < TOL
    ga, gb, gl = torch.autograd.grad((out * _t(g["G"])).sum(), [a, b, k.raw_lengthscale])
    assert rel_err(ga, g["gx1"]) < 1e-9 and rel_err(gb, g["gx2"]) < 1e-9
    assert rel_err(gl.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-9
    g = load_golden("hyperbolic_heat_H%d" % dim)
    k = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=dim, nb_points_integral=int(g["P"]))
    k.lengthscale = float(g["lengthscale"])
    a, b = _t(g["x1"]).requires_grad_(True), _t(g["x2"]).requires_grad_(True)
    out = k.forward(a, b)
    ga, gb, gl = torch.autograd.grad((out * _t(g["G"])).sum(), [a, b, k.raw_lengthscale])
    assert rel_err(ga, g["gx1"]) < 1e-9 and rel_err(gb, g["gx2"]) < 1e-9
    assert rel_err(gl.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-9


def test_euclidean_input_gradient_and_hvp():
    from materngabo_b200 import kernels_euclidean as ke

    g = load_golden("euclidean_intmatern_hvp")
    k = ke.EuclideanIntegratedMaternKernel(3, nu=float(g["nu"]))
    k.lengthscale = float(g["lengthscale"])
    _hvp_check(k, g)


def test_spd2_input_gradients():
    from materngabo_b200 import kernels_spd as ksp

    g = load_golden("spd2_matern_xgrad")
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=int(g["Pb"]), nb_points_integral_t=int(g["Pt"]))
    k.lengthscale = float(g["lengthscale"])
    a, b = _t(g["x1"]).requires_grad_(True), _t(g["x2"]).requires_grad_(True)
    out = k.forward(a, b)
    assert rel_err(out, g["K"]) < TOL
    ga, gb, gl = torch.autograd.grad((out * _t(g["G"])).sum(), [a, b, k.raw_lengthscale])
    assert rel_err(ga, g["gx1"]) < 1e-9 and rel_err(gb, g["gx2"]) < 1e-9
    assert torch.isfinite(gl).all()
    g = load_golden("spd2_heat_xgrad")
    k = ksp.SpdRiemannianGaussianKernel(2, nb_points_integral=int(g["Pb"]))
    k.lengthscale = float(g["lengthscale"])
    k.raw_lengthscale.requires_grad_(False)
    a, b = _t(g["x1"]).requires_grad_(True), _t(g["x2"]).requires_grad_(True)
    out = k.forward(a, b)
    assert rel_err(out, g["K"]) < TOL
    ga, gb = torch.autograd.grad((out * _t(g["G"])).sum(), [a, b])
    assert rel_err(ga, g["gx1"]) < 1e-9 and rel_err(gb, g["gx2"]) < 1e-9


def test_hyperbolic_acquisition_gradient_and_hvp():
    """EI value, gradient and Hessian-vector product at single candidates on H^3 (bo_hyperbolic.py uses exact
    Hessians: approx_hessian=False) against torch autograd through the oracle."""
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200.posterior import ExactGPPosterior, ExpectedImprovement
    from oracle import restated as R

    gen = torch.Generator().manual_seed(5)

    def lor(cnt):
        z = torch.randn(cnt, 3, generator=gen)
        return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)

    xt, xc = lor(40), lor(4)
    y = torch.tanh(xt[:, 1]) + 0.1 * xt[:, 2]
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=3, nu=2.5, nb_points_integral=32)
    k.lengthscale = 1.0
    gp = ExactGPPosterior(k, xt.to(DEV), y.to(DEV), outputscale=1.3, noise=1e-2, mean_const=0.05).fit()
    best = float(y.min())
    acq = ExpectedImprovement(gp, best_f=best, maximize=False)
    vec = torch.randn(4, 1, 4, generator=gen)
    kfn = lambda a, b: R.hyperbolic_integrated_matern(a, b, 3, 2.5, 1.0, 32)  # noqa: E731
    L, alpha = R.gp_fit(kfn, xt, y, 1.3, 1e-2, 0.05)

    def cost_ref(x):
        xs = x.squeeze(-2)
        ks = 1.3 * kfn(xs, xt)
        mean = 0.05 + ks @ alpha
        v = torch.linalg.solve_triangular(L, ks.t(), upper=False)
        var = 1.3 * kfn(xs[:1], xs[:1]).reshape(()).detach() - (v * v).sum(0)
        return -R.expected_improvement(mean, var, best, maximize=False).sum()

    outs = []
    for fn, dev in ((lambda x: -acq(x).sum(), DEV), (cost_ref, "cpu")):
        x = xc.unsqueeze(-2).to(dev).clone().requires_grad_(True)
        c = fn(x)
        (g,) = torch.autograd.grad(c, [x], create_graph=True)
        (h,) = torch.autograd.grad((g * vec.to(dev)).sum(), [x])
        outs.append((float(c.detach()), g.detach().cpu(), h.cpu()))
    assert abs(outs[0][0] - outs[1][0]) < 1e-9
    assert rel_err(outs[0][1], outs[1][1]) < 1e-8
    assert rel_err(outs[0][2], outs[1][2]) < 1e-7


def test_spd2_lattice_kernel_matches_direct_kernel_and_golden():
    """The lattice form (one exponential per lattice point of b_b / 2 t_a) against the node-by-node kernel and the
    golden vectors of the reference, default 50 x 50 grid and the 64 x 64 grid of BASELINE config 4."""
    from materngabo_b200 import _ops
    from materngabo_b200 import kernels_spd as ksp

    saved = _ops.SPD2_LATTICE_MIN_ENTRIES
    try:
        g = load_golden("spd2_matern")
        k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=64, nb_points_integral_t=64)
        k.lengthscale = float(g["lengthscale"])
        _ops.SPD2_LATTICE_MIN_ENTRIES = 1
        lat = k.forward(_t(g["x1"]), _t(g["x2"]))
        _ops.SPD2_LATTICE_MIN_ENTRIES = 1 << 60
        direct = k.forward(_t(g["x1"]), _t(g["x2"]))
        assert rel_err(lat, g["K"]) < TOL and rel_err(direct, g["K"]) < TOL
        assert rel_err(lat, direct) < 1e-12
        torch.manual_seed(3)
        a = torch.randn(300, 2, 2)
        s = a @ a.transpose(-1, -2) + 0.05 * torch.eye(2)        # includes ill-conditioned matrices (large Delta)
        v = torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1).to(DEV)
        for nb, nt, ls in ((50, 50, 0.6), (64, 64, 2.0), (40, 40, 1.0)):
            k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=1.5, nb_points_integral_b=nb, nb_points_integral_t=nt)
            k.lengthscale = ls
            assert k.nodes(DEV, with_lattice=True)[-1] is not None
            _ops.SPD2_LATTICE_MIN_ENTRIES = 1
            lat = k.forward(v, v[:77])
            _ops.SPD2_LATTICE_MIN_ENTRIES = 1 << 60
            direct = k.forward(v, v[:77])
            assert rel_err(lat, direct) < 1e-12
    finally:
        _ops.SPD2_LATTICE_MIN_ENTRIES = saved


@pytest.mark.parametrize("m,n", [(0, 4), (3, 0), (1, 1), (33, 7), (2, 65)])
def test_quadrature_kernels_ragged_and_empty_sizes(m, n):
    """Edge shapes through every quadrature Gram kernel, against the oracle (empty blocks return empty matrices)."""
    from materngabo_b200 import kernels_euclidean as ke
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200 import kernels_spd as ksp
    from oracle import restated as R

    gen = torch.Generator().manual_seed(100 + m + n)

    def lor(cnt, dim):
        z = torch.randn(cnt, dim, generator=gen)
        return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)

    def spd(cnt):
        a = torch.randn(cnt, 2, 2, generator=gen)
        s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
        return torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)

    for dim in (1, 2, 3):
        a, b = lor(m, dim), lor(n, dim)
        k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=1.5, nb_points_integral=24)
        k.lengthscale = 0.8
        out = k.forward(a.to(DEV), b.to(DEV))
        assert out.shape == (m, n)
        if m and n:
            assert rel_err(out, R.hyperbolic_integrated_matern(a, b, dim, 1.5, 0.8, 24)) < TOL
    a, b = spd(m), spd(n)
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=1.5, nb_points_integral_b=20, nb_points_integral_t=20)
    k.lengthscale = 1.1
    out = k.forward(a.to(DEV), b.to(DEV))
    assert out.shape == (m, n)
    if m and n:
        assert rel_err(out, R.spd2_integrated_matern(a, b, 1.5, 1.1, 20, 20)) < TOL
    a, b = torch.randn(m, 3, generator=gen), torch.randn(n, 3, generator=gen)
    k = ke.EuclideanIntegratedMaternKernel(3, nu=2.5)
    k.lengthscale = 0.9
    out = k.forward(a.to(DEV), b.to(DEV))
    assert out.shape == (m, n)
    if m and n:
        assert rel_err(out, R.euclidean_integrated_matern(a, b, 2.5, 0.9)) < TOL


def test_quadrature_kernels_botorch_batch_layout():
    """b x 1 x D candidates against the expanded training block and diag=True."""
    from materngabo_b200 import kernels_hyperbolic as kh
    from oracle import restated as R

    gen = torch.Generator().manual_seed(9)
    z = torch.randn(30, 2, generator=gen)
    x = torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)
    xc, xt = x[:7], x[7:]
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=32)
    k.lengthscale = 1.0
    out = k.forward(xc.unsqueeze(-2).to(DEV), xt.to(DEV).expand(7, *xt.shape))
    assert out.shape == (7, 1, 23)
    assert rel_err(out.squeeze(-2), R.hyperbolic_integrated_matern(xc, xt, 2, 2.5, 1.0, 32)) < TOL
    d = k.forward(x.to(DEV), x.to(DEV), diag=True)
    full = k.forward(x.to(DEV), x.to(DEV))
    assert d.shape == (30,) and rel_err(d, torch.diagonal(full)) < 1e-11     # per-row profile kernel vs fused Gram kernel
    x2 = x.roll(3, 0)
    d2 = k.forward(x.to(DEV), x2.to(DEV), diag=True)                          # genuinely different rows
    assert rel_err(d2, torch.diagonal(k.forward(x.to(DEV), x2.to(DEV)))) < 1e-11
    assert rel_err(d2, torch.diagonal(R.hyperbolic_integrated_matern(x, x2, 2, 2.5, 1.0, 32))) < TOL


def test_spd2_and_euclidean_diag_rows():
    from materngabo_b200 import kernels_euclidean as ke, kernels_spd as kspd

    gen = torch.Generator().manual_seed(10)
    a = torch.randn(40, 2, 2, generator=gen)
    sm = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
    x = torch.stack([sm[:, 0, 0], sm[:, 1, 1], sm[:, 0, 1] * 2 ** 0.5], -1).to(DEV)
    for k in (kspd.SpdRiemannianIntegratedMaternKernel(2, nu=2.5), kspd.SpdRiemannianGaussianKernel(2)):
        k.lengthscale = 1.0
        d = k.forward(x, x.roll(1, 0), diag=True)
        assert d.shape == (40,) and rel_err(d, torch.diagonal(k.forward(x, x.roll(1, 0)))) < 1e-10
    e = ke.EuclideanIntegratedMaternKernel(dim=3, nu=2.5)
    e.lengthscale = 0.8
    y = torch.randn(50, 3, generator=gen).to(DEV)
    d = e.forward(y, y.roll(2, 0), diag=True)
    assert rel_err(d, torch.diagonal(e.forward(y, y.roll(2, 0)))) < 1e-11


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_fused_radial_acquisition_hyperbolic(dim):
    """EI value, gradient and Hessian-vector product on H^dim from the radial acquisition kernel (profile derivatives
    from the device quadrature + pair geometry + closed-form EI chain) against torch autograd through the
    differentiable posterior path, incl. a candidate that IS a training point (distance floor, max(0, .) active)."""
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200.posterior import ExactGPPosterior, ExpectedImprovement

    gen = torch.Generator().manual_seed(40 + dim)

    def lor(cnt):
        z = torch.randn(cnt, dim, generator=gen)
        return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)

    xt = lor(35)
    xc = torch.cat([lor(5), xt[:1]])
    y = torch.tanh(xt[:, 1]) + 0.1 * xt[:, 0]
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=dim, nu=2.5, nb_points_integral=32).to(DEV)
    k.lengthscale = 1.0
    for p in k.parameters():
        p.requires_grad_(False)
    # the reference's H^2 quadrature kernel is not positive definite at these node counts (min eigenvalue about -0.5):
    # a large noise term makes K + noise I factorisable so that the derivative algebra can be compared
    noise = 1.5 if dim == 2 else 1e-2
    gp = ExactGPPosterior(k, xt.to(DEV), y.to(DEV), outputscale=1.3, noise=noise, mean_const=0.05).fit()
    acq = ExpectedImprovement(gp, best_f=float(y.min()), maximize=False)
    v = torch.randn(6, dim + 1, generator=gen).to(DEV)
    x = xc.to(DEV)
    ei, grad, hvp = acq.value_grad_hvp(x, v)
    xa = x.clone().requires_grad_(True)
    val = acq(xa.unsqueeze(-2))
    (ga,) = torch.autograd.grad(val.sum(), [xa], create_graph=True)
    (ha,) = torch.autograd.grad((ga * v).sum(), [xa])
    assert rel_err(ei, val) < 1e-11 and rel_err(grad, ga) < 1e-9 and rel_err(hvp, ha) < 1e-8
    ei2, grad2, none = acq.value_grad_hvp(x.unsqueeze(-2))
    assert none is None and rel_err(ei2, ei) < 1e-14 and rel_err(grad2.squeeze(-2), grad) < 1e-14


def test_fused_radial_acquisition_euclidean():
    from materngabo_b200 import kernels_euclidean as ke
    from materngabo_b200.posterior import ExactGPPosterior, ExpectedImprovement

    gen = torch.Generator().manual_seed(77)
    xt, xc = torch.randn(30, 3, generator=gen), torch.randn(5, 3, generator=gen)
    y = torch.sin(xt[:, 0]) + xt[:, 1] * 0.3
    k = ke.EuclideanIntegratedMaternKernel(3, nu=2.5).to(DEV)
    k.lengthscale = 0.9
    for p in k.parameters():
        p.requires_grad_(False)
    gp = ExactGPPosterior(k, xt.to(DEV), y.to(DEV), outputscale=1.1, noise=1e-2).fit()
    acq = ExpectedImprovement(gp, best_f=float(y.max()), maximize=True)
    v = torch.randn(5, 3, generator=gen).to(DEV)
    x = xc.to(DEV)
    ei, grad, hvp = acq.value_grad_hvp(x, v)
    xa = x.clone().requires_grad_(True)
    val = acq(xa.unsqueeze(-2))
    (ga,) = torch.autograd.grad(val.sum(), [xa], create_graph=True)
    (ha,) = torch.autograd.grad((ga * v).sum(), [xa])
    assert rel_err(ei, val) < 1e-11 and rel_err(grad, ga) < 1e-9 and rel_err(hvp, ha) < 1e-8


@pytest.mark.parametrize("which", ["matern", "heat"])
def test_fused_spd2_acquisition_value_and_gradient(which):
    """SPD(2): EI value + gradient from the fused kernel (closed-form derivative of the 2 x 2 log singular values,
    Cholesky chain for the heat kernel) against torch autograd through the differentiable posterior path."""
    from materngabo_b200 import kernels_spd as ksp
    from materngabo_b200.posterior import ExactGPPosterior, ExpectedImprovement

    gen = torch.Generator().manual_seed(55)

    def spd(cnt):
        a = torch.randn(cnt, 2, 2, generator=gen)
        s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
        return torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)

    xt, xc = spd(30).to(DEV), spd(6).to(DEV)
    y = torch.log(xt[:, 0]) - 0.3 * xt[:, 2]
    if which == "matern":
        k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=32, nb_points_integral_t=32).to(DEV)
    else:
        k = ksp.SpdRiemannianGaussianKernel(2, nb_points_integral=50).to(DEV)
    k.lengthscale = 1.2
    for p in k.parameters():
        p.requires_grad_(False)
    gp = ExactGPPosterior(k, xt, y, outputscale=1.1, noise=5e-2, mean_const=0.1).fit()
    acq = ExpectedImprovement(gp, best_f=float(y.min()), maximize=False)
    ei, grad, none = acq.value_grad_hvp(xc)
    assert none is None
    xa = xc.clone().requires_grad_(True)
    val = acq(xa.unsqueeze(-2))
    (ga,) = torch.autograd.grad(val.sum(), [xa])
    assert rel_err(ei, val) < 1e-11 and rel_err(grad, ga) < 1e-9
    with pytest.raises(NotImplementedError):
        acq.value_grad_hvp(xc, torch.zeros_like(xc))


@pytest.mark.parametrize("m,n,P", [(67, 71, 64), (130, 129, 50), (96, 80, 20), (257, 31, 33), (1024, 517, 64)])
def test_h2_four_entry_forward_kernel_matches_generic_kernel_and_oracle(m, n, P):
    """The forward Gram of H^2 blocks with >= 4096 entries runs the four-entries-per-warp kernel (h2_gram_fwd4_kernel);
    MGB_H2_PATH=generic keeps the two-entry kernel: same nodes, same tables - the values agree to rounding, for ragged
    sizes (partial last rounds / partial groups of four) and for one and two t nodes per lane."""
    import os
    from materngabo_b200 import kernels_hyperbolic as kh
    from oracle import restated as R

    torch.manual_seed(m + n)
    z1, z2 = 0.8 * torch.randn(m, 2), 0.8 * torch.randn(n, 2)
    x1 = torch.cat([torch.sqrt(1 + (z1 ** 2).sum(-1, keepdim=True)), z1], -1)
    x2 = torch.cat([torch.sqrt(1 + (z2 ** 2).sum(-1, keepdim=True)), z2], -1)
    x2[: min(m, n) // 3] = x1[: min(m, n) // 3]                  # repeated points: d = 2 asinh(5e-5)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=P)
    k.lengthscale = 0.9
    with torch.no_grad():
        fast = k.forward(x1.to(DEV), x2.to(DEV))
        os.environ["MGB_H2_PATH"] = "generic"
        try:
            slow = k.forward(x1.to(DEV), x2.to(DEV))
        finally:
            del os.environ["MGB_H2_PATH"]
    assert torch.isfinite(fast).all()
    assert rel_err(fast, slow) < 1e-13
    ref = R.hyperbolic_integrated_matern(x1[:40], x2[:36], 2, 2.5, 0.9, P)
    assert rel_err(fast[:40, :36], ref) < TOL


@pytest.mark.parametrize("m,n,nb,nt", [(67, 71, 64, 64), (33, 1, 50, 50), (130, 129, 40, 40), (1, 3, 64, 64), (200, 77, 64, 64)])
def test_spd2_two_entry_lattice_kernel_matches_one_entry_kernel(m, n, nb, nt):
    """spd2_lattice2p_kernel (two entries per iteration of a warp pair, the default) against spd2_lattice2_kernel (one
    warp per entry pair, MGB_SPD2_PATH=lattice2) and spd2_lattice_kernel (MGB_SPD2_PATH=lattice1): same lattice, same
    nodes; odd entry counts leave the second slot of the last pair empty."""
    import os
    from materngabo_b200 import _ops
    from materngabo_b200 import kernels_spd as ksp

    torch.manual_seed(m * 7 + n)
    a = torch.randn(max(m, n), 2, 2)
    s = a @ a.transpose(-1, -2) + 0.05 * torch.eye(2)
    v = torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1).to(DEV)
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=nb, nb_points_integral_t=nt)
    k.lengthscale = 1.3
    saved = _ops.SPD2_LATTICE_MIN_ENTRIES
    try:
        _ops.SPD2_LATTICE_MIN_ENTRIES = 1
        with torch.no_grad():
            two = k.forward(v[:m], v[:n])
            others = []
            for path in ("lattice1", "lattice2"):
                os.environ["MGB_SPD2_PATH"] = path
                try:
                    others.append(k.forward(v[:m], v[:n]))
                finally:
                    del os.environ["MGB_SPD2_PATH"]
            one, two_single_warp = others
            _ops.SPD2_LATTICE_MIN_ENTRIES = 1 << 60
            direct = k.forward(v[:m], v[:n])
    finally:
        _ops.SPD2_LATTICE_MIN_ENTRIES = saved
    assert torch.isfinite(two).all()
    assert rel_err(two, one) < 1e-13 and rel_err(two, two_single_warp) < 1e-13
    assert rel_err(two, direct) < 1e-12


def test_all_orders_profile_launch_matches_single_order_launches():
    """order = 3 / which = 3 of mgb_radial_profile / mgb_spd2_profile (one launch for the three quantities of the
    trust-region step) against the three single-order launches."""
    from materngabo_b200 import _ops
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200 import kernels_spd as ksp

    torch.manual_seed(5)
    z = (3.0 * torch.rand(257, dtype=torch.float64)).to(DEV)
    z[0] = 2e-4
    for dim in (1, 2, 3):
        for cls in (kh.HyperbolicRiemannianMillsonIntegratedMaternKernel, kh.HyperbolicRiemannianMillsonGaussianKernel):
            k = cls(dim=dim, nu=2.5) if "Matern" in cls.__name__ else cls(dim=dim)
            k.lengthscale = 0.8
            k = k.to(DEV)
            with torch.no_grad():
                tval, tcoef, s0, ds, ps = k.nodes(DEV)
            allp = _ops._radial_profile_raw(dim - 1, 3, z, tval, tcoef, float(s0), float(ds), int(ps))
            for o in range(3):
                one = _ops._radial_profile_raw(dim - 1, o, z, tval, tcoef, float(s0), float(ds), int(ps))
                assert rel_err(allp[o], one) < 1e-13, (dim, cls.__name__, o)
    delta = (4.0 * torch.rand(130, dtype=torch.float64)).to(DEV)
    r2 = (delta * delta * (0.5 + torch.rand(130, dtype=torch.float64).to(DEV)))
    k = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5).to(DEV)
    k.lengthscale = 1.1
    with torch.no_grad():
        tcoef, rscale, bscale, bv, bw = k.nodes(DEV)
    allp = _ops._spd2_profile_raw(3, delta, r2, tcoef, rscale, bscale, bv, bw)
    for wh in range(3):
        one = _ops._spd2_profile_raw(wh, delta, r2, tcoef, rscale, bscale, bv, bw)
        assert rel_err(allp[wh], one) < 1e-13, wh


@pytest.mark.parametrize("dim", [4, 5, 6, 7])
def test_hyperbolic_heat_high_dimension(dim):
    """H^n heat kernel, n > 3 (kernels_hyperbolic.py:118-180): -((1 / sinh d) d/dd)^m of the H^2 / H^3 kernel.  The
    reference normalises by the same expression at d = 1e-6 through nested autograd, which is numerical noise for
    n = 4, 6, 7 (normalisers of 6.6e5, -6.6e17, -4.2e9 at kappa = 1.3: repeated division by sinh(1e-6)); the CUDA path
    reproduces that number with the same torch operations on the host and evaluates the numerator - which IS well
    conditioned - from the d-derivatives of the device profile kernels.  Oracle = the reference's algorithm restated
    (oracle/restated.py), run on this machine's CPU."""
    from materngabo_b200 import kernels_hyperbolic as kh
    from oracle import restated as R

    gen = torch.Generator().manual_seed(60 + dim)
    z = 0.6 * torch.randn(23, dim, generator=gen)
    x = torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)
    k = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=dim, nb_points_integral=400)
    k.lengthscale = 1.3
    k = k.to(DEV)
    with pytest.raises(NotImplementedError):        # forward only: a silent zero lengthscale gradient would mislead a fit
        k.forward(x[:10].to(DEV), x[10:].to(DEV))
    k.raw_lengthscale.requires_grad_(False)
    out = k.forward(x[:10].to(DEV), x[10:].to(DEV)).cpu()
    num, nrm = R.hyperbolic_heat_high_dim(x[:10], x[10:], dim, 1.3, 400, return_parts=True)
    ref = num / nrm
    assert out.shape == (10, 13)
    assert float((out * nrm - num).abs().max()) < 1e-9 * float(num.abs().max())          # the well-conditioned part
    assert float((out - ref).abs().max()) < 1e-9 * float(ref.abs().max())
    d = k.forward(x[:10].to(DEV), x[10:20].to(DEV), diag=True).cpu()
    assert float((d - torch.diagonal(ref[:, :10])).abs().max()) < 1e-9 * float(ref.abs().max())
    with pytest.raises(NotImplementedError):
        kh.HyperbolicRiemannianMillsonGaussianKernel(dim=8).to(DEV).forward(torch.zeros(2, 9, device=DEV), torch.zeros(2, 9, device=DEV))
