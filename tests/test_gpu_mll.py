"""-m gpu: the fused single-CTA marginal-likelihood kernel (csrc/mll.cu) against the unfused path (Gram kernel +
torch.linalg + autograd) and against torch autograd through the oracle on the CPU."""
import pytest
import torch

from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
sp = torch.nn.functional.softplus


def _grads(loss, params):
    return torch.autograd.grad(loss, params)


def _close(a, b, rtol):
    """relative to max |b| with an absolute floor (a normalised kernel at a single point has an exactly zero
    lengthscale gradient: both paths return rounding noise there)"""
    a, b = a.detach().reshape(-1).cpu(), b.detach().reshape(-1).cpu()
    return float((a - b).abs().max()) <= rtol * float(b.abs().max()) + 1e-12


@pytest.mark.parametrize("n", [1, 2, 33, 37, 70, 90, 100, 113, 128, 129, 160])     # every size class of the sweep kernel + the Cholesky kernel beyond 128
def test_fused_mll_matches_unfused_and_oracle(n):
    from materngabo_b200 import kernels_sphere
    from materngabo_b200.gpytorch_compat import ScaleKernel
    from materngabo_b200.graphs import exact_mll
    from oracle import restated as R

    gen = torch.Generator().manual_seed(n)
    x = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)
    y = torch.sin(3 * x[:, 0]) + 0.2 * torch.randn(n, generator=gen)
    base = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=None)
    sk = ScaleKernel(base).to(DEV)
    with torch.no_grad():
        base.raw_lengthscale.fill_(-0.3)
        base.raw_nu.fill_(1.1)
        sk.raw_outputscale.fill_(0.4)
    params = [base.raw_lengthscale, base.raw_nu, sk.raw_outputscale]
    noise = torch.tensor(0.02, device=DEV, requires_grad=True)
    mean = torch.tensor(0.15, device=DEV, requires_grad=True)
    xd, yd = x.to(DEV), y.to(DEV)
    lf, info = exact_mll(sk, xd, yd, noise, mean)
    assert int(info) == 0
    gf = _grads(lf, params + [noise, mean])
    lu, _ = exact_mll(sk, xd, yd, noise, mean, fused=False)
    gu = _grads(lu, params + [noise, mean])
    assert abs(float(lf.detach()) - float(lu.detach())) < 1e-11 * max(1.0, abs(float(lu.detach())))
    for a, b in zip(gf, gu):
        assert _close(a, b, 1e-9)
    # oracle: torch autograd through the restated reference kernel on the CPU
    rl = torch.full((1, 1), -0.3, requires_grad=True)
    rn = torch.full((1, 1), 1.1, requires_grad=True)
    ro = torch.tensor(0.4, requires_grad=True)
    nz = torch.tensor(0.02, requires_grad=True)
    mn = torch.tensor(0.15, requires_grad=True)
    k = sp(ro) * R.sphere_matern(x, x, 3, sp(rn), sp(rl))
    lo, _ = exact_mll(k, x, y, nz, mn)
    go = _grads(lo, [rl, rn, ro, nz, mn])
    assert abs(float(lf.detach()) - float(lo.detach())) < 1e-9 * max(1.0, abs(float(lo.detach())))
    for a, b in zip(gf, go):
        assert _close(a, b, 1e-7)


def test_fused_mll_so3_and_circle_and_failure_flag():
    from materngabo_b200 import kernels_so, kernels_sphere
    from materngabo_b200.graphs import exact_mll

    gen = torch.Generator().manual_seed(3)
    q, r = torch.linalg.qr(torch.randn(64, 3, 3, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    q[:, :, 0] *= torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    xr = q.reshape(64, 9).to(DEV)
    y = torch.randn(64, generator=gen).to(DEV)
    k = kernels_so.SORiemannianMaternKernel(dim=3, nu=2.5).to(DEV)
    k.lengthscale = 0.6
    lf, info = exact_mll(k, xr, y, 0.05, 0.0)
    lu, _ = exact_mll(k, xr, y, 0.05, 0.0, fused=False)
    gf, gu = _grads(lf, [k.raw_lengthscale]), _grads(lu, [k.raw_lengthscale])
    assert int(info) == 0 and abs(float(lf) - float(lu)) < 1e-11 and rel_err(gf[0], gu[0]) < 1e-9
    for ls in (0.5, 0.15):                               # monomial and Chebyshev (55+ terms) circle series
        c = kernels_sphere.CircleRiemannianIntegratedMaternKernel(nu=2.5).to(DEV)
        c.lengthscale = ls
        xc = torch.nn.functional.normalize(torch.randn(50, 2, generator=gen), dim=-1).to(DEV)
        yc = torch.randn(50, generator=gen).to(DEV)
        lf, info = exact_mll(c, xc, yc, 0.05, 0.1)
        lu, _ = exact_mll(c, xc, yc, 0.05, 0.1, fused=False)
        gf, gu = _grads(lf, [c.raw_lengthscale]), _grads(lu, [c.raw_lengthscale])
        assert int(info) == 0 and abs(float(lf) - float(lu)) < 1e-10 and rel_err(gf[0], gu[0]) < 1e-8
    # not positive definite: negative "noise" larger than the smallest eigenvalue
    lf, info = exact_mll(k, xr, y, -5.0, 0.0)
    assert int(info) > 0 and not torch.isfinite(lf)


def test_device_coefficient_op_matches_the_torch_formulas():
    """mgb_spectral_coef (coefficients + Jacobian in one launch) against the torch definition evaluated on the CPU:
    values and gradients for the sphere (Matern with learnable nu, heat) and SO(3) kernels."""
    from materngabo_b200 import kernels_so, kernels_sphere

    def compare(make, raw):
        dev_k, cpu_k = make().to(DEV), make()
        outs = []
        for k in (dev_k, cpu_k):
            with torch.no_grad():
                for name, v in raw.items():
                    getattr(k, name).fill_(v)
            params = [getattr(k, name) for name in raw]
            for p in params:
                p.requires_grad_(True)
            spec, coef = k.series()
            w = torch.linspace(0.3, 1.7, coef.numel(), dtype=torch.float64, device=coef.device).reshape(coef.shape)
            g = torch.autograd.grad((coef * w).sum(), params)
            outs.append((spec.basis, coef.detach().cpu(), [t.detach().cpu() for t in g]))
        assert outs[0][0] == outs[1][0]
        assert rel_err(outs[0][1], outs[1][1]) < 1e-13
        for a, b in zip(outs[0][2], outs[1][2]):
            assert _close(a, b, 1e-11)

    compare(lambda: kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=None), {"raw_lengthscale": -0.4, "raw_nu": 0.9})
    compare(lambda: kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=None), {"raw_lengthscale": 0.3, "raw_nu": 2.0})
    compare(lambda: kernels_sphere.SphereRiemannianMaternKernel(dim=2, nu=None, serie_nb_terms=20), {"raw_lengthscale": 0.1, "raw_nu": 1.0})
    compare(lambda: kernels_sphere.SphereRiemannianGaussianKernel(dim=3), {"raw_lengthscale": -0.8})
    compare(lambda: kernels_so.SORiemannianMaternKernel(dim=3, nu=None), {"raw_lengthscale": -0.2, "raw_nu": 1.3})
    compare(lambda: kernels_so.SORiemannianGaussianKernel(dim=3), {"raw_lengthscale": 0.2})


def _hyp_points(n, dim, gen):
    z = 0.7 * torch.randn(n, dim, generator=gen)
    return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)


def _spd_points(n, gen):
    a = torch.randn(n, 2, 2, generator=gen)
    s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
    return torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)


@pytest.mark.parametrize("which,n", [("torus", 60), ("h2", 100), ("h3", 37), ("h3", 75), ("h3", 96), ("h3", 120), ("spd2", 100), ("spd2", 1), ("h2", 164)])
def test_dense_mll_matches_unfused_path_on_every_other_manifold(which, n):
    """Kernels without the single-factor series form go through their own Gram kernel + ONE launch of the CTA kernel
    in dense mode (mgb_dense_mll): value and gradients to every raw hyper-parameter, the noise and the mean against
    the Gram kernel + torch.linalg + autograd path."""
    from materngabo_b200 import kernels_hyperbolic as kh, kernels_spd as ksp, kernels_torus as kt
    from materngabo_b200 import _ops
    from materngabo_b200.gpytorch_compat import ScaleKernel
    from materngabo_b200.graphs import exact_mll

    gen = torch.Generator().manual_seed(n + len(which))
    if which == "torus":
        base, x = kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5), torch.rand(n, 3, generator=gen) * 6.28
    elif which == "h2":
        base, x = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5), _hyp_points(n, 2, gen)
    elif which == "h3":
        base, x = kh.HyperbolicRiemannianMillsonGaussianKernel(dim=3), _hyp_points(n, 3, gen)
    else:
        base, x = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5), _spd_points(n, gen)
    assert n <= _ops.dense_mll_max_n()
    y = torch.sin(2 * x[:, 0]) + 0.1 * torch.randn(n, generator=gen)
    sk = ScaleKernel(base).to(DEV)
    with torch.no_grad():
        sk.raw_outputscale.fill_(0.3)
    params = [p for p in sk.parameters() if p.requires_grad]
    # the reference's H^2 quadrature kernel is not positive definite at its default node counts: larger noise there
    noise = torch.tensor(1.5 if which == "h2" else 0.05, device=DEV, requires_grad=True)
    mean = torch.tensor(-0.1, device=DEV, requires_grad=True)
    xd, yd = x.to(DEV), y.to(DEV)
    lf, info = exact_mll(sk, xd, yd, noise, mean)
    assert int(info) == 0
    gf = torch.autograd.grad(lf, params + [noise, mean], allow_unused=True)
    lu, _ = exact_mll(sk, xd, yd, noise, mean, fused=False)
    gu = torch.autograd.grad(lu, params + [noise, mean], allow_unused=True)
    assert abs(float(lf.detach()) - float(lu.detach())) < 1e-11 * max(1.0, abs(float(lu.detach())))
    for a, b in zip(gf, gu):
        assert (a is None) == (b is None)
        if a is not None:
            assert _close(a, b, 1e-8)


def test_dense_mll_reports_a_non_positive_definite_matrix():
    from materngabo_b200 import _ops

    k = torch.tensor([[1.0, 2.0], [2.0, 1.0]], device=DEV)
    hyper = torch.tensor([1.0, 0.0, 0.0], device=DEV)
    _, info = _ops.dense_mll(k, torch.zeros(2, device=DEV), hyper)
    assert int(info) != 0


# ---- multi-CTA dense path (csrc/dense_spd.cu): 160 < n <= 16384 ---------------------------------------------------------
@pytest.mark.parametrize("n", [161, 500, 1000, 1057])
def test_large_mll_matches_unfused_and_oracle(n):
    """SURVEY.md 8f rank 4 beyond the single-CTA limit: Gram kernel -> cooperative blocked Cholesky -> triangular inverse
    by recursive doubling -> K^-1 = X^T X -> gradient contraction, against (a) the torch.linalg + autograd path on the
    same device and (b) torch autograd through the oracle kernel on the CPU (sizes the oracle finishes in seconds)."""
    from materngabo_b200 import _lib, kernels_sphere
    from materngabo_b200.gpytorch_compat import ScaleKernel
    from materngabo_b200.graphs import exact_mll
    from oracle import restated as R

    assert _lib.load().mgb_series_mll_max_n(4, 10) >= 1000
    gen = torch.Generator().manual_seed(n)
    x = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)
    y = torch.sin(3 * x[:, 0]) + 0.2 * torch.randn(n, generator=gen)
    base = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=None)
    sk = ScaleKernel(base).to(DEV)
    with torch.no_grad():
        base.raw_lengthscale.fill_(-0.3)
        base.raw_nu.fill_(1.1)
        sk.raw_outputscale.fill_(0.4)
    params = [base.raw_lengthscale, base.raw_nu, sk.raw_outputscale]
    noise = torch.tensor(0.02, device=DEV, requires_grad=True)
    mean = torch.tensor(0.15, device=DEV, requires_grad=True)
    xd, yd = x.to(DEV), y.to(DEV)
    lf, info = exact_mll(sk, xd, yd, noise, mean)
    assert int(info) == 0
    gf = _grads(lf, params + [noise, mean])
    lu, _ = exact_mll(sk, xd, yd, noise, mean, fused=False)
    gu = _grads(lu, params + [noise, mean])
    assert abs(float(lf.detach()) - float(lu.detach())) < 1e-11 * max(1.0, abs(float(lu.detach())))
    for a, b in zip(gf, gu):
        assert _close(a, b, 1e-8)
    if n <= 500:
        rl = torch.full((1, 1), -0.3, requires_grad=True)
        rn = torch.full((1, 1), 1.1, requires_grad=True)
        ro = torch.tensor(0.4, requires_grad=True)
        nz = torch.tensor(0.02, requires_grad=True)
        mn = torch.tensor(0.15, requires_grad=True)
        k = sp(ro) * R.sphere_matern(x, x, 3, sp(rn), sp(rl))
        lo, _ = exact_mll(k, x, y, nz, mn)
        go = _grads(lo, [rl, rn, ro, nz, mn])
        assert abs(float(lf.detach()) - float(lo.detach())) < 1e-9 * max(1.0, abs(float(lo.detach())))
        for a, b in zip(gf, go):
            assert _close(a, b, 1e-7)


def test_large_dense_mll_for_a_quadrature_kernel_and_failure_flag():
    from materngabo_b200 import kernels_hyperbolic as kh
    from materngabo_b200.graphs import exact_mll

    gen = torch.Generator().manual_seed(12)
    z = 0.8 * torch.randn(300, 3, generator=gen)          # H^3: closed-form heat factor, positive definite by construction
    x = torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1).to(DEV)
    y = torch.sin(2 * x[:, 1]).to(DEV)
    k = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=3, nu=2.5, nb_points_integral=24).to(DEV)
    k.lengthscale = 0.9
    lf, info = exact_mll(k, x, y, 0.05, 0.1)
    lu, _ = exact_mll(k, x, y, 0.05, 0.1, fused=False)
    assert int(info) == 0 and abs(float(lf.detach()) - float(lu.detach())) < 1e-10 * max(1.0, abs(float(lu.detach())))
    gf, gu = _grads(lf, [k.raw_lengthscale]), _grads(lu, [k.raw_lengthscale])
    assert _close(gf[0], gu[0], 1e-8)
    # not positive definite (negative "noise" larger than the smallest eigenvalue): the flag names a pivot, no exception
    lbad, ibad = exact_mll(k, x, y, -0.5, 0.0)
    assert int(ibad) > 0


@pytest.mark.parametrize("n", [161, 700, 2049])
def test_large_fit_matches_torch_linalg(n):
    """ExactGPPosterior.fit beyond 160 training points: this library's Cholesky + L^-1 (MGB_FIT_PATH=torch selects the
    cuSOLVER route) - L^-1 rows, alpha and the posterior agree; L^-1 L L^-T ... = identity residual at 1e-12."""
    import os

    from materngabo_b200 import kernels_sphere
    from materngabo_b200.posterior import ExactGPPosterior

    gen = torch.Generator().manual_seed(n)
    xt = torch.nn.functional.normalize(torch.randn(n, 11, generator=gen), dim=-1).to(DEV)
    xc = torch.nn.functional.normalize(torch.randn(513, 11, generator=gen), dim=-1).to(DEV)
    y = (torch.sin(3 * xt[:, 0]) + 0.3 * xt[:, 1]).to(DEV)
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
    k.lengthscale = 0.5
    own = ExactGPPosterior(k, xt, y, outputscale=1.2, noise=1e-2, mean_const=0.1).fit()
    os.environ["MGB_FIT_PATH"] = "torch"
    try:
        ref = ExactGPPosterior(k, xt, y, outputscale=1.2, noise=1e-2, mean_const=0.1).fit()
    finally:
        del os.environ["MGB_FIT_PATH"]
    assert float((own.rinv - ref.rinv).abs().max()) < 1e-10 * float(ref.rinv.abs().max())
    assert float((own.rinv.triu(1)).abs().max()) == 0.0
    assert float((own.alpha - ref.alpha).abs().max()) < 1e-9 * float(ref.alpha.abs().max())
    m0, v0 = own.posterior(xc)
    m1, v1 = ref.posterior(xc)
    assert float((m0 - m1).abs().max()) < 1e-10 * float(m1.abs().max())
    assert float((v0 - v1).abs().max()) < 1e-10 * float(v1.abs().max())
    if own.rinv_t is not None:
        assert torch.equal(own.rinv_t, own.rinv.t())
    kt = 1.2 * k.forward(xt, xt).detach() + 1e-2 * torch.eye(n, device=DEV)
    resid = own.rinv @ kt @ own.rinv.t() - torch.eye(n, device=DEV)
    assert float(resid.abs().max()) < 1e-10
