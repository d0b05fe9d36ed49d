"""-m gpu: SO(n) kernels for n != 3 (csrc/son_gram.cu) - the reference's character polynomial at the canonical
(descending) eigen-angles of R1 R2^T.

Parity with the UNMODIFIED reference holds on the pairs where LAPACK listed the conjugate pairs in that order
(golden vectors record which: oracle/make_golden_son.py); on ALL pairs the kernel is compared with the oracle evaluated
at the sorted angles - the well-defined function this path computes.  Also: hyper-parameter gradients, batch / diag
layouts and the positive-definiteness smoke test of kernel_utils/kernels_so.py:365-400."""
import numpy as np
import pytest
import torch

from tests.conftest import TOL, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _kernel(kind, n, ls=0.5):
    from materngabo_b200 import kernels_so as kso

    k = kso.SORiemannianMaternKernel(dim=n, nu=2.5) if kind == "matern" else kso.SORiemannianGaussianKernel(dim=n)
    k.lengthscale = ls
    return k.to(DEV)


def _rotations(cnt, n, gen):
    q, r = torch.linalg.qr(torch.randn(cnt, n, n, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    q[:, :, 0] = q[:, :, 0] * torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    return q.reshape(cnt, n * n)


@pytest.mark.parametrize("n", [2, 4, 5, 6])
@pytest.mark.parametrize("kind", ["matern", "gaussian"])
def test_parity_with_the_reference_on_canonical_pairs(n, kind):
    from oracle import restated as R

    g = load_golden("son_%s_SO%d" % (kind, n))
    x1, x2 = torch.from_numpy(g["x1"]), torch.from_numpy(g["x2"])
    k = _kernel(kind, n)
    out = k.forward(x1.to(DEV), x2.to(DEV)).detach().cpu()
    mask = torch.from_numpy(g["canonical"])
    ref = torch.from_numpy(g["K"])
    assert int(mask.sum()) >= 30
    assert float((out - ref)[mask].abs().max()) < TOL * float(ref.abs().max())      # the reference itself, canonical pairs
    # every pair: the oracle at the angles sorted descending (what this path defines)
    ang = torch.from_numpy(g["angles"]).abs().sort(dim=-1, descending=True).values
    full = R.son_matern(x1, x2, n, 2.5, 0.5, angles=ang) if kind == "matern" else R.son_gaussian(x1, x2, n, 0.5, angles=ang)
    assert rel_err(out, full) < TOL
    if n in (5, 6):      # the reference really is order dependent there: the non-canonical pairs differ
        assert float((full - ref)[~mask].abs().max()) > 1e-6
    else:                # n = 2 (one angle) and n = 4 (the A_2 polynomial is symmetric in its two arguments): every pair
        assert float((out - ref).abs().max()) < TOL * float(ref.abs().max())


@pytest.mark.parametrize("n", [4, 5])
def test_lengthscale_gradient_on_canonical_pairs(n):
    g = load_golden("son_matern_SO%d" % n)
    k = _kernel("matern", n)
    k.raw_lengthscale.requires_grad_(True)
    out = k.forward(torch.from_numpy(g["x1"]).to(DEV), torch.from_numpy(g["x2"]).to(DEV))
    (gl,) = torch.autograd.grad((out * torch.from_numpy(g["G"]).to(DEV)).sum(), [k.raw_lengthscale])
    assert rel_err(gl.reshape(-1), g["g_raw_lengthscale"].reshape(-1)) < 1e-9
    # the basis route (mode 1) and the fused route (mode 0) agree
    k.raw_lengthscale.requires_grad_(False)
    fused = k.forward(torch.from_numpy(g["x1"]).to(DEV), torch.from_numpy(g["x2"]).to(DEV))
    assert rel_err(out, fused) < 1e-13


@pytest.mark.parametrize("n", [2, 4, 5, 6, 7])
def test_function_of_the_pair_and_layouts(n):
    """Symmetric, invariant under conjugation (class function of R1 R2^T), unit diagonal, ragged sizes, batch, diag."""
    gen = torch.Generator().manual_seed(100 + n)
    x = _rotations(37, n, gen).to(DEV)
    k = _kernel("matern", n)
    full = k.forward(x, x)
    assert full.shape == (37, 37)
    assert float((full - full.t()).abs().max()) < 1e-10
    q = _rotations(1, n, gen).reshape(n, n).to(DEV)
    xc = (q @ x.reshape(-1, n, n) @ q.t()).reshape(-1, n * n)                        # simultaneous conjugation
    assert float((k.forward(xc, xc) - full).abs().max()) < 1e-9
    xr = (x.reshape(-1, n, n) @ q).reshape(-1, n * n)                                # right translation: R1 R2^T unchanged
    assert float((k.forward(xr, xr) - full).abs().max()) < 1e-9
    assert float((torch.diagonal(full) - 1.0).abs().max()) < 1e-9
    d = k.forward(x, x.roll(1, 0), diag=True)
    assert d.shape == (37,) and float((d - torch.diagonal(k.forward(x, x.roll(1, 0)))).abs().max()) < 1e-13
    b = k.forward(x[:20].reshape(4, 5, n * n), x[20:36].reshape(4, 4, n * n))
    assert b.shape == (4, 5, 4) and float((b[2] - k.forward(x[10:15], x[28:32])).abs().max()) < 1e-13
    assert k.forward(x[:0], x).shape == (0, 37)
    with pytest.raises(NotImplementedError):
        k.forward(x.clone().requires_grad_(True), x)


def test_positive_definiteness_smoke_test_of_the_reference():
    """kernels_so.py:365-400: Gram matrices of 2 ... 20 random rotations admit a Cholesky factorisation, n = 2 ... 6
    (n = 3 on the series kernel).  The reference only prints the outcome; here it is asserted with a 1e-10 jitter, and
    the smallest eigenvalue is reported through the assertion message."""
    gen = torch.Generator().manual_seed(1234)
    for n in (2, 3, 4, 5, 6):
        k = _kernel("matern", n)
        for size in (2, 5, 10, 20):
            x = _rotations(size, n, gen).to(DEV)
            gram = k.forward(x, x).detach().cpu().numpy()
            gram = 0.5 * (gram + gram.T)
            lo = float(np.linalg.eigvalsh(gram).min())
            assert lo > -1e-10, "SO(%d), %d points: smallest eigenvalue %.3e" % (n, size, lo)
