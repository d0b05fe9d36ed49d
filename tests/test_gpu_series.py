"""-m gpu: the CUDA series kernels (through the C-ABI) against the golden vectors generated from the unmodified
reference and against the restated oracle, forward and backward, plus the shapes / edge cases the drop-in
boundary has to accept.  Tolerance: 1e-10 relative to max|K_ref| (float64)."""
import numpy as np
import pytest
import torch

from tests.conftest import TOL, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device=DEV)


def _sphere(name, cls, g, **kw):
    from materngabo_b200 import kernels_sphere as ks

    k = getattr(ks, cls)(dim=int(g["dim"]), **kw)
    k.lengthscale = float(g["lengthscale"])
    return k


@pytest.mark.parametrize("name", ["sphere_matern_S2", "sphere_matern_S3", "sphere_matern_S10", "sphere_matern_S2_self",
                                  "sphere_matern_S3_example"])
def test_sphere_matern_forward(name):
    g = load_golden(name)
    k = _sphere(name, "SphereRiemannianMaternKernel", g, nu=float(g["nu"]))
    out = k.forward(_t(g["x1"]), _t(g["x2"]))
    assert out.shape == g["K"].shape
    assert rel_err(out, g["K"]) < TOL


@pytest.mark.parametrize("name", ["sphere_gaussian_S2", "sphere_gaussian_S3", "sphere_gaussian_S10"])
def test_sphere_gaussian_forward(name):
    g = load_golden(name)
    k = _sphere(name, "SphereRiemannianGaussianKernel", g)
    assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL


def test_sphere_integrated_matern():
    g = load_golden("sphere_intmatern_S3")
    k = _sphere("", "SphereRiemannianIntegratedMaternKernel", g, nu=2.5)
    assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL


def _grad_case(kernel, g, params):
    x1 = _t(g["x1"]).requires_grad_(True)
    x2 = _t(g["x2"]).requires_grad_(True)
    out = kernel.forward(x1, x2)
    plist = [getattr(kernel, p) for p in params]
    for p in plist:
        p.requires_grad_(True)
    res = torch.autograd.grad((out * _t(g["G"])).sum(), [x1, x2] + plist)
    assert rel_err(out, g["K"]) < TOL
    assert rel_err(res[0], g["gx1"]) < 1e-9
    assert rel_err(res[1], g["gx2"]) < 1e-9
    for name, val in zip(params, res[2:]):
        assert rel_err(val.reshape(-1), g["g_" + name].reshape(-1)) < 1e-9, name


@pytest.mark.parametrize("name", ["sphere_matern_S2", "sphere_matern_S3", "sphere_matern_S10"])
def test_sphere_matern_backward(name):
    g = load_golden(name)
    _grad_case(_sphere(name, "SphereRiemannianMaternKernel", g, nu=float(g["nu"])), g, ["raw_lengthscale"])


def test_sphere_matern_nu_gradient():
    from materngabo_b200 import kernels_sphere as ks

    g = load_golden("sphere_matern_S2_nugrad")
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=None)
    k.lengthscale = float(g["lengthscale"])
    k.nu = float(g["nu"])
    _grad_case(k, g, ["raw_lengthscale", "raw_nu"])


def test_sphere_gaussian_backward():
    g = load_golden("sphere_gaussian_S3")
    _grad_case(_sphere("", "SphereRiemannianGaussianKernel", g), g, ["raw_lengthscale"])


def test_circle_kernels():
    from materngabo_b200 import kernels_sphere as ks

    g = load_golden("circle_gaussian")
    k = ks.CircleRiemannianGaussianKernel()
    k.lengthscale = float(g["lengthscale"])
    _grad_case(k, g, ["raw_lengthscale"])
    g = load_golden("circle_intmatern")
    k = ks.CircleRiemannianIntegratedMaternKernel(nu=2.5)
    k.lengthscale = float(g["lengthscale"])
    _grad_case(k, g, ["raw_lengthscale"])


@pytest.mark.parametrize("dim", [3, 10])
def test_torus(dim):
    from materngabo_b200 import kernels_torus as kt

    for kind, cls in (("matern", "TorusProductOfManifoldsRiemannianMaternKernel"),
                      ("gaussian", "TorusProductOfManifoldsRiemannianGaussianKernel")):
        g = load_golden("torus_%s_T%d" % (kind, dim))
        k = getattr(kt, cls)(dim=dim, **({"nu": 2.5} if kind == "matern" else {}))
        for c, kk in enumerate(k.torus_kernel.kernels):
            kk.lengthscale = float(g["lengthscales"][c])
        out = k.forward(_t(g["x1"]), _t(g["x2"]))          # angle inputs
        assert rel_err(out, g["K"]) < TOL
        x1c = k._to_circles(_t(g["x1"]))                    # (cos, sin) inputs give the same matrix
        assert rel_err(k.forward(x1c, k._to_circles(_t(g["x2"]))), g["K"]) < TOL


def test_torus_backward_matches_oracle_autograd():
    from materngabo_b200 import kernels_torus as kt
    from oracle import restated as R

    torch.manual_seed(3)
    a1, a2 = torch.rand(9, 3) * 6.28, torch.rand(7, 3) * 6.28
    G = torch.randn(9, 7)
    ls = [0.4, 0.5, 0.6]
    k = kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5)
    for c, kk in enumerate(k.torus_kernel.kernels):
        kk.lengthscale = ls[c]
    c1 = k._to_circles(a1).to(DEV).requires_grad_(True)
    c2 = k._to_circles(a2).to(DEV)
    out = k.forward(c1, c2)
    params = [kk.raw_lengthscale for kk in k.torus_kernel.kernels]
    res = torch.autograd.grad((out * G.to(DEV)).sum(), [c1] + params)
    # oracle: same thing through torch autograd on CPU
    raw = [kk.raw_lengthscale.detach().clone().requires_grad_(True) for kk in k.torus_kernel.kernels]
    oc1 = k._to_circles(a1).requires_grad_(True)
    ref = R.torus_matern(oc1, k._to_circles(a2), 3, 2.5, [torch.nn.functional.softplus(r) for r in raw])
    rres = torch.autograd.grad((ref * G).sum(), [oc1] + raw)
    assert rel_err(out, ref) < TOL
    assert rel_err(res[0], rres[0]) < 1e-9
    for a, b in zip(res[1:], rres[1:]):
        assert rel_err(a, b) < 1e-9


@pytest.mark.parametrize("name,cls", [("so3_matern", "SORiemannianMaternKernel"), ("so3_gaussian", "SORiemannianGaussianKernel"),
                                      ("so3_matern_self", "SORiemannianMaternKernel")])
def test_so3(name, cls):
    from materngabo_b200 import kernels_so as kso

    g = load_golden(name)
    k = getattr(kso, cls)(dim=3, **({"nu": 2.5} if "Matern" in cls else {}))
    k.lengthscale = float(g["lengthscale"])
    if "G" in g:
        _grad_case(k, g, ["raw_lengthscale"])
    else:
        assert rel_err(k.forward(_t(g["x1"]), _t(g["x2"])), g["K"]) < TOL


def test_so_unsupported_dims_raise():
    """n = 2, 4 ... 7 run on csrc/son_gram.cu (tests/test_gpu_son.py); rank > 3 is refused loudly."""
    from materngabo_b200 import kernels_so as kso

    k = kso.SORiemannianMaternKernel(dim=8, nu=2.5)
    with pytest.raises(NotImplementedError):
        k.forward(torch.zeros(2, 64, device=DEV), torch.zeros(2, 64, device=DEV))
    k4 = kso.SORiemannianMaternKernel(dim=4, nu=2.5)
    with pytest.raises(RuntimeError):
        k4.forward(torch.zeros(2, 9, device=DEV), torch.zeros(2, 9, device=DEV))      # wrong row width


def test_circle_dimension_raises_like_reference():
    from materngabo_b200 import kernels_sphere as ks

    with pytest.raises(ValueError):  # math.gamma(0), as in the reference for S^1 via the Gegenbauer classes
        ks.SphereRiemannianMaternKernel(dim=1, nu=2.5)


# ---- shapes the gpytorch / botorch callers use ---------------------------------------------------
def test_batch_layouts_and_diag():
    from materngabo_b200 import kernels_sphere as ks
    from oracle import restated as R

    torch.manual_seed(0)
    k = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    k.lengthscale = 0.5
    xb = torch.nn.functional.normalize(torch.randn(17, 1, 4), dim=-1)
    xt = torch.nn.functional.normalize(torch.randn(23, 4), dim=-1)
    ref = R.sphere_matern(xb, xt.expand(17, 23, 4), 3, 2.5, 0.5)
    out = k.forward(xb.to(DEV), xt.to(DEV).expand(17, 23, 4))      # botorch posterior layout b x 1 x D vs b x N x D
    assert out.shape == (17, 1, 23) and rel_err(out, ref) < TOL
    x3 = torch.nn.functional.normalize(torch.randn(3, 5, 4), dim=-1)  # genuinely batched
    y3 = torch.nn.functional.normalize(torch.randn(3, 6, 4), dim=-1)
    assert rel_err(k.forward(x3.to(DEV), y3.to(DEV)), R.sphere_matern(x3, y3, 3, 2.5, 0.5)) < TOL
    d = k.forward(xt.to(DEV), xt.to(DEV), diag=True)
    assert d.shape == (23,) and rel_err(d, R.sphere_matern(xt, xt, 3, 2.5, 0.5, diag=True).reshape(-1)) < TOL
    kk = k(xt.to(DEV))                                             # Kernel.__call__ with x2 = None
    assert rel_err(kk, R.sphere_matern(xt, xt, 3, 2.5, 0.5)) < TOL
    # a MATERIALISED expansion of the training block (no stride-0 batch dimension): still one 2-D launch
    out2 = k.forward(xb.to(DEV), xt.to(DEV).expand(17, 23, 4).contiguous())
    assert out2.shape == (17, 1, 23) and torch.equal(out2, out)
    # diag of two different point lists (one thread per row)
    xs = xt.roll(5, 0)
    d2 = k.forward(xt.to(DEV), xs.to(DEV), diag=True)
    assert rel_err(d2, torch.diagonal(R.sphere_matern(xt, xs, 3, 2.5, 0.5))) < TOL


def test_batched_launch_and_diag_on_every_series_family():
    """gridDim.z batches (sphere small kernel, torus general kernel, SO(3)) and the per-row diag kernel (monomial and
    Chebyshev bases, multi-factor) against the oracle; a batch larger than one wave and ragged member sizes."""
    from materngabo_b200 import kernels_so as kso, kernels_sphere as ks, kernels_torus as kt
    from oracle import restated as R

    gen = torch.Generator().manual_seed(5)
    k = ks.SphereRiemannianMaternKernel(dim=10, nu=2.5)
    k.lengthscale = 0.5
    x3 = torch.nn.functional.normalize(torch.randn(40, 33, 11, generator=gen), dim=-1)
    y3 = torch.nn.functional.normalize(torch.randn(40, 70, 11, generator=gen), dim=-1)
    out = k.forward(x3.to(DEV), y3.to(DEV))
    assert out.shape == (40, 33, 70) and rel_err(out, R.sphere_matern(x3, y3, 10, 2.5, 0.5)) < TOL
    # two batch dimensions, non-contiguous members
    x4 = x3.reshape(5, 8, 33, 11).transpose(0, 1)
    y4 = y3.reshape(5, 8, 70, 11).transpose(0, 1)
    out4 = k.forward(x4.to(DEV), y4.to(DEV))
    assert out4.shape == (8, 5, 33, 70) and rel_err(out4, out.reshape(5, 8, 33, 70).transpose(0, 1)) < 1e-15
    # torus: multi-factor series (general kernel) batched + diag
    t = kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5)
    for kk in t.torus_kernel.kernels:
        kk.lengthscale = 0.5
    a = torch.rand(4, 9, 3, generator=gen) * 6.283185307179586
    b = torch.rand(4, 12, 3, generator=gen) * 6.283185307179586
    outt = t.forward(a.to(DEV), b.to(DEV))
    reft = torch.stack([R.torus_matern(R._to_circles(a[i], 3), R._to_circles(b[i], 3), 3, 2.5, [0.5] * 3) for i in range(4)])
    assert outt.shape == (4, 9, 12) and rel_err(outt, reft) < TOL
    dt = t.forward(a[0].to(DEV), b[0][:9].to(DEV), diag=True)
    assert rel_err(dt, torch.diagonal(reft[0][:, :9])) < TOL
    # SO(3)
    q, r = torch.linalg.qr(torch.randn(3, 20, 3, 3, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    q[..., 0] *= torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    rot = q.reshape(3, 20, 9)
    so = kso.SORiemannianMaternKernel(dim=3, nu=2.5)
    so.lengthscale = 0.5
    outs = so.forward(rot[:, :8].to(DEV), rot[:, 8:].to(DEV))
    refs = torch.stack([R.so3_matern(rot[i, :8], rot[i, 8:], 2.5, 0.5) for i in range(3)])
    assert rel_err(outs, refs) < TOL
    ds = so.forward(rot[0, :8].to(DEV), rot[1, :8].to(DEV), diag=True)
    assert rel_err(ds, torch.diagonal(R.so3_matern(rot[0, :8], rot[1, :8], 2.5, 0.5))) < TOL


def test_small_problem_kernel_matches_the_panel_kernels():
    """config1-sized problems take series_gram_fwd_small; MGB_SERIES_SMALL=0 is read once per process, so compare with
    the oracle and with a > 4 M-entry call (panel kernel) on the same rows."""
    from materngabo_b200 import kernels_so as kso, kernels_sphere as ks
    from oracle import restated as R

    gen = torch.Generator().manual_seed(8)
    for dim in (2, 3, 10):
        k = ks.SphereRiemannianMaternKernel(dim=dim, nu=2.5)
        k.lengthscale = 0.5
        x = torch.nn.functional.normalize(torch.randn(1000, dim + 1, generator=gen), dim=-1)
        x[5] = x[3]                                       # clamp active
        small = k.forward(x.to(DEV), x.to(DEV))           # 1 M entries: small kernel
        assert rel_err(small, R.sphere_matern(x, x, dim, 2.5, 0.5)) < TOL
        big = torch.nn.functional.normalize(torch.randn(5000, dim + 1, generator=gen), dim=-1)
        big[:1000] = x
        panel = k.forward(big.to(DEV), x.to(DEV))         # 5 M entries: persistent panel kernel
        assert float((panel[:1000] - small).abs().max()) < 1e-14
    for m, n in ((1, 1), (3, 130), (77, 65), (1001, 999)):
        k = ks.SphereRiemannianGaussianKernel(dim=3)
        k.lengthscale = 0.4
        a = torch.nn.functional.normalize(torch.randn(m, 4, generator=gen), dim=-1)
        b = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)
        assert rel_err(k.forward(a.to(DEV), b.to(DEV)), R.sphere_gaussian(a, b, 3, 0.4)) < TOL


@pytest.mark.parametrize("m,n", [(0, 5), (5, 0), (1, 1), (129, 65), (257, 3), (64, 127), (1, 1001)])
def test_ragged_sizes(m, n):
    from materngabo_b200 import kernels_sphere as ks
    from oracle import restated as R

    gen = torch.Generator().manual_seed(m * 1000 + n)
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    k.lengthscale = 0.5
    x1 = torch.nn.functional.normalize(torch.randn(max(m, 1), 3, generator=gen), dim=-1)[:m]
    x2 = torch.nn.functional.normalize(torch.randn(max(n, 1), 3, generator=gen), dim=-1)[:n]
    out = k.forward(x1.to(DEV), x2.to(DEV))
    assert out.shape == (m, n)
    if m and n:
        assert rel_err(out, R.sphere_matern(x1, x2, 2, 2.5, 0.5)) < TOL


def test_non_contiguous_and_strided_inputs():
    from materngabo_b200 import kernels_sphere as ks
    from oracle import restated as R

    torch.manual_seed(5)
    big = torch.nn.functional.normalize(torch.randn(40, 4), dim=-1)
    k = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    k.lengthscale = 0.5
    x1 = big.to(DEV)[::2]            # row-strided view
    x2 = big.to(DEV)[1:, :]          # offset view: base pointer not 16-byte aligned relative to the panel
    assert rel_err(k.forward(x1, x2), R.sphere_matern(big[::2], big[1:], 3, 2.5, 0.5)) < TOL


def test_generic_width_and_many_terms_use_general_kernel():
    from materngabo_b200 import kernels_sphere as ks
    from oracle import restated as R

    torch.manual_seed(6)
    x = torch.nn.functional.normalize(torch.randn(50, 8), dim=-1)   # S^7: no register-resident specialisation
    k = ks.SphereRiemannianMaternKernel(dim=7, nu=1.5)
    k.lengthscale = 0.7
    assert rel_err(k.forward(x.to(DEV), x[:30].to(DEV)), R.sphere_matern(x, x[:30], 7, 1.5, 0.7)) < TOL
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5, serie_nb_terms=20)  # T > 12 -> Chebyshev basis
    k.lengthscale = 0.3
    y = torch.nn.functional.normalize(torch.randn(40, 3), dim=-1)
    assert rel_err(k.forward(y.to(DEV), y.to(DEV)), R.sphere_matern(y, y, 2, 2.5, 0.3, terms=20)) < TOL


def test_cpu_tensors_are_rejected_not_silently_computed():
    from materngabo_b200 import kernels_sphere as ks
    from materngabo_b200._lib import MgbError

    k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    x = torch.nn.functional.normalize(torch.randn(4, 3), dim=-1)
    with pytest.raises(MgbError):
        k.forward(x, x)


def test_config1_full_size_vs_oracle():
    """BASELINE config 1: S^2 Matern nu=2.5 kappa=0.5, 1000 x 1000, seed 0."""
    from materngabo_b200 import kernels_sphere as ks
    from oracle import restated as R

    torch.manual_seed(0)
    x = torch.randn(1000, 3)
    x = x / x.norm(dim=-1, keepdim=True)
    k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    k.lengthscale = 0.5
    out = k.forward(x.to(DEV), x.to(DEV))
    assert rel_err(out, R.sphere_matern(x, x, 2, 2.5, 0.5)) < TOL
    assert rel_err(out, out.t()) < 1e-14                     # symmetry
    assert torch.linalg.eigvalsh(out + 1e-8 * torch.eye(1000, device=DEV)).min() > 0  # PD (kernels_so.py:365-400 style check)


def _forced_trace(kso, k, a, b):
    saved = kso.QUATERNION_MIN_ENTRIES
    try:
        kso.QUATERNION_MIN_ENTRIES = 1 << 62
        return k.forward(a, b)
    finally:
        kso.QUATERNION_MIN_ENTRIES = saved


def test_large_gram_properties():
    """Size-independent checks at a size the oracle cannot reach: row blocks of a 200k x 2k SO(3) Gram agree with
    the same rows computed alone, and a sampled set of entries agrees with the oracle."""
    from materngabo_b200 import kernels_so as kso
    from oracle import restated as R

    gen = torch.Generator().manual_seed(11)
    q, r = torch.linalg.qr(torch.randn(2000, 3, 3, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    q[:, :, 0] *= torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    base = q.reshape(2000, 9)
    idx = torch.randint(0, 2000, (200000,), generator=gen)
    xc = base[idx].to(DEV)
    k = kso.SORiemannianMaternKernel(dim=3, nu=2.5)
    k.lengthscale = 0.5
    big = k.forward(xc, base.to(DEV))
    assert big.shape == (200000, 2000)
    rows = torch.tensor([0, 1, 127, 128, 99999, 199999])
    small = k.forward(xc[rows.to(DEV)], base.to(DEV))
    # the large block goes through the quaternion form of the inner product, the 6-row block through the trace form
    assert rel_err(big[rows.to(DEV)], small) < 1e-13
    saved = kso.QUATERNION_MIN_ENTRIES
    try:
        kso.QUATERNION_MIN_ENTRIES = 1 << 62                  # force the 9-wide trace form
        big9 = k.forward(xc[:20000], base.to(DEV))
        assert torch.equal(big9[rows[:4].to(DEV)], small[:4])  # tiling does not change values
        assert rel_err(big[:20000], big9) < 1e-13
        kso.QUATERNION_MIN_ENTRIES = saved
        # inputs that are not rotations (here: drifted by 1e-6) must keep the reference's trace semantics exactly
        bad = xc[:20000].clone()
        bad[5] += 1e-6
        assert torch.equal(k.forward(bad, base.to(DEV)), _forced_trace(kso, k, bad, base.to(DEV)))
    finally:
        kso.QUATERNION_MIN_ENTRIES = saved
    ref = R.so3_matern(base[idx[rows]], base[:300], 2.5, 0.5)
    assert rel_err(big[rows.to(DEV)][:, :300], ref) < TOL
    assert torch.isfinite(big).all()


def test_second_order_input_gradient_matches_oracle():
    """Trust-region Hessian-vector products differentiate the input gradient again
    (pymanopt_addons/tools/autodiff/_pytorch.py:92-116): d/dx1 [ (d/dx1 sum(K*G)) . v ]."""
    from materngabo_b200 import kernels_sphere as ks
    from oracle import restated as R

    torch.manual_seed(8)
    x1 = torch.nn.functional.normalize(torch.randn(6, 4), dim=-1)
    x2 = torch.nn.functional.normalize(torch.randn(9, 4), dim=-1)
    G, v = torch.randn(6, 9), torch.randn(6, 4)
    k = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    k.lengthscale = 0.5

    def hvp(fn, a, b, g, vec):
        a = a.clone().requires_grad_(True)
        out = fn(a, b)
        (ga,) = torch.autograd.grad((out * g).sum(), [a], create_graph=True)
        (h,) = torch.autograd.grad((ga * vec).sum(), [a])
        return ga.detach(), h

    g_ref, h_ref = hvp(lambda a, b: R.sphere_matern(a, b, 3, 2.5, 0.5), x1, x2, G, v)
    g_dev, h_dev = hvp(lambda a, b: k.forward(a, b), x1.to(DEV), x2.to(DEV), G.to(DEV), v.to(DEV))
    assert rel_err(g_dev, g_ref) < 1e-9
    assert rel_err(h_dev, h_ref) < 1e-8


def test_gradient_through_scaled_sum_of_kernels():
    """Autograd composition as gpytorch uses it: ScaleKernel(base)(X) + noise, scalar loss, hyper-gradients."""
    from materngabo_b200 import kernels_sphere as ks
    from materngabo_b200.gpytorch_compat import ScaleKernel
    from oracle import restated as R

    torch.manual_seed(9)
    x = torch.nn.functional.normalize(torch.randn(50, 4), dim=-1)
    y = torch.sin(2 * x[:, 0])
    base = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    sk = ScaleKernel(base).to(DEV)          # as SingleTaskGP.to(train_X) does: hyper-parameters live on the GPU
    sk.outputscale = 1.7

    def nll(kmat, yv):
        L = torch.linalg.cholesky(kmat + 1e-2 * torch.eye(kmat.shape[0], device=kmat.device))
        a = torch.cholesky_solve(yv.unsqueeze(-1), L).squeeze(-1)
        return 0.5 * (yv * a).sum() + torch.log(torch.diagonal(L)).sum()

    loss = nll(sk(x.to(DEV)), y.to(DEV))
    grads = torch.autograd.grad(loss, [base.raw_lengthscale, sk.raw_outputscale])
    raw_ls = base.raw_lengthscale.detach().cpu().clone().requires_grad_(True)
    raw_os = sk.raw_outputscale.detach().cpu().clone().requires_grad_(True)
    kref = torch.nn.functional.softplus(raw_os) * R.sphere_matern(x, x, 3, 2.5, torch.nn.functional.softplus(raw_ls))
    ref = torch.autograd.grad(nll(kref, y), [raw_ls, raw_os])
    assert abs(float(loss) - float(nll(kref, y))) < 1e-9
    assert rel_err(grads[0], ref[0]) < 1e-8 and rel_err(grads[1].reshape(-1), ref[1].reshape(-1)) < 1e-8


def test_float32_construction_regime_on_the_device():
    """The reference's BO scripts build the kernels under a float32 default dtype and then cast the model to the
    float64 training data ("BO-loop dtype regime"): the constants stay float32-rounded inside
    float64 arithmetic.  The device path (incl. the one-launch coefficient op) must reproduce exactly that regime."""
    from materngabo_b200 import kernels_sphere as ks
    from oracle import restated as R

    torch.set_default_dtype(torch.float32)
    try:
        k = ks.SphereRiemannianMaternKernel(dim=2, nu=2.5)
        k.lengthscale = 0.5
    finally:
        torch.set_default_dtype(torch.float64)
    k = k.to(device=DEV, dtype=torch.float64)            # what SingleTaskGP does with the model (self.to(train_X))
    gen = torch.Generator().manual_seed(6)
    x1 = torch.nn.functional.normalize(torch.randn(50, 3, generator=gen), dim=-1)
    x2 = torch.nn.functional.normalize(torch.randn(40, 3, generator=gen), dim=-1)
    out = k.forward(x1.to(DEV), x2.to(DEV))
    # lengthscale and nu were set through float32 raw parameters: 0.5 and 2.5 only to float32 accuracy
    ls, nu = float(k.lengthscale.detach()), float(k.nu.detach())
    ref32 = R.sphere_matern(x1, x2, 2, nu, ls, const_dtype=torch.float32)
    ref64 = R.sphere_matern(x1, x2, 2, nu, ls)
    assert rel_err(out, ref32) < TOL
    assert rel_err(ref32, ref64) > 1e-9                   # the two regimes really differ


@pytest.mark.parametrize("ls,nu", [(0.5, 2.5), (0.15, 1.5), (1.3, 0.5)])
def test_circle_coefficients_on_the_device_match_the_torch_formulas(ls, nu):
    """mgb_circle_coef (one launch: 201 Chebyshev coefficients of a torus factor + Jacobian) against the torch
    formulas of _weights.py on the CPU: values, truncation length and the gradients to the lengthscale and nu."""
    from materngabo_b200 import _weights as w

    for heat in (False, True):
        outs = []
        for dev in ("cpu", DEV):
            l = torch.tensor(ls, dtype=torch.float64, device=dev, requires_grad=True)
            n = torch.tensor(nu, dtype=torch.float64, device=dev, requires_grad=True)
            c = w.circle_gaussian_coef(l) if heat else w.circle_integrated_matern_coef(l, n, 50)
            probe = torch.cos(0.37 * torch.arange(c.numel(), dtype=torch.float64, device=dev))
            g = torch.autograd.grad((c * probe).sum(), [l] if heat else [l, n])
            outs.append((c.detach().cpu(), [x.cpu() for x in g]))
        (c0, g0), (c1, g1) = outs
        assert c0.numel() == c1.numel()
        assert float((c0 - c1).abs().max()) < 1e-14 * float(c0.abs().max())
        for a, b in zip(g0, g1):
            assert abs(float(a) - float(b)) < 1e-11 * max(1.0, abs(float(a)))
