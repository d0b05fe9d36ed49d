"""The botorch-facing adapter (materngabo_b200/botorch_adapter.py) against doubles of the gpytorch / botorch objects
the reference builds (bo_sphere.py:211-233,265): neither package exists in this image, so SingleTaskGP, the Gaussian
likelihood, the constant mean and botorch's analytic ExpectedImprovement are restated here from their public
contracts.  CPU tests cover the hyper-parameter read-out; -m gpu tests run EI(model=adapter) end to end."""
import math

import numpy as np
import pytest
import torch

from materngabo_b200 import gpytorch_compat as gc


class _Likelihood(torch.nn.Module):
    def __init__(self, noise):
        super().__init__()
        self.raw_noise = torch.nn.Parameter(torch.zeros(1))
        self.lower = 1e-8
        with torch.no_grad():
            v = torch.tensor([noise - self.lower])
            self.raw_noise.copy_(v + torch.log(-torch.expm1(-v)))

    @property
    def noise(self):
        return torch.nn.functional.softplus(self.raw_noise) + self.lower


class _ConstantMean(torch.nn.Module):
    def __init__(self, c):
        super().__init__()
        self.constant = torch.nn.Parameter(torch.tensor([c]))


class SingleTaskGPDouble(torch.nn.Module):
    """Attribute layout of botorch.models.SingleTaskGP(x, y[:, None], covar_module=..., likelihood=...)."""

    def __init__(self, x, y, covar_module, likelihood, mean_const=0.0):
        super().__init__()
        self.train_inputs = (x,)
        self.train_targets = y.reshape(-1)
        self.covar_module = covar_module
        self.likelihood = likelihood
        self.mean_module = _ConstantMean(mean_const)


class BotorchEIDouble(torch.nn.Module):
    """botorch.acquisition.analytic.ExpectedImprovement.forward restated (0.4-0.6 era): model.posterior(X) ->
    mean / variance of shape b x 1 x 1 -> b values."""

    def __init__(self, model, best_f, maximize=True):
        super().__init__()
        self.add_module("model", model)            # botorch registers the model as a sub-module
        self.maximize = maximize
        self.register_buffer("best_f", torch.as_tensor(float(best_f)))

    def forward(self, X):
        self.best_f = self.best_f.to(X)
        posterior = self.model.posterior(X=X)
        mean = posterior.mean
        view_shape = mean.shape[:-2] if mean.shape[-2] == 1 else mean.shape[:-1]
        mean = mean.view(view_shape)
        sigma = posterior.variance.clamp_min(1e-9).sqrt().view(view_shape)
        u = (mean - self.best_f.expand_as(mean)) / sigma
        if not self.maximize:
            u = -u
        normal = torch.distributions.Normal(torch.zeros_like(u), torch.ones_like(u))
        return sigma * (torch.exp(normal.log_prob(u)) + u * normal.cdf(u))


def _model(dev="cpu", n=60, seed=0, noise=0.02, outputscale=1.7, mean_const=0.3):
    from materngabo_b200 import kernels_sphere as ks

    gen = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1).to(dev)
    y = (torch.sin(3 * x[:, 0]) + 0.2 * x[:, 1]).to(dev)
    base = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    base.lengthscale = 0.6
    k = gc.ScaleKernel(base)
    k.outputscale = outputscale
    model = SingleTaskGPDouble(x, y, k, _Likelihood(noise), mean_const).to(dev)
    return model, x, y


def test_read_model_extracts_the_fitted_hyperparameters():
    from materngabo_b200.botorch_adapter import FusedExactGP

    model, x, y = _model()
    kernel, tx, ty, s, noise, mean_const = FusedExactGP.read_model(model)
    assert kernel is model.covar_module.base_kernel
    assert tx.shape == (60, 4) and ty.shape == (60,)
    assert abs(s - 1.7) < 1e-12 and abs(noise - 0.02) < 1e-12 and abs(mean_const - 0.3) < 1e-12
    model.outcome_transform = object()
    with pytest.raises(NotImplementedError):
        FusedExactGP.read_model(model)


def test_adapter_refuses_cpu_models_loudly():
    from materngabo_b200 import _lib
    from materngabo_b200.botorch_adapter import FusedExactGP

    model, _, _ = _model()
    with pytest.raises(_lib.MgbError):
        FusedExactGP.from_model(model)            # CPU tensors: no fallback


@pytest.mark.gpu
def test_botorch_ei_runs_unchanged_on_the_adapter():
    from materngabo_b200.botorch_adapter import FusedExactGP
    from oracle import restated as R

    dev = "cuda:0"
    model, x, y = _model(dev)
    fused = FusedExactGP.from_model(model)
    acq = BotorchEIDouble(model=fused, best_f=float(y.min()), maximize=False).to(dev)
    gen = torch.Generator().manual_seed(9)
    X = torch.nn.functional.normalize(torch.randn(100, 1, 4, generator=gen), dim=-1).to(dev)   # b x 1 x D (manifold_optimize.py:323)
    with torch.no_grad():
        ei = acq(X)
    assert ei.shape == (100,)
    kfn = lambda a, b: R.sphere_matern(a, b, 3, 2.5, 0.6)  # noqa: E731
    mr, vr = R.gp_posterior(kfn, x.cpu(), y.cpu(), X.squeeze(-2).cpu(), 1.7, 0.02, 0.3)
    ref = R.expected_improvement(mr, vr, float(y.min()), maximize=False)
    assert float((ei.cpu() - ref).abs().max()) < 1e-10 * max(1.0, float(ref.abs().max()))
    post = fused.posterior(X)
    assert post.mean.shape == (100, 1, 1) and post.variance.shape == (100, 1, 1)
    assert float((post.mean.reshape(-1).cpu() - mr).abs().max()) < 1e-10 * float(mr.abs().max())
    # pymanopt cost / egrad closures: one point, gradient through the acquisition value (manifold_optimize.py:184-197)
    x1 = X[:1].clone().requires_grad_(True)
    val = -acq(x1).sum()
    (g,) = torch.autograd.grad(val, x1)
    xr = X[:1].squeeze(-2).cpu().clone().requires_grad_(True)
    mr1, vr1 = R.gp_posterior(kfn, x.cpu(), y.cpu(), xr, 1.7, 0.02, 0.3)
    (gr,) = torch.autograd.grad(-R.expected_improvement(mr1, vr1, float(y.min()), maximize=False).sum(), xr)
    assert float((g.reshape(-1).cpu() - gr.reshape(-1)).abs().max()) < 1e-8 * max(1.0, float(gr.abs().max()))
    # refit after the source model changed (next BO iteration)
    model.covar_module.outputscale = 0.9
    fused.refit()
    assert abs(fused.gp.outputscale - 0.9) < 1e-12
    post2 = fused.posterior(X, observation_noise=True)
    mr2, vr2 = R.gp_posterior(kfn, x.cpu(), y.cpu(), X.squeeze(-2).cpu(), 0.9, 0.02, 0.3)
    assert float((post2.variance.reshape(-1).cpu() - (vr2 + 0.02)).abs().max()) < 1e-10


def _same(a, b):
    """Same numbers up to the summation order of the split partial sums (block size changes the split count)."""
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return a.shape == b.shape and (a.numel() == 0 or float((a - b).abs().max()) <= 1e-13)


@pytest.mark.gpu
def test_persistent_handle_create_eval_many_destroy():
    """mgb_posterior_create (host state) / _create_from_device / _eval_host x N / _eval_device / _destroy through ctypes."""
    import ctypes as C

    from materngabo_b200 import _lib
    from materngabo_b200.posterior import ExactGPPosterior
    from oracle import restated as R

    dev = "cuda:0"
    model, x, y = _model(dev, n=150, seed=4)
    gp = ExactGPPosterior.from_model(model)
    gen = torch.Generator().manual_seed(2)
    xc = torch.nn.functional.normalize(torch.randn(20000, 4, generator=gen), dim=-1)
    m_ref, v_ref = (t.cpu() for t in gp.posterior(xc.to(dev)))
    kfn = lambda a, b: R.sphere_matern(a, b, 3, 2.5, 0.6)  # noqa: E731
    mr, vr = R.gp_posterior(kfn, x.cpu(), y.cpu(), xc[:500], 1.7, 0.02, 0.3)
    assert float((m_ref[:500] - mr).abs().max()) < 1e-10 * float(mr.abs().max())
    with gp.handle(max_chunk=4096) as h:
        assert h.chunk == 4096
        for b in (1, 5, 100, 500):                       # the reference's acq(X) call sizes: latency path
            for rep in range(3):
                mh, vh = h.eval_host(xc[rep * b:(rep + 1) * b].contiguous())
                assert _same(mh, m_ref[rep * b:(rep + 1) * b]) and _same(vh, v_ref[rep * b:(rep + 1) * b])
        mh, vh = h.eval_host(xc.pin_memory())            # throughput path: 5 blocks of 4096 rows, ragged tail
        assert _same(mh, m_ref) and _same(vh, v_ref)
        md, vd = h.eval_device(xc[:777].to(dev))
        torch.cuda.synchronize()
        assert _same(md.cpu(), m_ref[:777]) and _same(vd.cpu(), v_ref[:777])
        me, ve = h.eval_host(xc[:0].contiguous())
        assert me.numel() == 0
    # host-state constructor: plain numpy in, numpy out (what a reference maintainer's binding would call)
    lib = _lib.load()
    d = gp.spec.desc(gp.coef.shape[-1])
    xt_h = np.ascontiguousarray(gp.train_prepared.cpu().numpy())
    coef_h = np.ascontiguousarray(gp.coef.cpu().numpy())
    rinv_h = np.ascontiguousarray(gp.rinv.cpu().numpy())
    alpha_h = np.ascontiguousarray(gp.alpha.cpu().numpy())
    hp = C.c_void_p()
    rc = lib.mgb_posterior_create(C.byref(d), xt_h.ctypes.data, 150, coef_h.ctypes.data, rinv_h.ctypes.data,
                                  alpha_h.ctypes.data, float(gp.kss_value), float(gp.mean_const), 0, 0, C.byref(hp))
    assert rc == 0, lib.mgb_last_error()
    xq = np.ascontiguousarray(xc[:300].numpy())
    mo, vo = np.empty(300), np.empty(300)
    for _ in range(4):
        assert lib.mgb_posterior_eval_host(hp, xq.ctypes.data, 300, mo.ctypes.data, vo.ctypes.data) == 0
    assert _same(mo, m_ref[:300]) and _same(vo, v_ref[:300])
    assert lib.mgb_posterior_destroy(hp) == 0
    assert lib.mgb_posterior_destroy(None) == 0
    assert lib.mgb_posterior_eval_host(None, xq.ctypes.data, 300, mo.ctypes.data, vo.ctypes.data) == 1


@pytest.mark.gpu
def test_persistent_handle_in_int8_mode(monkeypatch):
    """The handle's opt-in INT8 contraction (mgb_posterior_use_int8 -> csrc/int8_fused.cu): same answers as the Python
    host's INT8 path bit for bit, 1e-10 of max var from the FP64 kernel and from the oracle; switchable back; and the
    host-state constructor picks it up from MGB_POSTERIOR_PATH=int8."""
    import ctypes as C

    from materngabo_b200 import _lib
    from materngabo_b200.posterior import ExactGPPosterior
    from oracle import restated as R

    dev = "cuda:0"
    model, x, y = _model(dev, n=300, seed=9)
    gen = torch.Generator().manual_seed(5)
    xc = torch.nn.functional.normalize(torch.randn(9000, 4, generator=gen), dim=-1)
    xc[:200] = torch.nn.functional.normalize(x[:200].cpu() + 1e-3 * torch.randn(200, 4, generator=gen), dim=-1)
    ref = ExactGPPosterior.from_model(model)
    m64, v64 = (t.cpu() for t in ref.posterior(xc.to(dev)))
    gp = ExactGPPosterior.from_model(model, int8=True)
    assert gp.int8 and gp.int8_impl == "fused"
    m8, v8 = (t.cpu() for t in gp.posterior(xc.to(dev)))
    vmax, mmax = float(v64.abs().max()), float(m64.abs().max())
    assert float((v8 - v64).abs().max()) < 1e-10 * vmax and float((m8 - m64).abs().max()) < 1e-12 * mmax
    kfn = lambda a, b: R.sphere_matern(a, b, 3, 2.5, 0.6)  # noqa: E731
    mr, vr = R.gp_posterior(kfn, x.cpu(), y.cpu(), xc[:400], 1.7, 0.02, 0.3)
    assert float((v8[:400] - vr).abs().max()) < 1e-10 * float(vr.abs().max())
    assert float((m8[:400] - mr).abs().max()) < 1e-10 * float(mr.abs().max())
    lib = _lib.load()
    with gp.handle(max_chunk=4096) as h:
        assert h.int8 and lib.mgb_posterior_is_int8(h._h) == 1
        mh, vh = h.eval_host(xc[:137].contiguous())                      # latency path
        assert float((vh - v64[:137]).abs().max()) < 1e-10 * vmax and float((mh - m64[:137]).abs().max()) < 1e-12 * mmax
        mh, vh = h.eval_host(xc.pin_memory())                            # throughput path, ragged tail
        assert float((vh - v64).abs().max()) < 1e-10 * vmax and float((mh - m64).abs().max()) < 1e-12 * mmax
        assert lib.mgb_posterior_use_int8(h._h, None, None) == 0         # back to the FP64 kernel
        assert lib.mgb_posterior_is_int8(h._h) == 0
        mh, vh = h.eval_host(xc[:500].contiguous())
        assert _same(mh, m64[:500]) and _same(vh, v64[:500])
    # host-state constructor under MGB_POSTERIOR_PATH=int8
    monkeypatch.setenv("MGB_POSTERIOR_PATH", "int8")
    d = ref.spec.desc(ref.coef.shape[-1])
    xt_h = np.ascontiguousarray(ref.train_prepared.cpu().numpy())
    coef_h = np.ascontiguousarray(ref.coef.cpu().numpy())
    rinv_h = np.ascontiguousarray(ref.rinv.cpu().numpy())
    alpha_h = np.ascontiguousarray(ref.alpha.cpu().numpy())
    hp = C.c_void_p()
    rc = lib.mgb_posterior_create(C.byref(d), xt_h.ctypes.data, 300, coef_h.ctypes.data, rinv_h.ctypes.data, alpha_h.ctypes.data,
                                  float(ref.kss_value), float(ref.mean_const), 2048, 0, C.byref(hp))
    assert rc == 0, lib.mgb_last_error()
    assert lib.mgb_posterior_is_int8(hp) == 1
    xq = np.ascontiguousarray(xc[:3000].numpy())
    mo, vo = np.empty(3000), np.empty(3000)
    assert lib.mgb_posterior_eval_host(hp, xq.ctypes.data, 3000, mo.ctypes.data, vo.ctypes.data) == 0
    assert np.abs(vo - v64[:3000].numpy()).max() < 1e-10 * vmax and np.abs(mo - m64[:3000].numpy()).max() < 1e-12 * mmax
    assert lib.mgb_posterior_destroy(hp) == 0
