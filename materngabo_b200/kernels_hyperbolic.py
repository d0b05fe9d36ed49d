"""Hyperbolic kernels: drop-in for BoManifolds/kernel_utils/kernels_hyperbolic.py
(HyperbolicRiemannianMillsonGaussianKernel :18-183, HyperbolicRiemannianMillsonIntegratedMaternKernel :186-389),
dimensions 1, 2, 3, evaluated by the warp-per-entry quadrature kernel (csrc/quad_gram.cu).
The heat kernel for H^n, n > 3 (nested autograd, :118-180) is out of scope.
"""
from __future__ import annotations

import torch

from . import _ops, _weights
from .gpytorch_compat import Kernel
from .kernels_sphere import _NuMixin, _check_batch

F64 = torch.float64


def _check_lorentz(x1, x2, dim):
    if x1.shape[-1] != dim + 1 or x2.shape[-1] != dim + 1:
        raise RuntimeError("H^%d points must have %d Lorentz coordinates" % (dim, dim + 1))


def _gram_fn(dim, x1, x2, tval, tcoef, s0, ds, p):
    """Fused Gram kernel, or - when an input requires grad (acquisition gradient / Hessian-vector product of the
    trust-region inner loop) - distance in torch + device profile with its d-derivatives (``_ops.RadialProfile``)."""
    if _ops.inputs_need_grad(x1, x2):
        return lambda a, b: _ops.hyperbolic_gram_autograd(dim, a, b, tval, tcoef, s0, ds, p)
    return lambda a, b: _ops.hyperbolic_gram(dim, a, b, tval, tcoef, s0, ds, p)


class HyperbolicRiemannianMillsonGaussianKernel(Kernel):
    has_lengthscale = True

    def __init__(self, dim, nb_points_integral=1000, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        self.nb_points_integral = nb_points_integral

    def nodes(self, dev):
        """(tval, tcoef, s0, ds, Ps): exp(-s^2 / (2 kappa^2)) = exp(-s^2 / (4 t)) with t = kappa^2 / 2 and the
        normaliser of the H^2 quadrature at d = 0 (kernels_hyperbolic.py:58-88)."""
        ls = self.lengthscale.reshape(()).to(device=dev, dtype=F64)
        tval = (ls ** 2 / 2).reshape(1)          # differentiable: the device op returns dL/dt
        p = self.nb_points_integral
        s0, ds = 1e-2, (100.0 - 1e-2) / (p - 1)
        if self.dim == 2:
            norm = _ops.hyperbolic_heat_nodes(2, torch.zeros(1, dtype=F64, device=dev), tval, s0, ds, p).reshape(1)
        else:
            norm = torch.ones(1, dtype=F64, device=dev)
        return tval, 1.0 / norm, s0, ds, p

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        if self.dim not in (1, 2, 3):
            raise NotImplementedError("H^n heat kernel for n > 3 is out of scope (nested autograd in the reference)")
        _check_lorentz(x1, x2, self.dim)
        tval, tcoef, s0, ds, p = self.nodes(x1.device)
        return _ops.pairwise(_gram_fn(self.dim, x1, x2, tval, tcoef, s0, ds, p), x1, x2, diag=diag,
                             rowwise=lambda a, b: _ops.hyperbolic_gram_autograd(self.dim, a, b, tval, tcoef, s0, ds, p))


class HyperbolicRiemannianMillsonIntegratedMaternKernel(_NuMixin, Kernel):
    has_lengthscale = True

    def __init__(self, dim, nu=None, nu_prior=None, nb_points_integral=50, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._register_nu(nu, nu_prior)
        self.dim = dim
        self.nb_points_integral = nb_points_integral

    def nodes(self, device):
        """(tval, tcoef, s0, ds, Ps): t grid and normalised, differentiable t weights (kernels_hyperbolic.py:367-383)."""
        p = self.nb_points_integral
        s0, ds = 1.5e-1, (100.0 - 1.5e-1) / (p - 1)
        ls = self.lengthscale.to(device)
        if ls.is_cuda:      # one launch for the normalised weights + Jacobian; grid and zero-distance integral cached
            h0 = lambda t: _ops.hyperbolic_heat_nodes(self.dim, torch.zeros(1, dtype=F64, device=t.device), t, s0, ds, p)[0]  # noqa: E731
            tval, tcoef, _ = _weights.matern_t_coef(ls, self.nu.to(device), -2.5, 1.5, p, 1.0, ("hyp", int(self.dim), p), h0)
            return tval, tcoef, s0, ds, p
        tval, w = _weights.matern_t_nodes(ls, self.nu.to(device), -2.5, 1.5, p, power_offset=1.0)
        heat0 = _ops.hyperbolic_heat_nodes(self.dim, torch.zeros(1, dtype=F64, device=device), tval, s0, ds, p)[0]
        return tval, w / (w * heat0).sum(), s0, ds, p

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        if self.dim not in (1, 2, 3):
            raise NotImplementedError
        _check_lorentz(x1, x2, self.dim)
        tval, tcoef, s0, ds, p = self.nodes(x1.device)
        return _ops.pairwise(_gram_fn(self.dim, x1, x2, tval, tcoef, s0, ds, p), x1, x2, diag=diag,
                             rowwise=lambda a, b: _ops.hyperbolic_gram_autograd(self.dim, a, b, tval, tcoef, s0, ds, p))
