"""Hyperbolic kernels: drop-in for BoManifolds/kernel_utils/kernels_hyperbolic.py
(HyperbolicRiemannianMillsonGaussianKernel :18-183, HyperbolicRiemannianMillsonIntegratedMaternKernel :186-389),
dimensions 1, 2, 3, evaluated by the warp-per-entry quadrature kernel (csrc/quad_gram.cu).
The heat kernel for H^n, n = 4 ... 7 (:118-180: the operator (1 / sinh d) d/dd applied (n - 2) / 2 times to the H^2
profile, or (n - 3) / 2 times to the H^3 profile, through nested torch autograd) uses the d-derivatives of the device
profile kernels (csrc/quad_profile.cu, orders 0 - 2); n >= 8 would need third derivatives and raises.
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

    # ---- n = 4 ... 7 (kernels_hyperbolic.py:118-180) ----------------------------------------------------------------
    MIN_DISTANCE = 1e-6      # :124,152: distances are clamped from below before differentiating

    def _high_dim_normaliser(self, base_dim, order):
        """-D^order base(d) at d = 1e-6, D = (1 / sinh d) d/dd, evaluated EXACTLY as the reference does: torch autograd
        on a 1 x 1 float64 CPU tensor through forward_dim2 / forward_dim3 (:135-143,163-171).  The value is numerically
        ill-conditioned there (for the H^3 base the derivative of (d + 1e-8) / sinh(d + 1e-8) is a difference of two
        numbers of size 1e6: ~1e-4 relative noise), so it is reproduced operation for operation rather than evaluated
        by a stable formula - the reference's kernel is DEFINED with that number."""
        ls = self.lengthscale.detach().reshape(()).to(device="cpu", dtype=F64)
        d = torch.full((1, 1), self.MIN_DISTANCE, dtype=F64, requires_grad=True)
        if base_dim == 3:
            part = torch.exp(-torch.pow(d, 2) / (2 * ls ** 2)) * (d + 1e-8) / torch.sinh(d + 1e-8)
        else:
            base = torch.linspace(1e-2, 100.0, self.nb_points_integral, dtype=F64)
            sv = base + d.unsqueeze(-1)
            vals = sv * torch.exp(-sv ** 2 / (2 * ls ** 2)) / torch.sqrt(torch.cosh(sv) - torch.cosh(d.unsqueeze(-1)))
            nv = base * torch.exp(-base ** 2 / (2 * ls ** 2)) / torch.sqrt(torch.cosh(base) - 1.0)
            part = torch.trapz(vals, sv, dim=-1) / torch.trapz(nv, base, dim=-1)      # forward_dim2 incl. its own normaliser
        for _ in range(order):
            (grad,) = torch.autograd.grad(part, d, grad_outputs=torch.ones_like(d), create_graph=True)
            part = grad / torch.sinh(d)
        return float(-part.detach().reshape(()))

    def _forward_high_dim(self, x1, x2, diag):
        n = self.dim
        base_dim = 2 if n % 2 == 0 else 3
        order = (n - base_dim) // 2
        if order > 2:
            raise NotImplementedError("H^n heat kernel for n >= 8 needs third derivatives of the profile (not provided)")
        if _ops.inputs_need_grad(x1, x2, self.raw_lengthscale):
            raise NotImplementedError("H^n heat kernel, n > 3: forward evaluation only (no input / lengthscale gradients)")
        dev = x1.device
        ls = self.lengthscale.detach().reshape(()).to(device=dev, dtype=F64)
        tval = (ls ** 2 / 2).reshape(1)
        p = self.nb_points_integral
        s0, ds = 1e-2, (100.0 - 1e-2) / (p - 1)
        if base_dim == 2:       # forward_dim2 divides by its zero-distance integral (:75-81)
            tcoef = 1.0 / _ops.hyperbolic_heat_nodes(2, torch.zeros(1, dtype=F64, device=dev), tval, s0, ds, p).reshape(1)
        else:
            tcoef = torch.ones(1, dtype=F64, device=dev)
        norm = self._high_dim_normaliser(base_dim, order)

        def fn2d(a, b):
            dist = _ops.lorentz_distance(a, b).clamp_min(self.MIN_DISTANCE)
            prof = _ops._radial_profile_raw(base_dim - 1, 3, dist, tval, tcoef, s0, ds, p)     # orders 0, 1, 2 in one launch
            sh = torch.sinh(dist)
            if order == 1:
                num = -prof[1] / sh
            else:       # D^2 P = (P'' - P' coth d) / sinh^2 d
                num = -(prof[2] - prof[1] * torch.cosh(dist) / sh) / (sh * sh)
            return num / norm

        with torch.no_grad():
            if diag:
                if x1.shape != x2.shape:
                    raise RuntimeError("diag=True requires x1 and x2 of the same shape")
                f1, f2 = x1.reshape(-1, 1, x1.shape[-1]), x2.reshape(-1, 1, x2.shape[-1])
                return fn2d(f1, f2).reshape(x1.shape[:-1])
            return _ops.pairwise(fn2d, x1, x2)

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        _check_lorentz(x1, x2, self.dim)
        if self.dim > 3:
            return self._forward_high_dim(x1, x2, diag)
        if self.dim not in (1, 2, 3):
            raise NotImplementedError("H^n: n >= 1")
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
        if ls.is_cuda:      # one launch for the normalised weights + Jacobian
            h0 = lambda t: _ops.hyperbolic_heat_nodes(self.dim, torch.zeros(1, dtype=F64, device=t.device), t, s0, ds, p)[0]  # noqa: E731
            if _weights.device_grids_enabled():       # no host read of the lengthscale: CUDA-graph capturable
                tval, tcoef = _weights.matern_t_coef_device(ls, self.nu.to(device), -2.5, 1.5, p, 1.0, h0)
                return tval, tcoef, s0, ds, p
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
