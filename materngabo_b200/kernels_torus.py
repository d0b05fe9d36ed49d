"""Torus kernels: drop-in for BoManifolds/kernel_utils/kernels_torus.py
(TorusProductOfManifoldsRiemannianGaussianKernel :15-90, ...MaternKernel :93-177).

T^d is a product of d circles.  The parameter layout of the reference is kept (``self.torus_kernel`` is a
ProductKernel of circle kernels with ``active_dims=(2i, 2i+1)``, each with its own lengthscale / nu), but
the product is evaluated by ONE launch of the multi-factor series kernel instead of d separate Gram
matrices multiplied together.
"""
from __future__ import annotations

import torch

from . import _lib, _ops, _weights
from .gpytorch_compat import Kernel, ProductKernel
from .kernels_sphere import (CircleRiemannianGaussianKernel, CircleRiemannianIntegratedMaternKernel, _CLAMP,
                             _check_batch, _series_on)


SYNC_FREE_MAX_ENTRIES = 1 << 16      # Gram blocks below this size (GP fit) take the sync-free coefficient chain


class _TorusKernel(Kernel):
    has_lengthscale = True

    def _to_circles(self, x):
        """angles (width dim) -> interleaved (cos, sin) pairs (kernels_torus.py:162-173)."""
        if x.shape[-1] == self.dim:
            return torch.stack([torch.cos(x), torch.sin(x)], dim=-1).reshape(*x.shape[:-1], 2 * self.dim)
        return x

    def series(self, sync_free=False):
        if self.dim > _lib.MAX_FACTORS:
            raise NotImplementedError("T^d with d > %d" % _lib.MAX_FACTORS)
        kernels = list(self.torus_kernel.kernels)
        if sync_free and _weights.device_grids_enabled() and all(k.raw_lengthscale.is_cuda for k in kernels):   # raw: no softplus launch
            # GP-fit sizes: every factor's 201 Chebyshev coefficients straight from the device (no truncation, no basis
            # decision, no lengthscale.item()): the marginal-likelihood step is then CUDA-graph capturable
            rows = [_weights.circle_coef_sync_free(k._coef_mode, k.lengthscale, getattr(k, "nu", None) if k._coef_mode == 0 else None,
                                                   getattr(k, "nb_points_integral", 0)) for k in kernels]
            return _ops.SeriesSpec(self.dim, 2, _lib.BASIS_CHEBYSHEV, 1.0, 0.0, -_CLAMP, _CLAMP), torch.stack(rows).contiguous()
        coef = _weights.stack_factor_coefs([k._coef() for k in kernels])
        coef, basis = _weights.chebyshev_or_monomial(coef)
        return _ops.SeriesSpec(self.dim, 2, basis, 1.0, 0.0, -_CLAMP, _CLAMP), coef

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        x1c = self._to_circles(x1)
        x2c = x1c if x2 is x1 else self._to_circles(x2)          # K(X, X) of the fit: one conversion
        if x1c.shape[-1] != 2 * self.dim or x2c.shape[-1] != 2 * self.dim:
            raise RuntimeError("T^%d points must be %d angles or %d circle coordinates" % (self.dim, self.dim, 2 * self.dim))
        fit_sized = (x1c.dim() == 2 and x2c.dim() == 2 and x1c.shape[0] * x2c.shape[0] < SYNC_FREE_MAX_ENTRIES
                     and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
                     and not (x1c.requires_grad or x2c.requires_grad))
        if fit_sized:
            spec, coef = self.series(sync_free=True)
            coef = coef.to(x1c.device)
        else:
            spec, coef = _series_on(self, x1c.device)
        if torch.is_grad_enabled() and (x1c.requires_grad or x2c.requires_grad):
            # Input gradients (and the Hessian-vector products of the trust-region solver, bo_torus.py:278): evaluate
            # the product factor by factor - every single-factor op is twice differentiable with respect to its inputs.
            one = _ops.SeriesSpec(1, 2, spec.basis, spec.scale, spec.shift, spec.lo, spec.hi)

            def fn(a, b):
                out = None
                for f in range(self.dim):
                    kf = _ops.series_gram(one, a[:, 2 * f:2 * f + 2], b[:, 2 * f:2 * f + 2], coef[f:f + 1])
                    out = kf if out is None else out * kf
                return out
            return _ops.pairwise(fn, x1c, x2c, diag=diag)
        return _ops.pairwise(lambda a, b: _ops.series_gram(spec, a, b, coef), x1c, x2c, diag=diag, series=(spec, coef))


class TorusProductOfManifoldsRiemannianGaussianKernel(_TorusKernel):
    def __init__(self, dim, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        kernels = [CircleRiemannianGaussianKernel(active_dims=torch.tensor(list(range(2 * i, 2 * i + 2))))
                   for i in range(self.dim)]
        self.torus_kernel = ProductKernel(*kernels)


class TorusProductOfManifoldsRiemannianMaternKernel(_TorusKernel):
    def __init__(self, dim, nu=None, nu_prior=None, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        if not isinstance(nu, (list, tuple)):
            nu = [nu for _ in range(self.dim)]
        if not isinstance(nu_prior, (list, tuple)):
            nu_prior = [nu_prior for _ in range(self.dim)]
        kernels = [CircleRiemannianIntegratedMaternKernel(nu=nu[i], nu_prior=nu_prior[i],
                                                          active_dims=torch.tensor(list(range(2 * i, 2 * i + 2))))
                   for i in range(self.dim)]
        self.torus_kernel = ProductKernel(*kernels)
