"""Euclidean kernels: drop-in for BoManifolds/kernel_utils/kernels_euclidean.py (EuclideanHeatKernel :20-75,
link_function :82-89, EuclideanIntegratedMaternKernel :92-207) - the Euclidean factor of the SPD product kernels."""
from __future__ import annotations

import numpy as np
import torch

from . import _ops, _weights
from .gpytorch_compat import Kernel
from .kernels_sphere import _NuMixin, _check_batch

F64 = torch.float64


class EuclideanHeatKernel(Kernel):
    """exp(-|x - y|^2 / (4 t)) / (4 pi t)^(d/2) at a given time t (kernels_euclidean.py:46-75)."""

    def __init__(self, **kwargs):
        self.has_lengthscale = False
        super().__init__(has_lengthscale=False, ard_num_dims=None, **kwargs)

    def forward(self, x1, x2, t, diag=False, **params):
        dim = x1.shape[1]
        tval = torch.as_tensor(t, dtype=F64, device=x1.device).reshape(1)
        tcoef = torch.pow(4 * np.pi * tval, -dim / 2)
        fn = _ops.euclidean_gram_autograd if _ops.inputs_need_grad(x1, x2) else _ops.euclidean_gram
        return _ops.pairwise(lambda a, b: fn(a, b, tval, tcoef), x1, x2, diag=diag,
                             rowwise=lambda a, b: _ops.euclidean_gram_autograd(a, b, tval, tcoef))


def unnormalized_heat_kernel(dist_sq, lengthscale):
    return torch.exp(-dist_sq / (2 * torch.pow(lengthscale, 2)))


def link_function(dist_sq, t, lengthscale, smoothness, dimension):
    """t^(nu-1) exp(-2 nu t / l^2) exp(-d^2 / (4 t)) (kernels_euclidean.py:82-89); host helper."""
    return torch.pow(t, smoothness - 1.0) * torch.exp(-2.0 * smoothness / lengthscale ** 2 * t) \
        * unnormalized_heat_kernel(dist_sq, torch.pow(2 * t, 0.5))


class EuclideanIntegratedMaternKernel(_NuMixin, Kernel):
    has_lengthscale = True

    def __init__(self, dim, nu=None, nu_prior=None, nb_points_integral=100, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._register_nu(nu, nu_prior)
        self.dim = dim
        self.nb_points_integral = nb_points_integral

    def nodes(self, device):
        """t grid logspace(-3 + log10 l, 3 + log10 l, P) and the normalised trapezoid weights (:186-207)."""
        ls = self.lengthscale.to(device)
        if ls.is_cuda:
            if _weights.device_grids_enabled():       # no host read of the lengthscale: CUDA-graph capturable
                return _weights.matern_t_coef_device(ls, self.nu.to(device), -3.0, 3.0, self.nb_points_integral, 1.0, None)
            t, tcoef, _ = _weights.matern_t_coef(ls, self.nu.to(device), -3.0, 3.0, self.nb_points_integral, 1.0, "euclid", None)
            return t, tcoef
        t, w = _weights.matern_t_nodes(ls, self.nu.to(device), -3.0, 3.0, self.nb_points_integral, power_offset=1.0)
        return t, w / w.sum()

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        tval, tcoef = self.nodes(x1.device)
        fn = _ops.euclidean_gram_autograd if _ops.inputs_need_grad(x1, x2) else _ops.euclidean_gram
        return _ops.pairwise(lambda a, b: fn(a, b, tval, tcoef), x1, x2, diag=diag,
                             rowwise=lambda a, b: _ops.euclidean_gram_autograd(a, b, tval, tcoef))
