"""SPD kernels: drop-in for the intrinsic SPD(2) kernels of BoManifolds/kernel_utils/kernels_spd.py
(SpdRiemannianGaussianKernel :24-136, SpdRiemannianIntegratedMaternKernel :139-289), evaluated by the
warp-per-entry quadrature kernel (csrc/quad_gram.cu).  Inputs are Mandel 3-vectors (s11, s22, sqrt2 s12).
As in the reference, dim != 2 raises NotImplementedError.  The product kernels (R x S, R x SO) and the
approximate / log-Euclidean kernels of the reference file are not part of this round.
"""
from __future__ import annotations

import torch

from . import _ops, _weights
from .gpytorch_compat import Kernel
from .kernels_sphere import _NuMixin, _check_batch

F64 = torch.float64


def _b_nodes(nb, device):
    b = _weights._logspace(-5.0, 1.0, nb, device)      # kernels_spd.py:118,265
    return b, _weights.trapz_weights(b)


class SpdRiemannianGaussianKernel(Kernel):
    has_lengthscale = True

    def __init__(self, dim, nb_points_integral=50, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        if self.dim != 2:
            raise NotImplementedError
        self.nb_points_integral = nb_points_integral

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        dev = x1.device
        ls = self.lengthscale.reshape(()).to(device=dev, dtype=F64).detach()
        b, bw = _b_nodes(self.nb_points_integral, dev)
        rscale = (1.0 / (2 * ls ** 2)).reshape(1)
        bscale = (1.0 / ls ** 2).reshape(1)
        # normaliser: the b-integral at Delta = 0 (kernels_spd.py:128-134)
        norm = (bw * 2.0 * b * torch.exp(-b * b * bscale) / torch.sinh(b)).sum()
        tcoef = (1.0 / norm).reshape(1)
        fn = lambda a, c: _ops.spd2_gram(1, a, c, tcoef, rscale, bscale, b, bw)  # noqa: E731
        return _ops.pairwise(fn, x1, x2, diag=diag)


class SpdRiemannianIntegratedMaternKernel(_NuMixin, Kernel):
    has_lengthscale = True

    def __init__(self, dim, nu=None, nu_prior=None, nb_points_integral_b=50, nb_points_integral_t=50, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._register_nu(nu, nu_prior)
        self.dim = dim
        if self.dim != 2:
            raise NotImplementedError
        self.nb_points_integral_b = nb_points_integral_b
        self.nb_points_integral_t = nb_points_integral_t

    def nodes(self, device):
        ls = self.lengthscale.to(device)
        t, w = _weights.matern_t_nodes(ls, self.nu.to(device), -2.0, 1.5, self.nb_points_integral_t, power_offset=1.0)
        b, bw = _b_nodes(self.nb_points_integral_b, device)
        rscale, bscale = 1.0 / (4 * t), 1.0 / (2 * t)
        inner0 = (bw * 2.0 * b / torch.sinh(b) * torch.exp(-(b * b).unsqueeze(0) * bscale.unsqueeze(-1))).sum(-1)
        tcoef = w / (w * inner0).sum()                   # kernels_spd.py:277-289
        return tcoef, rscale, bscale, b, bw

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        tcoef, rscale, bscale, b, bw = self.nodes(x1.device)
        fn = lambda a, c: _ops.spd2_gram(0, a, c, tcoef, rscale, bscale, b, bw)  # noqa: E731
        return _ops.pairwise(fn, x1, x2, diag=diag)
