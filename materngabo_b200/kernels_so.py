"""SO(n) kernels: drop-in for BoManifolds/kernel_utils/kernels_so.py (SORiemannianGaussianKernel :28-177,
SORiemannianMaternKernel :180-362).

SO(3) runs on the sm_100a series kernel: the reference's sympy-built character sum
sum_l c_l chi_l(e^{i theta}) with chi_l = sum_{m=-l}^{l} x^m is the Chebyshev series
sum_m v_m T_m(cos theta), cos theta = clip((tr(R1 R2^T) - 1)/2, +-(1 - 1e-8)) (kernels_so.py:323-327), and
tr(R1 R2^T) is the inner product of the two flattened 9-vectors.  No sympy is needed at construction.
SO(n), n >= 4 (eigenvalues of R1 R2^T + conjugate pairing, kernels_so.py:329-346) is not implemented yet.
"""
from __future__ import annotations

import torch

from . import _ops, _weights
from .gpytorch_compat import Kernel
from .kernels_sphere import _NuMixin, _check_batch, _series_on

_SO3_CLIP = 1.0 - 1e-8  # kernels_so.py:325
QUATERNION_MIN_ENTRIES = 1 << 22  # Gram blocks at least this large go through the quaternion form (one stream sync)


class _SOSeriesKernel(Kernel):
    has_lengthscale = True

    def _coef(self):
        raise NotImplementedError

    def series(self):
        coef, basis = self._coef()
        return _ops.SeriesSpec(1, 9, basis, 0.5, -0.5, -_SO3_CLIP, _SO3_CLIP), coef

    def forward(self, x1, x2, diag=False, **params):
        # the reference ignores ``diag`` (kernels_so.py:305); the gpytorch contract is honoured here
        _check_batch(self)
        if self.dim != 3:
            raise NotImplementedError("SO(n) for n != 3 is not implemented on the CUDA path yet")
        if x1.shape[-1] != 9 or x2.shape[-1] != 9:
            raise RuntimeError("SO(3) points must be flattened 3x3 rotation matrices (9 coordinates)")
        spec, coef = _series_on(self, x1.device)
        big = (not diag and x1.dim() == 2 and x2.dim() == 2
               and x1.shape[0] * x2.shape[0] >= QUATERNION_MIN_ENTRIES
               and not _ops.inputs_need_grad(x1, x2, coef))
        if big:
            # (tr(R1 R2^T) - 1) / 2 = 2 <q1, q2>^2 - 1: rows of 4 instead of 9 for the inner products (the FP64 pipe,
            # not HBM, bounds the 9-wide form).  Only for inputs that are rotations to 1e-13; otherwise the trace form.
            quats = _ops.so3_quaternions(x1, x2)
            if quats is not None:
                spec4 = _ops.SeriesSpec(1, 4, spec.basis, 2.0, -1.0, spec.lo, spec.hi, 1)
                with torch.no_grad():
                    return _ops._series_fwd_raw(spec4, quats[0], quats[1], coef.detach().contiguous())
        return _ops.pairwise(lambda a, b: _ops.series_gram(spec, a, b, coef), x1, x2, diag=diag, series=(spec, coef))


class SORiemannianGaussianKernel(_SOSeriesKernel):
    def __init__(self, dim, sigma_squared=1, summands_number_bound=10, lambda_coeff_bound=10 ** -10, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        self.summands_number_bound = summands_number_bound

    def _coef(self):
        return _weights.so3_gaussian_coef(self.lengthscale, self.summands_number_bound)


class SORiemannianMaternKernel(_NuMixin, _SOSeriesKernel):
    def __init__(self, dim, nu=None, nu_prior=None, sigma_squared=1, summands_number_bound=10,
                 lambda_coeff_bound=10 ** -10, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._register_nu(nu, nu_prior)
        self.dim = dim
        self.summands_number_bound = summands_number_bound

    def _coef(self):
        return _weights.so3_matern_coef(self.lengthscale, self.nu, self.summands_number_bound)
