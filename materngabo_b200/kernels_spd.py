"""SPD kernels: drop-in for the intrinsic SPD(2) kernels of BoManifolds/kernel_utils/kernels_spd.py
(SpdRiemannianGaussianKernel :24-136, SpdRiemannianIntegratedMaternKernel :139-289), evaluated by the
warp-per-entry quadrature kernel (csrc/quad_gram.cu).  Inputs are Mandel 3-vectors (s11, s22, sqrt2 s12).
As in the reference, dim != 2 raises NotImplementedError.

The product kernels (SpdProductOfRSManifolds... :292-549, SpdProductOfRSOManifolds... :552-774) are compositions
of the Euclidean factor with a sphere / circle / SO(3) factor through ProductKernel, exactly as in the reference;
they accept the pre-decomposed inputs (``spd_input=False``: eigenvalues followed by the sphere point or the
flattened rotation).  ``spd_input=True`` needs the eigen-decomposition front-end built on the removed
``torch.symeig`` (Riemannian_utils/spd_utils_torch.py:293-333) and raises NotImplementedError.  The approximate /
log-Euclidean kernels are out of scope.
"""
from __future__ import annotations

import math
from fractions import Fraction

import torch

from . import _ops, _weights
from .gpytorch_compat import Kernel, RBFKernel
from .kernels_euclidean import EuclideanIntegratedMaternKernel
from .kernels_so import SORiemannianGaussianKernel, SORiemannianMaternKernel
from .kernels_sphere import (CircleRiemannianGaussianKernel, CircleRiemannianIntegratedMaternKernel,
                             SphereRiemannianGaussianKernel, SphereRiemannianMaternKernel, _NuMixin, _check_batch)

F64 = torch.float64


def _b_nodes(nb, device):
    """b grid logspace(-5, 1, nb) (kernels_spd.py:118,265) and its trapezoid weights, cached per device."""
    key = ("spd_b_nodes", int(nb), str(device), str(torch.get_default_dtype()))
    hit = _weights._CONST_CACHE.get(key)
    if hit is None:
        b = _weights._logspace(-5.0, 1.0, nb, device)
        hit = (None, (b, _weights.trapz_weights(b)))
        _weights._CONST_CACHE[key] = hit
    return hit[1]


SYNC_FREE_MAX_ENTRIES = 1 << 16      # Gram blocks below this size take the sync-free node chain (direct forward kernel)


class SpdRiemannianGaussianKernel(Kernel):
    has_lengthscale = True

    def __init__(self, dim, nb_points_integral=50, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        if self.dim != 2:
            raise NotImplementedError
        self.nb_points_integral = nb_points_integral

    use_cholesky = 1      # singular values of chol(X1) chol(X2)^-1 (kernels_spd.py:90-102)

    def nodes(self, dev):
        """(tcoef, rscale, bscale, b, bw) of the heat kernel: one outer node with scales 1/(2 kappa^2), 1/kappa^2."""
        ls = self.lengthscale.reshape(()).to(device=dev, dtype=F64)
        b, bw = _b_nodes(self.nb_points_integral, dev)
        rscale = (1.0 / (2 * ls ** 2)).reshape(1)
        bscale = (1.0 / ls ** 2).reshape(1)
        # normaliser: the b-integral at Delta = 0 (kernels_spd.py:128-134)
        norm = (bw * 2.0 * b * torch.exp(-b * b * bscale) / torch.sinh(b)).sum()
        return (1.0 / norm).reshape(1), rscale, bscale, b, bw

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        tcoef, rscale, bscale, b, bw = self.nodes(x1.device)
        op = _ops.spd2_gram_autograd if _ops.inputs_need_grad(x1, x2) else _ops.spd2_gram
        fn = lambda a, c: op(1, a, c, tcoef, rscale, bscale, b, bw)  # noqa: E731
        return _ops.pairwise(fn, x1, x2, diag=diag,
                             rowwise=lambda a, c: _ops.spd2_gram_autograd(1, a, c, tcoef, rscale, bscale, b, bw))


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

    use_cholesky = 0      # singular values of X1 X2^-1 (kernels_spd.py:230-243)

    def nodes(self, device, with_lattice=False, sync_free=False):
        ls = self.lengthscale.to(device)
        b, bw = _b_nodes(self.nb_points_integral_b, device)
        if ls.is_cuda and sync_free and _weights.device_grids_enabled():
            # launch-bound sizes (GP fit): no host read of the lengthscale - CUDA-graph capturable; the lattice form of
            # the forward kernel needs g_min(kappa) on the host and is left to the large blocks
            inner = lambda t: (bw * 2.0 * b / torch.sinh(b) * torch.exp(-(b * b).unsqueeze(0) / (2 * t).unsqueeze(-1))).sum(-1)  # noqa: E731
            t, tcoef = _weights.matern_t_coef_device(ls, self.nu.to(device), -2.0, 1.5, self.nb_points_integral_t, 1.0, inner)
            out = (tcoef, 1.0 / (4 * t), 1.0 / (2 * t), b, bw)
            return out + (None,) if with_lattice else out
        if ls.is_cuda:      # one launch for the normalised weights + Jacobian; grid and zero-distance integral cached
            inner = lambda t: (bw * 2.0 * b / torch.sinh(b) * torch.exp(-(b * b).unsqueeze(0) / (2 * t).unsqueeze(-1))).sum(-1)  # noqa: E731
            t, tcoef, shift = _weights.matern_t_coef(ls, self.nu.to(device), -2.0, 1.5, self.nb_points_integral_t, 1.0,
                                                     ("spd2", self.nb_points_integral_b), inner)
            key = ("spd2_scales", float(shift), self.nb_points_integral_t, str(ls.device), str(torch.get_default_dtype()))
            hit = _weights._CONST_CACHE.get(key)
            if hit is None:
                hit = (None, (1.0 / (4 * t), 1.0 / (2 * t)))
                _weights._CONST_CACHE[key] = hit
            rscale, bscale = hit[1]
            if with_lattice:
                return tcoef, rscale, bscale, b, bw, self._lattice(shift)
            return tcoef, rscale, bscale, b, bw
        t, w, shift = _weights.matern_t_nodes(ls, self.nu.to(device), -2.0, 1.5, self.nb_points_integral_t,
                                              power_offset=1.0, return_shift=True)
        rscale, bscale = 1.0 / (4 * t), 1.0 / (2 * t)
        inner0 = (bw * 2.0 * b / torch.sinh(b) * torch.exp(-(b * b).unsqueeze(0) * bscale.unsqueeze(-1))).sum(-1)
        tcoef = w / (w * inner0).sum()                   # kernels_spd.py:277-289
        if with_lattice:
            return tcoef, rscale, bscale, b, bw, self._lattice(shift)
        return tcoef, rscale, bscale, b, bw

    def _lattice(self, shift):
        """Both quadrature grids are geometric (kernels_spd.py:262-265): log b steps by 6 ln10 / (Pb - 1), log(1 / 2t) by
        -3.5 ln10 / (Pt - 1), so b_b / (2 t_a) = g_min tau^(sb b + sa (Pt - 1 - a)) with sb : sa = 12 (Pt - 1) : 7 (Pb - 1).
        Returns (sb, sa, g_min, log tau) for mgb_spd2_gram_fwd_lattice, or None when the lattice would be too long."""
        pb, pt = self.nb_points_integral_b, self.nb_points_integral_t
        if pt < 2 or pt > 64 or pb < 2:
            return None
        fr = Fraction(12 * (pt - 1), 7 * (pb - 1))
        sb, sa = fr.numerator, fr.denominator
        nk = sb * (pb - 1) + sa * (pt - 1) + 1
        if nk > pb * pt // 2 or 8 * (pb * pt + nk + 4 * pb + 12 * (nk + pb)) > 200 * 1024:
            return None
        log_tau = 6.0 * math.log(10.0) / (pb - 1) / sb
        g_min = 1e-5 * 0.5 * 10.0 ** (-(1.5 + shift))     # b_0 / (2 t_max)
        return sb, sa, g_min, log_tau

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        small = x1.shape[-2] * x2.shape[-2] < SYNC_FREE_MAX_ENTRIES
        tcoef, rscale, bscale, b, bw, lattice = self.nodes(x1.device, with_lattice=True, sync_free=small)
        if _ops.inputs_need_grad(x1, x2):
            fn = lambda a, c: _ops.spd2_gram_autograd(0, a, c, tcoef, rscale, bscale, b, bw)  # noqa: E731
        else:
            fn = lambda a, c: _ops.spd2_gram(0, a, c, tcoef, rscale, bscale, b, bw, lattice)  # noqa: E731
        return _ops.pairwise(fn, x1, x2, diag=diag,
                             rowwise=lambda a, c: _ops.spd2_gram_autograd(0, a, c, tcoef, rscale, bscale, b, bw))


class _SpdProductKernel(Kernel):
    has_lengthscale = True

    decompose_target = 1      # 1: [eigenvalues | sphere point] (R x S kernels); 0: [eigenvalues | rotation] (R x SO)

    def forward(self, x1, x2, diag=False, **params):
        if self.spd_input:
            # Mandel vectors -> eigenvalues + rotation / sphere coordinates on the device (kernels_spd.py:520-533,746-757;
            # the reference's torch.symeig is gone, the solver and its sign convention are csrc/spd_eigh.cu's)
            x1 = _ops.spd_decompose(x1, self.dim, self.decompose_target)
            x2 = _ops.spd_decompose(x2, self.dim, self.decompose_target)
        return self.spd_kernel.forward(x1, x2, diag=diag)


def _sphere_dim(dim):
    if dim == 2:
        return 1
    if dim == 3:
        return 3
    raise NotImplementedError


class SpdProductOfRSManifoldsRiemannianGaussianKernel(_SpdProductKernel):
    def __init__(self, dim, spd_input=False, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        self.dim_sphere = _sphere_dim(dim)
        self.spd_input = spd_input
        self.Euclidean_kernel = RBFKernel(active_dims=torch.tensor(list(range(0, self.dim))))
        sdims = torch.tensor(list(range(self.dim, self.dim + self.dim_sphere + 1)))
        if self.dim_sphere == 1:
            self.sphere_kernel = CircleRiemannianGaussianKernel(active_dims=sdims)
        else:
            self.sphere_kernel = SphereRiemannianGaussianKernel(dim=self.dim_sphere, active_dims=sdims)
        self.spd_kernel = self.Euclidean_kernel * self.sphere_kernel


class SpdProductOfRSManifoldsRiemannianMaternKernel(_SpdProductKernel):
    def __init__(self, dim, nu=None, nu_sphere=None, nu_prior_sphere=None, nu_euclidean=None, nu_prior_euclidean=None,
                 spd_input=False, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        self.dim_sphere = _sphere_dim(dim)
        if nu:
            nu_euclidean = nu
            nu_sphere = nu
        self.spd_input = spd_input
        self.Euclidean_kernel = EuclideanIntegratedMaternKernel(self.dim, nu=nu_euclidean, nu_prior=nu_prior_euclidean,
                                                                active_dims=torch.tensor(list(range(0, self.dim))))
        sdims = torch.tensor(list(range(self.dim, self.dim + self.dim_sphere + 1)))
        if self.dim_sphere == 1:
            self.sphere_kernel = CircleRiemannianIntegratedMaternKernel(nu=nu_sphere, nu_prior=nu_prior_sphere,
                                                                        active_dims=sdims)
        else:
            self.sphere_kernel = SphereRiemannianMaternKernel(dim=self.dim_sphere, nu=nu_sphere, nu_prior=nu_prior_sphere,
                                                              active_dims=sdims)
        self.spd_kernel = self.Euclidean_kernel * self.sphere_kernel


class SpdProductOfRSOManifoldsRiemannianGaussianKernel(_SpdProductKernel):
    decompose_target = 0

    def __init__(self, dim, spd_input=False, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        if self.dim > 6:
            raise NotImplementedError
        self.spd_input = spd_input
        self.Euclidean_kernel = RBFKernel(active_dims=torch.tensor(list(range(0, self.dim))))
        self.so_kernel = SORiemannianGaussianKernel(dim=self.dim,
                                                    active_dims=torch.tensor(list(range(self.dim, self.dim + self.dim ** 2))))
        self.spd_kernel = self.Euclidean_kernel * self.so_kernel


class SpdProductOfRSOManifoldsRiemannianMaternKernel(_SpdProductKernel):
    decompose_target = 0

    def __init__(self, dim, nu=None, nu_so=None, nu_prior_so=None, nu_euclidean=None, nu_prior_euclidean=None,
                 spd_input=False, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        if self.dim > 6:
            raise NotImplementedError
        if nu:
            nu_euclidean = nu
            nu_so = nu
        self.spd_input = spd_input
        self.Euclidean_kernel = EuclideanIntegratedMaternKernel(self.dim, nu=nu_euclidean, nu_prior=nu_prior_euclidean,
                                                                active_dims=torch.tensor(list(range(0, self.dim))))
        self.so_kernel = SORiemannianMaternKernel(dim=self.dim, nu=nu_so, nu_prior=nu_prior_so,
                                                  active_dims=torch.tensor(list(range(self.dim, self.dim + self.dim ** 2))))
        self.spd_kernel = self.Euclidean_kernel * self.so_kernel
