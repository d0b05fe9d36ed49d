"""Sphere / circle kernels: drop-in for BoManifolds/kernel_utils/kernels_sphere.py (same class names,
constructor signatures, parameters and ``forward(x1, x2, diag=False, **params)`` contract), evaluated by the
sm_100a series kernel (csrc/series_gram.cu) instead of a chain of torch ops.

Reference classes and lines:
  SphereRiemannianMaternKernel            kernels_sphere.py:24-139
  SphereRiemannianGaussianKernel          kernels_sphere.py:142-224
  SphereRiemannianIntegratedMaternKernel  kernels_sphere.py:227-384
  CircleRiemannianGaussianKernel          kernels_sphere.py:409-472
  CircleRiemannianIntegratedMaternKernel  kernels_sphere.py:475-614
  SphereApproximatedGaussianKernel        kernels_sphere.py:617-722   (geodesic RBF, not PD in general)
  SphereRiemannianLaplaceKernel           kernels_sphere.py:725-762
"""
from __future__ import annotations

import torch

from . import _lib, _ops, _weights
from .gpytorch_compat import GreaterThan, Kernel, Positive

_CLAMP = 1.0 - 1e-15  # Riemannian_utils/sphere_utils_torch.py:57


class _NuMixin:
    """raw_nu parameter with a Positive (softplus) constraint; fixed when ``nu`` is given (kernels_sphere.py:62-75)."""

    def _register_nu(self, nu, nu_prior):
        self.register_parameter(name="raw_nu", parameter=torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, 1)))
        if nu_prior is not None:
            self.register_prior("nu_prior", nu_prior, lambda module: module.nu,
                                lambda module, value: module._set_nu(value))
        self.register_constraint("raw_nu", Positive())
        if nu is not None:
            self.nu = nu
            self.raw_nu.requires_grad = False

    @property
    def nu(self):
        return self.raw_nu_constraint.transform(self.raw_nu)

    @nu.setter
    def nu(self, value):
        self._set_nu(value)

    def _set_nu(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_nu)
        self.initialize(raw_nu=self.raw_nu_constraint.inverse_transform(value))


def _series_on(kernel, device):
    """(spec, coef on ``device``).  When no hyper-parameter needs a gradient the pair is cached against the
    parameters' version counters, so repeated posterior / acquisition calls with fixed hyper-parameters do not
    recompute the weights on the host nor pay a synchronising host-to-device copy per call."""
    params = list(kernel.parameters())
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        spec, coef = kernel.series()
        return spec, coef.to(device)
    key = (str(device), torch.get_default_dtype(),) + tuple((p._version, p.data_ptr()) for p in params)
    cached = kernel.__dict__.get("_series_cache")
    if cached is not None and cached[0] == key:
        return cached[1]
    with torch.no_grad():
        spec, coef = kernel.series()
        val = (spec, coef.to(device).contiguous())
    kernel.__dict__["_series_cache"] = (key, val)
    return val


def _check_batch(kernel):
    if len(kernel.batch_shape) != 0:
        raise NotImplementedError("batched hyper-parameters (batch_shape != []) are not supported by the CUDA path")


class _SphereSeriesKernel(Kernel):
    """Common forward: coefficients on the host (autograd), series on the device."""

    has_lengthscale = True

    def _init_tables(self, dim, serie_nb_terms):
        self.dim = dim
        self.serie_nb_terms = serie_nb_terms
        # Python lists of tensors in the default dtype, as in the reference (kernels_sphere.py:84,87)
        self.cst_nd, self.zero_gpolynomials = _weights.sphere_constants(dim, serie_nb_terms)

    def _coef(self):
        raise NotImplementedError

    def series(self):
        """(spec, coef) currently defined by the hyper-parameters: what the posterior engine consumes."""
        coef, basis = self._coef()
        spec = _ops.SeriesSpec(1, self.dim + 1, basis, 1.0, 0.0, -_CLAMP, _CLAMP)
        return spec, coef

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        if x1.shape[-1] != self.dim + 1 or x2.shape[-1] != self.dim + 1:
            raise RuntimeError("S^%d points must have %d coordinates" % (self.dim, self.dim + 1))
        spec, coef = _series_on(self, x1.device)
        return _ops.pairwise(lambda a, b: _ops.series_gram(spec, a, b, coef), x1, x2, diag=diag, series=(spec, coef))


class SphereRiemannianMaternKernel(_NuMixin, _SphereSeriesKernel):
    def __init__(self, dim, nu=None, serie_nb_terms=10, nu_prior=None, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._register_nu(nu, nu_prior)
        self._init_tables(dim, serie_nb_terms)

    def _coef(self):
        return _weights.sphere_matern_coef(self.lengthscale, self.nu, self.dim, self.cst_nd, self.zero_gpolynomials)


class SphereRiemannianGaussianKernel(_SphereSeriesKernel):
    def __init__(self, dim, serie_nb_terms=10, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._init_tables(dim, serie_nb_terms)

    def _coef(self):
        return _weights.sphere_gaussian_coef(self.lengthscale, self.dim, self.cst_nd, self.zero_gpolynomials)


class SphereRiemannianIntegratedMaternKernel(_NuMixin, _SphereSeriesKernel):
    def __init__(self, dim, nu=None, nu_prior=None, serie_nb_terms=10, nb_points_integral=100, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._register_nu(nu, nu_prior)
        self.nb_points_integral = nb_points_integral
        self._init_tables(dim, serie_nb_terms)

    def _coef(self):
        return _weights.sphere_integrated_matern_coef(self.lengthscale, self.nu, self.dim, self.cst_nd,
                                                      self.zero_gpolynomials, self.nb_points_integral)


def compute_riemannian_matern_kernel_constant(n, d):
    """c_{n,d} of the S^d Matern series (kernels_sphere.py:387-406), 1-element tensor in the default dtype."""
    cst, _ = _weights.sphere_constants(d, n + 1)
    return cst[n]


class _CircleSeriesKernel(Kernel):
    has_lengthscale = True
    dim = 1

    def _coef(self):
        raise NotImplementedError

    def series(self):
        coef, basis = _weights.chebyshev_or_monomial(self._coef().reshape(1, -1).contiguous())
        return _ops.SeriesSpec(1, 2, basis, 1.0, 0.0, -_CLAMP, _CLAMP), coef

    def forward(self, x1, x2, diag=False, **params):
        _check_batch(self)
        if x1.shape[-1] != 2 or x2.shape[-1] != 2:
            raise RuntimeError("S^1 points must have 2 coordinates")
        spec, coef = _series_on(self, x1.device)
        return _ops.pairwise(lambda a, b: _ops.series_gram(spec, a, b, coef), x1, x2, diag=diag, series=(spec, coef))


class CircleRiemannianGaussianKernel(_CircleSeriesKernel):
    def __init__(self, serie_nb_terms=100, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.serie_nb_terms = serie_nb_terms  # unused, as in the reference (theta_3 always sums 200 terms)

    _coef_mode = 1        # mgb_circle_coef mode: heat

    def _coef(self):
        return _weights.circle_gaussian_coef(self.lengthscale)


class CircleRiemannianIntegratedMaternKernel(_NuMixin, _CircleSeriesKernel):
    def __init__(self, nu=None, nu_prior=None, serie_nb_terms=20, nb_points_integral=50, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = 1
        self._register_nu(nu, nu_prior)
        self.nb_points_integral = nb_points_integral
        self.serie_nb_terms = serie_nb_terms

    _coef_mode = 0        # mgb_circle_coef mode: integrated Matern

    def _coef(self):
        return _weights.circle_integrated_matern_coef(self.lengthscale, self.nu, self.nb_points_integral)


def _sphere_distance(x1, x2, diag=False):
    """Geodesic distance on the device from the series kernel with p(u) = u (T = 2)."""
    d = x1.shape[-1]
    spec = _ops.SeriesSpec(1, d, _lib.BASIS_MONOMIAL, 1.0, 0.0, -_CLAMP, _CLAMP)
    coef = torch.tensor([[0.0, 1.0]], dtype=torch.float64, device=x1.device)
    u = _ops.pairwise(lambda a, b: _ops.series_gram(spec, a, b, coef), x1, x2, diag=diag, series=(spec, coef))
    return torch.acos(u)


class SphereApproximatedGaussianKernel(Kernel):
    """exp(-beta d^2) with the geodesic distance (kernels_sphere.py:617-722)."""

    def __init__(self, dim, beta_min=None, beta_prior=None, **kwargs):
        super().__init__(has_lengthscale=False, **kwargs)
        if beta_min is None:
            for bound, val in ((2, 6.5), (3, 2.), (4, 1.2), (5, 1.0), (10, 0.6), (15, 0.42), (50, 0.35), (160, 0.21)):
                if dim <= bound:
                    beta_min = val
                    break
        self.dim = dim
        self.beta_min = beta_min
        self.register_parameter(name="raw_beta", parameter=torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, 1)))
        if beta_prior is not None:
            self.register_prior("beta_prior", beta_prior, lambda m: m.beta, lambda m, v: m._set_beta(v))
        self.register_constraint("raw_beta", GreaterThan(self.beta_min))

    @property
    def beta(self):
        return self.raw_beta_constraint.transform(self.raw_beta)

    @beta.setter
    def beta(self, value):
        self._set_beta(value)

    def _set_beta(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_beta)
        self.initialize(raw_beta=self.raw_beta_constraint.inverse_transform(value))

    def forward(self, x1, x2, diag=False, **params):
        distance = _sphere_distance(x1, x2, diag=diag)
        return torch.exp(-(distance * distance).mul(self.beta.double().reshape(())))


class SphereRiemannianLaplaceKernel(Kernel):
    """exp(-d / kappa^2) (kernels_sphere.py:725-762)."""

    def __init__(self, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)

    def forward(self, x1, x2, diag=False, **params):
        distance = _sphere_distance(x1, x2, diag=diag)
        ls = self.lengthscale.double().reshape(())
        return torch.exp(-distance.div(ls * ls))
