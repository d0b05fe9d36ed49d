"""Host-side stand-in for the slice of gpytorch the manifold kernels sit behind.

The reference kernels are ``gpytorch.kernels.Kernel`` subclasses (e.g.
/root/reference BoManifolds/kernel_utils/kernels_sphere.py:24, :61-75). gpytorch is not part of
this image, on the dev container or on the GPU box, so the drop-in classes of this package
derive from ``Kernel`` below when ``import gpytorch`` fails and from the real class when it
succeeds (``HAVE_GPYTORCH``).  Only the contract the hot path touches is provided:

* ``register_parameter / register_constraint / register_prior / initialize`` with softplus
  (``Positive``) and shifted-softplus (``GreaterThan``) constraints,
* ``raw_lengthscale`` of shape ``(*batch_shape, 1, 1)`` + ``lengthscale`` property/setter,
* ``Kernel.__call__`` applying ``active_dims`` and defaulting ``x2 = x1`` (dense result, no
  lazy tensors), ``covar_dist``,
* ``ScaleKernel`` and ``ProductKernel``.

This is plumbing, not compute: nothing here evaluates a manifold kernel.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn.functional import softplus

try:  # pragma: no cover - exercised only where gpytorch exists
    import gpytorch as _real_gpytorch  # noqa: F401

    HAVE_GPYTORCH = True
except Exception:  # ModuleNotFoundError in this image
    _real_gpytorch = None
    HAVE_GPYTORCH = False


def _inv_softplus(y: torch.Tensor) -> torch.Tensor:
    return y + torch.log(-torch.expm1(-y))


class Interval(nn.Module):
    """Generic constraint object: ``transform`` maps the raw parameter to its constrained value."""

    def __init__(self, lower_bound=-math.inf, upper_bound=math.inf):
        super().__init__()
        self.lower_bound = torch.as_tensor(float(lower_bound))
        self.upper_bound = torch.as_tensor(float(upper_bound))

    def transform(self, raw):
        raise NotImplementedError

    def inverse_transform(self, value):
        raise NotImplementedError


class GreaterThan(Interval):
    def __init__(self, lower_bound, **_):
        super().__init__(lower_bound=lower_bound)

    def transform(self, raw):
        return softplus(raw) + self.lower_bound.to(raw)

    def inverse_transform(self, value):
        return _inv_softplus(value - self.lower_bound.to(value))


class Positive(GreaterThan):
    def __init__(self, **_):
        super().__init__(lower_bound=0.0)

    def transform(self, raw):
        return softplus(raw)

    def inverse_transform(self, value):
        return _inv_softplus(value)


class Module(nn.Module):
    """gpytorch.Module subset: named constraints / priors and ``initialize``."""

    def __init__(self):
        super().__init__()
        self._priors = {}
        self._constraint_names = {}

    def register_parameter(self, name, parameter):  # same signature as gpytorch.Module
        super().register_parameter(name, parameter)

    def register_constraint(self, param_name, constraint, replace=True):
        if param_name not in self._parameters:
            raise RuntimeError("Attempting to register constraint for nonexistent parameter.")
        cname = param_name + "_constraint"
        if cname in self._modules and not replace:
            return
        self.add_module(cname, constraint)
        self._constraint_names[param_name] = cname

    def constraint_for_parameter_name(self, param_name):
        return getattr(self, param_name + "_constraint", None)

    def register_prior(self, name, prior, param_or_closure, setting_closure=None):
        self._priors[name] = (prior, param_or_closure, setting_closure)

    def named_priors(self):
        for name, (prior, closure, setter) in self._priors.items():
            yield name, self, prior, closure, setter

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if not hasattr(self, name):
                raise AttributeError("Unknown parameter {p} for {c}".format(p=name, c=type(self).__name__))
            if name in self._parameters:
                param = self._parameters[name]
                with torch.no_grad():
                    if torch.is_tensor(val):
                        param.copy_(val.to(param).expand_as(param))
                    else:
                        param.fill_(float(val))
            else:  # property with a setter (e.g. ``lengthscale``)
                setattr(self, name, val)
        return self


class Kernel(Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                 lengthscale_prior=None, lengthscale_constraint=None, eps=1e-6, **kwargs):
        has_ls = kwargs.pop("has_lengthscale", None)  # old-style ctor kwarg used by the reference
        own = self.__dict__.get("has_lengthscale", type(self).has_lengthscale)
        super().__init__()
        self.__dict__["has_lengthscale"] = bool(own if has_ls is None else (has_ls or own))
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(active_dims, dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self._batch_shape = torch.Size(batch_shape)
        self.eps = eps
        if self.has_lengthscale:
            n = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale", nn.Parameter(torch.zeros(*self._batch_shape, 1, n)))
            self.register_constraint("raw_lengthscale",
                                     Positive() if lengthscale_constraint is None else lengthscale_constraint)
            if lengthscale_prior is not None:
                self.register_prior("lengthscale_prior", lengthscale_prior, lambda m: m.lengthscale,
                                    lambda m, v: m._set_lengthscale(v))

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale) if self.has_lengthscale else None

    @lengthscale.setter
    def lengthscale(self, value):
        self._set_lengthscale(value)

    def _set_lengthscale(self, value):
        if not self.has_lengthscale:
            raise RuntimeError("Kernel has no lengthscale.")
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_lengthscale)
        self.initialize(raw_lengthscale=self.raw_lengthscale_constraint.inverse_transform(value))

    def covar_dist(self, x1, x2, diag=False, last_dim_is_batch=False, square_dist=False, **_):
        if diag:
            d2 = ((x1 - x2) ** 2).sum(-1)
        else:
            diff = x1.unsqueeze(-2) - x2.unsqueeze(-3)
            d2 = (diff ** 2).sum(-1)
        return d2 if square_dist else d2.clamp_min(1e-30).sqrt()

    def forward(self, x1, x2, diag=False, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        x1_, x2_ = x1, x2
        if self.active_dims is not None:
            dims = self.active_dims.to(x1_.device)
            x1_ = x1_.index_select(-1, dims)
            if x2_ is not None:
                x2_ = x2_.index_select(-1, dims)
        if x1_.ndimension() == 1:
            x1_ = x1_.unsqueeze(1)
        if x2_ is None:
            x2_ = x1_
        elif x2_.ndimension() == 1:
            x2_ = x2_.unsqueeze(1)
        return super().__call__(x1_, x2_, diag=diag, **params)

    def __mul__(self, other):
        return ProductKernel(self, other)


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        kwargs.setdefault("batch_shape", getattr(base_kernel, "batch_shape", torch.Size([])))
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", nn.Parameter(torch.zeros(self.batch_shape)))
        self.register_constraint("raw_outputscale",
                                 Positive() if outputscale_constraint is None else outputscale_constraint)
        if outputscale_prior is not None:
            self.register_prior("outputscale_prior", outputscale_prior, lambda m: m.outputscale,
                                lambda m, v: m._set_outputscale(v))

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        self._set_outputscale(value)

    def _set_outputscale(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_outputscale)
        self.initialize(raw_outputscale=self.raw_outputscale_constraint.inverse_transform(value))

    def forward(self, x1, x2, diag=False, **params):
        k = self.base_kernel(x1, x2, diag=diag, **params)
        s = self.outputscale
        s = s.view(*s.shape, 1) if diag else s.view(*s.shape, 1, 1)
        return k * s


class RBFKernel(Kernel):
    """exp(-|x - y|^2 / (2 l^2)): gpytorch's own RBF kernel, used by the reference only as the Euclidean factor of
    the SPD product heat kernels (kernel_utils/kernels_spd.py:346,606).  Plain torch ops on the inputs' device."""

    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        ls = self.lengthscale.to(x1)
        d2 = self.covar_dist(x1 / ls, x2 / ls, diag=diag, square_dist=True)
        return torch.exp(-0.5 * d2)


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)

    def forward(self, x1, x2, diag=False, **params):
        res = None
        for k in self.kernels:
            term = k(x1, x2, diag=diag, **params)
            res = term if res is None else res * term
        return res


if HAVE_GPYTORCH:  # pragma: no cover
    from gpytorch.constraints import GreaterThan, Interval, Positive  # noqa: F811,F401
    from gpytorch.kernels import Kernel, ProductKernel, RBFKernel, ScaleKernel  # noqa: F811,F401
    from gpytorch import Module  # noqa: F811,F401
