"""Drop-in model object for botorch's analytic acquisition functions.

The reference builds ``botorch.models.SingleTaskGP(x, y[:, None], covar_module=ScaleKernel(manifold_kernel),
likelihood=GaussianLikelihood(...))``, fits it with ``botorch.fit_gpytorch_model`` and hands it to
``ExpectedImprovement(model=model, best_f=..., maximize=False)``
(/root/reference/examples/bo_manifolds_benchmark_examples/bo_sphere.py:211-233,262,265); the acquisition function then
calls ``model.posterior(X)`` with ``X`` of shape ``b x 1 x D`` and reads ``.mean`` / ``.variance``
(manifold_optimization/manifold_optimize.py:195,254,337).

``FusedExactGP.from_model(model)`` reads the fitted hyper-parameters and training data off such a model (anything
exposing ``covar_module`` [``.base_kernel``, ``.outputscale``], ``likelihood.noise``, ``mean_module.constant``,
``train_inputs``, ``train_targets``) and serves ``posterior(X)`` from the fused CUDA path
(``ExactGPPosterior``: csrc/posterior.cu for scoring, the differentiable path when ``X`` requires grad), so that

    acq = botorch.acquisition.ExpectedImprovement(model=FusedExactGP.from_model(model), best_f=..., maximize=False)

works unchanged.  gpytorch / botorch are not in this image: the contract is exercised against doubles in
tests/test_botorch_adapter.py; this module imports neither.
"""
from __future__ import annotations

from typing import Optional

import torch

from .posterior import ExactGPPosterior


class GPPosterior:
    """The slice of ``botorch.posteriors.GPyTorchPosterior`` analytic acquisition functions read: ``mean`` and
    ``variance`` of shape ``b x q x 1`` (single output), plus ``device`` / ``dtype``."""

    def __init__(self, mean: torch.Tensor, variance: torch.Tensor):
        self.mean = mean
        self.variance = variance

    @property
    def device(self):
        return self.mean.device

    @property
    def dtype(self):
        return self.mean.dtype

    @property
    def stddev(self):
        return self.variance.clamp_min(0.0).sqrt()


def _scalar(v, what) -> float:
    if v is None:
        raise ValueError("model has no %s" % what)
    if torch.is_tensor(v):
        if v.numel() != 1:
            raise NotImplementedError("%s with %d elements: batched / multi-output models are not supported" % (what, v.numel()))
        return float(v.detach().reshape(()).cpu())
    return float(v)


class FusedExactGP(torch.nn.Module):
    """Single-output exact GP with a constant mean behind botorch's ``Model.posterior`` call."""

    num_outputs = 1

    def __init__(self, kernel, train_x: torch.Tensor, train_y: torch.Tensor, outputscale: float, noise: float,
                 mean_const: float, **posterior_kwargs):
        super().__init__()
        self.kernel = kernel                                 # registered: acq.to(device) reaches its parameters
        self.gp = ExactGPPosterior(kernel, train_x, train_y, outputscale=outputscale, noise=noise, mean_const=mean_const,
                                   **posterior_kwargs).fit()
        self._source = None

    # ---- construction from a fitted gpytorch / botorch model ------------------------------------
    @staticmethod
    def read_model(model):
        """(kernel, train_x, train_y, outputscale, noise, mean_const) of a fitted SingleTaskGP-like model."""
        if getattr(model, "outcome_transform", None) is not None or getattr(model, "input_transform", None) is not None:
            raise NotImplementedError("models with input / outcome transforms are not supported (the reference uses none)")
        covar = model.covar_module
        if hasattr(covar, "base_kernel"):                    # gpytorch.kernels.ScaleKernel
            kernel, outputscale = covar.base_kernel, _scalar(covar.outputscale, "outputscale")
        else:
            kernel, outputscale = covar, 1.0
        noise = _scalar(model.likelihood.noise, "likelihood.noise")
        mean_module = getattr(model, "mean_module", None)
        const = getattr(mean_module, "constant", None) if mean_module is not None else None
        mean_const = _scalar(const, "mean_module.constant") if const is not None else 0.0      # ZeroMean
        train_x = model.train_inputs[0] if isinstance(model.train_inputs, (tuple, list)) else model.train_inputs
        train_y = model.train_targets
        if train_x.dim() != 2:
            raise NotImplementedError("batched training inputs are not supported")
        return kernel, train_x.detach(), train_y.detach().reshape(-1), outputscale, noise, mean_const

    @classmethod
    def from_model(cls, model, **posterior_kwargs):
        kernel, x, y, s, noise, mean_const = cls.read_model(model)
        out = cls(kernel, x, y, s, noise, mean_const, **posterior_kwargs)
        out._source = model
        return out

    def refit(self):
        """Re-read the source model (after another ``fit_gpytorch_model`` or new training data) and redo the N x N work."""
        if self._source is None:
            raise RuntimeError("refit() needs a model: construct with from_model")
        kernel, x, y, s, noise, mean_const = self.read_model(self._source)
        gp = self.gp
        gp.kernel, gp.train_x, gp.train_y = kernel, x.contiguous(), y.contiguous()
        gp.outputscale, gp.noise, gp.mean_const = s, noise, mean_const
        gp.fit()
        return self

    # ---- botorch Model contract -------------------------------------------------------------------
    def posterior(self, X: torch.Tensor, output_indices=None, observation_noise=False, posterior_transform=None,
                  **kwargs) -> GPPosterior:
        if output_indices is not None or posterior_transform is not None:
            raise NotImplementedError("output_indices / posterior_transform: single-output analytic acquisition only")
        if X.dim() < 2:
            raise ValueError("X must be (..., q, D)")
        if X.dim() == 2:
            X = X.unsqueeze(-2)                             # botorch's t-batch convention: n x D -> n x 1 x D
        if X.shape[-2] != 1:
            raise NotImplementedError("q > 1 joint posteriors are not provided by the analytic acquisition path")
        batch = X.shape[:-2]
        flat = X.reshape(-1, X.shape[-1])
        if torch.is_grad_enabled() and X.requires_grad:
            mean, var = self.gp.posterior_autograd(flat)    # trust-region inner loop: autograd to X (first / second order)
        else:
            mean, var = self.gp.posterior(flat.detach().contiguous())
        if observation_noise is True:
            var = var + self.gp.noise
        elif torch.is_tensor(observation_noise):
            var = var + observation_noise.reshape(-1)
        return GPPosterior(mean.reshape(*batch, 1, 1), var.reshape(*batch, 1, 1))

    def forward(self, X):
        post = self.posterior(X)
        return post.mean, post.variance

    @property
    def train_inputs(self):
        return (self.gp.train_x,)

    @property
    def train_targets(self):
        return self.gp.train_y

    def handle(self, max_chunk: int = 0):
        """Persistent C-ABI handle over this model's fitted state (see ExactGPPosterior.handle)."""
        return self.gp.handle(max_chunk)
