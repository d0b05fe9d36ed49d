"""CUDA-graph replay of the launch-bound calls of the BO loop.

At the sizes of the reference's BO scripts (5 ... 100 training points: examples/bo_manifolds_benchmark_examples/
bo_sphere.py:154,351) one marginal-likelihood evaluation or one trust-region step is a chain of ~100 tiny launches
(hyper-parameter -> series coefficients in torch, the Gram kernel, Cholesky, solves, the backward of all of it), and
the time goes to Python and launch overhead, not to the GPU.  Both are called hundreds of times with fixed shapes:
by the L-BFGS fit (``botorch.fit_gpytorch_model``, bo_sphere.py:262) and by the trust-region solver
(manifold_optimization/manifold_optimize.py:184-236).  ``GraphedValueAndGrad`` captures value + gradient
(+ optionally a Hessian-vector product) once into a CUDA graph and replays it: one launch per call.

Everything captured must be free of host synchronisation: the libmgb entry points are (they never synchronise or
allocate), torch ops are, ``torch.linalg.cholesky`` is not (it reads its ``info`` back) - use
``torch.linalg.cholesky_ex(..., check_errors=False)`` inside captured functions, as ``exact_mll`` does.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import torch

from . import _lib

F64 = torch.float64


class GraphedValueAndGrad:
    """``value = fn(*inputs)`` (a scalar) and its gradient with respect to ``wrt``, replayed from a CUDA graph.

    ``inputs``: example tensors fixing shapes / dtypes; per call their values are copied into static buffers.
    ``wrt``: tensors to differentiate with respect to - module parameters (static by nature: update them in place)
    and/or integer indices into ``inputs``.  ``hvp_index``: index into ``inputs`` of a direction ``v``; when given the
    graph also returns d/d wrt[0] <grad[0], v> (Hessian-vector product for the first ``wrt`` entry, what pymanopt's
    ``ehess`` asks for: pymanopt_addons/tools/autodiff/_pytorch.py:92-116).
    ``fn`` may return a tuple ``(value, aux...)``: the auxiliary tensors (e.g. the Cholesky ``info`` flag of
    ``exact_mll``) are kept in ``self.aux`` after every replay.
    """

    def __init__(self, fn: Callable, inputs: Sequence[torch.Tensor], wrt: Sequence, hvp_index: Optional[int] = None,
                 warmup: int = 3):
        dev = _lib.require_cuda_f64(*[t for t in inputs if t.dtype == F64]) if inputs else None
        if dev is None:
            tensors = [w for w in wrt if torch.is_tensor(w)]
            if not tensors or not tensors[0].is_cuda:
                raise _lib.MgbError("GraphedValueAndGrad needs CUDA tensors (no CPU fallback)")
            dev = tensors[0].device
        self.device = dev
        self.fn = fn
        self.static_inputs = [t.detach().clone() for t in inputs]
        self.wrt_spec = list(wrt)
        for i, w in enumerate(self.wrt_spec):
            if isinstance(w, int):
                self.static_inputs[w].requires_grad_(True)
        self.hvp_index = hvp_index
        self.graph = torch.cuda.CUDAGraph()
        self.value = None
        self.grads = None
        self.hvp = None
        self.aux = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # allocator warm-up, lazy initialisations, table uploads
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.cuda.graph(self.graph):
            self._run()

    def _wrt(self):
        return [self.static_inputs[w] if isinstance(w, int) else w for w in self.wrt_spec]

    def _run(self):
        out = self.fn(*self.static_inputs)
        value, aux = (out[0], out[1:]) if isinstance(out, tuple) else (out, ())
        wrt = self._wrt()
        second = self.hvp_index is not None
        grads = torch.autograd.grad(value, wrt, create_graph=second, allow_unused=True)
        hvp = None
        if second:
            v = self.static_inputs[self.hvp_index]
            (hvp,) = torch.autograd.grad((grads[0] * v).sum(), [wrt[0]])
            grads = tuple(g.detach() if g is not None else None for g in grads)
        self.value, self.grads, self.hvp, self.aux = value.detach(), grads, hvp, aux

    def __call__(self, *inputs):
        """Copies ``inputs`` into the static buffers, replays, returns (value, grads[, hvp]) - views of static output
        buffers, valid until the next call."""
        if len(inputs) != len(self.static_inputs):
            raise ValueError("expected %d inputs" % len(self.static_inputs))
        with torch.no_grad():
            for dst, src in zip(self.static_inputs, inputs):
                if src is not dst:
                    dst.copy_(src)
        self.graph.replay()
        if self.hvp_index is not None:
            return self.value, self.grads, self.hvp
        return self.value, self.grads


def _scalar(v, like):
    """Python number or tensor -> 0-dim float64 device tensor without a host-to-device copy (capturable)."""
    if torch.is_tensor(v):
        return v.reshape(()).to(device=like.device, dtype=F64)
    return torch.full((), float(v), dtype=F64, device=like.device)


_CONST_TAILS = {}


def _hyper(scale, noise, mean_const, jitter, like):
    """[outputscale, noise + jitter, mean] as one float64 device vector for the MLL kernels.  When noise and mean are plain
    numbers their pair is built once per (values, device) and reused (no fill / add / stack launches per evaluation; the
    first call happens in the warm-up before a graph capture)."""
    if not torch.is_tensor(noise) and not torch.is_tensor(mean_const):
        key = (float(noise) + float(jitter), float(mean_const), str(like.device))
        tail = _CONST_TAILS.get(key)
        if tail is None:
            if len(_CONST_TAILS) > 256:
                _CONST_TAILS.clear()
            tail = torch.tensor([key[0], key[1]], dtype=F64).to(like.device)
            _CONST_TAILS[key] = tail
        return torch.cat([_scalar(scale, like).reshape(1), tail])
    return torch.stack([_scalar(scale, like), _scalar(noise, like) + jitter, _scalar(mean_const, like)])


def _fused_mll(kernel, x, y, noise, mean_const, jitter):
    """(loss, info) through the single-CTA kernel csrc/mll.cu, or None when the problem does not qualify
    (not a single-factor series kernel, active_dims, batch dimensions, n above the shared-memory limit)."""
    from . import _ops
    from .kernels_sphere import _series_on

    base, scaled = kernel, False
    if hasattr(kernel, "base_kernel") and hasattr(type(kernel), "outputscale"):   # type(): hasattr on the instance would run the softplus
        base, scaled = kernel.base_kernel, True
        if getattr(kernel, "active_dims", None) is not None:
            return None
    if not hasattr(base, "series") or getattr(base, "active_dims", None) is not None or x.dim() != 2:
        return None
    if hasattr(base, "_to_circles"):
        return None
    scale = kernel.outputscale if scaled else 1.0
    spec, coef = _series_on(base, x.device)
    if spec.n_factors != 1 or x.shape[-1] != spec.factor_width:
        return None
    if x.shape[0] > _ops.dense_large_max_n():
        return None
    hyper = _hyper(scale, noise, mean_const, jitter, x)
    nll, info = _ops.series_mll(spec, x, y, coef, hyper)
    return nll / x.shape[0], info


def _dense_mll(kernel, x, y, noise, mean_const, jitter):
    """(loss, info) for any kernel: its own Gram kernel (with its own backward), then ONE launch for Cholesky, solves,
    log-determinant and dnll/dK (csrc/mll.cu, dense mode).  None when n exceeds the single-CTA limit or for batches."""
    from . import _ops

    if x.dim() != 2 or x.shape[0] > _ops.dense_large_max_n():
        return None
    base, scale = kernel, 1.0
    if hasattr(kernel, "base_kernel") and hasattr(type(kernel), "outputscale") and getattr(kernel, "active_dims", None) is None:
        base, scale = kernel.base_kernel, kernel.outputscale
    k = base(x, x)
    if not torch.is_tensor(k):          # gpytorch lazy tensors
        k = k.to_dense() if hasattr(k, "to_dense") else k.evaluate()
    hyper = _hyper(scale, noise, mean_const, jitter, x)
    nll, info = _ops.dense_mll(k, y, hyper)
    return nll / x.shape[0], info


def exact_mll(kernel, x, y, noise, mean_const=0.0, jitter: float = 0.0, fused: bool = True):
    """Negative exact-GP marginal log likelihood divided by N (gpytorch's ExactMarginalLogLikelihood convention) for
    ``kernel`` (typically ScaleKernel(manifold kernel)) + Gaussian noise + constant mean: the quantity
    ``botorch.fit_gpytorch_model`` minimises (bo_sphere.py:262).  Capturable: no host synchronisation.
    Single-factor series kernels (sphere, SO(3), circle) with N up to ~160 go through ONE fused launch (csrc/mll.cu:
    Gram, Cholesky, solves, log-determinant and all gradients in shared memory); every other kernel at those sizes
    through its own Gram kernel + one launch of the same CTA kernel in dense mode (dnll/dK handed to the kernel's
    backward); 160 < N <= 16384 through the multi-CTA path (csrc/dense_spd.cu: cooperative blocked Cholesky, triangular
    inverse by recursive doubling, gradient contraction); torch.linalg only beyond that or with fused=False.  Returns (loss, info) - ``info`` is the Cholesky status flag (non-zero: K + noise I not
    positive definite)."""
    if fused and hasattr(kernel, "forward"):
        res = _fused_mll(kernel, x, y, noise, mean_const, jitter)
        if res is None:
            res = _dense_mll(kernel, x, y, noise, mean_const, jitter)
        if res is not None:
            return res
    n = x.shape[-2]
    k = kernel(x, x) if hasattr(kernel, "forward") else kernel
    diag = (noise + jitter) if torch.is_tensor(noise) else float(noise) + jitter     # python scalars: no H2D copy
    kk = k + diag * torch.eye(n, dtype=F64, device=x.device)
    L, info = torch.linalg.cholesky_ex(kk, check_errors=False)
    r = (y - mean_const).unsqueeze(-1)
    alpha = torch.cholesky_solve(r, L)
    nll = 0.5 * (r * alpha).sum() + torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum() + 0.5 * n * math.log(2 * math.pi)
    return nll / n, info
