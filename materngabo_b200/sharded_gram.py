"""Row-sharded Gram matrix with the backward collective of SURVEY.md 8e.

Rows of x1 are independent, so ``K(X1, X2)`` is sharded by contiguous row blocks with no forward collective.  In the
backward pass of a loss that sums over all rows, each rank only sees its block of ``G = dL/dK``: the hyper-parameter
gradient ``dL/dcoef`` (T doubles) and the gradient of the replicated x2 are therefore partial sums that one
``all_reduce`` completes; ``dL/dx1`` stays local.  The reference has no counterpart (single device); the collective is
NCCL over NVLink on the GPU box, gloo in the CPU test (tests/test_sharding_gloo.py), with the per-rank Gram function
injectable for that purpose - the product default is the CUDA series kernel (``_ops.series_gram``).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class _AllReduceGrad(torch.autograd.Function):
    """Identity in the forward pass; sums the incoming gradient over the group in the backward pass.  Applied to a
    replicated operand (coefficients, x2) before the rank-local computation."""

    @staticmethod
    def forward(ctx, t, group):
        ctx.group = group
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist

        g = g.contiguous().clone()
        if dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def replicated(t: torch.Tensor, group=None) -> torch.Tensor:
    """Mark ``t`` as replicated over ``group``: its gradient becomes the sum of every rank's contribution."""
    return _AllReduceGrad.apply(t, group)


def sharded_series_gram(spec, x1_local: torch.Tensor, x2: torch.Tensor, coef: torch.Tensor, group=None,
                        gram_fn: Optional[Callable] = None) -> torch.Tensor:
    """This rank's row block of K(X1, X2) for a series kernel; autograd completes dL/dcoef (and dL/dx2 when x2
    requires grad) with one all-reduce of T (resp. n x D) doubles, as SURVEY.md 8e prescribes."""
    if gram_fn is None:
        from . import _ops

        gram_fn = lambda a, b, c: _ops.series_gram(spec, a, b, c)  # noqa: E731
    c = replicated(coef, group) if coef.requires_grad else coef
    b = replicated(x2, group) if x2.requires_grad else x2
    return gram_fn(x1_local, b, c)
