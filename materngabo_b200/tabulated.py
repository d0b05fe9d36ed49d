"""Opt-in tabulated form of the radial quadrature kernels (hyperbolic H^1..H^3, Euclidean integrated Matern).

The reference integrates a quadrature for EVERY Gram entry (kernel_utils/kernels_hyperbolic.py:281-304,367-383:
64 x 64 nodes per entry at BASELINE config 4), and so do the default CUDA kernels of this package
(csrc/quad_gram.cu) - that is the path parity, the benchmark configs and the acquisition gradients use.
But the quadrature does not depend on the pair: K_ij = P(z_ij) with ONE fixed profile P per set of
hyper-parameters.  For large candidate sets ``TabulatedRadialKernel`` samples P once with the same device
quadrature (``mgb_radial_profile``) at Chebyshev nodes of a piecewise partition, validates the interpolant against
the quadrature, and evaluates a degree-15 polynomial per entry (``mgb_radial_table_gram``): the Gram matrix becomes
bandwidth-bound instead of costing ~10^4 FP64 operations per entry.

Not a drop-in for ``Kernel.forward``: no autograd, values agree with the quadrature to the validated tolerance
(default 1e-12 of max |P|), and ``bench.py`` reports it under its own workload name.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib, _ops, _weights

F64 = torch.float64


class TabulatedRadialKernel:
    def __init__(self, kernel, v_max: float, degree=(7, 11, 15), sub_bits: int = 3, tol: float = 1e-12):
        """``kernel``: a hyperbolic kernel of this package with dim in {1, 2, 3} or ``EuclideanIntegratedMaternKernel``;
        ``v_max``: upper bound of z + 1 over all pairs that will be evaluated (z = 2 sinh(d / 2) for H^n, |x - y|
        for R^d) - see ``for_inputs``."""
        dev = next(kernel.parameters()).device
        if dev.type != "cuda":
            raise _lib.MgbError("TabulatedRadialKernel needs the kernel's parameters on a CUDA device (no CPU fallback)")
        self.device = dev
        with torch.no_grad():
            nodes = kernel.nodes(dev)
        if len(nodes) == 5:                      # hyperbolic: (tval, tcoef, s0, ds, Ps)
            self.geom, self.dim = 0, int(kernel.dim)
            self.kind = self.dim - 1
            self.tval, self.tcoef, self.s0, self.ds, self.ps = nodes
            if self.dim not in (1, 2, 3):
                raise NotImplementedError
        else:                                    # Euclidean: (tval, tcoef)
            self.geom, self.dim = 1, int(kernel.dim)
            self.kind = _ops.KIND_EUCLID
            self.tval, self.tcoef = nodes
            self.s0 = self.ds = 0.0
            self.ps = 0
            if self.dim > 4:
                raise NotImplementedError("Euclidean rows wider than 4")
        self.tval, self.tcoef = self.tval.detach().contiguous(), self.tcoef.detach().contiguous()
        if not (v_max >= 1.0) or not math.isfinite(v_max):
            raise ValueError("v_max must be a finite number >= 1")
        self.v_max = float(v_max)
        # Lowest degree first (the per-entry cost is one shared-memory read + one FMA per coefficient), refining the
        # partition until the interpolant reproduces the quadrature.  The H^2 heat kernel's first trapezoid node
        # (s0 = 0.01) puts structure at the scale d ~ 0.005 near d = 0 and needs 2^8 intervals per binade.
        degrees = (degree,) if isinstance(degree, int) else tuple(degree)
        found, best = None, float("inf")
        for deg in degrees:
            self.degree = int(deg)
            for bits in range(int(sub_bits), 9):
                if ((int(math.floor(math.log2(self.v_max))) + 1) << bits) * (self.degree + 2) * 8 > 200 * 1024:
                    break
                coef, err, scale = self._build(bits)
                best = min(best, err / scale)
                if err <= tol * scale:
                    found = (bits, coef, err, scale)
                    break
            if found:
                break
        if not found:
            raise _lib.MgbError("tabulated profile did not reach %.1e of max |P| (best %.3e)" % (tol, best))
        self.sub_bits, self.coef, self.max_abs_error, self.scale = found
        self.n_int = self.coef.shape[0]

    # ---- construction ------------------------------------------------------------------------------
    def profile(self, z: torch.Tensor) -> torch.Tensor:
        """P at pair scalars z (device quadrature, the same nodes as the Gram kernels)."""
        arg = 2.0 * torch.asinh(0.5 * z) if self.geom == 0 else z * z
        return _ops._radial_profile_raw(self.kind, 0, arg.contiguous(), self.tval, self.tcoef, float(self.s0), float(self.ds),
                                        int(self.ps))

    def _intervals(self, bits):
        n_int = (int(math.floor(math.log2(self.v_max))) + 1) << bits
        idx = np.arange(n_int)
        e, sub = idx >> bits, idx & ((1 << bits) - 1)
        lo = np.ldexp(1.0 + sub / float(1 << bits), e)
        hi = np.ldexp(1.0 + (sub + 1) / float(1 << bits), e)
        return lo, hi

    def _build(self, bits):
        n = self.degree + 1
        lo, hi = self._intervals(bits)
        j = np.arange(n)
        s = np.cos(np.pi * (j + 0.5) / n)                                   # Chebyshev nodes in (-1, 1)
        v = lo[:, None] + 0.5 * (s[None, :] + 1.0) * (hi - lo)[:, None]
        f = self.profile(torch.as_tensor(v - 1.0, dtype=F64, device=self.device).reshape(-1)).cpu().numpy().reshape(v.shape)
        k = np.arange(n)
        dct = np.cos(np.pi * np.outer(k, j + 0.5) / n) * (2.0 / n)          # c_k = 2/n sum_j f_j cos(k pi (j + 1/2) / n)
        dct[0] *= 0.5
        cheb = f @ dct.T
        mono = cheb @ _weights.chebyshev_to_monomial(n)                     # monomials in s
        # validation at points that are not interpolation nodes
        rng = np.random.default_rng(0)
        st = np.concatenate([rng.uniform(-1.0, 1.0, 6), [-1.0, 0.0, 1.0 - 1e-12]])
        vt = lo[:, None] + 0.5 * (st[None, :] + 1.0) * (hi - lo)[:, None]
        ft = self.profile(torch.as_tensor(vt - 1.0, dtype=F64, device=self.device).reshape(-1)).cpu().numpy().reshape(vt.shape)
        approx = np.zeros_like(vt)
        for kk in range(n - 1, -1, -1):
            approx = approx * st[None, :] + mono[:, kk:kk + 1]
        err = float(np.abs(approx - ft).max())
        scale = float(max(np.abs(f).max(), 1e-300))
        coef = torch.as_tensor(mono, dtype=F64, device=self.device).contiguous()
        return coef, err, scale

    @classmethod
    def for_inputs(cls, kernel, x1: torch.Tensor, x2: torch.Tensor, **kw):
        """Table covering every pair of rows of ``x1`` and ``x2`` (one reduction + host read to bound the range)."""
        _lib.require_cuda_f64(x1, x2)
        with torch.no_grad():
            if hasattr(kernel, "nodes") and len(kernel.nodes(x1.device)) == 5:
                # d(x, y) <= d(x, o) + d(o, y), d(x, o) = acosh(x_0);  z = 2 sinh(d / 2)
                d = torch.acosh(x1[..., 0].max().clamp_min(1.0)) + torch.acosh(x2[..., 0].max().clamp_min(1.0))
                v_max = float(2.0 * torch.sinh(0.5 * d) + 1.0) * (1.0 + 1e-9) + 1e-3
            else:
                r = x1.norm(dim=-1).max() + x2.norm(dim=-1).max()
                v_max = float(r + 1.0) * (1.0 + 1e-9)
        return cls(kernel, v_max, **kw)

    # ---- evaluation --------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda_f64(x1, x2)
        width = self.dim + 1 if self.geom == 0 else self.dim
        if x1.dim() != 2 or x2.dim() != 2 or x1.shape[-1] != width or x2.shape[-1] != width:
            raise RuntimeError("expected 2-D inputs with %d coordinates" % width)
        x1, x2 = x1.contiguous(), x2.contiguous()
        m, n = x1.shape[0], x2.shape[0]
        k = torch.empty((m, n), dtype=F64, device=x1.device)
        lib = _lib.load()
        with torch.cuda.device(x1.device):
            rc = lib.mgb_radial_table_gram(self.geom, self.dim, _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                           self.sub_bits, self.n_int, self.degree, _lib.ptr(self.coef), _lib.ptr(k),
                                           max(n, 1), _lib.stream_ptr(x1.device))
        _lib.check(rc, "mgb_radial_table_gram")
        return k

    __call__ = forward
