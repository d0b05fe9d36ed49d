"""Exact-GP posterior on the device + analytic EI + row-sharding over the GPUs of one box.

What this replaces in the reference's BO loop (all unvendored gpytorch / botorch arithmetic):
``SingleTaskGP(x, y, covar_module=ScaleKernel(manifold_kernel))`` in eval mode called through
``ExpectedImprovement(model, best_f, maximize=False)(X)`` with ``X`` of shape ``b x 1 x D``
(examples/bo_manifolds_benchmark_examples/bo_sphere.py:211-233,265; manifold_optimization/
manifold_optimize.py:195,254,337).  The arithmetic is

    mean = m + s K(X*, X) alpha,            alpha = (s K + noise I)^-1 (y - m)
    var  = s k(x*, x*) - || L^-1 s K(X*, X)^T ||^2,   L = chol(s K + noise I)

``fit`` does the one-off N x N work with torch.linalg (cuSOLVER) and packs L^-1 and alpha for the fused
kernel; ``posterior`` streams candidates through ``mgb_series_posterior`` (Gram chunk -> DMMA triangular
product + square-sum, csrc/posterior.cu) and never materialises K(X*, X) beyond one L2/HBM-resident chunk.
"""
from __future__ import annotations

import os

import contextlib
import ctypes as C
import math
from typing import Callable, Optional, Union

import torch

from . import _lib, _ops

F64 = torch.float64
DEFAULT_CHUNK = 37888  # 2 waves of 128-candidate tiles on 148 SMs


def _self_kernel_value(spec: _ops.SeriesSpec, coef: torch.Tensor) -> torch.Tensor:
    """k(x, x) for on-manifold x: every factor's pair scalar clamps to ``hi`` (unit vectors, rotations)."""
    c = coef.detach().to("cpu", F64)
    u = spec.hi
    val = 1.0
    for f in range(c.shape[0]):
        if spec.basis == _lib.BASIS_MONOMIAL:
            s = 0.0
            for k in range(c.shape[1] - 1, -1, -1):
                s = s * u + float(c[f, k])
        else:
            theta = math.acos(u)
            s = sum(float(c[f, k]) * math.cos(k * theta) for k in range(c.shape[1]))
        val *= s
    return val


@contextlib.contextmanager
def _frozen_parameters(module):
    params = [p for p in module.parameters() if p.requires_grad] if hasattr(module, "parameters") else []
    for p in params:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in params:
            p.requires_grad_(True)


INT8_CHUNK = 32768      # candidates per block of the library INT8 path: 640 GEMM tiles per launch = 9 waves of 2-SM tile pairs
INT8_FUSED_CHUNK = 37888   # fused tcgen05 kernel: 296 row blocks of 128 = two per SM
# MGB_POSTERIOR_PATH=auto (the default): the INT8 tensor-core contraction serves calls with at least INT8_AUTO_MIN_M
# candidates against INT8_AUTO_MIN_N .. INT8_AUTO_MAX_N training points (series kernels, fused implementation); smaller
# calls - the BO loop's hundreds of candidates against ~100 points - stay on the FP64 DMMA kernel, which has no
# per-fit slicing pass.  int32 accumulation of the digit products is exact up to n = 16384.  Measured ratios INT8 / FP64:
# 2.8 at 200 k x 2048, 2.5 at 1 M x 2048 per GPU, 2.7 at 10 M x 5000 (bench.py fp64_path); the lower edges are
# conservative round numbers, not tuned break-even points.
INT8_AUTO_MIN_N, INT8_AUTO_MAX_N, INT8_AUTO_MIN_M = 1024, 16384, 16384


def _int8_mode_from_env() -> str:
    v = os.environ.get("MGB_POSTERIOR_PATH", "auto").lower()
    if v not in ("auto", "fp64", "int8"):
        raise ValueError("MGB_POSTERIOR_PATH must be auto, fp64 or int8 (got %r)" % v)
    return {"auto": "auto", "fp64": "off", "int8": "on"}[v]


class ExactGPPosterior:
    """Constant-mean exact GP over one of the drop-in kernels.

    Series kernels (sphere, SO(3), circle, torus: anything exposing ``series() -> (SeriesSpec, coef)``) use the
    fully fused path (Gram chunk written directly in the operand layout).  Any other kernel of this package
    (hyperbolic, SPD(2), Euclidean, product kernels) goes through ``kernel.forward`` + a re-tiling kernel and the
    same fused triangular reduction; those kernels are stationary and normalised, so k(x, x) is evaluated once.
    """

    def __init__(self, kernel, train_x: torch.Tensor, train_y: torch.Tensor, outputscale: float = 1.0,
                 noise: float = 1e-2, mean_const: float = 0.0, chunk: int = DEFAULT_CHUNK, tabulate: bool = False,
                 int8: Optional[Union[bool, str]] = None):
        _lib.require_cuda_f64(train_x, train_y)
        self.kernel = kernel
        self.train_x = train_x.contiguous()
        self.train_y = train_y.contiguous().reshape(-1)
        self.outputscale = float(outputscale)
        self.noise = float(noise)
        self.mean_const = float(mean_const)
        self.chunk = int(chunk)
        # opt-in: candidate rows of the hyperbolic / Euclidean kernels through the tabulated profile
        # (materngabo_b200/tabulated.py) instead of one quadrature per entry; the fit always uses the quadrature
        self.tabulate = bool(tabulate)
        self.device = train_x.device
        # The N^2 contraction on the INT8 tensor cores by error-free slicing (csrc/int8_fused.cu: 56-bit operands, exact
        # int32 digit products, float64 recombination; ~3e-13 of max var from the FP64 kernel's ~1e-14, parity bar 1e-10).
        # int8 = True / False forces it on / off, "auto" or None (-> MGB_POSTERIOR_PATH, default auto) lets the call
        # size decide (INT8_AUTO_* above).
        if int8 is None:
            self._int8_mode = _int8_mode_from_env()
        elif isinstance(int8, str):
            self._int8_mode = "auto" if int8 == "auto" else "on"
        else:
            self._int8_mode = "on" if int8 else "off"
        # which INT8 implementation: "fused" = this repo's tcgen05 kernel (csrc/int8_fused.cu), "library" = CUTLASS int8
        # GEMMs + recombination pass (csrc/int8_posterior.cu); int8="library" or MGB_INT8_IMPL=library selects the latter
        self.int8_impl = int8 if isinstance(int8, str) and int8 != "auto" else os.environ.get("MGB_INT8_IMPL", "fused")
        if self.int8_impl not in ("fused", "library"):
            raise ValueError("int8 implementation must be 'fused' or 'library', got %r" % (self.int8_impl,))
        self._packed = None
        self._packed_i8 = None
        self._ws = None
        self._ws_i8 = None
        self.spec = None
        self.coef = None
        self.kss_value = None

    # ---- one-off N x N work ------------------------------------------------------------------
    @property
    def int8(self) -> bool:
        """True when large calls go through the INT8 tensor-core contraction (always, in the forced mode)."""
        if self._int8_mode != "auto":
            return self._int8_mode == "on"
        n = self.train_x.shape[0]
        return self.int8_impl == "fused" and self.is_series and INT8_AUTO_MIN_N <= n <= INT8_AUTO_MAX_N

    @int8.setter
    def int8(self, value):
        self._int8_mode = "auto" if value == "auto" else ("on" if value else "off")

    def _int8_for(self, m: int) -> bool:
        """Does a call with m candidates take the INT8 path?"""
        if not self.int8:
            return False
        if self._int8_mode == "on":
            return True
        # auto: large calls only, and only where the sliced factor exists or can be made (a peer rank that received the
        # FP64 state alone keeps the FP64 kernel)
        return m >= INT8_AUTO_MIN_M and (self._packed_i8 is not None or getattr(self, "rinv", None) is not None)

    @property
    def is_series(self):
        return hasattr(self.kernel, "series")

    @torch.no_grad()
    def fit(self):
        x = self._prepared(self.train_x)
        n = x.shape[0]
        if self.is_series:
            spec, coef = self.kernel.series()
            coef = (coef.detach().to(self.device, F64) * self._factor_scale(spec)).contiguous()
            if self._fused_fit(spec, coef, x):
                return self
            k = _ops.series_gram(spec, x, x, coef)
        else:
            spec, coef = None, None
            k = self.outputscale * self.kernel.forward(x, x)
        if not self._dense_fit(k.detach()) and not self._dense_fit_large(k.detach()):
            k = k.detach().clone()                        # n beyond mgb_dense_max_n(): torch.linalg (cuSOLVER)
            k.diagonal().add_(self.noise)
            L = torch.linalg.cholesky(k)
            self.alpha = torch.cholesky_solve((self.train_y - self.mean_const).unsqueeze(-1), L).squeeze(-1).contiguous()
            self.rinv = torch.linalg.solve_triangular(L, torch.eye(n, dtype=F64, device=self.device), upper=False).contiguous()
            self.rinv_t = self.rinv.t().contiguous()      # row access to the columns of L^-1 (fused acquisition kernel)
        self.spec, self.coef = spec, coef
        if self.is_series:
            self.kss_value = _self_kernel_value(spec, coef)
        else:
            self.kss_value = float(self.outputscale * self.kernel.forward(x[:1], x[:1]).reshape(()))
        self.train_prepared = x
        # quadrature node weights of the non-series kernels: fixed once the hyper-parameters are (acquisition path)
        nodes = getattr(self.kernel, "nodes", None)
        self._nodes = tuple(t.detach().contiguous() if torch.is_tensor(t) else t for t in nodes(self.device)) \
            if (nodes is not None and not self.is_series) else None
        self._pack()
        return self

    def _fused_fit(self, spec, coef, x) -> bool:
        """Launch-bound sizes (single-factor series kernel, n <= ~160): Gram, Cholesky, L^-1 and alpha from ONE launch of
        the single-CTA kernel (csrc/mll.cu) instead of the Gram kernel + four torch.linalg calls."""
        n = x.shape[0]
        if spec.n_factors != 1 or spec.pair_square or n > _ops.series_mll_max_n(spec, coef.shape[-1]):
            return False
        lib = _lib.load()
        self.rinv = torch.empty((n, n), dtype=F64, device=self.device)
        self.rinv_t = torch.empty((n, n), dtype=F64, device=self.device)
        self.alpha = torch.empty(n, dtype=F64, device=self.device)
        nll = torch.empty(1, dtype=F64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        # coef already carries the outputscale: hyper = (1, noise, mean)
        hyper = torch.tensor([1.0, self.noise, self.mean_const], dtype=F64).to(self.device)
        d = spec.desc(coef.shape[-1])
        with torch.cuda.device(self.device):
            rc = lib.mgb_series_fit(C.byref(d), _lib.ptr(x), n, x.stride(0), _lib.ptr(self.train_y), _lib.ptr(coef),
                                    _lib.ptr(hyper), _lib.ptr(self.rinv), _lib.ptr(self.rinv_t), n, _lib.ptr(self.alpha),
                                    _lib.ptr(nll), _lib.ptr(status), _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_series_fit")
        if int(status) != 0:
            raise torch.linalg.LinAlgError("ExactGPPosterior.fit: s K + noise I is not positive definite (pivot %d)" % int(status))
        self.spec, self.coef = spec, coef
        self.kss_value = _self_kernel_value(spec, coef)
        self.train_prepared = x
        self._nodes = None
        self._pack()
        return True

    def _dense_fit(self, k) -> bool:
        """Launch-bound sizes, any kernel: Cholesky, L^-1 and alpha of k + noise I from ONE launch of the single-CTA kernel
        in dense mode (mgb_dense_fit) instead of four torch.linalg calls."""
        n = k.shape[0]
        if k.dim() != 2 or n > _ops.dense_mll_max_n():
            return False
        k = k.contiguous()
        lib = _lib.load()
        self.rinv = torch.empty((n, n), dtype=F64, device=self.device)
        self.rinv_t = torch.empty((n, n), dtype=F64, device=self.device)
        self.alpha = torch.empty(n, dtype=F64, device=self.device)
        nll = torch.empty(1, dtype=F64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        hyper = torch.tensor([1.0, self.noise, self.mean_const], dtype=F64).to(self.device)      # k carries the outputscale
        with torch.cuda.device(self.device):
            rc = lib.mgb_dense_fit(_lib.ptr(k), n, k.stride(0), _lib.ptr(self.train_y), _lib.ptr(hyper), _lib.ptr(self.rinv),
                                   _lib.ptr(self.rinv_t), n, _lib.ptr(self.alpha), _lib.ptr(nll), _lib.ptr(status),
                                   _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_dense_fit")
        if int(status) != 0:
            raise torch.linalg.LinAlgError("ExactGPPosterior.fit: s K + noise I is not positive definite (pivot %d)" % int(status))
        return True

    def _dense_fit_large(self, k) -> bool:
        """160 < n <= mgb_dense_max_n(): Cholesky, L^-1 and alpha from this library's multi-CTA path (csrc/dense_spd.cu:
        one cooperative launch for the factorisation, recursive-doubling triangular inverse) instead of torch.linalg."""
        lib = _lib.load()
        n = k.shape[0]
        if k.dim() != 2 or n > lib.mgb_dense_max_n() or os.environ.get("MGB_FIT_PATH", "") == "torch":
            return False
        k = k.contiguous()
        ld = (n + 1) // 2 * 2
        rinv = torch.empty((n, ld), dtype=F64, device=self.device)
        self.alpha = torch.empty(n, dtype=F64, device=self.device)
        nll = torch.empty(1, dtype=F64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        hyper = torch.tensor([1.0, self.noise, self.mean_const], dtype=F64).to(self.device)      # k carries the outputscale
        ws = _ops._dense_workspace(self.device, lib.mgb_dense_fit_workspace_bytes(n))
        want_t = n <= 4096                               # the transposed copy only serves the fused acquisition kernel
        rinv_t = torch.empty((n, ld), dtype=F64, device=self.device) if want_t else None
        with torch.cuda.device(self.device):
            rc = lib.mgb_dense_fit_large(_lib.ptr(k), n, k.stride(0), _lib.ptr(self.train_y), _lib.ptr(hyper), _lib.ptr(rinv),
                                         _lib.ptr(rinv_t), ld, _lib.ptr(self.alpha), _lib.ptr(nll), _lib.ptr(status),
                                         _lib.ptr(ws), ws.numel() * 8, _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_dense_fit_large")
        if int(status) != 0:
            raise torch.linalg.LinAlgError("ExactGPPosterior.fit: s K + noise I is not positive definite (pivot %d)" % int(status))
        self.rinv = rinv[:, :n]
        self.rinv_t = rinv_t[:, :n] if rinv_t is not None else None
        self.fit_nll = nll
        return True

    def _factor_scale(self, spec):
        # outputscale multiplies the product of factors once: put it on factor 0 only
        s = torch.ones(spec.n_factors, 1, dtype=F64, device=self.device)
        s[0, 0] = self.outputscale
        return s

    def _prepared(self, x):
        """Torus kernels accept angles: convert exactly as their forward does."""
        prep = getattr(self.kernel, "_to_circles", None)
        x = prep(x) if prep is not None else x
        return x.contiguous()

    def _pack(self):
        lib = _lib.load()
        n = self.rinv.shape[0]
        nbytes = lib.mgb_posterior_pack_bytes(n)
        self._packed = torch.empty(nbytes // 8, dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.mgb_posterior_pack(_lib.ptr(self.rinv), n, self.rinv.stride(0), _lib.ptr(self.alpha),
                                        _lib.ptr(self._packed), _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_posterior_pack")
        self._packed_i8 = None                     # sliced lazily by the first INT8 evaluation after a (re)fit

    # ---- state exchange for the sharded path ---------------------------------------------------
    def state(self):
        """Tensors a peer rank needs to evaluate the posterior (what NCCL broadcasts after a fit)."""
        st = {"train": self.train_prepared, "packed": self._packed, "coef": self.coef}
        if self.int8:
            self._ensure_packed_i8()
            st["packed_i8"] = self._packed_i8
            st["alpha"] = self.alpha
        return st

    def load_state(self, spec, train, packed, coef, kss_value, packed_i8=None, alpha=None):
        self.spec, self.train_prepared, self._packed, self.coef, self.kss_value = spec, train, packed, coef, kss_value
        if packed_i8 is not None:
            self._packed_i8, self.alpha = packed_i8, alpha
        return self

    # ---- candidates ----------------------------------------------------------------------------
    @torch.no_grad()
    def posterior(self, x: torch.Tensor):
        """x: (M, D) or botorch's (b, 1, D).  Returns (mean, var), each (M,) / (b,)."""
        if self._packed is None:
            raise RuntimeError("call fit() (or load_state) first")
        _lib.require_cuda_f64(x)
        if x.dim() == 3:
            if x.shape[-2] != 1:
                raise NotImplementedError("q > 1 joint posteriors are not provided by the analytic acquisition path")
            x = x.squeeze(-2)
        x = self._prepared(x)
        m, n = x.shape[0], self.train_prepared.shape[0]
        mean = torch.empty(m, dtype=F64, device=self.device)
        var = torch.empty(m, dtype=F64, device=self.device)
        if m == 0:
            return mean, var
        lib = _lib.load()
        chunk = min(self.chunk, (m + 127) // 128 * 128)
        if self._int8_for(m):
            return self._posterior_int8(x, mean, var, chunk)
        if not self.is_series:
            return self._posterior_generic(x, mean, var, chunk)
        need = lib.mgb_posterior_workspace_bytes(chunk, n)
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = torch.empty((need + 7) // 8, dtype=F64, device=self.device)
        d = self.spec.desc(self.coef.shape[-1])
        with torch.cuda.device(self.device):
            rc = lib.mgb_series_posterior(C.byref(d), _lib.ptr(x), m, x.stride(0), _lib.ptr(self.train_prepared), n,
                                          self.train_prepared.stride(0), _lib.ptr(self.coef), _lib.ptr(self._packed),
                                          self.kss_value, self.mean_const, _lib.ptr(mean), _lib.ptr(var),
                                          _lib.ptr(self._ws), self._ws.numel() * 8, chunk,
                                          _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_series_posterior")
        return mean, var


    @torch.no_grad()
    def posterior_host(self, x_host: torch.Tensor, mean_host: Optional[torch.Tensor] = None,
                       var_host: Optional[torch.Tensor] = None, rows_per_copy: int = 1 << 20):
        """``posterior`` for candidates in (pinned) HOST memory with the results returned to the host: rows are uploaded
        in pieces on the current stream, evaluated by whichever path is selected (FP64 or INT8) and read back
        asynchronously; one synchronisation at the end.  The C-ABI twin for the FP64 path is mgb_series_posterior_host."""
        if x_host.is_cuda or x_host.dtype != F64 or x_host.dim() != 2:
            raise RuntimeError("posterior_host expects a 2-D float64 host tensor")
        m = x_host.shape[0]
        mean_host = mean_host if mean_host is not None else torch.empty(m, dtype=F64).pin_memory()
        var_host = var_host if var_host is not None else torch.empty(m, dtype=F64).pin_memory()
        for c0 in range(0, m, rows_per_copy):
            c1 = min(m, c0 + rows_per_copy)
            xd = x_host[c0:c1].to(self.device, non_blocking=True)
            mean, var = self.posterior(xd)
            mean_host[c0:c1].copy_(mean, non_blocking=True)
            var_host[c0:c1].copy_(var, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return mean_host, var_host

    def handle(self, max_chunk: int = 0) -> "PosteriorHandle":
        """Persistent C-ABI handle over the fitted state (no upload, no allocation per evaluation)."""
        return PosteriorHandle(self, max_chunk)

    @classmethod
    def from_model(cls, model, **kwargs):
        """Fitted SingleTaskGP-like model (covar_module[.base_kernel, .outputscale], likelihood.noise,
        mean_module.constant, train_inputs, train_targets; bo_sphere.py:211-233) -> fitted ExactGPPosterior."""
        from .botorch_adapter import FusedExactGP

        kernel, x, y, s, noise, mean_const = FusedExactGP.read_model(model)
        return cls(kernel, x, y, outputscale=s, noise=noise, mean_const=mean_const, **kwargs).fit()

    def _ensure_packed_i8(self):
        """Sliced L^-1 of the INT8 path: cut from ``rinv`` on the rank that did the fit, received through ``load_state``
        on its peers."""
        if getattr(self, "alpha", None) is None:
            raise RuntimeError("the INT8 path needs alpha (fit() locally, or load_state(..., packed_i8=, alpha=))")
        if self._packed_i8 is not None:
            return
        if getattr(self, "rinv", None) is None:
            raise RuntimeError("the INT8 path needs L^-1 of a local fit() or the sliced factor of a peer (load_state)")
        lib = _lib.load()
        n = self.rinv.shape[0]
        pre = self._i8_prefix()
        self._packed_i8 = torch.empty(getattr(lib, pre + "pack_bytes")(n), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(getattr(lib, pre + "pack")(_lib.ptr(self.rinv), n, self.rinv.stride(0), _lib.ptr(self._packed_i8),
                                                  _lib.stream_ptr(self.device)), pre + "pack")

    def _i8_prefix(self):
        if self.int8_impl == "library":
            if not _lib.load().mgb_int8_available():
                raise _lib.MgbError("libmgb was built without the CUTLASS headers: the library INT8 path is not available")
            return "mgb_posterior_int8_"
        return "mgb_posterior_i8f_"

    def _posterior_int8(self, x, mean, var, chunk):
        """Opt-in INT8 tensor-core path: row-major float64 kernel rows from the Gram kernel, cut into seven int8 slices,
        slice products against the sliced L^-1 on the INT8 tensor cores, recombined in float64 (fused: 8-bit slices,
        ~2e-13 of max var; library: 7-bit slices, ~2e-11).
        int8_impl "fused": one hand-written tcgen05 kernel per block (csrc/int8_fused.cu); "library": seven CUTLASS GEMMs
        per column block + a recombination pass (csrc/int8_posterior.cu)."""
        lib = _lib.load()
        pre = self._i8_prefix()
        m, n = x.shape[0], self.train_prepared.shape[0]
        st = _lib.stream_ptr(self.device)
        self._ensure_packed_i8()
        if self.int8_impl == "fused":
            chunk = min(INT8_FUSED_CHUNK, (m + 127) // 128 * 128)
        else:
            chunk = min(INT8_CHUNK, (m + 255) // 256 * 256)
        need = getattr(lib, pre + "workspace_bytes")(chunk, n)
        if self._ws_i8 is None or self._ws_i8.numel() < need:
            self._ws_i8 = torch.empty(need, dtype=torch.uint8, device=self.device)
        reduce = getattr(lib, pre + "reduce")
        forward = None
        if not self.is_series:
            forward = self.kernel.forward
        if self.is_series and self.int8_impl == "fused":
            # candidates -> kernel rows -> slices -> contraction in one C call per chunk (mgb_series_posterior_i8f)
            need = lib.mgb_series_posterior_i8f_workspace_bytes(chunk, n)
            if self._ws_i8 is None or self._ws_i8.numel() < need:
                self._ws_i8 = torch.empty(need, dtype=torch.uint8, device=self.device)
            d = self.spec.desc(self.coef.shape[-1])
            with torch.cuda.device(self.device):
                for c0 in range(0, m, chunk):
                    mc = min(chunk, m - c0)
                    xs = x[c0:c0 + mc]
                    rc = lib.mgb_series_posterior_i8f(C.byref(d), _lib.ptr(xs), mc, xs.stride(0), _lib.ptr(self.train_prepared), n,
                                                      self.train_prepared.stride(0), _lib.ptr(self.coef), _lib.ptr(self._packed_i8),
                                                      _lib.ptr(self.alpha), float(self.kss_value), self.mean_const,
                                                      _lib.ptr(mean[c0:]), _lib.ptr(var[c0:]), _lib.ptr(self._ws_i8),
                                                      self._ws_i8.numel(), st)
                    _lib.check(rc, "mgb_series_posterior_i8f")
            return mean, var
        with torch.cuda.device(self.device):
            for c0 in range(0, m, chunk):
                mc = min(chunk, m - c0)
                if forward is None:
                    k = _ops.series_gram(self.spec, x[c0:c0 + mc], self.train_prepared, self.coef)   # carries the outputscale
                else:
                    k = (self.outputscale * forward(x[c0:c0 + mc], self.train_prepared)).contiguous()
                rc = reduce(_lib.ptr(k), mc, n, k.stride(0), _lib.ptr(self._packed_i8), _lib.ptr(self.alpha),
                            float(self.kss_value), self.mean_const, _lib.ptr(mean[c0:]), _lib.ptr(var[c0:]),
                            _lib.ptr(self._ws_i8), self._ws_i8.numel(), st)
                _lib.check(rc, pre + "reduce")
                del k
        return mean, var

    def _posterior_generic(self, x, mean, var, chunk):
        """Row-major Gram chunk from ``kernel.forward`` -> re-tile -> fused reduction."""
        lib = _lib.load()
        m, n = x.shape[0], self.train_prepared.shape[0]
        chunk = min(chunk, 8192)                      # quadrature Gram chunks are expensive to hold row-major
        kb = torch.zeros(lib.mgb_posterior_tiled_bytes(chunk, n) // 8, dtype=F64, device=self.device)
        part = torch.empty(max(1, lib.mgb_posterior_reduce_workspace_bytes(chunk, n) // 8), dtype=F64, device=self.device)
        kss = torch.full((chunk,), self.kss_value, dtype=F64, device=self.device)
        st = _lib.stream_ptr(self.device)
        forward = self.kernel.forward
        if self.tabulate:
            from .tabulated import TabulatedRadialKernel
            forward = TabulatedRadialKernel.for_inputs(self.kernel, x, self.train_prepared).forward
        with torch.cuda.device(self.device):
            for c0 in range(0, m, chunk):
                mc = min(chunk, m - c0)
                k = (self.outputscale * forward(x[c0:c0 + mc], self.train_prepared)).contiguous()
                _lib.check(lib.mgb_posterior_retile(_lib.ptr(k), mc, n, k.stride(0), _lib.ptr(kb), st), "mgb_posterior_retile")
                _lib.check(lib.mgb_posterior_reduce(_lib.ptr(kb), mc, n, _lib.ptr(self._packed), _lib.ptr(kss),
                                                    self.mean_const, _lib.ptr(mean[c0:]), _lib.ptr(var[c0:]),
                                                    _lib.ptr(part), part.numel() * 8, st), "mgb_posterior_reduce")
        return mean, var

    # ---- differentiable path for the trust-region inner loop -------------------------------------
    def posterior_autograd(self, x: torch.Tensor):
        """Mean / variance for a SMALL block of candidates with autograd to ``x`` (first and, for single-factor
        kernels, second order).  This is what the reference's pymanopt cost / egrad / ehess closures differentiate
        (manifold_optimization/manifold_optimize.py:184-197; pymanopt_addons/tools/autodiff/_pytorch.py:89-116):
        the kernel row goes through the CUDA series op (custom backward + double backward), the N x N algebra
        through torch.  Use ``posterior`` (fused, no autograd) for scoring large candidate sets."""
        if self._packed is None:
            raise RuntimeError("call fit() (or load_state) first")
        _lib.require_cuda_f64(x)
        squeeze = x.dim() == 3
        if squeeze:
            if x.shape[-2] != 1:
                raise NotImplementedError("q > 1 joint posteriors are not provided by the analytic acquisition path")
            x = x.squeeze(-2)
        xp = self._prepared(x)
        if self.is_series and self.spec.n_factors == 1:
            ks = _ops.series_gram(self.spec, xp, self.train_prepared, self.coef)  # already scaled by outputscale
        else:   # multi-factor (torus) kernels differentiate factor by factor inside their own forward
            with _frozen_parameters(self.kernel):     # the fit is fixed here: differentiate with respect to x only
                ks = self.outputscale * self.kernel.forward(xp, self.train_prepared)
        mean = self.mean_const + ks @ self.alpha
        v = ks @ self.rinv.t()                                                    # (L^-1 k*)^T, rinv is lower triangular
        var = self.kss_value - (v * v).sum(-1)
        return mean, var


    # ---- fused EI value + gradient + Hessian-vector product (trust-region inner loop) ----------------------
    @torch.no_grad()
    def ei_value_grad_hvp(self, x: torch.Tensor, best_f: float, maximize: bool = True, v: Optional[torch.Tensor] = None):
        """Analytic EI, dEI/dx and (with directions ``v``) the Hessian-vector product d2EI/dx2 v at every candidate of
        ``x`` ((b, D) or botorch's (b, 1, D)) in ONE launch (csrc/acquisition.cu) - the three quantities the pymanopt
        cost / egrad / ehess closures of the reference obtain from botorch + torch autograd
        (manifold_optimize.py:184-217).  Series kernels (sphere, SO(3), circle, torus in (cos, sin) coordinates); other
        kernels go through ``posterior_autograd``.  Returns (ei, grad, hvp or None) shaped like ``x``."""
        if self._packed is None or getattr(self, "rinv", None) is None:
            raise RuntimeError("call fit() first")
        _lib.require_cuda_f64(x, v)
        if not self.is_series:
            return self._radial_ei_value_grad_hvp(x, best_f, maximize, v)
        shape = x.shape
        xf = x.reshape(-1, shape[-1]).contiguous()
        if x.dim() == 3 and shape[-2] != 1:
            raise NotImplementedError("q > 1 joint posteriors are not provided by the analytic acquisition path")
        vf = v.reshape(-1, shape[-1]).contiguous() if v is not None else None
        b, w = xf.shape
        if w != self.spec.factor_width * self.spec.n_factors:
            raise RuntimeError("expected %d coordinates (torus points as (cos, sin) pairs, not angles)"
                               % (self.spec.factor_width * self.spec.n_factors))
        n = self.train_prepared.shape[0]
        ei = torch.empty(b, dtype=F64, device=self.device)
        grad = torch.empty((b, w), dtype=F64, device=self.device)
        hvp = torch.empty((b, w), dtype=F64, device=self.device) if vf is not None else None
        lib = _lib.load()
        d = self.spec.desc(self.coef.shape[-1])
        with torch.cuda.device(self.device):
            rc = lib.mgb_series_acquisition_ei(C.byref(d), _lib.ptr(xf), b, xf.stride(0), _lib.ptr(vf),
                                               vf.stride(0) if vf is not None else 0, _lib.ptr(self.train_prepared), n,
                                               self.train_prepared.stride(0), _lib.ptr(self.coef), _lib.ptr(self.rinv),
                                               self.rinv.stride(0), _lib.ptr(self.rinv_t), self.rinv_t.stride(0),
                                               _lib.ptr(self.alpha), float(self.kss_value), self.mean_const, float(best_f),
                                               1 if maximize else 0, _lib.ptr(ei), _lib.ptr(grad), _lib.ptr(hvp), None, None,
                                               _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_series_acquisition_ei")
        return ei.reshape(shape[:-1] if x.dim() == 2 else shape[:-2]), grad.reshape(shape), \
            (hvp.reshape(shape) if hvp is not None else None)


    def _scratch(self, doubles):
        """Cached device scratch for the b x n intermediates of the fused acquisition step (grows, never shrinks)."""
        buf = getattr(self, "_scratch_buf", None)
        if buf is None or buf.numel() < doubles or buf.device != self.device:
            buf = torch.empty(max(int(doubles), 1), dtype=F64, device=self.device)
            self._scratch_buf = buf
        return buf

    def _pair_scalars(self, kind, xf, xt):
        """Flat (b * n) pair scalars from one launch (mgb_pair_scalars)."""
        lib = _lib.load()
        b, n = xf.shape[0], xt.shape[0]
        out0 = torch.empty(b * n, dtype=F64, device=self.device)
        out1 = torch.empty(b * n, dtype=F64, device=self.device) if kind >= 2 else None
        with torch.cuda.device(self.device):
            rc = lib.mgb_pair_scalars(kind, xf.shape[-1], _lib.ptr(xf), b, xf.stride(0), _lib.ptr(xt), n, xt.stride(0),
                                      _lib.ptr(out0), _lib.ptr(out1), _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_pair_scalars")
        return out0, out1

    def _spd2_ei_value_grad(self, x, best_f, maximize, v):
        """SPD(2) kernels: value + gradient (the reference's SPD trust region uses a finite-difference Hessian)."""
        if v is not None:
            raise NotImplementedError("SPD(2): Hessian-vector products are not provided (bo_spd.py uses approx_hessian=True)")
        shape = x.shape
        if x.dim() == 3 and shape[-2] != 1:
            raise NotImplementedError("q > 1 joint posteriors are not provided by the analytic acquisition path")
        xf = x.reshape(-1, shape[-1]).contiguous()
        xt = self.train_prepared
        if xf.shape[-1] != 3 or xt.shape[-1] != 3:
            raise RuntimeError("SPD(2) points are Mandel 3-vectors")
        chol = int(self.kernel.use_cholesky)
        tcoef, rscale, bscale, bv, bw = self._nodes
        b, n = xf.shape[0], xt.shape[0]
        ei = torch.empty(b, dtype=F64, device=self.device)
        grad = torch.empty((b, 3), dtype=F64, device=self.device)
        work = self._scratch(5 * b * n)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.mgb_spd2_acquisition_ei_step(chol, _lib.ptr(xf), b, xf.stride(0), _lib.ptr(xt), n, xt.stride(0),
                                                  _lib.ptr(tcoef), _lib.ptr(rscale), _lib.ptr(bscale), tcoef.numel(),
                                                  _lib.ptr(bv), _lib.ptr(bw), bv.numel(), self.outputscale,
                                                  _lib.ptr(self.rinv), self.rinv.stride(0), _lib.ptr(self.rinv_t),
                                                  self.rinv_t.stride(0), _lib.ptr(self.alpha), float(self.kss_value),
                                                  self.mean_const, float(best_f), 1 if maximize else 0, _lib.ptr(work),
                                                  work.numel(), _lib.ptr(ei), _lib.ptr(grad), _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_spd2_acquisition_ei_step")
        return ei.reshape(shape[:-1] if x.dim() == 2 else shape[:-2]), grad.reshape(shape), None

    def _radial_ei_value_grad_hvp(self, x, best_f, maximize, v):
        """Hyperbolic (H^1..H^3) and Euclidean kernels: profile and its derivatives at the b x n pair scalars from the
        device quadrature (mgb_radial_profile), then one launch of the radial acquisition kernel."""
        nodes = getattr(self.kernel, "nodes", None)
        if nodes is None:
            raise NotImplementedError("fused acquisition derivatives: series, hyperbolic, Euclidean and SPD(2) kernels")
        if hasattr(self.kernel, "use_cholesky"):
            return self._spd2_ei_value_grad(x, best_f, maximize, v)
        nd = self._nodes
        shape = x.shape
        if x.dim() == 3 and shape[-2] != 1:
            raise NotImplementedError("q > 1 joint posteriors are not provided by the analytic acquisition path")
        xf = x.reshape(-1, shape[-1]).contiguous()
        vf = v.reshape(-1, shape[-1]).contiguous() if v is not None else None
        xt = self.train_prepared
        if len(nd) == 5:                                   # hyperbolic: (tval, tcoef, s0, ds, Ps)
            tval, tcoef, s0, ds, ps = nd
            kind = int(self.kernel.dim) - 1
            if self.kernel.dim not in (1, 2, 3):
                raise NotImplementedError
        elif len(nd) == 2:                                 # Euclidean: (tval, tcoef)
            tval, tcoef = nd
            s0 = ds = 0.0
            ps = 0
            kind = _ops.KIND_EUCLID
        else:
            raise NotImplementedError("fused acquisition derivatives: series, hyperbolic and Euclidean kernels")
        b, w = xf.shape
        if w > 4 or w != xt.shape[-1]:
            raise NotImplementedError("radial acquisition kernel: rows up to 4 wide")
        n = xt.shape[0]
        ei = torch.empty(b, dtype=F64, device=self.device)
        grad = torch.empty((b, w), dtype=F64, device=self.device)
        hvp = torch.empty((b, w), dtype=F64, device=self.device) if vf is not None else None
        work = self._scratch(4 * b * n)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.mgb_radial_acquisition_ei_step(kind, w, _lib.ptr(xf), b, xf.stride(0), _lib.ptr(vf),
                                                    vf.stride(0) if vf is not None else 0, _lib.ptr(xt), n, xt.stride(0),
                                                    _lib.ptr(tval), _lib.ptr(tcoef), tval.numel(), float(s0), float(ds), int(ps),
                                                    self.outputscale, _lib.ptr(self.rinv), self.rinv.stride(0),
                                                    _lib.ptr(self.rinv_t), self.rinv_t.stride(0), _lib.ptr(self.alpha),
                                                    float(self.kss_value), self.mean_const, float(best_f),
                                                    1 if maximize else 0, _lib.ptr(work), work.numel(), _lib.ptr(ei),
                                                    _lib.ptr(grad), _lib.ptr(hvp), _lib.stream_ptr(self.device))
        _lib.check(rc, "mgb_radial_acquisition_ei_step")
        return ei.reshape(shape[:-1] if x.dim() == 2 else shape[:-2]), grad.reshape(shape), \
            (hvp.reshape(shape) if hvp is not None else None)


class PosteriorHandle:
    """Persistent C-ABI handle (include/mgb.h: mgb_posterior_create* / _eval_host / _eval_device / _destroy) over a
    fitted series-kernel posterior.  The handle borrows the device tensors of the ExactGPPosterior it was made from
    (kept alive here); evaluations reuse its device buffers, workspace, pinned staging block and streams."""

    def __init__(self, gp: "ExactGPPosterior", max_chunk: int = 0):
        if not gp.is_series or gp._packed is None:
            raise RuntimeError("PosteriorHandle: fit() a series-kernel posterior first")
        lib = _lib.load()
        self._lib = lib
        self._keep = (gp.train_prepared, gp.coef, gp._packed)
        self.device = gp.device
        self.n = gp.train_prepared.shape[0]
        self.width = gp.train_prepared.shape[1]
        d = gp.spec.desc(gp.coef.shape[-1])
        h = C.c_void_p()
        index = gp.device.index if gp.device.index is not None else torch.cuda.current_device()
        torch.cuda.current_stream(gp.device).synchronize()      # the state was produced on torch's stream
        rc = lib.mgb_posterior_create_from_device(C.byref(d), _lib.ptr(gp.train_prepared), self.n, gp.train_prepared.stride(0),
                                                  _lib.ptr(gp.coef), _lib.ptr(gp._packed), float(gp.kss_value),
                                                  float(gp.mean_const), int(max_chunk), int(index), C.byref(h))
        _lib.check(rc, "mgb_posterior_create_from_device")
        self._h = h
        self.int8 = False
        self._i8_state = None
        self._i8_auto = gp._int8_mode == "auto"          # auto: calls below INT8_AUTO_MIN_M candidates use the FP64 kernel
        if gp.int8 and gp.int8_impl == "fused" and (not self._i8_auto or getattr(gp, "rinv", None) is not None
                                                    or gp._packed_i8 is not None):
            # the handle evaluates through csrc/int8_fused.cu as well: it borrows the sliced factor and alpha
            gp._ensure_packed_i8()
            torch.cuda.current_stream(gp.device).synchronize()
            self._i8_state = (gp._packed_i8, gp.alpha)
            self._keep = self._keep + self._i8_state
            self._set_int8(True)

    def _set_int8(self, on: bool):
        if on == self.int8:
            return
        if on:
            rc = self._lib.mgb_posterior_use_int8(self._h, _lib.ptr(self._i8_state[0]), _lib.ptr(self._i8_state[1]))
        else:
            rc = self._lib.mgb_posterior_use_int8(self._h, None, None)
        _lib.check(rc, "mgb_posterior_use_int8")
        self.int8 = on

    def _select_path(self, m: int):
        if self._i8_auto and self._i8_state is not None:
            self._set_int8(m >= INT8_AUTO_MIN_M)

    @property
    def chunk(self) -> int:
        return int(self._lib.mgb_posterior_chunk(self._h))

    def eval_host(self, x_host, mean_host=None, var_host=None):
        """x_host: (m, D) float64 CPU tensor or numpy array (pinned memory makes the large-m copies asynchronous)."""
        if self._h is None:
            raise RuntimeError("handle destroyed")
        x = torch.as_tensor(x_host)
        if x.is_cuda or x.dtype != F64 or x.dim() != 2 or x.shape[1] != self.width or not x.is_contiguous():
            raise ValueError("eval_host expects a contiguous (m, %d) float64 host array" % self.width)
        m = x.shape[0]
        self._select_path(m)
        mean = mean_host if mean_host is not None else torch.empty(m, dtype=F64)
        var = var_host if var_host is not None else torch.empty(m, dtype=F64)
        rc = self._lib.mgb_posterior_eval_host(self._h, C.c_void_p(x.data_ptr()), m, C.c_void_p(mean.data_ptr()),
                                               C.c_void_p(var.data_ptr()))
        _lib.check(rc, "mgb_posterior_eval_host")
        return mean, var

    def eval_device(self, x: torch.Tensor):
        if self._h is None:
            raise RuntimeError("handle destroyed")
        _lib.require_cuda_f64(x)
        m = x.shape[0]
        self._select_path(m)
        mean = torch.empty(m, dtype=F64, device=x.device)
        var = torch.empty(m, dtype=F64, device=x.device)
        with torch.cuda.device(x.device):
            rc = self._lib.mgb_posterior_eval_device(self._h, _lib.ptr(x), m, x.stride(0), _lib.ptr(mean), _lib.ptr(var),
                                                     _lib.stream_ptr(x.device))
        _lib.check(rc, "mgb_posterior_eval_device")
        return mean, var

    def destroy(self):
        if getattr(self, "_h", None) is not None:
            self._lib.mgb_posterior_destroy(self._h)
            self._h = None
            self._keep = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()


class ExpectedImprovement:
    """Analytic EI with botorch's call convention: ``acq(X)`` with ``X`` of shape ``b x 1 x D`` -> ``b``."""

    def __init__(self, model: ExactGPPosterior, best_f, maximize: bool = True):
        self.model = model
        self.best_f = float(best_f)
        self.maximize = maximize

    def __call__(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            mean, var = self.model.posterior_autograd(x)     # trust-region inner loop: value + gradient (+ HVP)
        else:
            mean, var = self.model.posterior(x)              # scoring: fused kernels, no autograd
            ei = torch.empty_like(mean)
            lib = _lib.load()
            with torch.cuda.device(mean.device):
                rc = lib.mgb_expected_improvement(_lib.ptr(mean), _lib.ptr(var), mean.numel(), self.best_f,
                                                  1 if self.maximize else 0, _lib.ptr(ei), _lib.stream_ptr(mean.device))
            _lib.check(rc, "mgb_expected_improvement")
            return ei
        return expected_improvement(mean, var, self.best_f, self.maximize)


def _ei_value_grad_hvp(self, x, v=None):
    """(EI, dEI/dx, d2EI/dx2 v) at every candidate in one fused launch; see ExactGPPosterior.ei_value_grad_hvp."""
    return self.model.ei_value_grad_hvp(x, self.best_f, self.maximize, v)


ExpectedImprovement.value_grad_hvp = _ei_value_grad_hvp


def expected_improvement(mean, var, best_f, maximize=True):
    sigma = var.clamp_min(1e-9).sqrt()
    u = (mean - best_f) / sigma
    if not maximize:
        u = -u
    pdf = torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
    cdf = 0.5 * torch.erfc(-u / math.sqrt(2.0))
    return sigma * (pdf + u * cdf)


# --------------------------------------------------------------------------------------------------
# multi-GPU: rows of X* sharded over ranks, state broadcast once, results all-gathered
# --------------------------------------------------------------------------------------------------
def shard_bounds(m: int, world: int, rank: int):
    """Contiguous row block of rank ``rank``: sizes differ by at most one."""
    base, rem = divmod(m, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedPosterior:
    """One process per GPU.  The rank that did the fit (``src``) broadcasts X_train, the packed factor
    (L^-1 and alpha) and the series coefficients; every rank evaluates its contiguous block of candidate
    rows with no data-path collective; mean / variance are all-gathered (or left sharded).

    ``compute`` is injectable so that the partition / collective logic can be exercised with the gloo backend
    on CPU (tests/test_sharding_gloo.py) - the product default is the CUDA posterior."""

    def __init__(self, group=None, compute: Optional[Callable] = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.compute = compute

    def broadcast_state(self, tensors: dict, src: int = 0, meta=None):
        """Broadcast a dict of tensors from ``src``.

        ``meta`` = {name: (shape, dtype)} known on EVERY rank (e.g. derived from n and the kernel): the state then
        travels as ONE flat buffer per dtype - one NCCL broadcast over NVLink, no pickled shape exchange, no host
        synchronisation, receive buffers cached across calls (a refit re-broadcasts into the same memory).  Without
        ``meta`` the shapes are exchanged first (``broadcast_object_list``) and each tensor is broadcast on its own."""
        dist = self.dist
        if self.world == 1:
            return tensors
        if meta is not None:
            return self._broadcast_flat(tensors, src, meta)
        names = sorted(tensors.keys()) if self.rank == src else None
        meta = [names, {k: (tuple(tensors[k].shape), str(tensors[k].dtype)) for k in names}] if self.rank == src else [None, None]
        dist.broadcast_object_list(meta, src=src, group=self.group)
        names, shapes = meta
        out = {}
        for k in names:
            shape, dtype = shapes[k]
            if self.rank == src:
                t = tensors[k].contiguous()
            else:
                like = next(iter(tensors.values())) if tensors else None
                dev = like.device if like is not None else self._default_device()
                t = torch.empty(shape, dtype=getattr(torch, dtype.split(".")[-1]), device=dev)
            dist.broadcast(t, src=src, group=self.group)
            out[k] = t
        return out

    def _broadcast_flat(self, tensors, src, meta):
        dist = self.dist
        like = next(iter(tensors.values())) if tensors else None
        dev = like.device if like is not None else self._default_device()
        out = {}
        by_dtype = {}
        for name in sorted(meta):
            shape, dtype = meta[name]
            by_dtype.setdefault(dtype, []).append((name, tuple(shape)))
        cache = self.__dict__.setdefault("_flat_cache", {})
        for dtype, items in by_dtype.items():
            sizes = [int(math.prod(shape)) for _, shape in items]
            # every tensor starts 512 bytes apart at least: the packed operands are read with 16-byte bulk copies
            pads = [(sz + 511) // 512 * 512 for sz in sizes]
            total = sum(pads)
            key = (str(dtype), total, str(dev))
            flat = cache.get(key)
            if flat is None:
                flat = torch.empty(total, dtype=dtype, device=dev)
                cache[key] = flat
            if self.rank == src:
                off = 0
                for (name, shape), sz, pad in zip(items, sizes, pads):
                    flat[off:off + sz].copy_(tensors[name].reshape(-1))
                    off += pad
            dist.broadcast(flat, src=src, group=self.group)
            off = 0
            for (name, shape), sz, pad in zip(items, sizes, pads):
                out[name] = tensors[name] if self.rank == src else flat[off:off + sz].view(shape)
                off += pad
        return out

    def _default_device(self):
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    def evaluate(self, x_local: torch.Tensor, gather: bool = True, sizes=None):
        """x_local: this rank's rows.  Returns (mean, var) for all rows (gather=True) or the local block."""
        mean, var = self.compute(x_local)
        if not gather or self.world == 1:
            return mean, var
        dist = self.dist
        if sizes is None:
            cnt = torch.tensor([x_local.shape[0]], dtype=torch.int64, device=mean.device)
            allc = [torch.zeros_like(cnt) for _ in range(self.world)]
            dist.all_gather(allc, cnt, group=self.group)
            sizes = [int(c.item()) for c in allc]
        pad = max(sizes)
        cache = self.__dict__.setdefault("_gather_cache", {})
        key = (pad, self.world, str(mean.device), mean.dtype)
        bufs = cache.get(key)
        if bufs is None:
            bufs = (torch.empty((self.world, pad), dtype=mean.dtype, device=mean.device),
                    torch.empty((self.world, pad), dtype=mean.dtype, device=mean.device))
            cache[key] = bufs
        outs = []
        for src_t, buf in ((mean, bufs[0]), (var, bufs[1])):
            if src_t.shape[0] == pad:
                piece = src_t.contiguous()
            else:
                piece = torch.zeros(pad, dtype=src_t.dtype, device=src_t.device)
                piece[:src_t.shape[0]] = src_t
            dist.all_gather_into_tensor(buf.view(-1), piece, group=self.group)
            if all(sz == pad for sz in sizes):
                outs.append(buf.view(-1))                        # equal shards: the gathered buffer IS the result
            else:
                outs.append(torch.cat([buf[r, :sz] for r, sz in enumerate(sizes)]))
        return outs[0], outs[1]
