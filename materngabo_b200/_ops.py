"""torch.autograd bindings of the C-ABI kernels (device pointers + current stream via ctypes).

Autograd structure: the device ops are functions of (x1, x2, w) where ``w`` is the tiny coefficient /
node-weight tensor computed by ``_weights`` from the hyper-parameters.  ``backward`` returns dL/dw (a few
dozen numbers reduced on the device) and dL/dx1, dL/dx2; torch autograd carries dL/dw on to
``raw_lengthscale`` / ``raw_nu``.  First-order input gradients are themselves differentiable for
single-factor series kernels (``SeriesGramGradX``), which is what the trust-region Hessian-vector products
of the reference need (pymanopt_addons/tools/autodiff/_pytorch.py:92-116).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib

F64 = torch.float64


@dataclass(frozen=True)
class SeriesSpec:
    n_factors: int
    factor_width: int
    basis: int
    scale: float = 1.0
    shift: float = 0.0
    lo: float = -1.0
    hi: float = 1.0
    pair_square: int = 0

    def desc(self, n_terms: int) -> _lib.SeriesDesc:
        return _lib.SeriesDesc(self.n_factors, self.factor_width, n_terms, self.basis,
                               self.scale, self.shift, self.lo, self.hi, self.pair_square)


def _c2d(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    assert t.dim() == 2
    return t


# --------------------------------------------------------------------------------------------------
# spectral series
# --------------------------------------------------------------------------------------------------
# Which forward kernel serves a single-factor monomial series block (sphere, SO(3)): MGB_GRAM_PATH =
#   fp64 (default): the float64 kernels (DMMA / DFMA inner products);
#   int8: the INT8 tensor-core kernel (csrc/series_i8.cu) whenever it applies.  Opt-in: it is exact (inner products
#   rounded once; tests/test_gpu_series_i8.py) but measured SLOWER on B200 - 3.6e11 entries/s against 5.3e11 (S^10,
#   1 M x 2 k, sustained): unpacking seven int32 planes costs more issue slots than the inner product costs FP64 slots
#   (DESIGN.md section 11).
def _gram_path() -> str:
    v = os.environ.get("MGB_GRAM_PATH", "fp64").lower()
    if v not in ("fp64", "int8"):
        raise ValueError("MGB_GRAM_PATH must be fp64 or int8 (got %r)" % v)
    return v


def series_gram_i8(spec, x1, x2, coef, planes: bool = False):
    """K(x1, x2) of a single-factor monomial series kernel through mgb_series_gram_fwd_i8 (inner products on the INT8
    tensor cores, error-free digit slicing).  planes=True (tests): also returns the raw int32 accumulator planes
    [row panel][column tile][7][128][32]."""
    lib = _lib.load()
    m, n = x1.shape[0], x2.shape[0]
    d = spec.desc(coef.shape[-1])
    if not lib.mgb_series_gram_i8_supported(C.byref(d)):
        raise _lib.MgbError("series_gram_i8: single-factor monomial series with <= 12 terms and rows <= 16 wide only")
    k = torch.empty((m, n), dtype=F64, device=x1.device)
    ws = torch.empty(max(int(lib.mgb_series_gram_i8_workspace_bytes(n)), 256), dtype=torch.uint8, device=x1.device)
    with torch.cuda.device(x1.device):
        if planes:
            pl = torch.zeros(((m + 127) // 128, (n + 31) // 32, 7, 128, 32), dtype=torch.int32, device=x1.device)
            rc = lib.mgb_series_gram_fwd_i8_debug(C.byref(d), _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                                  _lib.ptr(coef), _lib.ptr(k), max(n, 1), _lib.ptr(ws), ws.numel(), _lib.ptr(pl),
                                                  _lib.stream_ptr(x1.device))
        else:
            rc = lib.mgb_series_gram_fwd_i8(C.byref(d), _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                            _lib.ptr(coef), _lib.ptr(k), max(n, 1), _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr(x1.device))
    _lib.check(rc, "mgb_series_gram_fwd_i8")
    return (k, pl) if planes else k


def _series_fwd_raw(spec, x1, x2, coef):
    lib = _lib.load()
    m, n = x1.shape[0], x2.shape[0]
    if _gram_path() == "int8" and m > 0 and n > 0:
        d8 = spec.desc(coef.shape[-1])
        if lib.mgb_series_gram_i8_supported(C.byref(d8)):
            return series_gram_i8(spec, x1, x2, coef)
    k = torch.empty((m, n), dtype=F64, device=x1.device)
    d = spec.desc(coef.shape[-1])
    with torch.cuda.device(x1.device):
        rc = lib.mgb_series_gram_fwd(C.byref(d), _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                     _lib.ptr(coef), _lib.ptr(k), max(n, 1), _lib.stream_ptr(x1.device))
    _lib.check(rc, "mgb_series_gram_fwd")
    return k


def _series_bwd_raw(spec, x1, x2, coef, g, want_coef, want_x1, want_x2):
    lib = _lib.load()
    m, n = x1.shape[0], x2.shape[0]
    dwidth = spec.n_factors * spec.factor_width
    gcoef = torch.zeros_like(coef) if want_coef else None
    # zeros, not empty: the library returns early for an empty opposite block without writing the outputs
    gx1 = torch.zeros((m, dwidth), dtype=F64, device=x1.device) if want_x1 else None
    gx2 = torch.zeros((n, dwidth), dtype=F64, device=x1.device) if want_x2 else None
    d = spec.desc(coef.shape[-1])
    with torch.cuda.device(x1.device):
        rc = lib.mgb_series_gram_bwd(C.byref(d), _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                     _lib.ptr(coef), _lib.ptr(g), g.stride(0), g.stride(1), _lib.ptr(gcoef),
                                     _lib.ptr(gx1), _lib.ptr(gx2), _lib.stream_ptr(x1.device))
    _lib.check(rc, "mgb_series_gram_bwd")
    return gcoef, gx1, gx2


def _series_dx_raw(spec, order, x1, x2, coef, g, want1, want2):
    lib = _lib.load()
    m, n = x1.shape[0], x2.shape[0]
    dwidth = spec.n_factors * spec.factor_width
    o1 = torch.zeros((m, dwidth), dtype=F64, device=x1.device) if want1 else None
    o2 = torch.zeros((n, dwidth), dtype=F64, device=x1.device) if want2 else None
    d = spec.desc(coef.shape[-1])
    with torch.cuda.device(x1.device):
        rc = lib.mgb_series_gram_dx(C.byref(d), order, _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                    _lib.ptr(coef), _lib.ptr(g), g.stride(0), g.stride(1), _lib.ptr(o1), _lib.ptr(o2),
                                    _lib.stream_ptr(x1.device))
    _lib.check(rc, "mgb_series_gram_dx")
    return o1, o2


class SeriesGramGradX(torch.autograd.Function):
    """(x1, x2, coef, G) -> (dL/dcoef, dL/dx1, dL/dx2) of L = sum(K * G); differentiable again with respect
    to x1 and G (single-factor kernels) so that Hessian-vector products through the kernel work."""

    @staticmethod
    def forward(ctx, spec, x1, x2, coef, g, want_coef, want_x1, want_x2):
        gcoef, gx1, gx2 = _series_bwd_raw(spec, x1, x2, coef, g, want_coef, want_x1, want_x2)
        ctx.spec = spec
        ctx.save_for_backward(x1, x2, coef, g)
        ctx.mark_non_differentiable(*[t for t in (gcoef, gx2) if t is not None])
        outs = tuple(t if t is not None else torch.zeros((), dtype=F64, device=x1.device)
                     for t in (gcoef, gx1, gx2))
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _v_coef, v1, _v2):
        spec = ctx.spec
        x1, x2, coef, g = ctx.saved_tensors
        if spec.n_factors != 1:
            raise NotImplementedError("second-order input gradients are implemented for single-factor kernels")
        if ctx.needs_input_grad[2]:
            # silently returning None here would drop terms of a Hessian-vector product (e.g. k(x, x) with x1 aliasing x2)
            raise NotImplementedError("second-order gradients are provided with respect to x1 and G only "
                                      "(x2 = the training points, constants of the acquisition path); detach x2")
        if v1 is None or v1.dim() != 2:
            return (None,) * 8
        v1 = v1.contiguous()
        proj = v1 @ x2.t()                                  # (x2_j . v_i)
        # d/dG_ij = p'(u_ij) mask scale (x2_j . v_i): first-derivative weights come from the dx kernel with
        # G = one-hot is wasteful; instead use the identity  sum_k v_ik gx1_ik = sum_j G_ij p' s (x2_j.v_i)
        # and differentiate it with respect to G through a derivative-Gram evaluated by the order-1 kernel.
        gG = _derivative_gram(spec, x1, x2, coef) * proj if ctx.needs_input_grad[4] else None
        gx1 = None
        if ctx.needs_input_grad[1]:
            gx1, _ = _series_dx_raw(spec, 2, x1, x2, coef, (g * proj).contiguous(), True, False)
        return None, gx1, None, None, gG, None, None, None


def _derivative_gram(spec, x1, x2, coef):
    """Matrix p'(u_ij) * scale * [lo <= raw u <= hi] built from the order-1 contraction kernel applied to the
    canonical basis of the (small) x2 block: cheap for the acquisition sizes (n = training points)."""
    # one column at a time would be n launches; instead exploit linearity in x2 rows: for D-wide rows the
    # contraction out1[i,:] = sum_j G_ij d_ij x2_j.  Choosing G = diag-selector is O(n) launches, so we use the
    # closed form on the host instead: d_ij via autograd-free polynomial derivative in torch (n is small).
    u_raw = (x1 @ x2.t()) * spec.scale + spec.shift
    u = u_raw.clamp(spec.lo, spec.hi)
    c = coef.reshape(-1)
    t = c.shape[0]
    if spec.basis == _lib.BASIS_MONOMIAL:
        d = torch.zeros_like(u)
        for k in range(t - 1, 0, -1):
            d = d * u + k * c[k]
    else:
        # sum_k c_k T_k'(u) = sum_k c_k k U_{k-1}(u)
        d = torch.zeros_like(u)
        u_prev, u_cur = torch.zeros_like(u), torch.ones_like(u)  # U_{-1}, U_0
        for k in range(1, t):
            d = d + c[k] * k * u_cur
            u_prev, u_cur = u_cur, 2 * u * u_cur - u_prev
    inside = (u_raw >= spec.lo) & (u_raw <= spec.hi)
    return d * spec.scale * inside


class SeriesGram(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, x1, x2, coef):
        ctx.spec = spec
        ctx.save_for_backward(x1, x2, coef)
        return _series_fwd_raw(spec, x1, x2, coef)

    @staticmethod
    def backward(ctx, g):
        spec = ctx.spec
        x1, x2, coef = ctx.saved_tensors
        need_x1, need_x2, need_c = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        if not (need_x1 or need_x2 or need_c):
            return None, None, None, None
        if not g.is_contiguous() and not (g.dim() == 2 and g.stride(1) in (0, 1)):
            g = g.contiguous()
        if g.stride(0) == 0 or g.stride(1) == 0:
            g = g.contiguous()
        gcoef, gx1, gx2 = SeriesGramGradX.apply(spec, x1, x2, coef, g, need_c, need_x1, need_x2)
        return None, (gx1 if need_x1 else None), (gx2 if need_x2 else None), (gcoef if need_c else None)


def series_gram(spec: SeriesSpec, x1: torch.Tensor, x2: torch.Tensor, coef: torch.Tensor) -> torch.Tensor:
    """K (m x n) for 2-D inputs."""
    _lib.require_cuda_f64(x1, x2, coef)
    return SeriesGram.apply(spec, _c2d(x1), _c2d(x2), coef.contiguous())


_DENSE_WS = {}


def _dense_workspace(device, nbytes: int) -> torch.Tensor:
    """Cached scratch for the multi-CTA dense path (csrc/dense_spd.cu), one buffer per device, grows, never shrinks."""
    key = str(device)
    buf = _DENSE_WS.get(key)
    if buf is None or buf.numel() * 8 < nbytes:
        buf = torch.empty((nbytes + 7) // 8, dtype=F64, device=device)
        _DENSE_WS[key] = buf
    return buf


class SeriesMLL(torch.autograd.Function):
    """Fused negative log marginal likelihood of an exact GP over a single-factor series kernel (csrc/mll.cu):
    (x, y, coef, hyper = [outputscale, noise, mean]) -> (nll, status).  The gradients with respect to ``coef`` and
    ``hyper`` come out of the same launch and are replayed in ``backward``."""

    @staticmethod
    def forward(ctx, spec, x, y, coef, hyper):
        lib = _lib.load()
        t = coef.shape[-1]
        out = torch.empty(4 + t, dtype=F64, device=x.device)
        d = spec.desc(t)
        n = x.shape[0]
        single = n <= lib.mgb_series_mll_single_cta_max_n(spec.factor_width, int(t))
        # the single-CTA kernels always write the flag: no fill launch for it
        status = (torch.empty if single else torch.zeros)(1, dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            if single:
                rc = lib.mgb_series_mll(C.byref(d), _lib.ptr(x), n, x.stride(0), _lib.ptr(y), _lib.ptr(coef),
                                        _lib.ptr(hyper), _lib.ptr(out), _lib.ptr(status), _lib.stream_ptr(x.device))
            else:       # multi-CTA path: Gram kernel + cooperative Cholesky + triangular inverse + gradient contraction
                ws = _dense_workspace(x.device, lib.mgb_dense_mll_workspace_bytes(n))
                rc = lib.mgb_series_mll_large(C.byref(d), _lib.ptr(x), n, x.stride(0), _lib.ptr(y), _lib.ptr(coef),
                                              _lib.ptr(hyper), _lib.ptr(out), _lib.ptr(status), _lib.ptr(ws),
                                              ws.numel() * 8, _lib.stream_ptr(x.device))
        _lib.check(rc, "mgb_series_mll")
        ctx.save_for_backward(out)
        ctx.coef_shape = coef.shape
        ctx.mark_non_differentiable(status)
        return out[0].clone(), status

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g, _gs):
        (out,) = ctx.saved_tensors
        scaled = g * out[1:]                       # one launch for both gradient blocks
        gcoef = scaled[3:].reshape(ctx.coef_shape) if ctx.needs_input_grad[3] else None
        ghyper = scaled[:3] if ctx.needs_input_grad[4] else None
        return None, None, None, gcoef, ghyper


def series_mll_max_n(spec: SeriesSpec, n_terms: int) -> int:
    """Largest n of the SINGLE-CTA fused kernel (everything in shared memory)."""
    return int(_lib.load().mgb_series_mll_single_cta_max_n(spec.factor_width, int(n_terms)))


def dense_large_max_n() -> int:
    """Largest n of the multi-CTA dense path (csrc/dense_spd.cu)."""
    return int(_lib.load().mgb_dense_max_n())


def series_mll(spec: SeriesSpec, x, y, coef, hyper):
    _lib.require_cuda_f64(x, y, coef, hyper)
    return SeriesMLL.apply(spec, _c2d(x), y.contiguous().reshape(-1), coef.contiguous(), hyper.contiguous())


class DenseMLL(torch.autograd.Function):
    """Fused negative log marginal likelihood for a Gram matrix any kernel produced (csrc/mll.cu, dense mode):
    (K, y, hyper = [outputscale, noise, mean]) -> (nll, status) with Kt = outputscale * K + noise I.  dnll/dK and
    dnll/dhyper come out of the same launch; ``backward`` hands dnll/dK to the kernel's own backward."""

    @staticmethod
    def forward(ctx, k, y, hyper):
        lib = _lib.load()
        n = k.shape[0]
        out = torch.empty(4, dtype=F64, device=k.device)
        g = torch.empty((n, n), dtype=F64, device=k.device)
        single = n <= lib.mgb_series_mll_single_cta_max_n(0, 0)
        status = (torch.empty if single else torch.zeros)(1, dtype=torch.int32, device=k.device)
        with torch.cuda.device(k.device):
            if single:
                rc = lib.mgb_dense_mll(_lib.ptr(k), n, k.stride(0), _lib.ptr(y), _lib.ptr(hyper), _lib.ptr(out), _lib.ptr(g), n,
                                       _lib.ptr(status), _lib.stream_ptr(k.device))
            else:
                ws = _dense_workspace(k.device, lib.mgb_dense_mll_workspace_bytes(n))
                rc = lib.mgb_dense_mll_large(_lib.ptr(k), n, k.stride(0), _lib.ptr(y), _lib.ptr(hyper), _lib.ptr(out),
                                             _lib.ptr(g), n, _lib.ptr(status), _lib.ptr(ws), ws.numel() * 8,
                                             _lib.stream_ptr(k.device))
        _lib.check(rc, "mgb_dense_mll")
        ctx.save_for_backward(out, g)
        ctx.mark_non_differentiable(status)
        return out[0].clone(), status

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad, _gs):
        out, g = ctx.saved_tensors
        gk = grad * g if ctx.needs_input_grad[0] else None
        ghyper = grad * out[1:4] if ctx.needs_input_grad[2] else None
        return gk, None, ghyper


def dense_mll_max_n() -> int:
    """Largest n of the single-CTA kernel in dense mode."""
    return int(_lib.load().mgb_series_mll_single_cta_max_n(0, 0))


def dense_mll(k, y, hyper):
    _lib.require_cuda_f64(k, y, hyper)
    if k.dim() != 2 or k.shape[0] != k.shape[1]:
        raise RuntimeError("dense_mll: square Gram matrix expected")
    return DenseMLL.apply(k.contiguous(), y.contiguous().reshape(-1), hyper.contiguous())


def so3_quaternions(*mats, tol: float = 1e-13):
    """Flattened rotations (m x 9) -> unit quaternions (m x 4) for the pair_square form of the SO(3) kernels.
    Returns None when any input is not a rotation to ``tol`` (max |R R^T - I|, |det R - 1| measured on the device):
    the caller then keeps the 9-wide trace form, which is defined for arbitrary matrices as in the reference.
    Reading the defect synchronises the stream once - this path is meant for large Gram blocks."""
    lib = _lib.load()
    dev = mats[0].device
    defect = torch.zeros(1, dtype=F64, device=dev)
    outs = []
    with torch.cuda.device(dev):
        for x in mats:
            x = _c2d(x)
            q = torch.empty((x.shape[0], 4), dtype=F64, device=dev)
            rc = lib.mgb_so3_to_quaternion(_lib.ptr(x), x.shape[0], x.stride(0), _lib.ptr(q), _lib.ptr(defect),
                                           _lib.stream_ptr(dev))
            _lib.check(rc, "mgb_so3_to_quaternion")
            outs.append(q)
    if float(defect) > tol:
        return None
    return outs


# --------------------------------------------------------------------------------------------------
# quadrature kernels
# --------------------------------------------------------------------------------------------------
class HyperbolicGram(torch.autograd.Function):
    """K = sum_a tcoef[a] heat(d_ij, tval[a]); differentiable with respect to tcoef and tval (hyper-parameters)."""

    @staticmethod
    def forward(ctx, dim, x1, x2, tval, tcoef, s0, ds, ps):
        lib = _lib.load()
        m, n = x1.shape[0], x2.shape[0]
        k = torch.empty((m, n), dtype=F64, device=x1.device)
        with torch.cuda.device(x1.device):
            rc = lib.mgb_hyperbolic_gram_fwd(dim, _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                             _lib.ptr(tval), _lib.ptr(tcoef), tval.shape[0], s0, ds, ps,
                                             _lib.ptr(k), max(n, 1), _lib.stream_ptr(x1.device))
        _lib.check(rc, "mgb_hyperbolic_gram_fwd")
        ctx.args = (dim, s0, ds, ps)
        ctx.save_for_backward(x1, x2, tval, tcoef)
        return k

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("input gradients of the hyperbolic quadrature kernels are not implemented yet")
        need_t, need_c = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        if not (need_t or need_c):
            return (None,) * 8
        lib = _lib.load()
        dim, s0, ds, ps = ctx.args
        x1, x2, tval, tcoef = ctx.saved_tensors
        g = g.contiguous()
        gc = gt = None
        with torch.cuda.device(x1.device):
            if need_c:
                gc = torch.zeros_like(tval)
                rc = lib.mgb_hyperbolic_gram_bwd_coef(dim, _lib.ptr(x1), x1.shape[0], x1.stride(0), _lib.ptr(x2),
                                                      x2.shape[0], x2.stride(0), _lib.ptr(tval), tval.shape[0], s0, ds,
                                                      ps, _lib.ptr(g), g.stride(0), _lib.ptr(gc),
                                                      _lib.stream_ptr(x1.device))
                _lib.check(rc, "mgb_hyperbolic_gram_bwd_coef")
            if need_t:
                gt = torch.zeros_like(tval)
                rc = lib.mgb_hyperbolic_gram_bwd_tval(dim, _lib.ptr(x1), x1.shape[0], x1.stride(0), _lib.ptr(x2),
                                                      x2.shape[0], x2.stride(0), _lib.ptr(tval), _lib.ptr(tcoef),
                                                      tval.shape[0], s0, ds, ps, _lib.ptr(g), g.stride(0), _lib.ptr(gt),
                                                      _lib.stream_ptr(x1.device))
                _lib.check(rc, "mgb_hyperbolic_gram_bwd_tval")
        return None, None, None, gt, gc, None, None, None


def hyperbolic_gram(dim, x1, x2, tval, tcoef, s0, ds, ps):
    _lib.require_cuda_f64(x1, x2, tval, tcoef)
    return HyperbolicGram.apply(dim, _c2d(x1), _c2d(x2), tval.contiguous(), tcoef.contiguous(), float(s0), float(ds), int(ps))


def _heat_nodes_raw(dim, derivative, dist, tval, s0, ds, ps):
    lib = _lib.load()
    out = torch.empty((dist.shape[0], tval.shape[0]), dtype=F64, device=dist.device)
    with torch.cuda.device(dist.device):
        rc = lib.mgb_hyperbolic_heat_nodes(dim, int(derivative), _lib.ptr(dist), dist.shape[0], _lib.ptr(tval), tval.shape[0],
                                           float(s0), float(ds), int(ps), _lib.ptr(out), _lib.stream_ptr(dist.device))
    _lib.check(rc, "mgb_hyperbolic_heat_nodes")
    return out


class HeatNodes(torch.autograd.Function):
    """heat(dist[e], tval[a]) -> (E, Pt); differentiable with respect to tval (normaliser of the heat kernels)."""

    @staticmethod
    def forward(ctx, dim, dist, tval, s0, ds, ps):
        ctx.args = (dim, s0, ds, ps)
        ctx.save_for_backward(dist, tval)
        return _heat_nodes_raw(dim, 0, dist, tval, s0, ds, ps)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        dim, s0, ds, ps = ctx.args
        dist, tval = ctx.saved_tensors
        gt = (g * _heat_nodes_raw(dim, 1, dist, tval, s0, ds, ps)).sum(0) if ctx.needs_input_grad[2] else None
        return None, None, gt, None, None, None


def hyperbolic_heat_nodes(dim, dist, tval, s0, ds, ps):
    _lib.require_cuda_f64(dist, tval)
    return HeatNodes.apply(dim, dist.contiguous().reshape(-1), tval.contiguous(), float(s0), float(ds), int(ps))


class EuclideanGram(torch.autograd.Function):
    """K = sum_a tcoef[a] exp(-|x1_i - x2_j|^2 / (4 tval[a])); differentiable with respect to tcoef."""

    @staticmethod
    def forward(ctx, x1, x2, tval, tcoef):
        lib = _lib.load()
        m, n = x1.shape[0], x2.shape[0]
        k = torch.empty((m, n), dtype=F64, device=x1.device)
        with torch.cuda.device(x1.device):
            rc = lib.mgb_euclidean_gram_fwd(_lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0), x1.shape[1],
                                            _lib.ptr(tval), _lib.ptr(tcoef), tval.shape[0], _lib.ptr(k), max(n, 1),
                                            _lib.stream_ptr(x1.device))
        _lib.check(rc, "mgb_euclidean_gram_fwd")
        ctx.save_for_backward(x1, x2, tval)
        return k

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            raise NotImplementedError("input gradients of the Euclidean integrated Matern kernel are not implemented yet")
        if not ctx.needs_input_grad[3]:
            return None, None, None, None
        lib = _lib.load()
        x1, x2, tval = ctx.saved_tensors
        g = g.contiguous()
        out = torch.zeros_like(tval)
        with torch.cuda.device(x1.device):
            rc = lib.mgb_euclidean_gram_bwd_coef(_lib.ptr(x1), x1.shape[0], x1.stride(0), _lib.ptr(x2), x2.shape[0],
                                                 x2.stride(0), x1.shape[1], _lib.ptr(tval), tval.shape[0], _lib.ptr(g),
                                                 g.stride(0), _lib.ptr(out), _lib.stream_ptr(x1.device))
        _lib.check(rc, "mgb_euclidean_gram_bwd_coef")
        return None, None, None, out


def euclidean_gram(x1, x2, tval, tcoef):
    _lib.require_cuda_f64(x1, x2, tval, tcoef)
    return EuclideanGram.apply(_c2d(x1), _c2d(x2), tval.contiguous(), tcoef.contiguous())


SPD2_LATTICE_MIN_ENTRIES = 4096   # below this the per-CTA table set-up of the lattice kernel does not pay


class Spd2Gram(torch.autograd.Function):
    @staticmethod
    def forward(ctx, use_chol, x1, x2, tcoef, rscale, bscale, bval, bw, lattice=None):
        lib = _lib.load()
        m, n = x1.shape[0], x2.shape[0]
        k = torch.empty((m, n), dtype=F64, device=x1.device)
        with torch.cuda.device(x1.device):
            if lattice is not None and m * n >= SPD2_LATTICE_MIN_ENTRIES:
                sb, sa, g_min, log_tau = lattice
                rc = lib.mgb_spd2_gram_fwd_lattice(use_chol, _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                                   _lib.ptr(tcoef), _lib.ptr(rscale), _lib.ptr(bscale), tcoef.shape[0],
                                                   _lib.ptr(bval), _lib.ptr(bw), bval.shape[0], int(sb), int(sa),
                                                   float(g_min), float(log_tau), _lib.ptr(k), max(n, 1),
                                                   _lib.stream_ptr(x1.device))
                _lib.check(rc, "mgb_spd2_gram_fwd_lattice")
            else:
                rc = lib.mgb_spd2_gram_fwd(use_chol, _lib.ptr(x1), m, x1.stride(0), _lib.ptr(x2), n, x2.stride(0),
                                           _lib.ptr(tcoef), _lib.ptr(rscale), _lib.ptr(bscale), tcoef.shape[0],
                                           _lib.ptr(bval), _lib.ptr(bw), bval.shape[0], _lib.ptr(k), max(n, 1),
                                           _lib.stream_ptr(x1.device))
                _lib.check(rc, "mgb_spd2_gram_fwd")
        ctx.use_chol = use_chol
        ctx.save_for_backward(x1, x2, tcoef, rscale, bscale, bval, bw)
        return k

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise NotImplementedError("input gradients of the SPD quadrature kernels are not implemented yet")
        need_c, need_s = ctx.needs_input_grad[3], (ctx.needs_input_grad[4] or ctx.needs_input_grad[5])
        if not (need_c or need_s):
            return (None,) * 9
        lib = _lib.load()
        x1, x2, tcoef, rscale, bscale, bval, bw = ctx.saved_tensors
        g = g.contiguous()
        gc = gr = gb = None
        with torch.cuda.device(x1.device):
            if need_c:
                gc = torch.zeros_like(rscale)
                rc = lib.mgb_spd2_gram_bwd_coef(ctx.use_chol, _lib.ptr(x1), x1.shape[0], x1.stride(0), _lib.ptr(x2),
                                                x2.shape[0], x2.stride(0), _lib.ptr(rscale), _lib.ptr(bscale),
                                                rscale.shape[0], _lib.ptr(bval), _lib.ptr(bw), bval.shape[0], _lib.ptr(g),
                                                g.stride(0), _lib.ptr(gc), _lib.stream_ptr(x1.device))
                _lib.check(rc, "mgb_spd2_gram_bwd_coef")
            if need_s:
                gr, gb = torch.zeros_like(rscale), torch.zeros_like(bscale)
                rc = lib.mgb_spd2_gram_bwd_scales(ctx.use_chol, _lib.ptr(x1), x1.shape[0], x1.stride(0), _lib.ptr(x2),
                                                  x2.shape[0], x2.stride(0), _lib.ptr(tcoef), _lib.ptr(rscale),
                                                  _lib.ptr(bscale), rscale.shape[0], _lib.ptr(bval), _lib.ptr(bw),
                                                  bval.shape[0], _lib.ptr(g), g.stride(0), _lib.ptr(gr), _lib.ptr(gb),
                                                  _lib.stream_ptr(x1.device))
                _lib.check(rc, "mgb_spd2_gram_bwd_scales")
        return None, None, None, gc, gr, gb, None, None, None


def spd2_gram(use_chol, x1, x2, tcoef, rscale, bscale, bval, bw, lattice=None):
    _lib.require_cuda_f64(x1, x2, tcoef, rscale, bscale, bval, bw)
    return Spd2Gram.apply(int(use_chol), _c2d(x1), _c2d(x2), tcoef.contiguous(), rscale.contiguous(),
                          bscale.contiguous(), bval.contiguous(), bw.contiguous(), lattice)


# --------------------------------------------------------------------------------------------------
# differentiable-input path of the quadrature kernels: cheap pair geometry in torch (any derivative order),
# quadrature profile and its derivatives in the pair scalar on the device (csrc/quad_profile.cu)
# --------------------------------------------------------------------------------------------------
KIND_H1, KIND_H2, KIND_H3, KIND_EUCLID = 0, 1, 2, 3


def inputs_need_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


def _radial_profile_raw(kind, order, z, tval, tcoef, s0, ds, ps):
    lib = _lib.load()
    z = z.contiguous()
    out = torch.empty_like(z) if order != 3 else torch.empty((3,) + tuple(z.shape), dtype=z.dtype, device=z.device)
    with torch.cuda.device(z.device):
        rc = lib.mgb_radial_profile(kind, order, _lib.ptr(z), z.numel(), _lib.ptr(tval), _lib.ptr(tcoef), tval.shape[0],
                                    s0, ds, ps, _lib.ptr(out), _lib.stream_ptr(z.device))
    _lib.check(rc, "mgb_radial_profile")
    return out


class RadialProfile(torch.autograd.Function):
    """P_order(z) = sum_a tcoef[a] d^order/dz^order heat(z, tval[a]) elementwise over ``z``.

    ``backward`` is g * P_{order+1}(z), itself an instance of this Function, so that first- and second-order
    derivatives with respect to the inputs (gradient and Hessian-vector product of the acquisition function)
    come from the same device kernel.  Hyper-parameter gradients: tcoef at every order, tval at order 0."""

    @staticmethod
    def forward(ctx, kind, order, z, tval, tcoef, s0, ds, ps):
        ctx.args = (kind, order, s0, ds, ps)
        ctx.save_for_backward(z, tval, tcoef)
        return _radial_profile_raw(kind, order, z, tval, tcoef, s0, ds, ps)

    @staticmethod
    def backward(ctx, g):
        kind, order, s0, ds, ps = ctx.args
        z, tval, tcoef = ctx.saved_tensors
        gz = gt = gc = None
        if ctx.needs_input_grad[2]:
            if order >= 2:
                raise NotImplementedError("third-order input derivatives of the quadrature kernels are not provided")
            gz = g * RadialProfile.apply(kind, order + 1, z, tval, tcoef, s0, ds, ps)
        if ctx.needs_input_grad[4]:
            lib = _lib.load()
            gd = g.detach().contiguous()
            gc = torch.zeros_like(tcoef)
            with torch.cuda.device(z.device):
                rc = lib.mgb_radial_profile_bwd_coef(kind, order, _lib.ptr(z), z.numel(), _lib.ptr(tval), tval.shape[0],
                                                     s0, ds, ps, _lib.ptr(gd), _lib.ptr(gc), _lib.stream_ptr(z.device))
            _lib.check(rc, "mgb_radial_profile_bwd_coef")
        if ctx.needs_input_grad[3]:
            if order != 0 or kind == KIND_EUCLID:
                raise NotImplementedError("mixed input / heat-time derivatives of the quadrature kernels are not provided")
            nodes = _heat_nodes_raw(kind + 1, 1, z.detach().reshape(-1), tval, s0, ds, ps)     # d heat / dt, (E, Pt)
            gt = (g.detach().reshape(-1, 1) * nodes).sum(0) * tcoef.detach()
        return None, None, gz, gt, gc, None, None, None


def radial_profile(kind, z, tval, tcoef, s0=0.0, ds=0.0, ps=0):
    _lib.require_cuda_f64(z, tval, tcoef)
    return RadialProfile.apply(int(kind), 0, z.contiguous(), tval.contiguous(), tcoef.contiguous(), float(s0), float(ds), int(ps))


def lorentz_distance(x1, x2):
    """Geodesic distance matrix on H^n (Lorentz model) in plain torch ops, operation order of
    Riemannian_utils/hyperbolic_utils_torch.py:29-44,59-60; differentiable to any order.  Used only on the
    differentiable-input path (the Gram kernels recompute it in registers)."""
    diff = x1.unsqueeze(-2) - x2.unsqueeze(-3)
    mink = -diff[..., 0] * diff[..., 0] + torch.sum(diff[..., 1:] * diff[..., 1:], dim=-1)
    sq = torch.maximum(torch.zeros_like(mink), mink)
    return 2 * torch.arcsinh(.5 * torch.sqrt(sq + 1e-8))


def hyperbolic_gram_autograd(dim, x1, x2, tval, tcoef, s0, ds, ps):
    return radial_profile(dim - 1, lorentz_distance(x1, x2), tval, tcoef, s0, ds, ps)


def euclidean_gram_autograd(x1, x2, tval, tcoef):
    diff = x1.unsqueeze(-2) - x2.unsqueeze(-3)
    return radial_profile(KIND_EUCLID, (diff * diff).sum(-1), tval, tcoef)


def _spd2_profile_raw(which, delta, r2, tcoef, rscale, bscale, bval, bw):
    lib = _lib.load()
    out = torch.empty_like(delta) if which != 3 else torch.empty((3,) + tuple(delta.shape), dtype=delta.dtype, device=delta.device)
    with torch.cuda.device(delta.device):
        rc = lib.mgb_spd2_profile(which, _lib.ptr(delta), _lib.ptr(r2), delta.numel(), _lib.ptr(tcoef), _lib.ptr(rscale),
                                  _lib.ptr(bscale), tcoef.shape[0], _lib.ptr(bval), _lib.ptr(bw), bval.shape[0],
                                  _lib.ptr(out), _lib.stream_ptr(delta.device))
    _lib.check(rc, "mgb_spd2_profile")
    return out


class Spd2Profile(torch.autograd.Function):
    """K(Delta, r2) = sum_a tcoef[a] h_a(Delta, r2) elementwise; first-order derivatives in (Delta, r2) and tcoef
    (the SPD trust region of the reference uses a finite-difference Hessian: bo_spd.py:383)."""

    @staticmethod
    def forward(ctx, delta, r2, tcoef, rscale, bscale, bval, bw):
        ctx.save_for_backward(delta, r2, tcoef, rscale, bscale, bval, bw)
        return _spd2_profile_raw(0, delta, r2, tcoef, rscale, bscale, bval, bw)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        delta, r2, tcoef, rscale, bscale, bval, bw = ctx.saved_tensors
        if ctx.needs_input_grad[3] or ctx.needs_input_grad[4]:
            raise NotImplementedError("simultaneous input and heat-lengthscale gradients of the SPD(2) kernel are not "
                                      "provided (detach one of them)")
        gd = gr = gc = None
        if ctx.needs_input_grad[0]:
            gd = g * _spd2_profile_raw(1, delta, r2, tcoef, rscale, bscale, bval, bw)
        if ctx.needs_input_grad[1]:
            gr = g * _spd2_profile_raw(2, delta, r2, tcoef, rscale, bscale, bval, bw)
        if ctx.needs_input_grad[2]:
            lib = _lib.load()
            gg = g.contiguous()
            gc = torch.zeros_like(tcoef)
            with torch.cuda.device(delta.device):
                rc = lib.mgb_spd2_profile_bwd_coef(0, _lib.ptr(delta), _lib.ptr(r2), delta.numel(), _lib.ptr(rscale),
                                                   _lib.ptr(bscale), rscale.shape[0], _lib.ptr(bval), _lib.ptr(bw),
                                                   bval.shape[0], _lib.ptr(gg), _lib.ptr(gc), _lib.stream_ptr(delta.device))
            _lib.check(rc, "mgb_spd2_profile_bwd_coef")
        return gd, gr, gc, None, None, None, None


def _mandel2(v, chol):
    """Mandel 3-vector -> entries (a, b, c, d) of the 2 x 2 matrix (spd_utils_torch.py:336-371) or of its lower
    Cholesky factor (kernels_spd.py:90-91)."""
    s11, s22, s12 = v[..., 0], v[..., 1], v[..., 2] / 2 ** 0.5
    if not chol:
        return s11, s12, s12, s22
    a = torch.sqrt(s11)
    c = s12 / a
    return a, torch.zeros_like(a), c, torch.sqrt(s22 - c * c)


def spd2_log_singular_values(x1, x2, chol):
    """H1 >= H2 = log singular values of A_i B_j^-1 (closed form for 2 x 2), differentiable torch ops."""
    a, b, c, d = (t.unsqueeze(-1) for t in _mandel2(x1, chol))
    e, f, g, h = (t.unsqueeze(-2) for t in _mandel2(x2, chol))
    det = e * h - f * g
    i11, i12, i21, i22 = h / det, -f / det, -g / det, e / det
    p11, p12 = a * i11 + b * i21, a * i12 + b * i22
    p21, p22 = c * i11 + d * i21, c * i12 + d * i22
    q1 = torch.sqrt((p11 + p22) ** 2 + (p12 - p21) ** 2)
    q2sq = (p11 - p22) ** 2 + (p12 + p21) ** 2
    q2 = torch.sqrt(torch.where(q2sq > 0, q2sq, torch.ones_like(q2sq))) * (q2sq > 0)   # zero (sub)gradient at q2 = 0
    s1 = 0.5 * (q1 + q2)
    s2 = torch.abs(p11 * p22 - p12 * p21) / s1
    return torch.log(s1), torch.log(s2)


def spd2_gram_autograd(use_chol, x1, x2, tcoef, rscale, bscale, bval, bw):
    _lib.require_cuda_f64(x1, x2, tcoef, rscale, bscale, bval, bw)
    h1, h2 = spd2_log_singular_values(x1, x2, bool(use_chol))
    return Spd2Profile.apply((h1 - h2).contiguous(), (h1 * h1 + h2 * h2).contiguous(), tcoef.contiguous(),
                             rscale.contiguous(), bscale.contiguous(), bval.contiguous(), bw.contiguous())


def spd_decompose(x: torch.Tensor, dim: int, target: int) -> torch.Tensor:
    """Mandel vectors (..., d(d+1)/2) of SPD(dim) matrices -> (..., dim + dim^2) [eigenvalues | vec(Q)] (target 0) or
    (..., dim + 2 | 4) [eigenvalues | sphere point] (target 1) on the device (csrc/spd_eigh.cu; conventions in mgb.h)."""
    _lib.require_cuda_f64(x)
    if inputs_need_grad(x):
        raise NotImplementedError("spd_input=True: input gradients through the eigen-decomposition are not provided")
    width = dim * (dim + 1) // 2
    if x.shape[-1] != width:
        raise RuntimeError("SPD(%d) points must be Mandel vectors with %d coordinates" % (dim, width))
    flat = x.detach().reshape(-1, width).contiguous()
    wout = dim + (dim * dim if target == 0 else (2 if dim == 2 else 4))
    out = torch.empty((flat.shape[0], wout), dtype=F64, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        rc = lib.mgb_spd_decompose(dim, target, _lib.ptr(flat), flat.shape[0], width, _lib.ptr(out), wout,
                                   _lib.stream_ptr(x.device))
    _lib.check(rc, "mgb_spd_decompose")
    return out.reshape(*x.shape[:-1], wout)


# --------------------------------------------------------------------------------------------------
# pair-batch plumbing shared by the Kernel.forward overrides
# --------------------------------------------------------------------------------------------------
def _is_broadcast_of_2d(t: torch.Tensor) -> bool:
    return t.dim() > 2 and all(s == 0 or sz == 1 for s, sz in zip(t.stride()[:-2], t.shape[:-2]))


def _series_diag_raw(spec, x1, x2, coef):
    lib = _lib.load()
    rows = x1.shape[0]
    out = torch.empty(rows, dtype=F64, device=x1.device)
    d = spec.desc(coef.shape[-1])
    with torch.cuda.device(x1.device):
        rc = lib.mgb_series_gram_diag(C.byref(d), _lib.ptr(x1), x1.stride(0), _lib.ptr(x2), x2.stride(0), rows,
                                      _lib.ptr(coef), _lib.ptr(out), _lib.stream_ptr(x1.device))
    _lib.check(rc, "mgb_series_gram_diag")
    return out


def _series_batched_raw(spec, f1, f2, coef):
    """(B, m, D) x (B, n, D) -> (B, m, n) from one launch (gridDim.z = B)."""
    lib = _lib.load()
    bsz, m, n = f1.shape[0], f1.shape[1], f2.shape[1]
    k = torch.empty((bsz, m, n), dtype=F64, device=f1.device)
    if bsz == 0 or m == 0 or n == 0:
        return k
    d = spec.desc(coef.shape[-1])
    with torch.cuda.device(f1.device):
        rc = lib.mgb_series_gram_fwd_batched(C.byref(d), _lib.ptr(f1), m, f1.stride(1), f1.stride(0), _lib.ptr(f2), n,
                                             f2.stride(1), f2.stride(0), _lib.ptr(coef), _lib.ptr(k), n, m * n, bsz,
                                             _lib.stream_ptr(f1.device))
    _lib.check(rc, "mgb_series_gram_fwd_batched")
    return k


def pairwise(fn2d, x1: torch.Tensor, x2: torch.Tensor, diag: bool = False, series=None, rowwise=None) -> torch.Tensor:
    """Apply a 2-D Gram function over gpytorch-style batch dimensions.

    * no batch dims: one launch;
    * x2 a broadcast of one N x D block (botorch's posterior layout b x q x D against the expanded training
      inputs) - by strides, or materialised with identical members: one launch on (b*q) x D against N x D;
    * otherwise one batched launch (``series=(spec, coef)``, no autograd needed) or one launch per batch element.
    ``diag=True`` returns the per-row kernel k(x1_i, x2_i) (shape ... x N): one thread per row for the series kernels,
    ``rowwise`` (the kernel's profile path on (rows, 1, D) batches) for the quadrature kernels."""
    plain_series = series is not None and not inputs_need_grad(x1, x2, series[1])
    if diag:
        if x1.shape != x2.shape:
            raise RuntimeError("diag=True requires x1 and x2 of the same shape")
        flat1 = x1.reshape(-1, x1.shape[-1]).contiguous()
        flat2 = x2.reshape(-1, x2.shape[-1]).contiguous()
        rows = flat1.shape[0]
        if plain_series:
            _lib.require_cuda_f64(flat1, flat2)
            return _series_diag_raw(series[0], flat1, flat2, series[1].contiguous()).reshape(x1.shape[:-1])
        if rowwise is not None and rows > 0:
            return rowwise(flat1.unsqueeze(-2), flat2.unsqueeze(-2)).reshape(x1.shape[:-1])
        blk = 32                       # block-diagonal sweep (differentiable series path): blk x blk entries per blk values
        pieces = [torch.diagonal(fn2d(flat1[s:s + blk], flat2[s:s + blk])) for s in range(0, rows, blk)]
        out = torch.cat(pieces) if pieces else torch.empty(0, dtype=F64, device=x1.device)
        return out.reshape(x1.shape[:-1])
    if x1.dim() == 2 and x2.dim() == 2:
        return fn2d(x1, x2)
    if x1.dim() < x2.dim():
        x1 = x1.expand(*x2.shape[:-2], *x1.shape[-2:])
    elif x2.dim() < x1.dim():
        x2 = x2.expand(*x1.shape[:-2], *x2.shape[-2:])
    batch = torch.broadcast_shapes(x1.shape[:-2], x2.shape[:-2])
    x1 = x1.expand(*batch, *x1.shape[-2:])
    x2 = x2.expand(*batch, *x2.shape[-2:])
    m, n = x1.shape[-2], x2.shape[-2]
    first = (0,) * len(batch)

    def _same_members(t):
        """A materialised expansion (gpytorch may hand train_inputs.expand(...).contiguous()): identical batch members."""
        if t.numel() == 0 or t.requires_grad or (t.is_cuda and torch.cuda.is_current_stream_capturing()):
            return False                    # the comparison reads a flag back: not during CUDA-graph capture
        return bool((t == t[first]).all())

    if _is_broadcast_of_2d(x2) or (not _is_broadcast_of_2d(x1) and _same_members(x2)):
        k = fn2d(x1.reshape(-1, x1.shape[-1]), x2[first])
        return k.reshape(*batch, m, n)
    if _is_broadcast_of_2d(x1):
        k = fn2d(x1[first], x2.reshape(-1, x2.shape[-1]))          # m x (B*n)
        return k.reshape(m, *batch, n).movedim(0, -2)
    f1 = x1.reshape(-1, m, x1.shape[-1])
    f2 = x2.reshape(-1, n, x2.shape[-1])
    if plain_series:
        _lib.require_cuda_f64(f1, f2)
        f1 = f1 if f1.stride(-1) == 1 else f1.contiguous()
        f2 = f2 if f2.stride(-1) == 1 else f2.contiguous()
        return _series_batched_raw(series[0], f1, f2, series[1].contiguous()).reshape(*batch, m, n)
    out = torch.stack([fn2d(f1[b], f2[b]) for b in range(f1.shape[0])]) if f1.shape[0] else \
        torch.empty((0, m, n), dtype=F64, device=x1.device)
    return out.reshape(*batch, m, n)
