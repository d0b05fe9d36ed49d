"""Host-side (torch, float64, autograd) computation of the tiny hyper-parameter-dependent vectors the CUDA
kernels consume: series coefficients and quadrature node weights.

Every reference kernel is ``K_ij = sum_n w_n(kappa, nu) * phi_n(pair scalar)``; all hyper-parameter
dependence lives in ``w`` (a few dozen numbers), so lengthscale / smoothness gradients come back from the
device as ``dL/dw`` and flow through the formulas below by ordinary torch autograd.  The formulas mirror
the reference line by line (cited per function, paths relative to /root/reference/BoManifolds),
including its ``.item()`` detachments of the quadrature grids and the dtype of its constants.
"""
from __future__ import annotations

import math
import os
from fractions import Fraction
from functools import lru_cache

import numpy as np
import torch

from . import _lib

F64 = torch.float64
MONOMIAL_MAX_TERMS = 12  # Horner in the monomial basis up to here (error <= ~1e-12 of max|K|), Chebyshev beyond


# --------------------------------------------------------------------------------------------------
# basis-change tables (pure numbers, cached)
# --------------------------------------------------------------------------------------------------
def gegenbauer_monomial_row(n: int, alpha: float):
    """Coefficients of u^k in C_n^alpha(u), the same float64 numbers the reference's power sum uses
    (math_utils/gegenbauer_polynomials.py:31-33).  math.gamma(0) raises ValueError for alpha = 0."""
    row = [0.0] * (n + 1)
    ga = math.gamma(alpha)
    for i in range(n // 2 + 1):
        coef = math.gamma(n - i + alpha) / (ga * math.factorial(i) * math.factorial(n - 2 * i))
        row[n - 2 * i] = math.pow(-1, i) * coef * math.pow(2.0, n - 2 * i)
    return row


@lru_cache(maxsize=None)
def gegenbauer_to_monomial(terms: int, alpha: float) -> np.ndarray:
    m = np.zeros((terms, terms))
    for n in range(terms):
        m[n, :n + 1] = gegenbauer_monomial_row(n, alpha)
    return m


@lru_cache(maxsize=None)
def gegenbauer_to_chebyshev(terms: int, two_alpha: int) -> np.ndarray:
    """C_n^alpha(cos t) = sum_j (alpha)_j (alpha)_{n-j} / (j! (n-j)!) cos((n-2j) t), exact rationals."""
    alpha = Fraction(two_alpha, 2)
    math.gamma(float(alpha))  # same ValueError as the reference for S^1

    def poch(a, k):
        r = Fraction(1)
        for i in range(k):
            r *= a + i
        return r

    m = np.zeros((terms, terms))
    for n in range(terms):
        acc = [Fraction(0)] * (n + 1)
        for j in range(n + 1):
            acc[abs(n - 2 * j)] += poch(alpha, j) * poch(alpha, n - j) / (math.factorial(j) * math.factorial(n - j))
        m[n, :n + 1] = [float(v) for v in acc]
    return m


@lru_cache(maxsize=None)
def chebyshev_to_monomial(terms: int) -> np.ndarray:
    """Integer coefficients of T_m(u)."""
    rows = [[1], [0, 1]]
    for m in range(2, terms):
        prev, prev2 = rows[m - 1], rows[m - 2]
        cur = [0] * (m + 1)
        for k, v in enumerate(prev):
            cur[k + 1] += 2 * v
        for k, v in enumerate(prev2):
            cur[k] -= v
        rows.append(cur)
    out = np.zeros((terms, terms))
    for m in range(terms):
        out[m, :m + 1] = rows[m]
    return out


_CONST_CACHE = {}


def _t(x, like):
    """numpy table -> float64 tensor on ``like``'s device; the lru-cached basis-change tables are uploaded once per
    device (a GP fit calls this hundreds of times with the same table)."""
    key = (id(x), str(like.device))
    hit = _CONST_CACHE.get(key)
    if hit is None or hit[0] is not x:
        hit = (x, torch.as_tensor(x, dtype=F64, device=like.device))
        _CONST_CACHE[key] = hit
    return hit[1]


def _const_vec(lst, device):
    """Python list of 1-element tensors (the reference's cst_nd / zero_gpolynomials) -> one float64 vector on
    ``device``, cached against the list object (values keep the dtype rounding they were created with)."""
    key = (id(lst), str(device))
    hit = _CONST_CACHE.get(key)
    if hit is None or hit[0] is not lst:
        hit = (lst, torch.cat([v.reshape(1) for v in lst]).to(device=device, dtype=F64))
        _CONST_CACHE[key] = hit
    return hit[1]


def choose_basis(terms: int) -> int:
    return _lib.BASIS_MONOMIAL if terms <= MONOMIAL_MAX_TERMS else _lib.BASIS_CHEBYSHEV


def _logspace(lo, hi, steps, device):
    """The reference builds its grids with torch.logspace on the CPU in the DEFAULT dtype and then moves
    them (e.g. kernels_hyperbolic.py:368-369): same here, so the node values agree bit for bit."""
    return torch.logspace(lo, hi, steps).to(device=device, dtype=F64)


def trapz_weights(nodes: torch.Tensor) -> torch.Tensor:
    """w with trapz(f, nodes) == sum(w * f)."""
    d = nodes[1:] - nodes[:-1]
    w = torch.zeros_like(nodes)
    w[:-1] += 0.5 * d
    w[1:] += 0.5 * d
    return w


# --------------------------------------------------------------------------------------------------
# sphere S^d  (kernels_sphere.py)
# --------------------------------------------------------------------------------------------------
def sphere_constants(dim: int, terms: int):
    """cst_nd and zero_gpolynomials as Python lists of 1-element tensors IN THE DEFAULT DTYPE, exactly as
    kernels_sphere.py:84,87 / :387-406 build them (float32-rounded unless the caller set float64)."""
    from .math_utils import gegenbauer_polynomial

    alpha = (dim - 1.0) / 2.0
    zero_g = [gegenbauer_polynomial(n, alpha, torch.ones(1)) for n in range(terms)]
    cst = []
    for n in range(terms):
        dn = (2 * n + dim - 1) * math.gamma(n + dim - 1) / math.gamma(dim) / math.gamma(n + 1)
        cst.append(dn * math.gamma(alpha) / (2 * math.pow(math.pi, alpha) * zero_g[n]))
    return cst, zero_g


def _sphere_finish(a: torch.Tensor, zero_g: torch.Tensor, dim: int):
    """a_n (Gegenbauer weights) -> normalised device coefficients + basis."""
    terms = a.shape[-1]
    w = a / (a * zero_g).sum()
    basis = choose_basis(terms)
    if basis == _lib.BASIS_MONOMIAL:
        mat = gegenbauer_to_monomial(terms, (dim - 1.0) / 2.0)
    else:
        mat = gegenbauer_to_chebyshev(terms, dim - 1)
    return (w.reshape(1, terms) @ _t(mat, w)).contiguous(), basis


class _SpectralCoef(torch.autograd.Function):
    """coef(kappa, nu) of the sphere / SO(3) series and its Jacobian from ONE device launch (mgb_spectral_coef);
    the torch formulas below stay the definition (and the CPU path the host tests exercise)."""

    @staticmethod
    def forward(ctx, mode, lam, c, g, mat, half_dim, kappa, nu):
        lib = _lib.load()
        t, tout = mat.shape
        coef = torch.empty((1, tout), dtype=F64, device=kappa.device)
        jac = torch.empty((2, tout), dtype=F64, device=kappa.device)
        kappa = kappa.detach().contiguous()
        nu_c = nu.detach().contiguous() if nu is not None else None
        with torch.cuda.device(kappa.device):
            rc = lib.mgb_spectral_coef(int(mode), int(t), int(tout), _lib.ptr(lam), _lib.ptr(c), _lib.ptr(g), _lib.ptr(mat),
                                       float(half_dim), _lib.ptr(kappa), _lib.ptr(nu_c), _lib.ptr(coef), _lib.ptr(jac),
                                       _lib.stream_ptr(kappa.device))
        _lib.check(rc, "mgb_spectral_coef")
        ctx.save_for_backward(jac)
        ctx.has_nu = nu is not None
        return coef

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcoef):
        (jac,) = ctx.saved_tensors
        gv = jac @ gcoef.reshape(-1)                       # (2,)
        gk = gv[0].reshape(()) if ctx.needs_input_grad[6] else None
        gn = gv[1].reshape(()) if (ctx.has_nu and ctx.needs_input_grad[7]) else None
        return None, None, None, None, None, None, gk, gn


def _so3_matrix(terms: int, basis: int) -> np.ndarray:
    """lam -> series coefficients as one matrix: v_m = (2 - delta_m0) sum_{l >= m} lam_l, then Chebyshev -> monomial."""
    key = ("so3mat", terms, basis)
    hit = _CONST_CACHE.get(key)
    if hit is None:
        tail = np.zeros((terms, terms))
        for l in range(terms):
            for m in range(l + 1):
                tail[l, m] = 1.0 if m == 0 else 2.0
        mat = tail @ chebyshev_to_monomial(terms) if basis == _lib.BASIS_MONOMIAL else tail
        hit = (None, np.ascontiguousarray(mat))
        _CONST_CACHE[key] = hit
    return hit[1]


def _device_vec(key, device, make):
    k = (key, str(device))
    hit = _CONST_CACHE.get(k)
    if hit is None:
        hit = (None, make().to(device=device, dtype=F64).contiguous())
        _CONST_CACHE[k] = hit
    return hit[1]


def _eigenvalues(terms: int, dim: int, device):
    """n (n + dim - 1), n < terms: Laplace-Beltrami eigenvalues of S^dim, cached per device."""
    key = ("lb", terms, dim, str(device))
    hit = _CONST_CACHE.get(key)
    if hit is None:
        n = torch.arange(terms, dtype=F64, device=device)
        hit = (None, n * (n + dim - 1))
        _CONST_CACHE[key] = hit
    return hit[1]


def _sphere_device(mode, ls, nu, dim, cst, zero_g):
    terms = len(cst)
    basis = choose_basis(terms)
    mat = gegenbauer_to_monomial(terms, (dim - 1.0) / 2.0) if basis == _lib.BASIS_MONOMIAL else \
        gegenbauer_to_chebyshev(terms, dim - 1)
    coef = _SpectralCoef.apply(mode, _eigenvalues(terms, dim, ls.device), _const_vec(cst, ls.device),
                               _const_vec(zero_g, ls.device), _t(mat, ls), dim / 2.0, ls, nu)
    return coef, basis


def _so3_device(mode, ls, nu, summands):
    basis = choose_basis(summands)
    dev = ls.device
    lam = _device_vec(("so3lam", summands), dev, lambda: torch.arange(summands, dtype=F64) * (torch.arange(summands, dtype=F64) + 1))
    dims = _device_vec(("so3dim", summands), dev, lambda: 2 * torch.arange(summands, dtype=F64) + 1)
    coef = _SpectralCoef.apply(mode, lam, dims, dims, _t(_so3_matrix(summands, basis), ls), 1.5, ls, nu)
    return coef, basis


def sphere_matern_coef(lengthscale, nu, dim, cst, zero_g):
    """kernels_sphere.py:126-139."""
    ls, nu = lengthscale.reshape(()).to(F64), nu.reshape(()).to(F64)
    terms = len(cst)
    if ls.is_cuda:
        return _sphere_device(0, ls, nu, dim, cst, zero_g)
    base = 2 * nu / ls ** 2
    power = -(nu + dim / 2)
    e = torch.pow(base + _eigenvalues(terms, dim, ls.device), power) / torch.pow(base, power)
    c, g = _const_vec(cst, ls.device), _const_vec(zero_g, ls.device)
    return _sphere_finish(e * c, g, dim)


def sphere_gaussian_coef(lengthscale, dim, cst, zero_g):
    """kernels_sphere.py:212-224."""
    ls = lengthscale.reshape(()).to(F64)
    terms = len(cst)
    if ls.is_cuda:
        return _sphere_device(1, ls, None, dim, cst, zero_g)
    e = torch.exp(-ls ** 2 / 2.0 * _eigenvalues(terms, dim, ls.device))
    c, g = _const_vec(cst, ls.device), _const_vec(zero_g, ls.device)
    return _sphere_finish(e * c, g, dim)


def sphere_integrated_matern_coef(lengthscale, nu, dim, cst, zero_g, nb_points):
    """kernels_sphere.py:318-384: the t-trapezoid of the heat series collapses into the series weights."""
    ls, nu = lengthscale.reshape(()).to(F64), nu.reshape(()).to(F64)
    terms = len(cst)
    shift = torch.log10(ls).item()
    t = _logspace(-4 + shift, 3 + shift, nb_points, ls.device)
    wt = trapz_weights(t)
    n = torch.arange(terms, dtype=F64, device=ls.device)
    front = wt * torch.pow(t, nu + dim / 2 - 1.0) * torch.exp(-2.0 * nu / ls ** 2 * t)      # (P,)
    heat = torch.exp(-t.unsqueeze(-1) * (n * (n + dim - 1)).unsqueeze(0))                    # (P, T)
    c, g = _const_vec(cst, ls.device), _const_vec(zero_g, ls.device)
    return _sphere_finish((front.unsqueeze(-1) * heat).sum(0) * c, g, dim)


# --------------------------------------------------------------------------------------------------
# circle S^1 / torus  (kernels_sphere.py:443-472, 551-614; kernels_torus.py)
# --------------------------------------------------------------------------------------------------
THETA3_TERMS = 200  # hard-wired in the reference (math_utils/jacobi_theta_functions.py:16)
_TRUNC = 1e-16      # dropped tail: sum |c_n| below this fraction of sum |c| (|T_n| <= 1, so the kernel value moves by
                    # less than half an ulp of its own scale - far inside the 1e-10 parity bar and the rounding of the
                    # reference's own 201-term float64 sum)


def _truncate(c: torch.Tensor) -> int:
    """Terms to keep: the tail sum below _TRUNC of the total is dropped.  Device coefficients (mgb_circle_coef) arrive
    with that rule already applied - their negligible tail is written as exact zeros by the kernel - so only the last
    non-zero is looked up."""
    if c.is_cuda:
        nz = torch.nonzero(c.detach() != 0)
        return max(int(nz.max().item()) + 1 if nz.numel() else 0, 2)
    mag = c.detach().abs().cpu()
    tail = torch.flip(torch.cumsum(torch.flip(mag, [0]), 0), [0])     # tail[k] = sum_{n >= k} |c_n|
    keep = int(torch.nonzero(tail > _TRUNC * tail[0]).max().item()) + 1
    return max(keep, 2)


class _CircleCoef(torch.autograd.Function):
    """coef(kappa, nu) of a circle kernel (201 Chebyshev coefficients, normalised at phi = 0) and its Jacobian from ONE
    device launch (mgb_circle_coef); the torch formulas below stay the definition and the CPU path."""

    @staticmethod
    def forward(ctx, mode, grid, kappa, nu):
        lib = _lib.load()
        terms = THETA3_TERMS + 1
        coef = torch.empty(terms, dtype=F64, device=kappa.device)
        jac = torch.empty((2, terms), dtype=F64, device=kappa.device)
        kappa = kappa.detach().contiguous()
        nu_c = nu.detach().contiguous() if nu is not None else None
        with torch.cuda.device(kappa.device):
            rc = lib.mgb_circle_coef(int(mode), terms, _lib.ptr(grid[0]) if grid is not None else None,
                                     _lib.ptr(grid[1]) if grid is not None else None,
                                     int(grid.shape[1]) if grid is not None else 0, _lib.ptr(kappa), _lib.ptr(nu_c),
                                     _lib.ptr(coef), _lib.ptr(jac), _lib.stream_ptr(kappa.device))
        _lib.check(rc, "mgb_circle_coef")
        ctx.save_for_backward(jac)
        ctx.has_nu = nu is not None
        return coef

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcoef):
        (jac,) = ctx.saved_tensors
        gv = jac @ gcoef.reshape(-1)
        gk = gv[0].reshape(()) if ctx.needs_input_grad[2] else None
        gn = gv[1].reshape(()) if (ctx.has_nu and ctx.needs_input_grad[3]) else None
        return None, None, gk, gn


def _circle_grid(shift: float, nb_points: int, device):
    """(2, P) device tensor [t; trapezoid weights] of the reference's logspace grid for one lengthscale, built on the
    host exactly as the torch path does and cached (the acquisition phase evaluates the kernel hundreds of times with
    the hyper-parameters fixed)."""
    key = ("circle_grid", float(shift), int(nb_points), str(device), str(torch.get_default_dtype()))
    hit = _CONST_CACHE.get(key)
    if hit is None:
        t = torch.logspace(-3 + shift, 1 + shift, nb_points).to(F64)
        if len(_CONST_CACHE) > 4096:
            _CONST_CACHE.clear()
        hit = (None, torch.stack([t, trapz_weights(t)]).to(device).contiguous())
        _CONST_CACHE[key] = hit
    return hit[1]


def circle_coef_sync_free(mode, lengthscale, nu, nb_points):
    """All 201 Chebyshev coefficients of a circle kernel (mode 0: integrated Matern, 1: heat) WITHOUT any host read:
    the t grid comes from mgb_matern_t_grid (device lengthscale), nothing is truncated and no basis decision is taken
    (both need the values on the host).  For launch-bound problems (GP fit on a torus: CUDA-graph capturable); large
    Gram matrices use the truncated / monomial form of ``circle_*_coef``."""
    ls = lengthscale.reshape(()).to(F64)
    grid = None
    if mode == 0:
        lib = _lib.load()
        grid = torch.empty((2, int(nb_points)), dtype=F64, device=ls.device)
        kd = ls.detach().contiguous()
        with torch.cuda.device(ls.device):
            rc = lib.mgb_matern_t_grid(_lib.ptr(kd), -3.0, 1.0, int(nb_points), _lib.ptr(grid[0]), _lib.ptr(grid[1]),
                                       _lib.stream_ptr(ls.device))
        _lib.check(rc, "mgb_matern_t_grid")
    return _CircleCoef.apply(int(mode), grid, ls, nu.reshape(()).to(F64) if nu is not None else None)


def circle_gaussian_coef(lengthscale):
    """theta_3(phi/2, q) / theta_3(0, q), q = exp(-2 pi^2 kappa^2)  ->  Chebyshev series in cos(phi):
    c_0 = 1, c_n = 2 q^(n^2)."""
    ls = lengthscale.reshape(()).to(F64)
    if ls.is_cuda:
        c = _CircleCoef.apply(1, None, ls, None)
        return c[:_truncate(c)]
    q = torch.exp(-2 * math.pi ** 2 * ls ** 2)
    n = torch.arange(0, THETA3_TERMS + 1, dtype=F64, device=ls.device)
    c = 2.0 * torch.pow(q, n ** 2)
    c = torch.cat([torch.ones(1, dtype=F64, device=ls.device), c[1:]])
    c = c / c.sum()
    return c[:_truncate(c)]


def circle_integrated_matern_coef(lengthscale, nu, nb_points):
    """trapz_t t^(nu-1/2) exp(-2 nu t/kappa^2) theta_3(phi/2, exp(-4 pi^2 t)), normalised at phi = 0."""
    ls, nu = lengthscale.reshape(()).to(F64), nu.reshape(()).to(F64)
    shift = torch.log10(ls).item()
    if ls.is_cuda:
        c = _CircleCoef.apply(0, _circle_grid(shift, nb_points, ls.device), ls, nu)
        return c[:_truncate(c)]
    t = _logspace(-3 + shift, 1 + shift, nb_points, ls.device)
    wt = trapz_weights(t)
    front = wt * torch.pow(t, nu + 0.5 - 1.0) * torch.exp(-2.0 * nu / ls ** 2 * t)          # (P,)
    n = torch.arange(0, THETA3_TERMS + 1, dtype=F64, device=ls.device)
    qn = torch.exp(-4 * math.pi ** 2 * t.unsqueeze(-1) * (n ** 2).unsqueeze(0))              # (P, 201)
    c = (front.unsqueeze(-1) * qn).sum(0)
    c = torch.cat([c[:1], 2.0 * c[1:]])
    c = c / c.sum()
    return c[:_truncate(c)]


MONOMIAL_CONDITION_MAX = 100.0   # sum |a_k| / sum |c_n| above which the monomial form would amplify rounding


def chebyshev_or_monomial(c: torch.Tensor):
    """(F, T) Chebyshev coefficients -> (coefficients, basis) for the device.

    Horner in the monomial basis costs one FMA per term, the Chebyshev recurrence two.  The monomial re-expansion
    a = c @ [T_n]_k is safe exactly when it does not blow up: the rounding error of Horner on [-1, 1] is bounded by
    2^-53 * sum |a_k|, and for the circle / torus series the coefficients decay so fast (c_n ~ exp(-const n^2)) that
    sum |a_k| stays O(1) for lengthscales >= ~0.3 (33 terms at kappa = 0.5: sum |a| = 1.0001; 55 terms at kappa = 0.2:
    1.8e4 - kept in the Chebyshev basis).  Reads the coefficients back once (they live where the hyper-parameters
    live)."""
    terms = c.shape[-1]
    if terms > _lib.MAX_TERMS or terms > 64:
        return c, _lib.BASIS_CHEBYSHEV
    a = c @ _t(chebyshev_to_monomial(terms), c)
    ratio = a.detach().abs().sum(-1) / c.detach().abs().sum(-1).clamp_min(1e-300)
    if float(ratio.max()) <= MONOMIAL_CONDITION_MAX:
        return a.contiguous(), _lib.BASIS_MONOMIAL
    return c, _lib.BASIS_CHEBYSHEV


def stack_factor_coefs(coefs):
    """Per-factor Chebyshev coefficient vectors -> zero-padded (F, T) matrix."""
    t = max(int(c.shape[0]) for c in coefs)
    rows = [torch.cat([c, torch.zeros(t - c.shape[0], dtype=F64, device=c.device)]) for c in coefs]
    return torch.stack(rows).contiguous()


# --------------------------------------------------------------------------------------------------
# SO(3)  (kernels_so.py with dim = 3)
# --------------------------------------------------------------------------------------------------
def _so3_finish(lam: torch.Tensor):
    """lam_l already holds the representation dimension (2l+1).  Character chi_l = 1 + 2 sum_{m<=l} cos(m theta):
    Chebyshev weights v_0 = sum_l lam_l, v_m = 2 sum_{l>=m} lam_l; normaliser = series at theta = 0 =
    sum_l lam_l (2l+1) (the 1/sum lam^2 prefactor of kernels_so.py:283 cancels against :360-362)."""
    terms = lam.shape[0]
    l = torch.arange(terms, dtype=F64, device=lam.device)
    tail = torch.flip(torch.cumsum(torch.flip(lam, [0]), 0), [0])
    v = torch.cat([tail[:1], 2.0 * tail[1:]]) / (lam * (2 * l + 1)).sum()
    basis = choose_basis(terms)
    if basis == _lib.BASIS_MONOMIAL:
        v = v.reshape(1, terms) @ _t(chebyshev_to_monomial(terms), v)
    return v.reshape(1, terms).contiguous(), basis


def so3_matern_coef(lengthscale, nu, summands):
    """(2 nu/kappa^2 + l(l+1))^(-nu - 3/2) (2l+1); kernels_so.py:264-273 (so_dim = 3)."""
    ls, nu = lengthscale.reshape(()).to(F64), nu.reshape(()).to(F64)
    if ls.is_cuda:
        return _so3_device(0, ls, nu, summands)
    l = torch.arange(summands, dtype=F64, device=ls.device)
    return _so3_finish(torch.pow(2 * nu / ls ** 2 + l * (l + 1), -nu - 1.5) * (2 * l + 1))


def so3_gaussian_coef(lengthscale, summands):
    """exp(-l(l+1) kappa^2 / 2) (2l+1); kernels_so.py:101."""
    ls = lengthscale.reshape(()).to(F64)
    if ls.is_cuda:
        return _so3_device(1, ls, None, summands)
    l = torch.arange(summands, dtype=F64, device=ls.device)
    return _so3_finish(torch.exp(-l * (l + 1) * ls ** 2 / 2) * (2 * l + 1))


# --------------------------------------------------------------------------------------------------
# quadrature kernels: node grids (detached, as the reference's .item()) and differentiable node weights
# --------------------------------------------------------------------------------------------------
class _MaternTWeights(torch.autograd.Function):
    """Normalised t-node weights tcoef(kappa, nu) on a fixed grid and their Jacobian from ONE device launch
    (mgb_matern_t_weights); the torch formulas of ``matern_t_nodes`` + the kernels' ``nodes`` stay the definition."""

    @staticmethod
    def forward(ctx, grid, has_h0, power_offset, kappa, nu):
        lib = _lib.load()
        p = grid.shape[1]
        tcoef = torch.empty(p, dtype=F64, device=kappa.device)
        jac = torch.empty((2, p), dtype=F64, device=kappa.device)
        kappa, nu = kappa.detach().contiguous(), nu.detach().contiguous()
        with torch.cuda.device(kappa.device):
            rc = lib.mgb_matern_t_weights(_lib.ptr(grid[0]), _lib.ptr(grid[1]), _lib.ptr(grid[2]) if has_h0 else None, int(p),
                                          float(power_offset), _lib.ptr(kappa), _lib.ptr(nu), _lib.ptr(tcoef), _lib.ptr(jac),
                                          _lib.stream_ptr(kappa.device))
        _lib.check(rc, "mgb_matern_t_weights")
        ctx.save_for_backward(jac)
        return tcoef

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (jac,) = ctx.saved_tensors
        gv = jac @ g.reshape(-1)
        return None, None, None, (gv[0].reshape(()) if ctx.needs_input_grad[3] else None), \
            (gv[1].reshape(()) if ctx.needs_input_grad[4] else None)


def device_grids_enabled() -> bool:
    """Sync-free (CUDA-graph capturable) node grids: only in the float64 regime - in the reference's float32-default
    regime torch.logspace rounds the grid to float32 on the host, which a device kernel cannot reproduce bit for bit."""
    return torch.get_default_dtype() == F64 and os.environ.get("MGB_DEVICE_GRIDS", "1") != "0"


def matern_t_coef_device(lengthscale, nu, lo_exp, hi_exp, nb_points, power_offset, h0_fn):
    """``matern_t_coef`` without the host read of the lengthscale: grid and trapezoid weights from ``mgb_matern_t_grid``
    (device kappa), zero-distance inner integral recomputed on the device, weights + Jacobian from
    ``mgb_matern_t_weights``.  No cache, no synchronisation.  Returns (t, tcoef)."""
    ls, nu = lengthscale.reshape(()).to(F64), nu.reshape(()).to(F64)
    lib = _lib.load()
    p = int(nb_points)
    grid = torch.empty((3, p), dtype=F64, device=ls.device)
    kd = ls.detach().contiguous()
    with torch.cuda.device(ls.device):
        rc = lib.mgb_matern_t_grid(_lib.ptr(kd), float(lo_exp), float(hi_exp), p, _lib.ptr(grid[0]), _lib.ptr(grid[1]),
                                   _lib.stream_ptr(ls.device))
    _lib.check(rc, "mgb_matern_t_grid")
    if h0_fn is not None:
        with torch.no_grad():
            grid[2].copy_(h0_fn(grid[0]).reshape(-1))
    tcoef = _MaternTWeights.apply(grid, h0_fn is not None, power_offset, ls, nu)
    return grid[0], tcoef


def matern_t_coef(lengthscale, nu, lo_exp, hi_exp, nb_points, power_offset, h0_key, h0_fn):
    """Device path of ``matern_t_nodes`` + normalisation: (t, tcoef, shift) with tcoef = w / sum(w h0).
    ``h0_fn(t)`` evaluates the inner integral at zero distance on the grid (None: 1); the grid, its trapezoid weights
    and h0 depend on the lengthscale only through the detached shift and are cached per (shift, ...)."""
    ls, nu = lengthscale.reshape(()).to(F64), nu.reshape(()).to(F64)
    shift = torch.log10(ls).item()
    key = ("t_grid", float(shift), float(lo_exp), float(hi_exp), int(nb_points), h0_key, str(ls.device),
           str(torch.get_default_dtype()))
    hit = _CONST_CACHE.get(key)
    if hit is None:
        if len(_CONST_CACHE) > 4096:
            _CONST_CACHE.clear()
        t = torch.logspace(lo_exp + shift, hi_exp + shift, nb_points).to(F64)
        tw = trapz_weights(t)
        t_dev = t.to(ls.device)
        with torch.no_grad():
            h0 = h0_fn(t_dev).reshape(-1).to(F64) if h0_fn is not None else torch.ones_like(t_dev)
        grid = torch.stack([t_dev, tw.to(ls.device), h0]).contiguous()
        hit = (None, grid)
        _CONST_CACHE[key] = hit
    grid = hit[1]
    tcoef = _MaternTWeights.apply(grid, h0_fn is not None, power_offset, ls, nu)
    return grid[0], tcoef, shift


def matern_t_nodes(lengthscale, nu, lo_exp, hi_exp, nb_points, power_offset=1.0, return_shift=False):
    """t grid logspace(lo+log10 kappa, hi+log10 kappa, P) and the weights
    trapz_w * t^(nu - power_offset) * exp(-2 nu t / kappa^2)
    (kernels_hyperbolic.py:367-375, kernels_spd.py:262-264, kernels_euclidean.py:186-190)."""
    ls, nu = lengthscale.reshape(()).to(F64), nu.reshape(()).to(F64)
    shift = torch.log10(ls).item()
    t = _logspace(lo_exp + shift, hi_exp + shift, nb_points, ls.device)
    w = trapz_weights(t) * torch.pow(t, nu - power_offset) * torch.exp(-2.0 * nu / ls ** 2 * t)
    if return_shift:
        return t, w, shift
    return t, w
