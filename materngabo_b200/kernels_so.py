"""SO(n) kernels: drop-in for BoManifolds/kernel_utils/kernels_so.py (SORiemannianGaussianKernel :28-177,
SORiemannianMaternKernel :180-362).

SO(3) runs on the sm_100a series kernel: the reference's sympy-built character sum
sum_l c_l chi_l(e^{i theta}) with chi_l = sum_{m=-l}^{l} x^m is the Chebyshev series
sum_m v_m T_m(cos theta), cos theta = clip((tr(R1 R2^T) - 1)/2, +-(1 - 1e-8)) (kernels_so.py:323-327), and
tr(R1 R2^T) is the inner product of the two flattened 9-vectors.  No sympy is needed at construction.
SO(n), n in {2, 4, 5, 6, 7} (kernels_so.py:329-346: eigenvalues of R1 R2^T + conjugate pairing + character
polynomial) runs on csrc/son_gram.cu: the reference's polynomial (rebuilt exactly by _so_characters.py, no sympy)
evaluated at the eigen-angles of R1 R2^T in (0, pi) in descending order.  The reference substitutes the angles in the
order LAPACK lists the conjugate pairs, which makes its value order dependent; the two agree where that order is
descending.
"""
from __future__ import annotations

import torch

import ctypes as C

from . import _lib, _ops, _so_characters, _weights
from .gpytorch_compat import Kernel
from .kernels_sphere import _NuMixin, _check_batch, _series_on

_SO3_CLIP = 1.0 - 1e-8  # kernels_so.py:325
QUATERNION_MIN_ENTRIES = 1 << 22  # Gram blocks at least this large go through the quaternion form (one stream sync)


class _SOSeriesKernel(Kernel):
    has_lengthscale = True

    def _coef(self):
        raise NotImplementedError

    def series(self):
        coef, basis = self._coef()
        return _ops.SeriesSpec(1, 9, basis, 0.5, -0.5, -_SO3_CLIP, _SO3_CLIP), coef

    def forward(self, x1, x2, diag=False, **params):
        # the reference ignores ``diag`` (kernels_so.py:305); the gpytorch contract is honoured here
        _check_batch(self)
        if self.dim != 3:
            return self._forward_general(x1, x2, diag)
        if x1.shape[-1] != 9 or x2.shape[-1] != 9:
            raise RuntimeError("SO(3) points must be flattened 3x3 rotation matrices (9 coordinates)")
        spec, coef = _series_on(self, x1.device)
        big = (not diag and x1.dim() == 2 and x2.dim() == 2
               and x1.shape[0] * x2.shape[0] >= QUATERNION_MIN_ENTRIES
               and not _ops.inputs_need_grad(x1, x2, coef))
        if big:
            # (tr(R1 R2^T) - 1) / 2 = 2 <q1, q2>^2 - 1: rows of 4 instead of 9 for the inner products (the FP64 pipe,
            # not HBM, bounds the 9-wide form).  Only for inputs that are rotations to 1e-13; otherwise the trace form.
            quats = _ops.so3_quaternions(x1, x2)
            if quats is not None:
                spec4 = _ops.SeriesSpec(1, 4, spec.basis, 2.0, -1.0, spec.lo, spec.hi, 1)
                with torch.no_grad():
                    return _ops._series_fwd_raw(spec4, quats[0], quats[1], coef.detach().contiguous())
        return _ops.pairwise(lambda a, b: _ops.series_gram(spec, a, b, coef), x1, x2, diag=diag, series=(spec, coef))


    # ---- n != 3: character sum at the canonical eigen-angles (csrc/son_gram.cu) ---------------------------------
    def _lambda(self, lb_ev, rep_dim, so_dim):
        """lambda_pi(kappa, nu) of the summands (kernels_so.py:101 / :273), differentiable torch."""
        raise NotImplementedError

    def _general_tables(self, device):
        if self.dim not in (2, 4, 5, 6, 7):
            raise NotImplementedError("SO(%d): the CUDA path covers n = 2 ... 7 (rank <= 3)" % self.dim)
        key = (self.dim, self.summands_number_bound, str(device))
        cached = self.__dict__.get("_son_cache")
        if cached is None or cached[0] != key:
            tb = _so_characters.table(self.dim, self.summands_number_bound)
            j, maps = tb.grid_maps()
            order = sorted(maps)                                    # lexicographic sign patterns = device order
            mcat = torch.cat([torch.from_numpy(maps[s]) for s in order], dim=1).to(device=device, dtype=torch.float64)
            ev = torch.tensor(tb.lb_ev, dtype=torch.float64, device=device)
            dims = torch.tensor(tb.rep_dim, dtype=torch.float64, device=device)
            cached = (key, (tb, j, mcat, ev, dims))
            self.__dict__["_son_cache"] = cached
        return cached[1]

    def son_grids(self, device):
        """Normalised coefficient grids (flat, device) for the current hyper-parameters, differentiable, and J."""
        tb, j, mcat, ev, dims = self._general_tables(device)
        lam = self._lambda(ev, dims, float(tb.so_dim))
        return (lam @ mcat) / (lam * dims).sum(), j

    def _forward_general(self, x1, x2, diag):
        n2 = self.dim * self.dim
        if x1.shape[-1] != n2 or x2.shape[-1] != n2:
            raise RuntimeError("SO(%d) points must be flattened rotation matrices (%d coordinates)" % (self.dim, n2))
        if _ops.inputs_need_grad(x1, x2):
            raise NotImplementedError("SO(n), n != 3: input gradients are not provided (the reference's SO trust region "
                                      "uses approx_hessian=True, bo_so.py:266, and only SO(3) is exercised)")
        grids, j = self.son_grids(x1.device)
        lib = _lib.load()
        dim = self.dim

        def fn2d(a, b):
            a, b = _ops._c2d(a), _ops._c2d(b)
            _lib.require_cuda_f64(a, b)
            m, n = a.shape[0], b.shape[0]
            want_grad = torch.is_grad_enabled() and grids.requires_grad
            with torch.cuda.device(a.device):
                if not want_grad:
                    g = grids.detach().contiguous()
                    k = torch.empty((m, n), dtype=torch.float64, device=a.device)
                    rc = lib.mgb_son_gram(dim, 0, _lib.ptr(a), m, a.stride(0), _lib.ptr(b), n, b.stride(0), _lib.ptr(g), j,
                                          _lib.ptr(k), max(n, 1), _lib.stream_ptr(a.device))
                    _lib.check(rc, "mgb_son_gram")
                    return k
                # hyper-parameter gradients (GP fit): per-pair basis from the device, contraction with the grids in torch
                nb = int(lib.mgb_son_basis_size(dim, j))
                basis = torch.empty((m, n, nb), dtype=torch.float64, device=a.device)
                rc = lib.mgb_son_gram(dim, 1, _lib.ptr(a), m, a.stride(0), _lib.ptr(b), n, b.stride(0), None, j,
                                      _lib.ptr(basis), nb, _lib.stream_ptr(a.device))
                _lib.check(rc, "mgb_son_gram")
                return basis @ grids

        if diag:
            if x1.shape != x2.shape:
                raise RuntimeError("diag=True requires x1 and x2 of the same shape")
            f1, f2 = x1.reshape(-1, n2), x2.reshape(-1, n2)
            pieces = [torch.diagonal(fn2d(f1[s:s + 64], f2[s:s + 64])) for s in range(0, f1.shape[0], 64)]
            out = torch.cat(pieces) if pieces else torch.empty(0, dtype=torch.float64, device=x1.device)
            return out.reshape(x1.shape[:-1])
        return _ops.pairwise(fn2d, x1, x2)


class SORiemannianGaussianKernel(_SOSeriesKernel):
    def __init__(self, dim, sigma_squared=1, summands_number_bound=10, lambda_coeff_bound=10 ** -10, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self.dim = dim
        self.summands_number_bound = summands_number_bound

    def _coef(self):
        return _weights.so3_gaussian_coef(self.lengthscale, self.summands_number_bound)

    def _lambda(self, lb_ev, rep_dim, so_dim):
        ls = self.lengthscale.reshape(()).to(lb_ev)
        return torch.exp(-lb_ev * ls ** 2 / 2) * rep_dim                      # kernels_so.py:101


class SORiemannianMaternKernel(_NuMixin, _SOSeriesKernel):
    def __init__(self, dim, nu=None, nu_prior=None, sigma_squared=1, summands_number_bound=10,
                 lambda_coeff_bound=10 ** -10, **kwargs):
        self.has_lengthscale = True
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._register_nu(nu, nu_prior)
        self.dim = dim
        self.summands_number_bound = summands_number_bound

    def _coef(self):
        return _weights.so3_matern_coef(self.lengthscale, self.nu, self.summands_number_bound)

    def _lambda(self, lb_ev, rep_dim, so_dim):
        ls, nu = self.lengthscale.reshape(()).to(lb_ev), self.nu.reshape(()).to(lb_ev)
        return torch.pow(2 * nu / ls ** 2 + lb_ev, -nu - so_dim / 2) * rep_dim    # kernels_so.py:273
