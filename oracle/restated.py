"""TEST INFRASTRUCTURE ONLY - CPU restatement (plain torch, float64) of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker or the timed CPU baseline.  The product
(``materngabo_b200``) never imports it.

The reference (NoemieJaquier/MaternGaBO) is pure Python/torch, so the restatement uses the same
torch operator sequence as the reference (pair expansion + batched product, Python loops over series
terms and quadrature nodes), which also makes it the honest CPU baseline.  Every function cites the
reference lines it follows (paths relative to /root/reference/BoManifolds).

Pinning of the posterior / EI arithmetic (gpytorch / botorch: unvendored, not installable here): PARITY UNPINNED
against gpytorch itself; pinned against scikit-learn's GaussianProcessRegressor and scipy's normal distribution
(tests/test_oracle_posterior_sklearn.py) and against an extended-precision restatement (oracle/truth_ld.c).

Pinning: the reference ships no tests or golden vectors.  This file is pinned
against the UNMODIFIED reference imported in the dev container (``oracle/reference_loader.py``) by
``tests/test_oracle_vs_reference.py`` (runs only where /root/reference exists) and by the committed
fixtures ``tests/golden/*.npz`` that ``oracle/make_golden.py`` generated from the reference itself.

Regime: float64 everywhere, i.e. the reference constructed after
``torch.set_default_dtype(torch.float64)`` (otherwise its constants are float32-rounded, SURVEY.md
section 0).  ``const_dtype=torch.float32`` reproduces that other regime for the sphere constants.
"""
from __future__ import annotations

import math

import torch

F64 = torch.float64


# ----------------------------------------------------------------------------------------------
# pair expansion used by every reference distance helper
# ----------------------------------------------------------------------------------------------
def _pairs(x1: torch.Tensor, x2: torch.Tensor):
    """All (i, j) row pairs materialised, as the reference does with torch.cat
    (Riemannian_utils/sphere_utils_torch.py:35-40, hyperbolic_utils_torch.py:27-32)."""
    m, n = x1.shape[-2], x2.shape[-2]
    a = x1.unsqueeze(-2).expand(*x1.shape[:-1], n, x1.shape[-1]).contiguous()
    b = x2.unsqueeze(-3).expand(*x2.shape[:-2], m, n, x2.shape[-1]).contiguous()
    return a, b


# ----------------------------------------------------------------------------------------------
# sphere
# ----------------------------------------------------------------------------------------------
def sphere_distance(x1, x2, diag=False):
    """acos(clamp(<x1_i, x2_j>, -1+1e-15, 1-1e-15)); sphere_utils_torch.py:16-59."""
    if not diag:
        a, b = _pairs(x1, x2)
        d = a.shape[-1]
        ip = torch.bmm(a.reshape(-1, 1, d), b.reshape(-1, d, 1)).reshape(a.shape[:-1])
    else:
        ip = torch.bmm(x1.unsqueeze(-2), x2.unsqueeze(-1)).squeeze(-1)  # (..., N, 1) as the reference
    return torch.acos(ip.clamp(-1.0 + 1e-15, 1.0 - 1e-15))


def gegenbauer(n: int, alpha: float, z: torch.Tensor):
    """C_n^alpha(z) by the explicit power sum; math_utils/gegenbauer_polynomials.py:12-43.
    math.gamma(0) raises ValueError for alpha = 0 (S^1), as in the reference."""
    ga = math.gamma(alpha)
    out = 0.0
    for i in range(n // 2 + 1):
        c = math.gamma(n - i + alpha) / (ga * math.factorial(i) * math.factorial(n - 2 * i))
        out = out + math.pow(-1, i) * torch.pow(2 * z, n - 2 * i) * c
    return out


def sphere_constant(n: int, d: int, dtype=F64):
    """c_{n,d}; kernel_utils/kernels_sphere.py:387-406 (built in the *default dtype* there)."""
    dn = (2 * n + d - 1) * math.gamma(n + d - 1) / math.gamma(d) / math.gamma(n + 1)
    g1 = gegenbauer(n, (d - 1.0) / 2.0, torch.ones(1, dtype=dtype))
    return dn * math.gamma((d - 1.0) / 2.0) / (2 * math.pow(math.pi, (d - 1.0) / 2.0) * g1)


def _sphere_tables(dim, terms, const_dtype):
    cst = [sphere_constant(n, dim, const_dtype) for n in range(terms)]
    g1 = [gegenbauer(n, (dim - 1.0) / 2.0, torch.ones(1, dtype=const_dtype)) for n in range(terms)]
    return cst, g1


def _as_param(v):
    return v if torch.is_tensor(v) else torch.tensor([[float(v)]], dtype=F64)


def sphere_matern(x1, x2, dim, nu, lengthscale, terms=10, diag=False, const_dtype=F64):
    """kernels_sphere.py:102-139."""
    nu, ls = _as_param(nu), _as_param(lengthscale)
    cst, g1 = _sphere_tables(dim, terms, const_dtype)
    cosd = torch.cos(sphere_distance(x1, x2, diag=diag))
    k = torch.zeros_like(cosd)
    norm = torch.zeros((1, 1), dtype=F64)
    e0 = torch.pow(2 * nu / ls ** 2, -(nu + dim / 2))
    for n in range(terms):
        e = torch.pow(2 * nu / ls ** 2 + n * (n + dim - 1), -(nu + dim / 2)) / e0
        k = k + e * cst[n] * gegenbauer(n, (dim - 1.0) / 2.0, cosd)
        norm = norm + e * cst[n] * g1[n]
    return k / norm


def sphere_gaussian(x1, x2, dim, lengthscale, terms=10, diag=False, const_dtype=F64):
    """kernels_sphere.py:188-224."""
    ls = _as_param(lengthscale)
    cst, g1 = _sphere_tables(dim, terms, const_dtype)
    cosd = torch.cos(sphere_distance(x1, x2, diag=diag))
    k = torch.zeros_like(cosd)
    norm = torch.zeros((1, 1), dtype=F64)
    for n in range(terms):
        e = torch.exp(-ls ** 2 / 2.0 * n * (n + dim - 1))
        k = k + e * cst[n] * gegenbauer(n, (dim - 1.0) / 2.0, cosd)
        norm = norm + e * cst[n] * g1[n]
    return k / norm


def sphere_integrated_matern(x1, x2, dim, nu, lengthscale, terms=10, nb_points=100, diag=False,
                             const_dtype=F64):
    """kernels_sphere.py:318-384 (t-trapezoid of the heat series)."""
    nu, ls = _as_param(nu), _as_param(lengthscale)
    cst, g1 = _sphere_tables(dim, terms, const_dtype)
    cosd = torch.cos(sphere_distance(x1, x2, diag=diag))
    gp = [gegenbauer(n, (dim - 1.0) / 2.0, cosd) for n in range(terms)]

    def link(g, t, like):
        heat = torch.zeros_like(like)
        for n in range(terms):
            heat = heat + torch.exp(-t * n * (n + dim - 1)) * cst[n] * g[n]
        return torch.pow(t, nu + dim / 2 - 1.0) * torch.exp(-2.0 * nu / ls ** 2 * t) * heat

    shift = torch.log10(ls).item()
    t_vals = torch.logspace(-4 + shift, 3 + shift, nb_points, dtype=F64)
    vals = torch.stack([link(gp, t_vals[i], cosd) for i in range(nb_points)])
    k = torch.trapz(vals, t_vals, dim=0)
    nvals = torch.stack([link(g1, t_vals[i], torch.zeros(1, 1, dtype=F64)).reshape(()) for i in range(nb_points)])
    return k / torch.trapz(nvals, t_vals, dim=0)


# ----------------------------------------------------------------------------------------------
# circle / torus
# ----------------------------------------------------------------------------------------------
def theta3(z, q, terms=200):
    """1 + 2 sum_{n=1}^{200} q^{n^2} cos(2 n z); math_utils/jacobi_theta_functions.py:16-20."""
    acc = torch.zeros_like(z)
    for n in range(1, terms + 1):
        acc = acc + torch.pow(q, n ** 2) * torch.cos(2 * n * z)
    return 2 * acc + 1.0


def circle_gaussian(x1, x2, lengthscale, diag=False):
    """kernels_sphere.py:443-472."""
    ls = _as_param(lengthscale)
    sd = sphere_distance(x1, x2, diag=diag) / (2 * math.pi)
    q = torch.exp(-2 * math.pi ** 2 * ls ** 2)
    return theta3(math.pi * sd, q) / theta3(torch.zeros(1, 1, dtype=F64), q)


def circle_integrated_matern(x1, x2, nu, lengthscale, nb_points=50, diag=False):
    """kernels_sphere.py:551-614 (dim = 1)."""
    nu, ls = _as_param(nu), _as_param(lengthscale)
    sd = sphere_distance(x1, x2, diag=diag) / (2 * math.pi)

    def link(s, t):
        heat = theta3(math.pi * s, torch.exp(-4 * math.pi ** 2 * t))
        return torch.pow(t, nu + 0.5 - 1.0) * torch.exp(-2.0 * nu / ls ** 2 * t) * heat

    shift = torch.log10(ls).item()
    t_vals = torch.logspace(-3 + shift, 1 + shift, nb_points, dtype=F64)
    k = torch.trapz(torch.stack([link(sd, t_vals[i]) for i in range(nb_points)]), t_vals, dim=0)
    z = torch.zeros(1, 1, dtype=F64)
    nrm = torch.trapz(torch.stack([link(z, t_vals[i]).reshape(()) for i in range(nb_points)]), t_vals, dim=0)
    return k / nrm


def _to_circles(x, dim):
    """angles -> interleaved (cos, sin); kernels_torus.py:162-173."""
    if x.shape[-1] != dim:
        return x
    out = torch.zeros(list(x.shape[:-1]) + [2 * dim], dtype=x.dtype)
    out[..., ::2] = torch.cos(x)
    out[..., 1::2] = torch.sin(x)
    return out


def torus_matern(x1, x2, dim, nu, lengthscales, nb_points=50):
    """kernels_torus.py:142-177: product over circles c of the S^1 integrated Matern on columns (2c, 2c+1)."""
    x1, x2 = _to_circles(x1, dim), _to_circles(x2, dim)
    nus = nu if isinstance(nu, (list, tuple)) else [nu] * dim
    lss = lengthscales if isinstance(lengthscales, (list, tuple)) else [lengthscales] * dim
    k = None
    for c in range(dim):
        f = circle_integrated_matern(x1[..., 2 * c:2 * c + 2], x2[..., 2 * c:2 * c + 2], nus[c], lss[c], nb_points)
        k = f if k is None else k * f
    return k


def torus_gaussian(x1, x2, dim, lengthscales):
    """kernels_torus.py:61-90."""
    x1, x2 = _to_circles(x1, dim), _to_circles(x2, dim)
    lss = lengthscales if isinstance(lengthscales, (list, tuple)) else [lengthscales] * dim
    k = None
    for c in range(dim):
        f = circle_gaussian(x1[..., 2 * c:2 * c + 2], x2[..., 2 * c:2 * c + 2], lss[c])
        k = f if k is None else k * f
    return k


# ----------------------------------------------------------------------------------------------
# SO(3)
# ----------------------------------------------------------------------------------------------
def _so3_series(x, coeff):
    """sum_l coeff_l * chi_l(x) / sum_l coeff_l^2 with chi_l(x) = sum_{m=-l}^{l} x^m, which is what the
    sympy construction of kernels_so.py:204-290 produces for dim = 3 (rank 1, highest weights 0, 2, 4, ...;
    Laplace-Beltrami eigenvalue l(l+1); coeff_l already contains the representation dimension 2l+1)."""
    num = 0
    den = 0
    for l, c in enumerate(coeff):
        chi = 0
        for m in range(-l, l + 1):
            chi = chi + x ** m
        num = num + c * chi
        den = den + c ** 2
    return num / den


def _so3_angle(x1, x2):
    """theta = arccos(clip((tr(R1 R2^T) - 1)/2, -1+1e-8, 1-1e-8)); kernels_so.py:305-326."""
    r1 = x1.reshape(*x1.shape[:-1], 3, 3)
    r2 = x2.reshape(*x2.shape[:-1], 3, 3)
    m, n = r1.shape[-3], r2.shape[-3]
    a = r1.unsqueeze(-3).expand(*r1.shape[:-3], m, n, 3, 3).contiguous()
    b = r2.unsqueeze(-4).expand(*r2.shape[:-3], m, n, 3, 3).contiguous()
    ratio = torch.bmm(a.reshape(-1, 3, 3), b.reshape(-1, 3, 3).transpose(-1, -2))
    tr = ratio.diagonal(dim1=-1, dim2=-2).sum(-1)
    return torch.arccos(torch.clip((tr - 1.0) / 2.0, -1.0 + 1e-8, 1.0 - 1e-8)), a.shape[:-2]


def so3_matern(x1, x2, nu, lengthscale, summands=10):
    """kernels_so.py:305-362 with dim = 3; coefficient (2 nu/k^2 + l(l+1))^(-nu - 3/2) (2l+1), :273."""
    nu, ls = _as_param(nu), _as_param(lengthscale)
    coeff = [(2 * nu / ls ** 2 + l * (l + 1)) ** (-nu - 1.5) * (2 * l + 1) for l in range(summands)]
    theta, shape = _so3_angle(x1, x2)
    x = torch.exp(torch.complex(torch.zeros_like(theta), theta))
    k = _so3_series(x, coeff).real
    nrm = _so3_series(torch.ones(1, dtype=F64), coeff)
    return k.reshape(shape) / nrm


def so3_gaussian(x1, x2, lengthscale, summands=10):
    """kernels_so.py:149-177 with dim = 3; coefficient exp(-l(l+1) k^2 / 2) (2l+1), :101."""
    ls = _as_param(lengthscale)
    coeff = [torch.exp(-l * (l + 1) * ls ** 2 / 2) * (2 * l + 1) for l in range(summands)]
    theta, shape = _so3_angle(x1, x2)
    x = torch.exp(torch.complex(torch.zeros_like(theta), theta))
    k = _so3_series(x, coeff).real
    nrm = _so3_series(torch.ones(1, dtype=F64), coeff)
    return k.reshape(shape) / nrm


# ----------------------------------------------------------------------------------------------
# SO(n), n != 3: eigenvalues of R1 R2^T, conjugate pairing in LAPACK order, Weyl character sum
# ----------------------------------------------------------------------------------------------
def _son_root_data(n):
    """Root-system data the reference builds with sympy (kernels_so.py:204-227, math_utils/so_root_systems.py), here in
    exact rational arithmetic: Gram matrix B of the simple roots (integer scaled), fundamental weights in the root
    basis, delta, the Weyl group as (matrix, determinant) pairs and the positive roots."""
    from fractions import Fraction as Fr
    from math import lcm

    rk = n // 2
    if n in (2, 3):
        cm = [[Fr(2)]]
    else:
        cm = [[Fr(2) if i == j else (Fr(-1) if abs(i - j) == 1 else Fr(0)) for j in range(rk)] for i in range(rk)]
        if n % 2:
            cm[rk - 2][rk - 1] = Fr(-2)                                  # B_k (so_root_systems.py:14-17)
        elif rk > 2:
            for (i, j), v in {(0, 1): -1, (0, 2): -1, (1, 0): -1, (1, 2): 0, (2, 0): -1, (2, 1): 0}.items():
                cm[rk - 3 + i][rk - 3 + j] = Fr(v)                       # D_k tail (:20-24)
    sq = [None] * rk                                                     # squared root lengths (symmetrize, :39-60)
    sq[0] = Fr(1)
    changed = True
    while changed:
        changed = False
        for x in range(rk):
            for y in range(rk):
                if sq[x] is not None and sq[y] is None and cm[x][y] != 0:
                    sq[y] = sq[x] * cm[y][x] / cm[x][y]
                    changed = True
    bm = [[cm[x][y] * sq[y] / 2 for y in range(rk)] for x in range(rk)]
    scale = 1
    for row in bm:
        for v in row:
            scale = lcm(scale, v.denominator)
    bm = [[v * scale for v in row] for row in bm]
    # fundamental weights: columns of (cm^-1)^T, by solving cm^T w = e_i with Gaussian elimination
    aug = [[cm[j][i] for j in range(rk)] + [Fr(int(i == c)) for c in range(rk)] for i in range(rk)]   # rows of cm^T | I
    for c in range(rk):
        piv = next(r for r in range(c, rk) if aug[r][c] != 0)
        aug[c], aug[piv] = aug[piv], aug[c]
        aug[c] = [v / aug[c][c] for v in aug[c]]
        for r in range(rk):
            if r != c and aug[r][c] != 0:
                f = aug[r][c]
                aug[r] = [a - f * b for a, b in zip(aug[r], aug[c])]
    inv_t = [row[rk:] for row in aug]                                     # (cm^T)^-1
    fw = inv_t                                                            # fw[i][j]: i-th root coordinate of weight j
    delta = [sum(fw[i][j] for j in range(rk)) for i in range(rk)]

    def bil(u, v):
        return sum(u[i] * bm[i][j] * v[j] for i in range(rk) for j in range(rk))

    unit = [[Fr(int(i == j)) for i in range(rk)] for j in range(rk)]

    def reflect(i, v):
        f = 2 * bil(v, unit[i]) / bil(unit[i], unit[i])
        return tuple(a - f * b for a, b in zip(v, unit[i]))

    # Weyl group through its action on the basis vectors (columns), closure under the simple reflections
    ident = tuple(tuple(u) for u in unit)
    group = {ident: 1}
    frontier = [ident]
    while frontier:
        nxt = []
        for g in frontier:
            for i in range(rk):
                h = tuple(reflect(i, col) for col in g)
                if h not in group:
                    group[h] = -group[g]
                    nxt.append(h)
        frontier = nxt
    roots = set()
    for g in group:
        for col in g:
            if all(c >= 0 for c in col):
                roots.add(col)
    return rk, bm, fw, delta, group, sorted(roots), bil


def _son_summands(n, bound):
    """Highest weights by height until >= bound summands (kernels_so.py:247-282): (shifted weight pi + delta,
    Laplace-Beltrami eigenvalue, dimension by Weyl's dimension formula)."""
    import itertools
    from fractions import Fraction as Fr

    rk, bm, fw, delta, group, roots, bil = _son_root_data(n)
    out = []
    for height in itertools.count(0):
        for pi in itertools.product(range(height + 1), repeat=rk):
            if sum(pi) != height:
                continue
            vec = [sum(fw[i][j] * pi[j] for j in range(rk)) for i in range(rk)]
            if any(sum(bm[i][j] * vec[j] for j in range(rk)).denominator != 1 for i in range(rk)):
                continue
            sh = [a + b for a, b in zip(vec, delta)]
            dim = Fr(1)
            for r in roots:
                dim *= bil(sh, r) / bil(delta, r)
            out.append((sh, float(bil(sh, sh) - bil(delta, delta)), float(dim)))
        if len(out) >= bound:
            break
    return rk, bm, delta, group, out


def son_reference_angles(x1, x2, n):
    """Arguments of the character polynomial as the reference picks them (kernels_so.py:329-346): eigvals of R1 R2^T,
    first member of each complex-conjugate pair in the order LAPACK lists them.  (m, k, n // 2) angles."""
    m, k = x1.shape[0], x2.shape[0]
    a = x1.reshape(m, 1, n, n).expand(m, k, n, n).reshape(-1, n, n)
    b = x2.reshape(1, k, n, n).expand(m, k, n, n).reshape(-1, n, n)
    ev = torch.linalg.eigvals(torch.bmm(a, b.transpose(-1, -2)))
    ang = torch.zeros(m * k, n // 2, dtype=F64)
    for p in range(ev.shape[0]):
        lst = list(ev[p])
        q = 0
        while lst:
            e = lst.pop(0)
            for i, other in enumerate(lst):
                if torch.abs(torch.conj(e) - other) < 1e-10:
                    ang[p, q] = torch.angle(e)
                    q += 1
                    del lst[i]
                    break
    return ang.reshape(m, k, n // 2)


def son_kernel_from_angles(angles, n, lengthscale, nu=None, summands=10):
    """sum_pi lambda_pi Re chi_pi / sum_pi lambda_pi dim_pi with chi_pi by Weyl's character formula evaluated
    NUMERICALLY as a ratio of alternating sums (kernels_so.py:234-239,269 build the same ratio symbolically and cancel
    it); x_i = exp(i angle_i) in the reference's simple-root coordinates.  nu = None: heat kernel coefficients (:101),
    else Matern (:273).  Valid away from the zeros of the Weyl denominator (generic pairs)."""
    rk, bm, delta, group, summ = _son_summands(n, summands)
    so_dim = 2 * rk * rk + rk if n % 2 else 2 * rk * rk - rk
    ls = _as_param(lengthscale).reshape(())
    th = angles.to(torch.complex128)

    def alt(weight):
        tot = 0
        for g, sign in group.items():
            w = [sum(g[j][i] * weight[j] for j in range(rk)) for i in range(rk)]          # g applied to the weight
            e = [float(sum(w[i] * bm[i][j] for i in range(rk))) for j in range(rk)]
            tot = tot + sign * torch.exp(1j * sum(e[j] * th[..., j] for j in range(rk)))
        return tot

    den = alt(delta)
    num_total, norm = 0, 0
    for sh, ev, dim in summ:
        lam = torch.exp(-ev * ls ** 2 / 2) * dim if nu is None else \
            torch.pow(2 * _as_param(nu).reshape(()) / ls ** 2 + ev, -_as_param(nu).reshape(()) - so_dim / 2) * dim
        num_total = num_total + lam * (alt(sh) / den).real
        norm = norm + lam * dim
    return num_total / norm


def son_matern(x1, x2, n, nu, lengthscale, summands=10, angles=None):
    """kernels_so.py:305-362, dim != 3.  ``angles`` overrides the LAPACK-order arguments (golden vectors record the ones
    the reference used; passing angles sorted descending gives the order-independent function the CUDA path evaluates)."""
    ang = son_reference_angles(x1, x2, n) if angles is None else angles
    return son_kernel_from_angles(ang, n, lengthscale, nu, summands)


def son_gaussian(x1, x2, n, lengthscale, summands=10, angles=None):
    """kernels_so.py:121-177, dim != 3."""
    ang = son_reference_angles(x1, x2, n) if angles is None else angles
    return son_kernel_from_angles(ang, n, lengthscale, None, summands)


# ----------------------------------------------------------------------------------------------
# hyperbolic (Lorentz model)
# ----------------------------------------------------------------------------------------------
def lorentz_distance(x1, x2, diag=False):
    """2 asinh(sqrt(max(0, <D, D>_L) + 1e-8) / 2), D = x1_i - x2_j;
    Riemannian_utils/hyperbolic_utils_torch.py:11-60."""
    if not diag:
        a, b = _pairs(x1, x2)
        diff = (a.reshape(-1, a.shape[-1]) - b.reshape(-1, b.shape[-1])).transpose(-1, -2)
        shape = a.shape[:-1]
    else:
        diff = (x1 - x2).transpose(-1, -2)
        shape = None
    mink = -diff[0] * diff[0] + torch.sum(diff[1:] * diff[1:], dim=0)
    nrm = torch.sqrt(torch.maximum(torch.zeros_like(mink), mink) + 1e-8)
    d = 2 * torch.arcsinh(0.5 * nrm)
    return d.reshape(shape) if shape is not None else d


def hyperbolic_heat(x1, x2, dim, lengthscale, nb_points=1000, diag=False):
    """kernels_hyperbolic.py:58-88 (dim 1, 2, 3 branches of forward :111-117)."""
    ls = _as_param(lengthscale)
    d = lorentz_distance(x1, x2, diag=diag)
    if dim == 1:
        return torch.exp(-torch.pow(d, 2) / (2 * ls ** 2))
    if dim == 3:
        k = torch.exp(-torch.pow(d, 2) / (2 * ls ** 2))
        return k * (d + 1e-8) / torch.sinh(d + 1e-8)
    if dim > 3:
        return hyperbolic_heat_high_dim(x1, x2, dim, lengthscale, nb_points, diag=diag)
    base = torch.linspace(1e-2, 100.0, nb_points, dtype=F64)
    s = base + d.unsqueeze(-1)
    vals = s * torch.exp(-s ** 2 / (2 * ls ** 2)) / torch.sqrt(torch.cosh(s) - torch.cosh(d.unsqueeze(-1)))
    k = torch.trapz(vals, s, dim=-1)
    nv = base * torch.exp(-base ** 2 / (2 * ls ** 2)) / torch.sqrt(torch.cosh(base) - 1.0)
    return k / torch.trapz(nv, base, dim=-1)


def hyperbolic_heat_high_dim(x1, x2, dim, lengthscale, nb_points=1000, diag=False, return_parts=False):
    """kernels_hyperbolic.py:118-180: H^n heat kernel for n > 3 as -((1 / sinh d) d/dd)^m applied to the H^2 (even n,
    m = (n - 2) / 2) or H^3 (odd n, m = (n - 3) / 2) kernel by nested torch autograd, distances clamped to >= 1e-6,
    normalised by the same expression at d = 1e-6.  return_parts: (numerator, normaliser) un-divided - the normaliser
    is ill-conditioned in the reference (~1e-4 relative for the H^3 base), the numerator is not."""
    ls = _as_param(lengthscale)
    even = dim % 2 == 0
    m = (dim - 2) // 2 if even else (dim - 3) // 2

    def base(d):
        if even:
            grid = torch.linspace(1e-2, 100.0, nb_points, dtype=F64)
            s = grid + d.unsqueeze(-1)
            vals = s * torch.exp(-s ** 2 / (2 * ls ** 2)) / torch.sqrt(torch.cosh(s) - torch.cosh(d.unsqueeze(-1)))
            k = torch.trapz(vals, s, dim=-1)
            nv = grid * torch.exp(-grid ** 2 / (2 * ls ** 2)) / torch.sqrt(torch.cosh(grid) - 1.0)
            return k / torch.trapz(nv, grid, dim=-1)
        k = torch.exp(-torch.pow(d, 2) / (2 * ls ** 2))
        return k * (d + 1e-8) / torch.sinh(d + 1e-8)

    def operator(d):
        part = base(d)
        for _ in range(m):
            (g,) = torch.autograd.grad(part, d, grad_outputs=torch.ones_like(d), create_graph=True)
            part = g / torch.sinh(d)
        return -part

    dist = lorentz_distance(x1, x2, diag=diag).detach().clone()
    dist[dist < 1e-6] = 1e-6
    dist.requires_grad_(True)
    num = operator(dist).detach()
    d0 = torch.full((1, 1), 1e-6, dtype=F64, requires_grad=True)
    nrm = operator(d0).detach()
    if return_parts:
        return num, nrm
    return num / nrm


def hyperbolic_integrated_matern(x1, x2, dim, nu, lengthscale, nb_points=50, diag=False):
    """kernels_hyperbolic.py:258-389."""
    nu, ls = _as_param(nu), _as_param(lengthscale)
    if dim not in (1, 2, 3):
        raise NotImplementedError
    d = lorentz_distance(x1, x2, diag=diag)

    def heat(dist, t):
        if dim == 1:
            return torch.exp(-torch.pow(dist, 2) / (4 * t))
        if dim == 3:
            h = torch.exp(-torch.pow(dist, 2) / (4 * t))
            return h * (dist + 1e-8) / torch.sinh(dist + 1e-8)
        s = torch.linspace(1.5e-1, 100.0, nb_points, dtype=F64) + dist.unsqueeze(-1)
        vals = s * torch.exp(-s ** 2 / (4 * t)) / torch.sqrt(torch.cosh(s) - torch.cosh(dist.unsqueeze(-1)))
        return torch.trapz(vals, s, dim=-1)

    def link(dist, t):
        return torch.pow(t, nu - 1.0) * torch.exp(-2.0 * nu / ls ** 2 * t) * heat(dist, t)

    shift = torch.log10(ls).item()
    t_vals = torch.logspace(-2.5 + shift, 1.5 + shift, nb_points, dtype=F64)
    k = torch.trapz(torch.stack([link(d, t_vals[i]) for i in range(nb_points)]), t_vals, dim=0)
    z = torch.zeros(1, 1, dtype=F64)
    nrm = torch.trapz(torch.stack([link(z, t_vals[i]).reshape(()) for i in range(nb_points)]), t_vals, dim=0)
    return k / nrm


# ----------------------------------------------------------------------------------------------
# SPD(2)
# ----------------------------------------------------------------------------------------------
def mandel_to_sym(v):
    """Mandel vector (diagonal first, then the k-th super-diagonals / sqrt 2) -> symmetric matrix;
    Riemannian_utils/spd_utils_torch.py:336-371."""
    dvec = v.shape[-1]
    n = int((-1.0 + (1.0 + 8.0 * dvec) ** 0.5) / 2.0)
    out = torch.zeros(*v.shape[:-1], n, n, dtype=v.dtype)
    idx = torch.arange(n)
    out[..., idx, idx] = v[..., :n]
    pos = n
    for k in range(1, n):
        for i in range(n - k):
            out[..., i, i + k] = v[..., pos] / 2.0 ** 0.5
            out[..., i + k, i] = v[..., pos] / 2.0 ** 0.5
            pos += 1
    return out


def _spd2_log_singular(a, b):
    """H1 >= H2 = log singular values of A_i B_j^{-1} for all pairs; kernels_spd.py:230-256."""
    m, n = a.shape[-3], b.shape[-3]
    ae = a.unsqueeze(-3).expand(*a.shape[:-3], m, n, 2, 2).contiguous()
    be = b.unsqueeze(-4).expand(*b.shape[:-3], m, n, 2, 2).contiguous()
    prod = torch.bmm(ae.reshape(-1, 2, 2), torch.linalg.inv(be).reshape(-1, 2, 2))
    sv = torch.linalg.svdvals(prod).reshape(*ae.shape[:-2], 2)
    return torch.log(sv[..., 0]), torch.log(sv[..., 1])


def spd2_heat(x1, x2, lengthscale, nb_points=50):
    """kernels_spd.py:65-136 (uses the Cholesky factors of the two matrices, :90-91)."""
    ls = _as_param(lengthscale)
    h1, h2 = _spd2_log_singular(torch.linalg.cholesky(mandel_to_sym(x1)), torch.linalg.cholesky(mandel_to_sym(x2)))
    r2, dlt = h1 ** 2 + h2 ** 2, h1 - h2

    def link(delta, b):
        return (2.0 * b + delta) * torch.exp(-b * (b + delta) / ls ** 2) \
            * torch.pow(torch.sinh(b) * torch.sinh(b + delta), -1 / 2)

    b_vals = torch.logspace(-5.0, 1.0, nb_points, dtype=F64)
    k = torch.trapz(torch.stack([link(dlt, b_vals[i]) for i in range(nb_points)]), b_vals, dim=0)
    k = k * torch.exp(-r2 / (2 * ls ** 2))
    z = torch.zeros(1, 1, dtype=F64)
    nrm = torch.trapz(torch.stack([link(z, b_vals[i]).reshape(()) for i in range(nb_points)]), b_vals, dim=0)
    return k / nrm


def spd2_integrated_matern(x1, x2, nu, lengthscale, nb_b=50, nb_t=50):
    """kernels_spd.py:200-289."""
    nu, ls = _as_param(nu), _as_param(lengthscale)
    h1, h2 = _spd2_log_singular(mandel_to_sym(x1), mandel_to_sym(x2))
    r2, dlt = h1 ** 2 + h2 ** 2, h1 - h2

    def link(delta, rr, b, t):
        heat = (2.0 * b + delta) * torch.exp(-b * (b + delta) / (2 * t)) \
            * torch.pow(torch.sinh(b) * torch.sinh(b + delta), -1 / 2) * torch.exp(-rr / (4 * t))
        return torch.pow(t, nu - 1.0) * torch.exp(-2.0 * nu / ls ** 2 * t) * heat

    shift = torch.log10(ls).item()
    t_vals = torch.logspace(-2 + shift, 1.5 + shift, nb_t, dtype=F64)
    b_vals = torch.logspace(-5.0, 1.0, nb_b, dtype=F64)
    rows = []
    for i in range(nb_t):
        rows.append(torch.stack([link(dlt, r2, b_vals[j], t_vals[i]) for j in range(nb_b)]))
    k = torch.trapz(torch.trapz(torch.stack(rows), b_vals, dim=1), t_vals, dim=0)
    z = torch.zeros(1, 1, dtype=F64)
    nrows = []
    for i in range(nb_t):
        nrows.append(torch.stack([link(z, z, b_vals[j], t_vals[i]).reshape(()) for j in range(nb_b)]))
    nrm = torch.trapz(torch.trapz(torch.stack(nrows), b_vals, dim=1), t_vals, dim=0)
    return k / nrm


# ----------------------------------------------------------------------------------------------
# Euclidean factor of the SPD product kernels
# ----------------------------------------------------------------------------------------------
def euclidean_integrated_matern(x1, x2, nu, lengthscale, nb_points=100):
    """kernels_euclidean.py:82-89, 162-207."""
    nu, ls = _as_param(nu), _as_param(lengthscale)
    d2 = ((x1.unsqueeze(-2) - x2.unsqueeze(-3)) ** 2).sum(-1)

    def link(q, t):
        return torch.pow(t, nu - 1.0) * torch.exp(-2.0 * nu / ls ** 2 * t) \
            * torch.exp(-q / (2 * torch.pow(torch.pow(2 * t, 0.5), 2)))

    shift = torch.log10(ls).item()
    t_vals = torch.logspace(-3 + shift, 3 + shift, nb_points, dtype=F64)
    k = torch.trapz(torch.stack([link(d2, t_vals[i]) for i in range(nb_points)]), t_vals, dim=0)
    z = torch.zeros(1, 1, dtype=F64)
    nrm = torch.trapz(torch.stack([link(z, t_vals[i]).reshape(()) for i in range(nb_points)]), t_vals, dim=0)
    return k / nrm


def spd_product_matern(x1, x2, dim, nu, ls_euclidean, ls_manifold, manifold):
    """Product kernels on pre-decomposed SPD inputs (spd_input=False): Euclidean integrated Matern on the first
    ``dim`` columns (eigenvalues) times the sphere / circle / SO(dim) Matern kernel on the remaining ones;
    kernels_spd.py:415-549 (R x S: S^1 for dim 2, S^3 for dim 3) and :658-774 (R x SO)."""
    ke = euclidean_integrated_matern(x1[..., :dim], x2[..., :dim], nu, ls_euclidean)
    a, b = x1[..., dim:], x2[..., dim:]
    if manifold == "so":
        if dim != 3:
            raise NotImplementedError
        km = so3_matern(a, b, nu, ls_manifold)
    elif dim == 2:
        km = circle_integrated_matern(a, b, nu, ls_manifold)
    elif dim == 3:
        km = sphere_matern(a, b, 3, nu, ls_manifold)
    else:
        raise NotImplementedError
    return ke * km


def rotation_to_quaternion(rot):
    """Riemannian_utils/sphere_utils_torch.py:127-190 (Corke's algorithm), rot: (..., 3, 3) -> (..., 4)."""
    shape = rot.shape[:-2]
    r = rot.reshape(-1, 3, 3)
    out = torch.zeros(r.shape[0], 4, dtype=rot.dtype)
    for i in range(r.shape[0]):
        m = r[i]
        qs = min(float(torch.sqrt(m[0, 0] + m[1, 1] + m[2, 2] + 1) / 2.0), 1.0)
        kx, ky, kz = m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1]
        if m[0, 0] >= m[1, 1] and m[0, 0] >= m[2, 2]:
            k1 = (m[0, 0] - m[1, 1] - m[2, 2] + 1, m[1, 0] + m[0, 1], m[2, 0] + m[0, 2])
            add = kx >= 0
        elif m[1, 1] >= m[2, 2]:
            k1 = (m[1, 0] + m[0, 1], m[1, 1] - m[0, 0] - m[2, 2] + 1, m[2, 1] + m[1, 2])
            add = ky >= 0
        else:
            k1 = (m[2, 0] + m[0, 2], m[2, 1] + m[1, 2], m[2, 2] - m[0, 0] - m[1, 1] + 1)
            add = kz >= 0
        sg = 1.0 if add else -1.0
        kv = torch.stack([kx + sg * k1[0], ky + sg * k1[1], kz + sg * k1[2]])
        nm = kv.norm()
        out[i, 0] = qs
        if nm != 0:
            out[i, 1:] = (1 - qs ** 2) ** 0.5 / nm * kv
    return out.reshape(*shape, 4)


def spd_decompose(x, dim, target):
    """spd_to_so_and_euclidean_torch / spd_to_sphere_and_euclidean_torch (Riemannian_utils/spd_utils_torch.py:293-333)
    with torch.linalg.eigh in place of the removed torch.symeig and the sign convention the CUDA front-end states
    (largest-magnitude component of every eigenvector positive), then Q = det(V) V (:309-311).
    target 0 -> [eigenvalues | vec(Q)], target 1 -> [eigenvalues | first row of Q (dim 2) or quaternion (dim 3)]."""
    mat = mandel_to_sym(x)
    ev, vec = torch.linalg.eigh(mat)
    big = vec.abs().argmax(dim=-2, keepdim=True)
    sign = torch.sign(torch.gather(vec, -2, big))
    vec = vec * sign
    q = torch.linalg.det(vec).unsqueeze(-1).unsqueeze(-1) * vec
    if target == 0:
        tail = q.reshape(*q.shape[:-2], dim * dim)
    elif dim == 2:
        tail = q[..., 0, :]
    else:
        tail = rotation_to_quaternion(q)
    return torch.cat([ev, tail], dim=-1)


# ----------------------------------------------------------------------------------------------
# exact GP posterior + analytic EI (gpytorch / botorch arithmetic; unvendored - see SURVEY 8c)
# ----------------------------------------------------------------------------------------------
def gp_fit(kernel_fn, x_train, y_train, outputscale=1.0, noise=1e-2, mean_const=0.0):
    """L = chol(s K + noise I), alpha = (s K + noise I)^-1 (y - m)."""
    ktt = outputscale * kernel_fn(x_train, x_train) + noise * torch.eye(x_train.shape[0], dtype=F64)
    L = torch.linalg.cholesky(ktt)
    alpha = torch.cholesky_solve((y_train - mean_const).unsqueeze(-1), L).squeeze(-1)
    return L, alpha


def gp_predict(kernel_fn, x_train, L, alpha, x_cand, outputscale=1.0, mean_const=0.0, chunk=4096):
    """mean = m + s K* alpha;  var = s k** - || L^-1 s K*^T ||^2."""
    means, vars_ = [], []
    for s in range(0, x_cand.shape[0], chunk):
        xc = x_cand[s:s + chunk]
        ks = outputscale * kernel_fn(xc, x_train)
        # test x test block in botorch's b x q x D layout with q = 1 (manifold_optimize.py:323)
        kss = outputscale * kernel_fn(xc.unsqueeze(-2), xc.unsqueeze(-2)).reshape(-1)
        v = torch.linalg.solve_triangular(L, ks.transpose(-1, -2), upper=False)
        means.append(mean_const + ks @ alpha)
        vars_.append(kss - (v * v).sum(0))
    return torch.cat(means), torch.cat(vars_)


def gp_posterior(kernel_fn, x_train, y_train, x_cand, outputscale=1.0, noise=1e-2, mean_const=0.0,
                 chunk=4096):
    """Exact-GP posterior mean / variance with a constant mean, as gpytorch's ExactGP prediction computes it.
    Call sites: examples/.../bo_sphere.py:215-233,265; manifold_optimization/manifold_optimize.py:195,254,337."""
    L, alpha = gp_fit(kernel_fn, x_train, y_train, outputscale, noise, mean_const)
    return gp_predict(kernel_fn, x_train, L, alpha, x_cand, outputscale, mean_const, chunk)


def expected_improvement(mean, var, best_f, maximize=False):
    """botorch analytic EI: sigma (phi(u) + u Phi(u)), u = (mean - best_f)/sigma, sign flipped when minimising."""
    sigma = var.clamp_min(1e-9).sqrt()
    u = (mean - best_f) / sigma
    if not maximize:
        u = -u
    normal = torch.distributions.Normal(torch.zeros_like(u), torch.ones_like(u))
    return sigma * (torch.exp(normal.log_prob(u)) + u * normal.cdf(u))
