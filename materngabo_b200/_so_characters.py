"""Host-side tables for the SO(n) kernels with n != 3: the character sum of the reference as an explicit cosine series.

What the reference does (kernel_utils/kernels_so.py:204-290, math_utils/so_root_systems.py): it builds, symbolically,

    kernel(x_1 .. x_r) = sum_pi  lambda_pi(kappa, nu) * chi_pi(x)          r = n // 2,

over the first >= 10 "analytically integral" highest weights pi ordered by height, with chi_pi from Weyl's character
formula written in SIMPLE-ROOT coordinates: a weight v (root basis) contributes the monomial prod_i x_i^((v^T B)_i), B
the symmetrised Cartan matrix scaled to integers.  ``forward`` (:329-362) then substitutes x_i = one eigenvalue of the
i-th conjugate pair of R1 R2^T *in the order torch.linalg.eigvals lists the pairs* and divides by the value at x = 1.

This module rebuilds the same Laurent polynomials without sympy: Cartan matrix -> B, fundamental weights, Weyl group
(closure of the simple reflections in the root basis), alternating sums as {exponent vector: integer} dictionaries and
an EXACT multivariate Laurent-polynomial division (the reference calls ``sympy.simplify(factor / factor)``), all in
integer / rational arithmetic.  On the unit torus x_i = exp(i theta_i) the real part of the kernel is

    sum_m C_m cos(<m, theta>),    C_m = sum_pi lambda_pi mult_pi(m),

which is regrouped into dense coefficient grids over products of cos(a theta_i) / sin(a theta_i) - the form the device
kernel evaluates with Chebyshev recurrences (csrc/son_gram.cu).

Order dependence: the reference's value depends on the order of the conjugate pairs (root coordinates are not
symmetric under permuting eigen-angles; DESIGN.md section 8, oracle/probe_so_n_symmetry.py).  The device kernel fixes
the order canonically - eigen-angles in (0, pi), DESCENDING - so that the result is a function of the pair; it equals
the reference wherever LAPACK happens to list the pairs in that order.
"""
from __future__ import annotations

import itertools
from fractions import Fraction
from functools import lru_cache
from math import lcm

import numpy as np


# ---- root-system data, following math_utils/so_root_systems.py:14-60 -------------------------------
def cartan_matrix_so(n: int):
    """B_k for odd n, the reference's "D_k" formula for even n (for k = 2 this is the A_2 matrix and for k = 3 the A_3
    = D_3 matrix with the reference's node numbering), [[2]] for n = 2, 3."""
    if n in (2, 3):
        return [[Fraction(2)]]
    k, odd = divmod(n, 2)
    c = [[Fraction(2 if i == j else (-1 if abs(i - j) == 1 else 0)) for j in range(k)] for i in range(k)]
    if odd:
        c[k - 2][k - 1] = Fraction(-2)
    elif k > 2:
        blk = [[2, -1, -1], [-1, 2, 0], [-1, 0, 2]]
        for i in range(3):
            for j in range(3):
                c[k - 3 + i][k - 3 + j] = Fraction(blk[i][j])
    return c


def symmetrised_gram(cm):
    """Gram matrix of the simple roots up to an integer scale (so_root_systems.py:39-60): root lengths propagated along
    the Dynkin diagram from node 0 with |a_y|^2 / |a_x|^2 = cm[y][x] / cm[x][y], then G[x][y] = cm[x][y] |a_y|^2 / 2,
    scaled by the lcm of the denominators."""
    rk = len(cm)
    length = [None] * rk
    for start in range(rk):
        if length[start] is not None:
            continue
        length[start] = Fraction(1)
        frontier = [start]
        while frontier:
            nxt = []
            for x in frontier:
                for y in range(rk):
                    if cm[x][y] != 0 and y != x and length[y] is None:
                        length[y] = length[x] * cm[y][x] / cm[x][y]
                        nxt.append(y)
            frontier = nxt
    g = [[cm[x][y] * length[y] / 2 for y in range(rk)] for x in range(rk)]
    scale = 1
    for row in g:
        for v in row:
            scale = lcm(scale, v.denominator)
    return [[v * scale for v in row] for row in g]


def _inverse(mat):
    n = len(mat)
    a = [list(row) + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(mat)]
    for col in range(n):
        piv = next(r for r in range(col, n) if a[r][col] != 0)
        a[col], a[piv] = a[piv], a[col]
        pv = a[col][col]
        a[col] = [v / pv for v in a[col]]
        for r in range(n):
            if r != col and a[r][col] != 0:
                f = a[r][col]
                a[r] = [v - f * w for v, w in zip(a[r], a[col])]
    return [row[n:] for row in a]


def _matvec(m, v):
    return tuple(sum(m[i][j] * v[j] for j in range(len(v))) for i in range(len(m)))


def _matmul(a, b):
    n = len(a)
    return tuple(tuple(sum(a[i][k] * b[k][j] for k in range(n)) for j in range(n)) for i in range(n))


def _det(m):
    n = len(m)
    if n == 1:
        return m[0][0]
    if n == 2:
        return m[0][0] * m[1][1] - m[0][1] * m[1][0]
    return sum((-1) ** j * m[0][j] * _det([row[:j] + row[j + 1:] for row in m[1:]]) for j in range(n))


def weyl_group(bm):
    """All products of the simple reflections w_i(v) = v - 2 b(v, a_i) / b(a_i, a_i) a_i as matrices in the root basis
    (so_root_systems.py:77-84; closure loop of kernels_so.py:216-227), with their determinants."""
    rk = len(bm)
    refl = []
    for i in range(rk):
        w = [[Fraction(int(r == c)) for c in range(rk)] for r in range(rk)]
        for j in range(rk):
            w[i][j] -= 2 * bm[j][i] / bm[i][i]
        refl.append(tuple(tuple(row) for row in w))
    group = set(refl)
    frontier = list(refl)
    while frontier:
        nxt = []
        for w in frontier:
            for s in refl:
                p = _matmul(s, w)
                if p not in group:
                    group.add(p)
                    nxt.append(p)
        frontier = nxt
    return [(w, int(_det([list(r) for r in w]))) for w in group]


# ---- Laurent polynomials as {exponent tuple: int} ---------------------------------------------------
def _alternating_sum(weight, group, bm, double):
    """sum_w det(w) x^((w weight)^T B); exponents multiplied by ``double`` (2 when half-integers occur)."""
    rk = len(bm)
    poly = {}
    for w, sign in group:
        v = _matvec(w, weight)
        e = tuple(sum(v[i] * bm[i][j] for i in range(rk)) * double for j in range(rk))
        if any(x.denominator != 1 for x in e):
            raise ValueError("non-integral exponent in the alternating sum")
        e = tuple(int(x) for x in e)
        poly[e] = poly.get(e, 0) + sign
    return {e: c for e, c in poly.items() if c != 0}


def _exact_quotient(num, den):
    """num / den for Laurent polynomials with integer coefficients (lexicographic leading terms; the leading
    coefficient of a Weyl denominator is +-1).  Raises if the division is not exact."""
    num = dict(num)
    lead_d = max(den)
    cd = den[lead_d]
    quo = {}
    while num:
        lead_n = max(num)
        cn = num[lead_n]
        if cn % cd != 0:
            raise ArithmeticError("character division is not exact")
        q = cn // cd
        shift = tuple(a - b for a, b in zip(lead_n, lead_d))
        quo[shift] = quo.get(shift, 0) + q
        for e, c in den.items():
            t = tuple(a + b for a, b in zip(e, shift))
            v = num.get(t, 0) - q * c
            if v:
                num[t] = v
            else:
                num.pop(t, None)
        if len(quo) > 100000:
            raise ArithmeticError("character division does not terminate")
    return quo


class SOCharacterTable:
    """Irreducible summands of the reference's SO(n) kernel: for each highest weight the Laplace-Beltrami eigenvalue
    (in the reference's scaled form), the dimension, and the character as {x-exponent vector: multiplicity}."""

    def __init__(self, n: int, summands_number_bound: int = 10):
        if n < 2:
            raise ValueError("SO(n) needs n >= 2")
        self.n = n
        self.rank = rk = n // 2
        self.so_dim = 2 * rk * rk + rk if n % 2 else 2 * rk * rk - rk          # kernels_so.py:210
        cm = cartan_matrix_so(n)
        bm = symmetrised_gram(cm)
        inv = _inverse(cm)
        fwsm = [[inv[j][i] for j in range(rk)] for i in range(rk)]            # (cm^-1)^T: fundamental weights, root basis
        delta = _matvec(fwsm, [Fraction(1)] * rk)
        group = weyl_group(bm)
        self.weyl_order = len(group)

        def form(v):
            return sum(v[i] * bm[i][j] * v[j] for i in range(rk) for j in range(rk))

        den2 = _alternating_sum(delta, group, bm, 2)
        self.weights, self.lb_ev, self.rep_dim, self.characters = [], [], [], []
        for height in itertools.count(0):
            for pi in sorted(p for p in itertools.product(range(height + 1), repeat=rk) if sum(p) == height):
                pi_vec = _matvec(fwsm, [Fraction(x) for x in pi])
                if any(x.denominator != 1 for x in _matvec(bm, pi_vec)):       # analytically integral? (:261)
                    continue
                shifted = tuple(a + b for a, b in zip(pi_vec, delta))
                quo2 = _exact_quotient(_alternating_sum(shifted, group, bm, 2), den2)
                if any(x % 2 for e in quo2 for x in e):
                    raise ArithmeticError("character with half-integral exponents")
                chi = {tuple(x // 2 for x in e): c for e, c in quo2.items()}
                self.weights.append(pi)
                self.lb_ev.append(float(form(shifted) - form(delta)))          # (:264)
                self.rep_dim.append(int(sum(chi.values())))                    # chi(1) (:271)
                self.characters.append(chi)
            if len(self.weights) >= summands_number_bound:                     # (:281)
                break
        self.max_exponent = max(abs(x) for chi in self.characters for e in chi for x in e)

    # ---- cosine-product grids ------------------------------------------------------------------
    def grid_maps(self):
        """(J, maps): the real part of sum_pi lambda_pi chi_pi(exp(i theta)) equals

            sum_{s in {0,1}^r, |s| even}  sum_{a in [0..J]^r}  G_s[a]  prod_i (cos(a_i theta_i) if s_i == 0 else sin(a_i theta_i))

        and ``maps[s]`` is the (n_summands x (J+1)^r) float64 matrix with G_s = lambda^T maps[s].  (cos(<m, theta>)
        expands into products with an even number of sines; sign bookkeeping from cos(A + B) = cc - ss.)"""
        rk, j1 = self.rank, self.max_exponent + 1
        patterns = [s for s in itertools.product((0, 1), repeat=rk) if sum(s) % 2 == 0]
        maps = {s: np.zeros((len(self.weights), j1 ** rk)) for s in patterns}
        for p, chi in enumerate(self.characters):
            for e, mult in chi.items():
                a = tuple(abs(x) for x in e)
                flat = 0
                for x in a:
                    flat = flat * j1 + x
                # cos(sum_i e_i t_i) = Re prod_i (cos(a_i t_i) + i sgn(e_i) sin(a_i t_i))
                for s in patterns:
                    k = sum(s)
                    sign = (-1) ** (k // 2)
                    for i in range(rk):
                        if s[i]:
                            if e[i] == 0:
                                sign = 0
                                break
                            sign *= 1 if e[i] > 0 else -1
                    if sign:
                        maps[s][p, flat] += sign * mult
        return self.max_exponent, maps


@lru_cache(maxsize=None)
def table(n: int, summands_number_bound: int = 10) -> SOCharacterTable:
    return SOCharacterTable(n, summands_number_bound)
