"""not gpu: the from-scratch character tables of the SO(n) kernels (materngabo_b200/_so_characters.py: exact Laurent-
polynomial division of Weyl's alternating sums) against the oracle's NUMERICAL evaluation of the same character formula
(oracle/restated.py) and against Weyl's dimension formula - two independent routes to the reference's polynomial
(kernel_utils/kernels_so.py:204-290)."""
import math

import numpy as np
import pytest
import torch

from materngabo_b200 import _so_characters as sc
from oracle import restated as R


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7])
def test_tables_match_numeric_weyl_formula(n):
    tb = sc.table(n)
    rk, bm, delta, group, summ = R._son_summands(n, 10)
    assert tb.rank == rk and tb.weyl_order == len(group) == {1: 2, 2: 6 if n == 4 else 8, 3: 24 if n == 6 else 48}[rk]
    assert len(summ) == len(tb.weights)
    assert sorted(tb.rep_dim) == sorted(int(round(d)) for _, _, d in summ)           # chi(1) vs Weyl's dimension formula
    assert np.allclose(sorted(tb.lb_ev), sorted(ev for _, ev, _ in summ), atol=1e-12)
    gen = torch.Generator().manual_seed(n)
    th = torch.rand(7, rk, generator=gen, dtype=torch.float64) * 3.0 + 0.05
    for kappa, nu in ((0.5, 2.5), (1.3, None)):
        ref = R.son_kernel_from_angles(th, n, kappa, nu)
        lam = np.array([(math.exp(-ev * kappa ** 2 / 2) if nu is None else (2 * nu / kappa ** 2 + ev) ** (-nu - tb.so_dim / 2)) * d
                        for ev, d in zip(tb.lb_ev, tb.rep_dim)])
        j, maps = tb.grid_maps()
        val = np.zeros(7)
        for s, mat in maps.items():
            grid = (lam @ mat).reshape((j + 1,) * rk)
            for idx in np.ndindex(*grid.shape):
                if grid[idx] == 0.0:
                    continue
                term = np.full(7, grid[idx])
                for q in range(rk):
                    term = term * (np.sin(idx[q] * th[:, q].numpy()) if s[q] else np.cos(idx[q] * th[:, q].numpy()))
                val += term
        val /= (lam * np.array(tb.rep_dim)).sum()
        assert np.abs(val - ref.numpy()).max() < 1e-10      # the numerical ratio of 48-term alternating sums carries ~1e-12


def test_reference_quirks_are_reproduced():
    """n = 4 uses the A_2 Cartan matrix (SU(3) dimensions 1, 3, 3, 6, 8, 6, ...), n = 6 the A_3 one incl. the spinor
    representations (4, 4): math_utils/so_root_systems.py:20-35 - reproduced, not corrected."""
    assert sorted(sc.table(4).rep_dim) == [1, 3, 3, 6, 6, 8, 10, 10, 15, 15]
    assert sorted(sc.table(6).rep_dim) == [1, 4, 4, 6, 10, 10, 15, 20, 20, 20]
    assert sc.table(5).rep_dim[:4] == [1, 5, 10, 14] and len(sc.table(5).weights) == 12
    assert sc.table(3).rep_dim == [2 * l + 1 for l in range(10)]
    with pytest.raises(ValueError):
        sc.SOCharacterTable(1)
