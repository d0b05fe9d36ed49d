"""not gpu: numpy models of the two digit-extraction schemes of the INT8 kernels (csrc/int8_fused.cu, csrc/series_i8.cu).

  * float64 scheme (f8_slice_kernel, g8_digits): seven times "round 256 r to nearest, keep the exact remainder", then a
    carry pass that rewrites a digit of +128 as -128 and increments its more significant neighbour;
  * integer scheme (f8_gram_slice_kernel): r 2^56 as two exact integers (the leading 32 bits and the following 24 of the
    exact remainder), then balanced base-256 digits from the least significant end ((v + 128) >> 8, digit = low byte).

Both must produce THE balanced representation (digits in [-128, 127]) of round(r 2^56): the same digits, and a value within
2^-57 of r - the premise of the error-free products on the tensor cores."""
import numpy as np


def digits_float_scheme(r):
    r = np.asarray(r, dtype=np.float64).copy()
    q = []
    for _ in range(7):
        t = np.rint(r * 256.0)
        q.append(t.astype(np.int64))
        r = r * 256.0 - t                      # exact: |r| <= 1/2
    for p in range(6, 0, -1):
        c = (q[p] == 128).astype(np.int64)
        q[p] = q[p] - 256 * c
        q[p - 1] = q[p - 1] + c
    return np.stack(q)


def digits_integer_scheme(r):
    r = np.asarray(r, dtype=np.float64)
    t1 = np.rint(r * 2.0 ** 32)
    hi4 = t1.astype(np.int64)
    rem = r * 2.0 ** 32 - t1                   # exact, |rem| <= 1/2
    lo3 = np.rint(rem * 2.0 ** 24).astype(np.int64)
    q = [None] * 7
    q[6] = lo3
    lo3 = (lo3 + 128) >> 8
    q[5] = lo3
    lo3 = (lo3 + 128) >> 8
    q[4] = lo3
    hi4 = hi4 + ((lo3 + 128) >> 8)
    q[3] = hi4
    hi4 = (hi4 + 128) >> 8
    q[2] = hi4
    hi4 = (hi4 + 128) >> 8
    q[1] = hi4
    q[0] = (hi4 + 128) >> 8
    return np.stack([((x & 255) ^ 128) - 128 for x in q])      # the low byte read as int8


def _samples():
    rng = np.random.default_rng(7)
    lim = 126.0 / 256.0
    edge = np.array([0.0, lim, -lim, 2.0 ** -30, -2.0 ** -30, 2.0 ** -56, 2.0 ** -57, 3 * 2.0 ** -57, -3 * 2.0 ** -57,
                     0.5 ** 9 + 0.5 ** 33, -(0.5 ** 9 + 0.5 ** 33), 127.5 / 256 ** 2, -127.5 / 256 ** 2, 0.5 / 256, -0.5 / 256,
                     128.0 / 256 ** 3, -128.0 / 256 ** 3, lim - 2.0 ** -54])
    return np.concatenate([rng.uniform(-lim, lim, 200000), rng.uniform(-lim, lim, 20000) * 2.0 ** -20, edge])


def test_both_schemes_give_the_same_balanced_digits():
    r = _samples()
    a, b = digits_float_scheme(r), digits_integer_scheme(r)
    assert a.min() >= -128 and a.max() <= 127 and b.min() >= -128 and b.max() <= 127
    assert np.array_equal(a, b)


def test_digits_reconstruct_the_value_to_half_a_unit_of_the_last_digit():
    r = _samples()
    d = digits_integer_scheme(r)
    rec = sum(d[p].astype(np.longdouble) * np.longdouble(2.0) ** (-8 * (p + 1)) for p in range(7))
    assert float(np.abs(rec - r.astype(np.longdouble)).max()) <= 2.0 ** -57
    assert int(np.abs(d[0]).max()) <= 127


def test_plane_bounds_leave_room_in_int32_and_in_the_recombination_words():
    """series_i8.cu: |P_s| <= (s + 1) W 2^14 for W <= 16 digits per chunk; hi = P0 2^16 + P1 2^8 + P2 and
    lo = P3 2^24 + P4 2^16 + P5 2^8 + P6 must stay below 2^51 (the 1.5 2^52 bit trick), the 32-bit partial sums below 2^31."""
    w = 16
    p = [(s + 1) * w * 128 * 128 for s in range(7)]
    assert max(p) < 2 ** 31
    assert p[1] * 256 + p[2] < 2 ** 31 and p[3] * 256 + p[4] < 2 ** 31 and p[5] * 256 + p[6] < 2 ** 31
    assert p[0] * 2 ** 16 + p[1] * 2 ** 8 + p[2] < 2 ** 51
    assert p[3] * 2 ** 24 + p[4] * 2 ** 16 + p[5] * 2 ** 8 + p[6] < 2 ** 51
    # int8_fused.cu: 7 n 128^2 < 2^31 up to n = 16384 (its stated limit)
    assert 7 * 16384 * 128 * 128 < 2 ** 31 <= 7 * 18725 * 128 * 128


def test_stacked_operand_rows_contract_to_the_planes_with_four_k_steps():
    """series_i8.cu's GEMM plan in numpy: an A row is its digit vectors side by side [d_0 | ... | d_6 | 0] (eight chunks of
    16), the B row of (plane s, column j) is [d_s(y_j) | d_(s-1)(y_j) | ... | d_0(y_j) | 0 ...]; K step a contracts the
    chunk pair (2a, 2a + 1) and is issued only against planes s >= 2a (the rows it skips hold zeros in that pair).  The
    accumulated result must be P_s = sum_{p + q = s} <d_p(x), d_q(y)> for every plane."""
    rng = np.random.default_rng(3)
    w, m, n = 11, 9, 13
    x = rng.uniform(-0.49, 0.49, (m, w))
    y = rng.uniform(-0.49, 0.49, (n, w))
    dx = np.stack([digits_integer_scheme(x[:, k]) for k in range(w)], axis=-1)      # [7][m][w]
    dy = np.stack([digits_integer_scheme(y[:, k]) for k in range(w)], axis=-1)      # [7][n][w]
    a_rows = np.zeros((m, 8, 16), dtype=np.int64)
    a_rows[:, :7, :w] = dx.transpose(1, 0, 2)
    b_rows = np.zeros((7, n, 8, 16), dtype=np.int64)                                # [plane][column][chunk][16]
    for s in range(7):
        for c in range(s + 1):
            b_rows[s, :, c, :w] = dy[s - c]
    acc = np.zeros((7, m, n), dtype=np.int64)
    for a in range(4):                                                              # one K = 32 MMA per chunk pair
        for s in range(2 * a, 7):                                                   # N = 32 (7 - 2a) stacked rows
            acc[s] += np.einsum("ick,jck->ij", a_rows[:, 2 * a:2 * a + 2], b_rows[s, :, 2 * a:2 * a + 2])
        assert not b_rows[:2 * a, :, 2 * a:2 * a + 2].any()                         # what the shortened MMAs skip is zero
    want = np.zeros_like(acc)
    for p in range(7):
        for q in range(7 - p):
            want[p + q] += dx[p] @ dy[q].T
    assert np.array_equal(acc, want)
    # and the planes recombine to the inner product: sum_s P_s 2^(-8 (s + 2)) = <x, y> up to the dropped products
    u = sum(want[s].astype(np.longdouble) * np.longdouble(2.0) ** (-8 * (s + 2)) for s in range(7))
    assert float(np.abs(u - x.astype(np.longdouble) @ y.astype(np.longdouble).T).max()) < 6 * w * 2.0 ** 14 * 2.0 ** -72 + w * 2.0 ** -56
