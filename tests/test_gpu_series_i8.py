"""-m gpu: the opt-in INT8 tensor-core series Gram kernel (csrc/series_i8.cu, MGB_GRAM_PATH=int8) - accumulator planes
against an exact integer matmul of the same digits, K against the golden vectors generated from the unmodified
reference (1e-10 relative to max|K_ref|) and against a long-double evaluation, shapes / edge cases of the boundary."""
import numpy as np
import pytest
import torch

from tests.conftest import TOL, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F64 = torch.float64
COEF10 = [0.11, 0.2, 0.15, 0.1, 0.09, 0.08, 0.07, 0.06, 0.05, 0.04]


def digits_np(x):
    """numpy restatement of the kernel's slicing: rows -> (exponent e, digits [7][rows][W]) with
    x = 2^e sum_p d_p 2^(-8 (p + 1)) up to 2^(e - 57), digits in [-128, 127] after the carry pass."""
    x = np.asarray(x, dtype=np.float64)
    mx = np.abs(x).max(axis=1)
    f, e = np.frexp(mx)
    e = np.where(mx > 0, e + np.where(f > 0.984375, 2, 1), 0)
    r = x * np.ldexp(1.0, -e)[:, None]
    q = []
    for _ in range(7):
        t = np.rint(r * 256.0)          # round half to even, as the magic-number add does
        q.append(t.astype(np.int64))
        r = r * 256.0 - t
    for p in range(6, 0, -1):
        c = (q[p] == 128).astype(np.int64)
        q[p] = q[p] - 256 * c
        q[p - 1] = q[p - 1] + c
    return e, np.stack(q)


def planes_np(x1, x2):
    _, da = digits_np(x1)
    _, db = digits_np(x2)
    pl = np.zeros((7, x1.shape[0], x2.shape[0]), dtype=np.int64)
    for p in range(7):
        for q in range(7 - p):
            pl[p + q] += da[p] @ db[q].T
    return pl


def truth_ld(spec, x1, x2, coef):
    u = x1.numpy().astype(np.longdouble) @ x2.numpy().astype(np.longdouble).T
    u = spec.scale * (u * u if spec.pair_square else u) + spec.shift
    u = np.clip(u, np.longdouble(spec.lo), np.longdouble(spec.hi))
    k = np.zeros_like(u)
    for ck in np.asarray(coef)[::-1]:
        k = k * u + np.longdouble(ck)
    return k


def _points(m, w, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(m, w, generator=g, dtype=F64), dim=-1) * scale


CASES = [
    # name, (n_factors, width, basis, scale, shift, lo, hi, pair_square), m, n, input scale, coefficients
    ("s10_ragged", (1, 11, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15, 0), 300, 100, 1.0, COEF10),
    ("s2", (1, 3, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15, 0), 256, 64, 1.0, COEF10),
    ("so3_trace", (1, 9, 0, 0.5, -0.5, -1.0 + 1e-8, 1.0 - 1e-8, 0), 130, 40, 3.0 ** 0.5, COEF10),
    ("quaternion_square", (1, 4, 0, 2.0, -1.0, -1.0 + 1e-8, 1.0 - 1e-8, 1), 257, 33, 1.0, COEF10),
    ("w16_t12", (1, 16, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15, 0), 128, 32, 1.0, COEF10 + [0.03, 0.02]),
    ("single_row_and_column", (1, 4, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15, 0), 1, 1, 1.0, COEF10[:3]),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_planes_are_exact_and_k_matches_long_double(case):
    from materngabo_b200 import _ops

    _, sp, m, n, scale, coef = case
    spec = _ops.SeriesSpec(*sp)
    x1, x2 = _points(m, sp[1], 1, scale), _points(n, sp[1], 2, scale)
    x2[: min(5, m, n)] = x1[: min(5, m, n)]                    # repeated points: the reference's clamp is active
    c = torch.tensor(coef, dtype=F64)
    k, pl = _ops.series_gram_i8(spec, x1.to(DEV), x2.to(DEV), c.to(DEV), planes=True)
    got = pl.cpu().numpy()                                     # [panel][tile][7][128][32]
    got = got.transpose(2, 0, 3, 1, 4).reshape(7, got.shape[0] * 128, got.shape[1] * 32)[:, :m, :n]
    assert np.array_equal(got, planes_np(x1.numpy(), x2.numpy()))      # int32 accumulators = exact integer matmul
    ref = truth_ld(spec, x1, x2, coef)
    scale_k = float(np.abs(ref).max())
    assert float(np.abs(k.cpu().numpy() - ref).max()) <= 1e-14 * scale_k
    k2 = _ops.series_gram_i8(spec, x1.to(DEV), x2.to(DEV), c.to(DEV))  # the production instantiation
    assert torch.equal(k, k2)


def test_more_than_4096_columns_and_unaligned_rows():
    """n > 4096: one launch per 4096-column block; odd ldk = n: the 8-byte store path."""
    from materngabo_b200 import _ops

    spec = _ops.SeriesSpec(1, 11, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15)
    x1, x2 = _points(200, 11, 3), _points(4131, 11, 4)
    c = torch.tensor(COEF10, dtype=F64)
    k = _ops.series_gram_i8(spec, x1.to(DEV), x2.to(DEV), c.to(DEV))
    ref = truth_ld(spec, x1, x2, COEF10)
    assert float(np.abs(k.cpu().numpy() - ref).max()) <= 1e-14 * float(np.abs(ref).max())


def test_against_the_float64_kernel_at_a_bandwidth_bound_size():
    from materngabo_b200 import _ops

    spec = _ops.SeriesSpec(1, 11, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15)
    x1, x2 = _points(40000, 11, 5).to(DEV), _points(2000, 11, 6).to(DEV)
    c = torch.tensor(COEF10, dtype=F64, device=DEV)
    k8 = _ops.series_gram_i8(spec, x1, x2, c)
    k64 = _ops._series_fwd_raw(spec, x1, x2, c)
    assert float((k8 - k64).abs().max()) <= 1e-14 * float(k64.abs().max())


def test_non_finite_rows_and_columns_become_nan_only_there():
    from materngabo_b200 import _ops

    spec = _ops.SeriesSpec(1, 3, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15)
    x1, x2 = _points(140, 3, 7), _points(70, 3, 8)
    x1[17, 1] = float("nan")
    x1[130, 0] = float("inf")
    x2[33, 2] = float("nan")
    c = torch.tensor(COEF10, dtype=F64)
    k = _ops.series_gram_i8(spec, x1.to(DEV), x2.to(DEV), c.to(DEV)).cpu()
    bad = torch.zeros(140, 70, dtype=torch.bool)
    bad[17, :] = bad[130, :] = True
    bad[:, 33] = True
    assert torch.isnan(k[bad]).all() and torch.isfinite(k[~bad]).all()
    clean1, clean2 = x1.clone(), x2.clone()
    clean1[17], clean1[130], clean2[33] = _points(1, 3, 9)[0], _points(1, 3, 10)[0], _points(1, 3, 11)[0]
    ref = truth_ld(spec, clean1, clean2, COEF10)
    assert float(np.abs(k.numpy() - ref)[~bad.numpy()].max()) <= 1e-14 * float(np.abs(ref).max())


@pytest.mark.parametrize("name,cls", [("sphere_matern_S10", "SphereRiemannianMaternKernel"), ("sphere_matern_S2", "SphereRiemannianMaternKernel"),
                                      ("sphere_gaussian_S3", "SphereRiemannianGaussianKernel")])
def test_drop_in_kernels_through_the_int8_path_match_the_reference_goldens(name, cls, monkeypatch):
    from materngabo_b200 import kernels_sphere as ks

    monkeypatch.setenv("MGB_GRAM_PATH", "int8")
    g = load_golden(name)
    kw = {"nu": float(g["nu"])} if "matern" in name else {}
    k = getattr(ks, cls)(dim=int(g["dim"]), **kw)
    k.lengthscale = float(g["lengthscale"])
    x1 = torch.tensor(np.asarray(g["x1"]), dtype=F64, device=DEV)
    x2 = torch.tensor(np.asarray(g["x2"]), dtype=F64, device=DEV)
    assert rel_err(k.forward(x1, x2), g["K"]) < TOL


def test_so3_through_the_int8_path_matches_the_reference_golden(monkeypatch):
    from materngabo_b200 import kernels_so

    monkeypatch.setenv("MGB_GRAM_PATH", "int8")
    g = load_golden("so3_matern_self")
    k = kernels_so.SORiemannianMaternKernel(dim=3, nu=2.5)
    k.lengthscale = float(g["lengthscale"])
    x1 = torch.tensor(np.asarray(g["x1"]), dtype=F64, device=DEV)
    x2 = torch.tensor(np.asarray(g["x2"]), dtype=F64, device=DEV)
    assert rel_err(k.forward(x1, x2), g["K"]) < TOL


def test_unsupported_descriptors_are_rejected():
    from materngabo_b200 import _lib, _ops

    torus = _ops.SeriesSpec(3, 2, 1, 1.0, 0.0, -1.0, 1.0)
    x = torch.zeros(4, 6, dtype=F64, device=DEV)
    with pytest.raises(_lib.MgbError):
        _ops.series_gram_i8(torus, x, x, torch.ones(3, 5, dtype=F64, device=DEV))
