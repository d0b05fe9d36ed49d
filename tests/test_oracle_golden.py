"""not gpu: the restated oracle (oracle/restated.py) against the golden vectors generated from the UNMODIFIED
reference (oracle/make_golden.py), values and autograd gradients.  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import restated as R
from tests.conftest import load_golden, rel_err

PIN = 1e-13  # restatement vs reference: rounding-level agreement


def _t(a, grad=False):
    return torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=grad)


def _ls(raw_free):  # softplus(raw) parametrisation used by the reference (gpytorch Positive)
    return torch.nn.functional.softplus(raw_free)


def _inv_softplus(v):
    v = torch.tensor(float(v))
    return (v + torch.log(-torch.expm1(-v))).reshape(1, 1).requires_grad_(True)


CASES = {
    "sphere_matern_S2": lambda g, ls: R.sphere_matern(g["x1"], g["x2"], 2, 2.5, ls),
    "sphere_matern_S3": lambda g, ls: R.sphere_matern(g["x1"], g["x2"], 3, 2.5, ls),
    "sphere_matern_S10": lambda g, ls: R.sphere_matern(g["x1"], g["x2"], 10, 2.5, ls),
    "sphere_gaussian_S2": lambda g, ls: R.sphere_gaussian(g["x1"], g["x2"], 2, ls),
    "sphere_gaussian_S3": lambda g, ls: R.sphere_gaussian(g["x1"], g["x2"], 3, ls),
    "sphere_gaussian_S10": lambda g, ls: R.sphere_gaussian(g["x1"], g["x2"], 10, ls),
    "sphere_intmatern_S3": lambda g, ls: R.sphere_integrated_matern(g["x1"], g["x2"], 3, 2.5, ls),
    "circle_gaussian": lambda g, ls: R.circle_gaussian(g["x1"], g["x2"], ls),
    "circle_intmatern": lambda g, ls: R.circle_integrated_matern(g["x1"], g["x2"], 2.5, ls),
    "so3_matern": lambda g, ls: R.so3_matern(g["x1"], g["x2"], 2.5, ls),
    "so3_gaussian": lambda g, ls: R.so3_gaussian(g["x1"], g["x2"], ls),
    "hyperbolic_matern_H1": lambda g, ls: R.hyperbolic_integrated_matern(g["x1"], g["x2"], 1, 2.5, ls, 64),
    "hyperbolic_matern_H2": lambda g, ls: R.hyperbolic_integrated_matern(g["x1"], g["x2"], 2, 2.5, ls, 64),
    "hyperbolic_matern_H3": lambda g, ls: R.hyperbolic_integrated_matern(g["x1"], g["x2"], 3, 2.5, ls, 64),
    "hyperbolic_heat_H1": lambda g, ls: R.hyperbolic_heat(g["x1"], g["x2"], 1, ls),
    "hyperbolic_heat_H2": lambda g, ls: R.hyperbolic_heat(g["x1"], g["x2"], 2, ls),
    "hyperbolic_heat_H3": lambda g, ls: R.hyperbolic_heat(g["x1"], g["x2"], 3, ls),
    "euclidean_intmatern": lambda g, ls: R.euclidean_integrated_matern(g["x1"], g["x2"], 2.5, ls),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_value_and_gradients(name):
    raw = load_golden(name)
    g = {"x1": _t(raw["x1"], True), "x2": _t(raw["x2"], True)}
    raw_ls = _inv_softplus(raw["lengthscale"])
    k = CASES[name](g, _ls(raw_ls))
    assert rel_err(k, raw["K"]) < PIN
    res = torch.autograd.grad((k * _t(raw["G"])).sum(), [g["x1"], g["x2"], raw_ls])
    assert rel_err(res[0], raw["gx1"]) < 1e-11
    assert rel_err(res[1], raw["gx2"]) < 1e-11
    assert rel_err(res[2].reshape(-1), raw["g_raw_lengthscale"].reshape(-1)) < 1e-11


VALUE_ONLY = {
    "spd3_product_rso_matern": lambda g: R.spd_product_matern(g["x1"], g["x2"], 3, 2.5, 1.2, 0.6, "so"),
    "spd3_product_rs_matern": lambda g: R.spd_product_matern(g["x1"], g["x2"], 3, 2.5, 0.9, 0.5, "sphere"),
    "spd2_product_rs_matern": lambda g: R.spd_product_matern(g["x1"], g["x2"], 2, 1.5, 1.1, 0.4, "sphere"),
    "sphere_matern_S3_example": lambda g: R.sphere_matern(g["x1"], g["x2"], 3, 2.5, 0.5),
    "sphere_matern_S2_self": lambda g: R.sphere_matern(g["x1"], g["x2"], 2, 2.5, 0.5),
    "so3_matern_self": lambda g: R.so3_matern(g["x1"], g["x2"], 2.5, 0.5),
    "hyperbolic_matern_H2_self": lambda g: R.hyperbolic_integrated_matern(g["x1"], g["x2"], 2, 2.5, 1.0, 64),
    "spd2_matern": lambda g: R.spd2_integrated_matern(g["x1"], g["x2"], 2.5, 1.0, 64, 64),
    "spd2_matern_self": lambda g: R.spd2_integrated_matern(g["x1"], g["x2"], 2.5, 1.0, 64, 64),
    "spd2_heat": lambda g: R.spd2_heat(g["x1"], g["x2"], 1.4, 50),
}


@pytest.mark.parametrize("name", sorted(VALUE_ONLY))
def test_value_only(name):
    raw = load_golden(name)
    g = {"x1": _t(raw["x1"]), "x2": _t(raw["x2"])}
    assert rel_err(VALUE_ONLY[name](g), raw["K"]) < PIN


@pytest.mark.parametrize("dim", [3, 10])
def test_torus(dim):
    raw = load_golden("torus_matern_T%d" % dim)
    k = R.torus_matern(_t(raw["x1"]), _t(raw["x2"]), dim, 2.5, [float(v) for v in raw["lengthscales"]])
    assert rel_err(k, raw["K"]) < PIN
    raw = load_golden("torus_gaussian_T%d" % dim)
    k = R.torus_gaussian(_t(raw["x1"]), _t(raw["x2"]), dim, [float(v) for v in raw["lengthscales"]])
    assert rel_err(k, raw["K"]) < PIN


def test_spd2_heat_lengthscale_gradient():
    raw = load_golden("spd2_heat_lsgrad")
    raw_ls = _inv_softplus(raw["lengthscale"])
    k = R.spd2_heat(_t(raw["x1"]), _t(raw["x2"]), _ls(raw_ls), 50)
    assert rel_err(k, raw["K"]) < PIN
    (gl,) = torch.autograd.grad((k * _t(raw["G"])).sum(), [raw_ls])
    assert rel_err(gl.reshape(-1), raw["g_raw_lengthscale"].reshape(-1)) < 1e-11


def test_nu_gradient():
    raw = load_golden("sphere_matern_S2_nugrad")
    raw_ls, raw_nu = _inv_softplus(raw["lengthscale"]), _inv_softplus(raw["nu"])
    k = R.sphere_matern(_t(raw["x1"]), _t(raw["x2"]), 2, _ls(raw_nu), _ls(raw_ls))
    assert rel_err(k, raw["K"]) < PIN
    res = torch.autograd.grad((k * _t(raw["G"])).sum(), [raw_ls, raw_nu])
    assert rel_err(res[0].reshape(-1), raw["g_raw_lengthscale"].reshape(-1)) < 1e-11
    assert rel_err(res[1].reshape(-1), raw["g_raw_nu"].reshape(-1)) < 1e-11


def test_known_structure():
    """Known answers independent of the fixtures: unit diagonal up to the reference's clamps (SURVEY.md section 5:
    S^d ~ 1, SO(3) diag = 1 - O(1e-8) because of the 1-1e-8 clip), symmetry, positive definiteness."""
    torch.manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(40, 3), dim=-1)
    k = R.sphere_matern(x, x, 2, 2.5, 0.5)
    assert float((k.diagonal() - 1).abs().max()) < 1e-12
    assert torch.linalg.eigvalsh(k).min() > 0
    q, _ = torch.linalg.qr(torch.randn(12, 3, 3))
    q[:, :, 0] *= torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    k3 = R.so3_matern(q.reshape(12, 9), q.reshape(12, 9), 2.5, 0.5)
    assert 1e-9 < float((1 - k3.diagonal()).max()) < 1e-6
    assert rel_err(k3, k3.t()) < 1e-12


def test_gp_posterior_identities():
    torch.manual_seed(1)
    xt = torch.nn.functional.normalize(torch.randn(30, 4), dim=-1)
    y = torch.sin(xt[:, 0] * 3)
    kfn = lambda a, b: R.sphere_matern(a, b, 3, 2.5, 0.5)  # noqa: E731
    mean, var = R.gp_posterior(kfn, xt, y, xt, outputscale=1.0, noise=1e-6)
    assert float((mean - y).abs().max()) < 1e-3          # interpolates when the noise vanishes
    assert float(var.abs().max()) < 1e-4
    far = R.gp_posterior(kfn, xt[:1], y[:1], -xt[:1], outputscale=2.0, noise=1e-2, mean_const=0.3)
    assert abs(float(far[1]) - 2.0) < 0.2                 # antipodal point: nearly the prior variance


HVP_CASES = {
    "hyperbolic_matern_H1_hvp": lambda a, b: R.hyperbolic_integrated_matern(a, b, 1, 2.5, 0.9, 32),
    "hyperbolic_matern_H2_hvp": lambda a, b: R.hyperbolic_integrated_matern(a, b, 2, 2.5, 0.9, 32),
    "hyperbolic_matern_H3_hvp": lambda a, b: R.hyperbolic_integrated_matern(a, b, 3, 2.5, 0.9, 32),
    "hyperbolic_heat_H1_hvp": lambda a, b: R.hyperbolic_heat(a, b, 1, 1.1, 200),
    "hyperbolic_heat_H2_hvp": lambda a, b: R.hyperbolic_heat(a, b, 2, 1.1, 200),
    "hyperbolic_heat_H3_hvp": lambda a, b: R.hyperbolic_heat(a, b, 3, 1.1, 200),
    "euclidean_intmatern_hvp": lambda a, b: R.euclidean_integrated_matern(a, b, 1.5, 0.8),
}


@pytest.mark.parametrize("name", sorted(HVP_CASES))
def test_input_gradient_and_hessian_vector_product(name):
    """What the trust-region inner loop differentiates (manifold_optimize.py:184-217): gradient and Hessian-vector
    product with respect to the candidate point, through the restated kernels."""
    raw = load_golden(name)
    a = _t(raw["x1"], True)
    k = HVP_CASES[name](a, _t(raw["x2"]))
    assert rel_err(k, raw["K"]) < PIN
    (ga,) = torch.autograd.grad((k * _t(raw["G"])).sum(), [a], create_graph=True)
    assert rel_err(ga, raw["gx1"]) < 1e-11
    (hv,) = torch.autograd.grad((ga * _t(raw["V"])).sum(), [a])
    assert rel_err(hv, raw["hvp_x1"]) < 1e-10


@pytest.mark.parametrize("name", ["spd2_matern_xgrad", "spd2_heat_xgrad"])
def test_spd2_input_gradients(name):
    raw = load_golden(name)
    a, b = _t(raw["x1"], True), _t(raw["x2"], True)
    if name == "spd2_matern_xgrad":
        k = R.spd2_integrated_matern(a, b, 2.5, 1.2, 40, 36)
    else:
        k = R.spd2_heat(a, b, 1.4, 50)
    assert rel_err(k, raw["K"]) < PIN
    ga, gb = torch.autograd.grad((k * _t(raw["G"])).sum(), [a, b])
    assert rel_err(ga, raw["gx1"]) < 1e-10
    assert rel_err(gb, raw["gx2"]) < 1e-10


@pytest.mark.parametrize("n", [2, 4, 5, 6])
@pytest.mark.parametrize("kind", ["matern", "gaussian"])
def test_son_oracle_against_reference_goldens(n, kind):
    """SO(n), n != 3 (oracle/make_golden_son.py): the restatement evaluates Weyl's character formula numerically at the
    arguments the reference used (recorded per pair: they depend on LAPACK's eigenvalue order) - value on ALL pairs and
    the raw_lengthscale gradient of the masked sum."""
    raw = load_golden("son_%s_SO%d" % (kind, n))
    ang = _t(raw["angles"])
    raw_ls = _inv_softplus(raw["lengthscale"])
    x1, x2 = _t(raw["x1"]), _t(raw["x2"])
    fn = (lambda ls: R.son_matern(x1, x2, n, 2.5, ls, angles=ang)) if kind == "matern" else \
        (lambda ls: R.son_gaussian(x1, x2, n, ls, angles=ang))
    k = fn(_ls(raw_ls))
    assert rel_err(k, raw["K"]) < 1e-12
    (gl,) = torch.autograd.grad((k * _t(raw["G"])).sum(), [raw_ls])
    assert rel_err(gl.reshape(-1), raw["g_raw_lengthscale"].reshape(-1)) < 1e-10
    assert raw["canonical"].sum() >= 0.3 * raw["canonical"].size       # a usable share of pairs is in canonical order
