"""BASELINE.json configs[1]: GaBO on S^3, 100 training points - latency of the calls the BO loop makes
(SURVEY.md 8d cfg 2: 'latency-bound: report microseconds per call, not roofline').
  (i)  K(X, X) forward + backward to raw_lengthscale (one MLL step), N = 100
  (ii) acquisition posterior / EI for b = 100 and b = 500 candidates in botorch's b x 1 x 4 layout
  (iii) trust-region inner step: EI value + gradient for ONE candidate (b = 1)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere  # noqa: E402
from materngabo_b200.gpytorch_compat import ScaleKernel  # noqa: E402
from materngabo_b200.posterior import ExactGPPosterior, ExpectedImprovement  # noqa: E402

dev = "cuda:0"


def timeit(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


gen = torch.Generator().manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(100, 4, generator=gen), dim=-1).to(dev)
y = torch.sin(3 * x[:, 0])
base = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5)
sk = ScaleKernel(base).to(dev)
out = {}


def mll_step():
    for p in sk.parameters():
        p.grad = None
    base.raw_lengthscale.requires_grad_(True)
    k = sk(x)
    L = torch.linalg.cholesky(k + 1e-2 * torch.eye(100, device=dev))
    a = torch.cholesky_solve(y.unsqueeze(-1), L).squeeze(-1)
    loss = 0.5 * (y * a).sum() + torch.log(torch.diagonal(L)).sum()
    loss.backward()


out["mll_step_fwd_bwd_us"] = timeit(mll_step, reps=100)
with torch.no_grad():
    out["gram_100x100_fwd_us"] = timeit(lambda: base.forward(x, x))
base.raw_lengthscale.requires_grad_(False)
gp = ExactGPPosterior(base, x, y, outputscale=1.0, noise=1e-2).fit()
acq = ExpectedImprovement(gp, best_f=float(y.min()), maximize=False)
for b in (100, 500):
    xc = torch.nn.functional.normalize(torch.randn(b, 1, 4, generator=gen), dim=-1).to(dev)
    with torch.no_grad():
        out["ei_b%d_us" % b] = timeit(lambda: acq(xc))
x1 = torch.nn.functional.normalize(torch.randn(1, 1, 4, generator=gen), dim=-1).to(dev)


def tr_step():
    xx = x1.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(-acq(xx).sum(), [xx])
    return g


out["ei_value_and_grad_b1_us"] = timeit(tr_step, reps=100)
out["fit_n100_us"] = timeit(lambda: gp.fit(), reps=50)
print(json.dumps(out))
