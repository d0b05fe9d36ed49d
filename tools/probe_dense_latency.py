"""Latency of the multi-CTA dense path (csrc/dense_spd.cu) against the torch.linalg route: marginal-likelihood step
(value + gradients) at N = 300 / 500 / 1000 and the posterior fit at n = 1000 / 2000 / 5000."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere  # noqa: E402
from materngabo_b200.gpytorch_compat import ScaleKernel  # noqa: E402
from materngabo_b200.graphs import exact_mll  # noqa: E402
from materngabo_b200.posterior import ExactGPPosterior  # noqa: E402

dev = "cuda:0"


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


gen = torch.Generator().manual_seed(0)
for n in (300, 500, 1000, 2000):
    x = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1).to(dev)
    y = torch.sin(3 * x[:, 0])
    base = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    sk = ScaleKernel(base).to(dev)
    wrt = [base.raw_lengthscale, sk.raw_outputscale]

    def step(fused):
        loss, _ = exact_mll(sk, x, y, 1e-2, 0.0, fused=fused)
        g = torch.autograd.grad(loss, wrt)
        return float(loss)

    print("MLL step n=%d: multi-CTA path %.3f ms, torch.linalg + autograd %.3f ms" % (n, timeit(lambda: step(True), 20), timeit(lambda: step(False), 20)), flush=True)

for n in (1000, 2000, 5000):
    xt = torch.nn.functional.normalize(torch.randn(n, 11, generator=gen), dim=-1).to(dev)
    y = torch.sin(3 * xt[:, 0])
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
    k.lengthscale = 0.5
    gp = ExactGPPosterior(k, xt, y, outputscale=1.0, noise=1e-2)
    own = timeit(lambda: gp.fit(), 5)
    os.environ["MGB_FIT_PATH"] = "torch"
    ref = timeit(lambda: gp.fit(), 5)
    del os.environ["MGB_FIT_PATH"]
    print("posterior fit n=%d: own Cholesky + L^-1 %.2f ms, torch.linalg (cuSOLVER / cuBLAS) %.2f ms" % (n, own, ref), flush=True)
