"""Launch list source for the multi-CTA dense path: 2 marginal-likelihood steps at n = 1000 and 2 posterior fits at
n = 5000 (run under ncu --metrics gpu__time_duration.sum)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere  # noqa: E402
from materngabo_b200.gpytorch_compat import ScaleKernel  # noqa: E402
from materngabo_b200.graphs import exact_mll  # noqa: E402
from materngabo_b200.posterior import ExactGPPosterior  # noqa: E402

dev = "cuda:0"
gen = torch.Generator().manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(1000, 4, generator=gen), dim=-1).to(dev)
y = torch.sin(3 * x[:, 0])
base = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5)
sk = ScaleKernel(base).to(dev)
for _ in range(2):
    loss, _ = exact_mll(sk, x, y, 1e-2, 0.0)
    torch.autograd.grad(loss, [base.raw_lengthscale, sk.raw_outputscale])
xt = torch.nn.functional.normalize(torch.randn(5000, 11, generator=gen), dim=-1).to(dev)
k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
k.lengthscale = 0.5
gp = ExactGPPosterior(k, xt, torch.sin(3 * xt[:, 0]), outputscale=1.0, noise=1e-2)
for _ in range(2):
    gp.fit()
torch.cuda.synchronize()
print("ok")
