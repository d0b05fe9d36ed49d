"""One marginal-likelihood evaluation at n = 1000 through the multi-CTA dense path (profile target: chol_coop_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere  # noqa: E402
from materngabo_b200.gpytorch_compat import ScaleKernel  # noqa: E402
from materngabo_b200.graphs import exact_mll  # noqa: E402

dev = "cuda:0"
n = int(os.environ.get("N", 1000))
gen = torch.Generator().manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1).to(dev)
y = torch.sin(3 * x[:, 0])
base = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5)
sk = ScaleKernel(base).to(dev)
for _ in range(3):
    loss, _ = exact_mll(sk, x, y, 1e-2, 0.0)
torch.cuda.synchronize()
print("ok", float(loss.detach()))
