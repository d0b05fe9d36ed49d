"""One dense-mode MLL step at N = 100 (H^3 kernel) for an ncu capture of series_mll_kernel: python tools/probe_mll_kernel.py [sphere]"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_hyperbolic as kh, kernels_sphere as ks
from materngabo_b200.gpytorch_compat import ScaleKernel
from materngabo_b200.graphs import exact_mll

gen = torch.Generator().manual_seed(0)
n = 100
if len(sys.argv) > 1 and sys.argv[1] == "sphere":      # series mode of the single-CTA kernel
    x = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1).to("cuda:0")
    base = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5)
else:                                                  # dense mode
    z = 0.7 * torch.randn(n, 3, generator=gen)
    x = torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1).to("cuda:0")
    base = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=3, nu=2.5)
y = torch.sin(3 * x[:, 0])
sk = ScaleKernel(base).to("cuda:0")
params = [p for p in sk.parameters() if p.requires_grad]
for _ in range(3):
    loss, info = exact_mll(sk, x, y, 1e-2, 0.0)
    torch.autograd.grad(loss, params, allow_unused=True)
torch.cuda.synchronize()
print("ok", float(loss))
