"""Ordered device launches of one MLL step (N = 100) on a chosen manifold: python tools/probe_launch_order.py sphere|torus|spd2|h2"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_hyperbolic as kh, kernels_spd as ksp, kernels_sphere as ks, kernels_torus as kt
from materngabo_b200.gpytorch_compat import ScaleKernel
from materngabo_b200.graphs import exact_mll
from torch.profiler import ProfilerActivity, profile

which = sys.argv[1] if len(sys.argv) > 1 else "torus"
gen = torch.Generator().manual_seed(0)
n = 100
if which == "sphere":
    base, x = ks.SphereRiemannianMaternKernel(dim=3, nu=2.5), torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)
elif which == "torus":
    base, x = kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5), torch.rand(n, 3, generator=gen) * 6.28
elif which == "spd2":
    a = torch.randn(n, 2, 2, generator=gen)
    s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
    base, x = ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5), torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)
else:
    z = 0.7 * torch.randn(n, 2, generator=gen)
    base, x = kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5), torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)
x = x.to("cuda:0")
y = torch.sin(3 * x[:, 0])
sk = ScaleKernel(base).to("cuda:0")
params = [p for p in sk.parameters() if p.requires_grad]


def step():
    loss, info = exact_mll(sk, x, y, 1e-2, 0.0)
    return torch.autograd.grad(loss, params, allow_unused=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as p:
    step()
    torch.cuda.synchronize()
evs = sorted((e for e in p.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
for e in evs:
    print("%7.1f us  %s" % (e.device_time, e.name[:110]))
print("total %.1f us in %d launches" % (sum(e.device_time for e in evs), len(evs)))
