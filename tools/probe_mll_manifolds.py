"""MLL step (value + gradient to the raw hyper-parameters) at N = 100 for one kernel of every manifold: eager wall time
per call and the kernel launches behind it (torch profiler).  usage (one GPU): python tools/probe_mll_manifolds.py"""
import collections
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch

torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_hyperbolic as kh, kernels_sphere as ks, kernels_spd as ksp, kernels_torus as kt
from materngabo_b200.gpytorch_compat import ScaleKernel
from materngabo_b200.graphs import exact_mll
from torch.profiler import ProfilerActivity, profile

dev = "cuda:0"
gen = torch.Generator().manual_seed(0)
n = 100


def hyp(dim):
    z = 0.7 * torch.randn(n, dim, generator=gen)
    return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)


def spd():
    a = torch.randn(n, 2, 2, generator=gen)
    s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
    return torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1)


cases = [
    ("S^3 Matern", ks.SphereRiemannianMaternKernel(dim=3, nu=2.5), torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)),
    ("T^3 Matern", kt.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5), torch.rand(n, 3, generator=gen) * 6.28),
    ("H^2 Matern", kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5), hyp(2)),
    ("H^3 Matern", kh.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=3, nu=2.5), hyp(3)),
    ("SPD(2) Matern", ksp.SpdRiemannianIntegratedMaternKernel(2, nu=2.5), spd()),
]
for name, base, x in cases:
    x = x.to(dev)
    y = torch.sin(3 * x[:, 0])
    sk = ScaleKernel(base).to(dev)
    params = [p for p in sk.parameters() if p.requires_grad]

    def step():
        loss, info = exact_mll(sk, x, y, 1e-2, 0.0)
        return torch.autograd.grad(loss, params, allow_unused=True)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    us = (time.perf_counter() - t0) / 20 * 1e6
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
        step()
        torch.cuda.synchronize()
    evs = [e for e in p.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = sum(getattr(e, "device_time", 0) for e in evs)
    graph_us = float("nan")
    try:
        from materngabo_b200.graphs import GraphedValueAndGrad
        gm = GraphedValueAndGrad(lambda: exact_mll(sk, x, y, 1e-2, 0.0), inputs=(), wrt=params)

        def gstep():
            value, grads = gm()
            return float(value)              # the L-BFGS driver reads the loss back every evaluation

        for _ in range(5):
            gstep()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            gstep()
        torch.cuda.synchronize()
        graph_us = (time.perf_counter() - t0) / 50 * 1e6
    except Exception as exc:                 # not capturable (host synchronisation inside the step)
        print("      graph capture failed: %s: %s" % (type(exc).__name__, str(exc)[:120]))
    print("%-14s %8.1f us per MLL step eager, %8.1f us under CUDA-graph replay; %3d device launches, %7.1f us of device time"
          % (name, us, graph_us, len(evs), tot))
    c = collections.Counter()
    for e in evs:
        c[e.name[:70]] += getattr(e, "device_time", 0)
    for k, v in c.most_common(8):
        print("      %7.1f us  %s" % (v, k))
