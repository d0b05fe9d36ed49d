"""FP64 tensor-core path against the opt-in INT8 path of the posterior at BASELINE size (n = 5000 training points):
candidates per second of ExactGPPosterior.posterior and the difference of the variances.
usage (one GPU): python tools/probe_int8_posterior.py [candidates]"""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch

torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere
from materngabo_b200.posterior import ExactGPPosterior

dev = "cuda:0"
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n = 5000
gen = torch.Generator().manual_seed(0)
xt = torch.nn.functional.normalize(torch.randn(n, 11, generator=gen), dim=-1).to(dev)
xc = torch.nn.functional.normalize(torch.randn(m, 11, generator=gen), dim=-1).to(dev)
xc[:1000] = torch.nn.functional.normalize(xt[:1000] + 1e-3 * torch.randn(1000, 11, device=dev), dim=-1)
y = torch.sin(3 * xt[:, 0])
k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
k.lengthscale = 0.5
k = k.to(dev)
out = {}
for name, flag in (("fp64", False), ("int8", True)):
    gp = ExactGPPosterior(k, xt, y, outputscale=1.0, noise=1e-2, int8=flag).fit()
    gp.posterior(xc[:65536])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mean, var = gp.posterior(xc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[name] = (mean, var)
    print("%s: %d candidates x %d training points: %.1f ms, %.4g candidates/s" % (name, m, n, ms, m / (ms * 1e-3)))
v0, v1 = out["fp64"][1], out["int8"][1]
print("max |var_int8 - var_fp64| / max var = %.3e ; mean: %.3e ; min var %.3e" % (
    float((v1 - v0).abs().max() / v0.abs().max()),
    float((out["int8"][0] - out["fp64"][0]).abs().max() / out["fp64"][0].abs().max()), float(v0.min())))
