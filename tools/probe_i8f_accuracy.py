"""Fused INT8 posterior path against the FP64 tensor-core kernel at the headline training size (S^10, n = 5000), incl.
candidates near / on training points: python tools/probe_i8f_accuracy.py   (one GPU)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere
from materngabo_b200.posterior import ExactGPPosterior

gen = torch.Generator().manual_seed(6)
n, m = 5000, 8192
xt = torch.nn.functional.normalize(torch.randn(n, 11, generator=gen), dim=-1)
xc = torch.nn.functional.normalize(torch.randn(m, 11, generator=gen), dim=-1)
xc[:512] = torch.nn.functional.normalize(xt[:512] + 1e-3 * torch.randn(512, 11, generator=gen), dim=-1)
xc[512:640] = xt[512:640]
y = torch.sin(3.0 * xt[:, 0]) + 0.5 * xt[:, 1] * xt[:, 2]
k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
k.lengthscale = 0.5
for noise in (1e-2, 1e-4, 1e-6):
    a = ExactGPPosterior(k, xt.cuda(), y.cuda(), outputscale=1.0, noise=noise).fit()
    m0, v0 = a.posterior(xc.cuda())
    for impl in ("fused", "library"):
        b = ExactGPPosterior(k, xt.cuda(), y.cuda(), outputscale=1.0, noise=noise, int8=impl).fit()
        try:
            m1, v1 = b.posterior(xc.cuda())
        except Exception as exc:
            print("noise %.0e %s: %s" % (noise, impl, exc))
            continue
        print("noise %.0e: int8 %-7s vs fp64 kernel: var %.2e of max var, mean %.2e of max mean"
              % (noise, impl, float((v1 - v0).abs().max() / v0.abs().max()), float((m1 - m0).abs().max() / m0.abs().max())), flush=True)
