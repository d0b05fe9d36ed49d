"""Small end-to-end pass over every CUDA entry point, sized for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from materngabo_b200 import (kernels_euclidean, kernels_hyperbolic, kernels_so, kernels_spd, kernels_sphere,  # noqa: E402
                             kernels_torus)
from materngabo_b200.posterior import ExactGPPosterior  # noqa: E402

dev = "cuda:0"
gen = torch.Generator().manual_seed(0)


def unit(m, d):
    return torch.nn.functional.normalize(torch.randn(m, d, generator=gen), dim=-1).to(dev)


for dim, m, n in ((2, 131, 67), (3, 40, 33), (10, 257, 129), (7, 50, 30)):
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=dim, nu=2.5)
    k.lengthscale = 0.5
    k.raw_lengthscale.requires_grad_(True)
    x1, x2 = unit(m, dim + 1).requires_grad_(True), unit(n, dim + 1)
    out = k.forward(x1, x2)
    out.sum().backward()
k = kernels_so.SORiemannianMaternKernel(dim=3, nu=2.5)
q, _ = torch.linalg.qr(torch.randn(70, 3, 3, generator=gen))
k.forward(q.reshape(70, 9).to(dev), q.reshape(70, 9)[:33].to(dev))
kt = kernels_torus.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5)
a = (torch.rand(45, 3, generator=gen) * 6.28).to(dev).requires_grad_(True)
kt.forward(a, a.detach()[:20]).sum().backward()
z = torch.randn(40, 2, generator=gen)
h = torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1).to(dev)
kh = kernels_hyperbolic.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=64)
kh.raw_lengthscale.requires_grad_(True)
kh.forward(h, h[:17]).sum().backward()
kg = kernels_hyperbolic.HyperbolicRiemannianMillsonGaussianKernel(dim=2, nb_points_integral=200)
kg.raw_lengthscale.requires_grad_(True)
kg.forward(h, h[:17]).sum().backward()
am = torch.randn(20, 2, 2, generator=gen)
s = am @ am.transpose(-1, -2) + 0.5 * torch.eye(2)
v = torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1).to(dev)
ks = kernels_spd.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=32, nb_points_integral_t=32)
ks.forward(v, v[:9])
kg2 = kernels_spd.SpdRiemannianGaussianKernel(2)
kg2.raw_lengthscale.requires_grad_(True)
kg2.forward(v, v[:9]).sum().backward()
ke = kernels_euclidean.EuclideanIntegratedMaternKernel(3, nu=2.5)
ke.forward(torch.randn(30, 3, generator=gen).to(dev), torch.randn(11, 3, generator=gen).to(dev))
for n, m in ((100, 300), (257, 1000), (130, 129)):
    xt, xc = unit(n, 4), unit(m, 4)
    kk = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    gp = ExactGPPosterior(kk, xt, torch.sin(3 * xt[:, 0]), noise=1e-2).fit()
    mean, var = gp.posterior(xc)
torch.cuda.synchronize()
print("sanitize smoke ok", float(mean.sum()), float(var.sum()))
