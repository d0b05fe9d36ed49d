"""Round-2 kernels in one process (profile target for ncu -k regex:...): series_gram_fwd_small (configs[0]),
son_gram_kernel<5> / <6> (4096 x 2048), dense_gemm_kernel + chol_coop_kernel (posterior fit at n = 5000), spd_decompose.
Prints CUDA-event timings when run plain."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from materngabo_b200 import _ops, kernels_so, kernels_sphere  # noqa: E402
from materngabo_b200.posterior import ExactGPPosterior  # noqa: E402

dev = "cuda:0"
gen = torch.Generator().manual_seed(0)


REPS = int(os.environ.get("REPS", "0"))


def timed(name, fn, reps=5):
    reps = REPS or reps
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-58s %10.3f ms per call" % (name, e0.elapsed_time(e1) / reps), flush=True)
    return out


def rotations(cnt, n):
    q, r = torch.linalg.qr(torch.randn(cnt, n, n, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    q[:, :, 0] = q[:, :, 0] * torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    return q.reshape(cnt, n * n).to(dev)


with torch.no_grad():
    x = torch.nn.functional.normalize(torch.randn(1000, 3, generator=gen), dim=-1).to(dev)
    k2 = kernels_sphere.SphereRiemannianMaternKernel(dim=2, nu=2.5)
    k2.lengthscale = 0.5
    timed("S^2 Gram 1000 x 1000 (series_gram_fwd_small)", lambda: k2.forward(x, x), 20)
    for n in (4, 5, 6):
        ks = kernels_so.SORiemannianMaternKernel(dim=n, nu=2.5).to(dev)
        ks.lengthscale = 0.5
        a, b = rotations(4096, n), rotations(2048, n)
        out = timed("SO(%d) Gram 4096 x 2048 (son_gram_kernel)" % n, lambda: ks.forward(a, b))
        print("      %.3e entries/s" % (4096 * 2048 / 1.0), end="\r")
    sm = torch.randn(1 << 20, 3, 3, generator=gen)
    sm = sm @ sm.transpose(-1, -2) + 0.3 * torch.eye(3)
    mand = torch.stack([sm[:, 0, 0], sm[:, 1, 1], sm[:, 2, 2], sm[:, 0, 1] * 2 ** 0.5, sm[:, 1, 2] * 2 ** 0.5, sm[:, 0, 2] * 2 ** 0.5], -1).to(dev)
    timed("SPD(3) eigen front-end, 1 M matrices (spd_decompose_kernel)", lambda: _ops.spd_decompose(mand, 3, 0))
    xt = torch.nn.functional.normalize(torch.randn(5000, 11, generator=gen), dim=-1).to(dev)
    k10 = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
    k10.lengthscale = 0.5
    gp = ExactGPPosterior(k10, xt, torch.sin(3 * xt[:, 0]), outputscale=1.0, noise=1e-2)
    timed("posterior fit n = 5000 (chol_coop + dense_gemm + pack)", lambda: gp.fit(), 3)
print("ok")
