"""Scratch GPU probe used during development (first contact with the hardware): FP64 peaks + raw kernel timings."""
import ctypes as C
import json
import time

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

torch.set_default_dtype(torch.float64)
from materngabo_b200 import _lib, _ops, kernels_so, kernels_sphere, kernels_torus  # noqa: E402
from materngabo_b200.posterior import ExactGPPosterior  # noqa: E402

lib = _lib.load()
out = {}
for kind, name in ((0, "dfma"), (1, "dmma")):
    tf, ms = C.c_double(), C.c_double()
    _lib.check(lib.mgb_fp64_peak(kind, 5, C.byref(tf), C.byref(ms)), "peak")
    out[name + "_tflops"] = tf.value
    out[name + "_ms"] = ms.value
print(json.dumps(out))
dev = "cuda:0"


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def unit(m, d):
    return torch.nn.functional.normalize(torch.randn(m, d, device=dev), dim=-1)


for dim, m, n in ((2, 200000, 2000), (10, 200000, 2000), (10, 100000, 5000)):
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=dim, nu=2.5)
    k.lengthscale = 0.5
    spec, coef = k.series()
    coef = coef.detach().to(dev)
    x1, x2 = unit(m, dim + 1), unit(n, dim + 1)
    ms = timeit(lambda: _ops._series_fwd_raw(spec, x1, x2, coef))
    print("sphere S%d %dx%d: %.3f ms  %.3e entries/s  %.1f GB/s" % (dim, m, n, ms, m * n / ms * 1e3, 8 * m * n / ms / 1e6))
k = kernels_so.SORiemannianMaternKernel(dim=3, nu=2.5)
k.lengthscale = 0.5
spec, coef = k.series()
coef = coef.detach().to(dev)
x1, x2 = torch.randn(200000, 9, device=dev), torch.randn(2000, 9, device=dev)
ms = timeit(lambda: _ops._series_fwd_raw(spec, x1, x2, coef))
print("SO3 200k x 2k: %.3f ms  %.3e entries/s  %.1f GB/s" % (ms, 4e8 / ms * 1e3, 8 * 4e8 / ms / 1e6))
k = kernels_torus.TorusProductOfManifoldsRiemannianMaternKernel(dim=10, nu=2.5)
for kk in k.torus_kernel.kernels:
    kk.lengthscale = 0.5
spec, coef = k.series()
coef = coef.detach().to(dev)
print("torus terms", coef.shape)
x1, x2 = k._to_circles(torch.rand(100000, 10, device=dev) * 6.28), k._to_circles(torch.rand(2000, 10, device=dev) * 6.28)
ms = timeit(lambda: _ops._series_fwd_raw(spec, x1, x2, coef))
print("T10 100k x 2k: %.3f ms  %.3e entries/s" % (ms, 2e8 / ms * 1e3))

# posterior S^10, n = 5000
n, m = 5000, 37888 * 2
xt, xc = unit(n, 11), unit(m, 11)
y = torch.sin(3 * xt[:, 0])
k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
k.lengthscale = 0.5
t0 = time.time()
gp = ExactGPPosterior(k, xt, y, outputscale=1.0, noise=1e-2).fit()
torch.cuda.synchronize()
print("fit n=5000: %.3f s" % (time.time() - t0))
ms = timeit(lambda: gp.posterior(xc), reps=3)
flops = m * (n * 78 + 2 * n + n * n + 2 * n)
print("posterior S10 %d x %d: %.2f ms  %.3e cands/s  %.2f TFLOP/s (algorithmic, triangular)" % (m, n, ms, m / ms * 1e3, flops / ms / 1e9))
