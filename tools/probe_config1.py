"""Where do the ~9 us of BASELINE configs[0] (S^2 Matern Gram 1000 x 1000, L2 flushed, CUDA-graph replay between two
events) go?  Floor = a one-element kernel under the same protocol; then the Gram kernel with warm / cold L2, and an 8 MB
fill (pure write of the same size)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(1000, 3, generator=gen, device=dev)
x = x / x.norm(dim=-1, keepdim=True)
k = kernels_sphere.SphereRiemannianMaternKernel(dim=2, nu=2.5)
k.lengthscale = 0.5
k = k.to(dev)
tiny = torch.zeros(32, device=dev)
out8 = torch.empty(1000, 1000, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def graphed(fn):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def measure(name, fn, cold, reps=300):
    g = graphed(fn)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        if cold:
            flush.fill_(1.0)
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    print("%-46s %s L2: median %.2f us  min %.2f us" % (name, "cold" if cold else "warm", ts[len(ts) // 2], ts[0]), flush=True)


with torch.no_grad():
    for cold in (False, True):
        measure("one-element kernel (protocol floor)", lambda: tiny.add_(1.0), cold)
        measure("8 MB fill (pure write, torch)", lambda: out8.fill_(0.5), cold)
        measure("S^2 Gram 1000 x 1000", lambda: k.forward(x, x), cold)
        measure("S^2 Gram 2000 x 2000 (small-kernel limit)", lambda: k.forward(torch.cat([x, x]), torch.cat([x, x])), cold)
