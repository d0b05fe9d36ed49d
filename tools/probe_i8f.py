"""Bring-up probe for csrc/int8_fused.cu (hand-written tcgen05 INT8 posterior kernel): raw accumulator planes against an
integer matmul of the same slices on the host, variances against float64, then timing at the headline size.
Run on a GPU box: python tools/probe_i8f.py [small|time|all]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

from materngabo_b200 import _lib

F64 = torch.float64
S, BITS = 7, 8


def slices_np(x):
    """Rows of x -> (exponent e, S int slices) exactly as f8_slice_kernel does (8-bit digits, carry pass)."""
    mx = np.abs(x).max(axis=1)
    e = np.zeros(x.shape[0], dtype=np.int64)
    f, ex = np.frexp(mx[mx > 0])
    e[mx > 0] = ex + np.where(f > 0.984375, 2, 1)
    r = np.ldexp(x, -e[:, None])
    out = []
    for _ in range(S):
        t = r * (1 << BITS)
        q = np.rint(t)
        out.append(q.astype(np.int64))
        r = t - q
    for p in range(S - 1, 0, -1):
        c = out[p] == 128
        out[p][c] = -128
        out[p - 1][c] += 1
    return e, out


def run(m, n, debug, seed=0, reps=1):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    ks = torch.rand(m, n, generator=g, dtype=F64) * 2 - 1
    ks = (ks * torch.rand(m, 1, generator=g, dtype=F64)).to(dev)
    rinv = torch.tril(torch.randn(n, n, generator=g, dtype=F64)).to(dev) * 3.0
    alpha = torch.randn(n, generator=g, dtype=F64).to(dev)
    pk = torch.empty(lib.mgb_posterior_i8f_pack_bytes(n), dtype=torch.uint8, device=dev)
    ws = torch.empty(lib.mgb_posterior_i8f_workspace_bytes(m, n), dtype=torch.uint8, device=dev)
    mean = torch.empty(m, dtype=F64, device=dev)
    var = torch.empty(m, dtype=F64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.mgb_posterior_i8f_pack(rinv.data_ptr(), n, n, pk.data_ptr(), st), 'pack')
    torch.cuda.synchronize()
    kss, mc = 2.5, 0.3
    planes = None
    if debug:
        planes = torch.zeros(lib.mgb_posterior_i8f_planes_bytes(m, n) // 4, dtype=torch.int32, device=dev)
        _lib.check(lib.mgb_posterior_i8f_reduce_debug(ks.data_ptr(), m, n, n, pk.data_ptr(), alpha.data_ptr(), kss, mc,
                                                      mean.data_ptr(), var.data_ptr(), ws.data_ptr(), ws.numel(), planes.data_ptr(), st), 'reduce_debug')
    else:
        _lib.check(lib.mgb_posterior_i8f_reduce(ks.data_ptr(), m, n, n, pk.data_ptr(), alpha.data_ptr(), kss, mc,
                                                mean.data_ptr(), var.data_ptr(), ws.data_ptr(), ws.numel(), st), 'reduce')
    torch.cuda.synchronize()
    v = ks @ rinv.T
    ref_var = kss - (v * v).sum(-1)
    ref_mean = mc + ks @ alpha
    scale = float(ref_var.abs().max())
    print("m %d n %d: var err / max|var| %.3e   mean err %.3e" % (m, n, float((var - ref_var).abs().max()) / scale,
                                                                 float((mean - ref_mean).abs().max())), flush=True)
    if debug:
        n_rb, CT = (m + 127) // 128, (n + 63) // 64
        pl = planes.view(n_rb, CT, S, 128, 64).cpu().numpy().astype(np.int64)
        ea, A = slices_np(ks.cpu().numpy())
        eb, B = slices_np(rinv.cpu().numpy())
        worst = 0
        for d in range(S):
            C = sum(A[p] @ B[d - p].T for p in range(d + 1))          # m x n
            Cp = np.zeros((n_rb * 128, CT * 64), dtype=np.int64)
            Cp[:m, :n] = C
            got = pl[:, :, d].transpose(0, 2, 1, 3).reshape(n_rb * 128, CT * 64)
            diff = np.abs(got - Cp)
            worst = max(worst, int(diff[:m, :n].max()))
            if diff[:m, :n].max() != 0:
                bad = np.argwhere(diff[:m, :n] != 0)
                print("  plane %d: %d wrong entries, first %s got %d want %d" % (d, len(bad), bad[0], got[tuple(bad[0])], Cp[tuple(bad[0])]))
        print("  planes: max |difference| = %d" % worst, flush=True)
    if reps > 1:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            lib.mgb_posterior_i8f_reduce(ks.data_ptr(), m, n, n, pk.data_ptr(), alpha.data_ptr(), kss, mc, mean.data_ptr(),
                                         var.data_ptr(), ws.data_ptr(), ws.numel(), st)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        print("  %.3f ms per block -> %.3e candidates/s (reduce only, K* resident)" % (ms, m / ms * 1e3), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("small", "all"):
        run(128, 64, True)
        run(128, 128, True)
        run(200, 300, True)
        run(700, 1000, True)
    if what in ("time", "all"):
        run(4096, 5000, False, reps=1)
        run(37888, 5000, False, reps=5)
    if what == "one":
        run(37888, 5000, False, reps=3)
