"""Bring-up probe of the INT8 tensor-core series Gram kernel (csrc/series_i8.cu): accumulator planes against an exact
integer matmul of the same digits, K against the float64 kernel and a long-double evaluation, then timings.

    python tools/probe_series_i8.py [--time] [--seconds 2.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from materngabo_b200 import _lib, _ops  # noqa: E402

F64 = torch.float64


def digits_np(x):
    """numpy restatement of g8_exponent / g8_digits: rows -> (e, digits [7][rows][W] int64)."""
    x = np.asarray(x, dtype=np.float64)
    mx = np.abs(x).max(axis=1)
    f, e = np.frexp(mx)
    e = np.where(mx > 0, e + np.where(f > 0.984375, 2, 1), 0)
    r = x * np.ldexp(1.0, -e)[:, None]
    q = []
    for _ in range(7):
        t = np.rint(r * 256.0)          # round half to even, as the magic-number add does
        q.append(t.astype(np.int64))
        r = r * 256.0 - t
    for p in range(6, 0, -1):
        c = (q[p] == 128).astype(np.int64)
        q[p] = q[p] - 256 * c
        q[p - 1] = q[p - 1] + c
    return e, np.stack(q)


def planes_np(x1, x2):
    ea, da = digits_np(x1)
    eb, db = digits_np(x2)
    pl = np.zeros((7, x1.shape[0], x2.shape[0]), dtype=np.int64)
    for p in range(7):
        for q in range(7 - p):
            pl[p + q] += da[p] @ db[q].T
    return ea, eb, pl


def check(name, spec, x1, x2, coef, report):
    dev = torch.device("cuda")
    a, b, c = x1.to(dev), x2.to(dev), coef.to(dev)
    k8, pl = _ops.series_gram_i8(spec, a, b, c, planes=True)
    torch.cuda.synchronize()
    m, n = x1.shape[0], x2.shape[0]
    ea, eb, ref = planes_np(x1.numpy(), x2.numpy())
    got = pl.cpu().numpy()                               # [panel][tile][7][128][32]
    got = got.transpose(2, 0, 3, 1, 4).reshape(7, got.shape[0] * 128, got.shape[1] * 32)[:, :m, :n]
    plane_bad = int((got != ref).sum())
    os.environ["MGB_GRAM_PATH"] = "fp64"
    k64 = _ops._series_fwd_raw(spec, a, b, c)
    os.environ["MGB_GRAM_PATH"] = "fp64"
    k8n = _ops.series_gram_i8(spec, a, b, c)            # non-debug instantiation
    # long-double evaluation of the same formula
    u = (x1.numpy().astype(np.longdouble) @ x2.numpy().astype(np.longdouble).T)
    u = spec.scale * (u * u if spec.pair_square else u) + spec.shift
    u = np.clip(u, np.longdouble(spec.lo), np.longdouble(spec.hi))
    kk = np.zeros_like(u)
    for ck in coef.numpy()[::-1]:
        kk = kk * u + np.longdouble(ck)
    scale = float(np.abs(kk).max())
    report[name] = {"m": m, "n": n, "plane_mismatches": plane_bad,
                    "i8_vs_truth": float(np.abs(k8.cpu().numpy() - kk).max() / scale),
                    "i8_nodebug_vs_truth": float(np.abs(k8n.cpu().numpy() - kk).max() / scale),
                    "fp64_vs_truth": float(np.abs(k64.cpu().numpy() - kk).max() / scale),
                    "i8_vs_fp64": float((k8 - k64).abs().max().item() / scale)}


def bench(name, spec, w, m, n, coef, seconds, report, quat=False):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(1)
    x1 = torch.nn.functional.normalize(torch.randn(m, w, generator=g, dtype=F64), dim=-1).to(dev)
    x2 = torch.nn.functional.normalize(torch.randn(n, w, generator=g, dtype=F64), dim=-1).to(dev)
    c = coef.to(dev)
    res = {}
    for path in ("fp64", "int8"):
        os.environ["MGB_GRAM_PATH"] = path
        out = None
        for _ in range(3):
            del out
            out = _ops._series_fwd_raw(spec, x1, x2, c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            del out
            out = _ops._series_fwd_raw(spec, x1, x2, c)
        e1.record()
        torch.cuda.synchronize()
        burst = e0.elapsed_time(e1) / 5
        steps = max(5, int(seconds * 1e3 / burst))
        e0.record()
        for _ in range(steps):
            del out
            out = _ops._series_fwd_raw(spec, x1, x2, c)
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / steps
        res[path] = {"burst_ms": burst, "sustained_ms": sus, "burst_entries_per_s": m * n / burst * 1e3,
                     "sustained_entries_per_s": m * n / sus * 1e3, "sustained_frac_hbm_6542": 8.0 * m * n / sus * 1e3 / 6542.4e9,
                     "checksum": float(out[:4, :4].sum())}
        del out
    os.environ["MGB_GRAM_PATH"] = "fp64"
    report[name] = res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--out", default="gpurun_out/probe_series_i8.json")
    args = ap.parse_args()
    _lib.load()
    report = {}
    g = torch.Generator().manual_seed(0)
    coef10 = torch.tensor([0.11, 0.2, 0.15, 0.1, 0.09, 0.08, 0.07, 0.06, 0.05, 0.04], dtype=F64)
    cases = [
        ("s10_ragged", _ops.SeriesSpec(1, 11, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15), 11, 300, 100),
        ("s2", _ops.SeriesSpec(1, 3, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15), 3, 256, 64),
        ("so3_trace", _ops.SeriesSpec(1, 9, 0, 0.5, -0.5, -1.0 + 1e-8, 1.0 - 1e-8), 9, 130, 40),
        ("quat_square", _ops.SeriesSpec(1, 4, 0, 2.0, -1.0, -1.0 + 1e-8, 1.0 - 1e-8, 1), 4, 257, 33),
        ("w16_t12", _ops.SeriesSpec(1, 16, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15), 16, 128, 32),
    ]
    for name, spec, w, m, n in cases:
        x1 = torch.nn.functional.normalize(torch.randn(m, w, generator=g, dtype=F64), dim=-1)
        x2 = torch.nn.functional.normalize(torch.randn(n, w, generator=g, dtype=F64), dim=-1)
        if name == "so3_trace":
            x1, x2 = x1 * np.sqrt(3.0), x2 * np.sqrt(3.0)          # rotation-matrix scale (Frobenius norm sqrt 3)
        x2[: min(5, n)] = x1[: min(5, n)]                           # repeated points: the clamp is active
        coef = coef10 if name != "w16_t12" else torch.cat([coef10, torch.tensor([0.03, 0.02], dtype=F64)])
        try:
            check(name, spec, x1, x2, coef, report)
        except Exception as exc:  # noqa: BLE001
            report[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        print(name, json.dumps(report[name]), flush=True)
    if args.time:
        for name, spec, w, m, n in [
            ("time_s10", cases[0][1], 11, 1000000, 2000),
            ("time_s2", cases[1][1], 3, 1000000, 2000),
            ("time_so3_trace", cases[2][1], 9, 1000000, 2000),
            ("time_quat", cases[3][1], 4, 1000000, 2000),
        ]:
            try:
                bench(name, spec, w, m, n, coef10, args.seconds, report)
            except Exception as exc:  # noqa: BLE001
                report[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            print(name, json.dumps(report[name]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
