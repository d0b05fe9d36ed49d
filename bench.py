#!/usr/bin/env python
"""bench.py - headline benchmark of the hot path (see BASELINE.json / SURVEY.md section 8d).

Default workload (BASELINE.json configs[4]): S^10 Riemannian Matern (nu=2.5, kappa=0.5, 10 terms) exact-GP
posterior mean/variance for 10M candidates x 5k training points, candidate rows sharded over the N GPUs of
one box (strong scaling, one process per GPU, NCCL only for the state broadcast and the mean/var all-gather).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (CUDA, sm_100a)
  python bench.py --impl reference ...                          the reference's CPU algorithm (oracle port)
  python bench.py --workload gram-so3|gram-so5|gram-s2|gram-s10|gram-t10|gram-h2|gram-spd2|gram-spd3   Gram-matrix workloads
  python bench.py --workload gram-h2-table   opt-in tabulated H^2 profile (not the per-entry quadrature)
  python bench.py --workload config1-s2      BASELINE configs[0]: S^2 Matern Gram 1000 x 1000 (L2 flushed between steps)
  python bench.py --workload config2-bo-s3   BASELINE configs[1]: S^3 GaBO call latencies (N = 100)
  python bench.py --workload tr-step         trust-region inner step on every manifold: fused kernels vs autograd

Rank 0 prints ONE JSON line.  ``value`` = candidates/s with inputs resident in HBM; ``e2e`` = the same through
the host-buffer C-ABI call (pinned host inputs, H2D + D2H inside the timed region).  The posterior runs on the
path the library picks by itself (--posterior-path auto: at the default size the INT8 tensor-core contraction,
``dtype`` "f64+i8"); the line then also carries ``fp64_path`` - the whole workload on the FP64 DMMA kernel and
the largest difference between the two result sets over every candidate - and ``roofline.fp64_kernel``.
--posterior-path fp64 | int8 force one path.  The oracle is used only by the cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F64 = torch.float64
NOMINAL_FP64_TFLOPS = 37.2   # 148 SM x 64 FMA/clk x 2 x 1.965 GHz


# ---------------------------------------------------------------------------------------------------
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="posterior-s10")
    ap.add_argument("--cands", dest="m", type=int, default=None, help="override candidate / row count")
    ap.add_argument("--train", dest="n", type=int, default=None, help="override training / column count")
    ap.add_argument("--posterior-path", default="auto", choices=["auto", "fp64", "int8"],
                    help="default workload: auto = what ExactGPPosterior picks by itself (MGB_POSTERIOR_PATH=auto: the INT8 "
                         "tensor-core contraction from 16384 candidates x 1024..16384 training points on, else the FP64 DMMA "
                         "kernel); fp64 / int8 force one of them")
    ap.add_argument("--int8-impl", default="fused", choices=["fused", "library"],
                    help="--posterior-path int8: fused = this repo's tcgen05 kernel (csrc/int8_fused.cu); library = CUTLASS "
                         "int8 GEMMs + recombination pass (csrc/int8_posterior.cu)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gram", action="store_true", help="default workload: skip the Gram-matrix leg (configs[0], [2], [3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # samples under load only (idle samples before/after the region would drag the median down)
        loaded = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] if power else sm
        return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def gpu_temperature(index):
    """Die temperature in degrees Celsius from nvidia-smi, or None."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=temperature.gpu", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
        return float(out[0])
    except Exception:
        return None


def cool_down(index, start_c, margin_c=4.0, max_wait_s=30.0):
    """Idle until the die is back within margin_c of the temperature it had before the first kernel of this process (or
    max_wait_s): the Gram workloads run at the board's power cap, where a die heated by the preceding posterior steps
    clocks ~2 % lower for the same 1 kW.  Returns what happened, for the JSON line."""
    t0 = time.time()
    before = gpu_temperature(index)
    now = before
    while start_c is not None and now is not None and now > start_c + margin_c and time.time() - t0 < max_wait_s:
        time.sleep(1.0)
        now = gpu_temperature(index)
    return {"start_c": start_c, "before_c": before, "after_c": now, "waited_s": round(time.time() - t0, 1),
            "rule": "idle until the die is within %.0f C of its temperature at process start, at most %.0f s" % (margin_c, max_wait_s)}


def unit_rows(m, d, gen, device):
    x = torch.randn(m, d, generator=gen, device=device, dtype=F64)
    return x / x.norm(dim=-1, keepdim=True)


# ---------------------------------------------------------------------------------------------------
# posterior workload
# ---------------------------------------------------------------------------------------------------
def posterior_problem(n, device, seed=0):
    """Training set + fit for S^10 (SURVEY.md 8d cfg 5): kappa=0.5, nu=2.5, T=10, noise 1e-2, outputscale 1."""
    from materngabo_b200 import kernels_sphere
    from materngabo_b200.posterior import ExactGPPosterior

    gen = torch.Generator(device=device).manual_seed(seed)
    xt = unit_rows(n, 11, gen, device)
    y = torch.sin(3.0 * xt[:, 0]) + 0.5 * xt[:, 1] * xt[:, 2]
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
    k.lengthscale = 0.5
    gp = ExactGPPosterior(k, xt, y, outputscale=1.0, noise=1e-2, mean_const=0.0)
    return gp, k


# posterior_reduce_kernel at the default chunk (37888 candidates x 5000 training points): 12.251238 GB read +
# 10.569472 MB written per launch (ncu --set full, profiles/r01b_ncu_posterior_reduce.txt).  The kernel is bound by
# the FP64 tensor pipe (93 % active); this DRAM traffic is 0.44 TB/s = 7 % of the HBM peak.
NCU_REDUCE_TRAFFIC_BYTES = 12.251238e9 + 10.569472e6


def flops_per_candidate(n):
    """N*78 (kernel row) + 2N (mean) + N^2 (triangular L^-1 k*, N^2/2 FMA) + 2N (norm)."""
    return n * 78.0 + 2.0 * n + float(n) * n + 2.0 * n


def run_posterior(args, rank, world, local):
    import torch.distributed as dist
    from materngabo_b200 import _lib
    from materngabo_b200 import posterior as posterior_mod
    from materngabo_b200.posterior import ExactGPPosterior, ShardedPosterior, shard_bounds

    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    temp_start = gpu_temperature(local) if rank == 0 else None
    lib = _lib.load()
    m_total = args.m or 10_000_000
    n = args.n or 5000
    lo, hi = shard_bounds(m_total, world, rank)
    m_local = hi - lo

    # rank 0 owns the fit; the other ranks receive the state every step (as after a refit in the BO loop)
    gp, kernel = posterior_problem(n, device)
    gp.int8_impl = args.int8_impl
    if args.posterior_path == "auto":
        gp.int8 = "auto"
        # the rule of ExactGPPosterior._int8_for on the per-rank share, evaluated identically on every rank
        use_int8 = bool(gp.int8) and (m_total // world) >= posterior_mod.INT8_AUTO_MIN_M and args.int8_impl == "fused"
        gp.int8 = use_int8
    else:
        use_int8 = args.posterior_path == "int8"
        gp.int8 = use_int8
    if use_int8 and args.int8_impl == "library" and not lib.mgb_int8_available():
        raise SystemExit("--posterior-path int8 --int8-impl library: libmgb was built without the CUTLASS headers")
    if rank == 0:
        gp.fit()
    sh = ShardedPosterior(compute=None) if world > 1 else None
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    xc = unit_rows(m_local, 11, gen, device)                      # this rank's candidate rows, HBM resident
    spec, coef0 = kernel.series()
    # shapes of the fitted state, known on every rank from n and the kernel: the broadcast is one flat NCCL transfer
    meta_fp64 = {"train": ((n, 11), F64), "coef": (tuple(coef0.shape), F64),
                 "packed": ((lib.mgb_posterior_pack_bytes(n) // 8,), F64)}
    meta_int8 = dict(meta_fp64)
    if use_int8:
        pre = "mgb_posterior_i8f_" if args.int8_impl == "fused" else "mgb_posterior_int8_"
        meta_int8["packed_i8"] = ((int(getattr(lib, pre + "pack_bytes")(n)),), torch.uint8)
        meta_int8["alpha"] = ((n,), F64)

    def step():
        if world > 1:
            # gp.int8 is set identically on every rank (headline path, then the FP64 pass): the state that travels follows it
            state = sh.broadcast_state(gp.state() if rank == 0 else {"like": xc[:0]}, src=0,
                                       meta=meta_int8 if gp.int8 else meta_fp64)
            if rank != 0:
                gp.load_state(spec, state["train"], state["packed"], state["coef"],
                              _self_k(spec, state["coef"]), state.get("packed_i8"), state.get("alpha"))
        mean, var = gp.posterior(xc)
        if world > 1:
            sizes = [shard_bounds(m_total, world, r)[1] - shard_bounds(m_total, world, r)[0] for r in range(world)]
            sh.compute = lambda _x: (mean, var)
            mean, var = sh.evaluate(xc, gather=True, sizes=sizes)
        return mean, var

    def _self_k(spec_, coef_):
        from materngabo_b200.posterior import _self_kernel_value
        return _self_kernel_value(spec_, coef_)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    launches0 = lib.mgb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mean, var = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.mgb_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, float(launches)], dtype=F64, device=device)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(tsum[1])
    ms_per_step = ms / args.steps
    value = m_total / (ms_per_step * 1e-3)
    checksum = float(mean[: min(1000, mean.numel())].sum() + var[: min(1000, var.numel())].sum())

    # ---- INT8 headline: the SAME workload on the FP64 DMMA kernel (all ranks), a bounded number of steps, and the
    # difference between the two paths' results over every candidate of the timed step ---------------------------
    fp64_path = None
    if use_int8:
        mean_i8, var_i8 = mean.clone(), var.clone()
        gp.int8 = False
        fsteps = max(1, min(args.steps, 3))
        step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(fsteps):
            mean_f, var_f = step()
        f1.record()
        barrier()
        tf_ = torch.tensor([f0.elapsed_time(f1)], dtype=F64, device=device)
        if world > 1:
            dist.all_reduce(tf_, op=dist.ReduceOp.MAX)
        fms = float(tf_[0]) / fsteps
        fp64_path = {"value": m_total / (fms * 1e-3), "unit": "candidates/s", "steps": fsteps, "ms_per_step": fms,
                     "ratio_headline_to_fp64_path": value / (m_total / (fms * 1e-3)),
                     "max_var_diff_rel_to_max_var": float((var_i8 - var_f).abs().max() / var_f.abs().max()),
                     "max_mean_diff_rel_to_max_mean": float((mean_i8 - mean_f).abs().max() / mean_f.abs().max()),
                     "compared_candidates": int(mean_f.numel()),
                     "note": "the whole workload of the headline (same candidates, same sharding, same step()) on "
                             "posterior_reduce_kernel (FP64 DMMA m8n8k4), timed right after it; the differences are taken over "
                             "every candidate of the last timed step of each path (parity bar of the tests: 1e-10)"}
        del mean_i8, var_i8, mean_f, var_f
        gp.int8 = True

    # ---- per-kernel roofline pass (rank 0): events around every posterior_reduce launch -----------
    roof = None
    if rank == 0:
        roof = posterior_roofline(gp, xc, n, lib)
        if use_int8:
            fused = args.int8_impl == "fused"
            # executed k extent per column: the exact triangle at 128 x 64 granularity (fused) or five column blocks
            tri = sum((ct // 2 + 1) * 128 * 64 for ct in range((n + 63) // 64)) / float(n * n) if fused else 0.6
            roof = {"bound": "tensor",
                    "kernel": ("f8_posterior_kernel (this repo: tcgen05.mma kind::i8, planes in tensor memory, fused float64 "
                               "recombination), 1 launch per block of 37888 candidates") if fused else
                              "CUTLASS int8 GEMM (tcgen05 kind::i8), 35 launches per block of 32768 candidates",
                    "achieved": value / world * 28 * tri * 2.0 * n * n / 1e12, "unit": "TOP/s (int8, executed slice products over the "
                    "whole step time of one GPU: slicing, recombination and the Gram kernel included)",
                    "peak": 2.0 * float(measured_peaks().get("bf16_tflops_sustained") or 1371.4),
                    "peak_source": "2 x the sustained bf16 rate of MEASURED_PEAKS.json (int8 dense = 2 x bf16 dense)",
                    "traffic": (6.524901e9 + 28.624128e6) if (fused and n == 5000) else None,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one f8_posterior_kernel launch (37888 candidates "
                                      "x 5000 training points; ncu --set full, profiles/r02_ncu_f8_posterior.txt); algorithmic "
                                      "operand bytes per launch: A slices 1.36e9 + triangular B slices 0.9e8",
                    "ncu": "tensor pipe active 79 %, 81 % of the int8 tensor peak at the realised clock (1 kW power cap)",
                    "fp64_kernel": roof}
            roof["frac"] = roof["achieved"] / roof["peak"]
            # the same achieved figure against the other denominators one could choose (MEASURED_PEAKS.json has no int8 entry):
            # 2 x the measured bf16 BURST rate, and the hardware's int8 issue rate (ncu: 16384 int8 ops / cycle / SM,
            # sm__ops_path_tensor_op_utcimma_src_int8 peak_sustained) at the SM clock realised under this run's power cap
            try:
                burst = measured_peaks().get("bf16_tflops")
                sms = torch.cuda.get_device_properties(device).multi_processor_count
                mhz = (clocks or {}).get("sm_mhz")
                roof["other_denominators"] = {
                    "2x_measured_bf16_burst": {"peak": 2.0 * burst, "frac": roof["achieved"] / (2.0 * burst)} if burst else None,
                    "hw_int8_rate_at_realised_clock": {"peak": 16384.0 * sms * mhz * 1e6 / 1e12, "sm_mhz": mhz,
                                                       "frac": roof["achieved"] / (16384.0 * sms * mhz * 1e6 / 1e12)} if mhz else None,
                    "note": "achieved counts the executed slice products over the WHOLE step (slicing kernel and launch gaps included); "
                            "f8_posterior_kernel alone: 81 % of the hardware rate at its realised clock (ncu)"}
            except Exception as exc:      # extra information only: never lose the line over it
                roof["other_denominators"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            # north_star asks for the fraction of the FP64 roof: the float64 work of the step (SURVEY.md 8d flops per
            # candidate) over the FP64 tensor peak measured in this run - above 1 because the contraction is emulated
            # exactly on the int8 tensor cores instead of running on the FP64 pipe
            roof["fp64_equivalent"] = {"achieved": value / world * flops_per_candidate(n) / 1e12, "unit": "TFLOP/s (float64-equivalent)",
                                       "peak": roof["fp64_kernel"]["peak"],
                                       "frac": value / world * flops_per_candidate(n) / 1e12 / roof["fp64_kernel"]["peak"]}

    # ---- opt-in INT8 tensor-core path on a bounded sample (rank 0, one GPU), with its parity against `value`'s path ---
    int8 = None
    if rank == 0 and world == 1 and m_local >= 65536 and not use_int8:
        int8 = posterior_int8_sample(gp, xc[: min(m_local, 2 * 1024 * 1024)], value, "fused")
        if lib.mgb_int8_available():
            int8["library_variant"] = posterior_int8_sample(gp, xc[: min(m_local, 1024 * 1024)], value, "library")

    # ---- end to end through the host-buffer C-ABI call -------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = posterior_e2e(args, gp, rank, world, local, m_total, m_local, n, xc, dist)

    # ---- Gram-matrix half of the metric (rank 0, one GPU): configs[0], [2], [3] with sustained timing --------------
    gram, cooled = None, None
    if rank == 0 and world == 1 and not args.no_gram and args.m is None and args.n is None:
        del xc
        gp._ws = None
        gp._ws_i8 = None
        torch.cuda.empty_cache()
        cooled = cool_down(local, temp_start)
        gram = gram_leg(args, rank, world, local)

    out = None
    if rank == 0:
        out = {
            "metric": "posterior_cands_per_sec", "value": value, "unit": "candidates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64+i8" if use_int8 else "f64", "data": "synthetic",
            "dtype_detail": ("inputs, kernel rows and outputs float64; the N^2 contraction as exact products of 8-bit digits "
                             "(56 bits per operand, int32 accumulation, tcgen05 kind::i8) recombined in float64 - "
                             "see fp64_path for the difference to the FP64 kernel on every candidate") if use_int8 else
                            "float64 throughout (FP64 DMMA tensor-core kernel)",
            "posterior_path": "int8" if use_int8 else "fp64", "posterior_path_requested": args.posterior_path,
            "config": {"workload": "S^10 Riemannian Matern nu=2.5 kappa=0.5 T=10 exact-GP posterior mean/var, "
                                   "%d candidates x %d training points, candidate rows sharded over %d GPU(s); "
                                   "BASELINE.json configs[4]" % (m_total, n, world),
                       "candidates": m_total, "train": n, "chunk": gp.chunk,
                       "l2_policy": "inputs larger than L2: every step streams %.1f GB of Gram chunks through HBM "
                                    "(chunk buffer %.2f GB > 126 MB L2)" % (8e-9 * m_total * n / world,
                                                                            8e-9 * gp.chunk * n)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "e2e": e2e,
            "checksum": checksum, "int8_path": int8, "fp64_path": fp64_path, "gram": gram, "gram_cool_down": cooled,
        }
    return out


def posterior_int8_sample(gp, xs, fp64_value, impl):
    """The opt-in INT8 path on a sample of the same candidates: candidates/s and the difference to the FP64 path's results
    on that sample.  Reported beside the headline, not as the headline (north_star quotes the FP64 roof; error ~2e-13 of max var for the fused
    kernel, ~2e-11 for the library variant, against ~1e-14 of the FP64 kernel).
    impl "fused": this repo's tcgen05 kernel (csrc/int8_fused.cu); "library": CUTLASS int8 GEMMs between this repo's
    slicing and recombination kernels (csrc/int8_posterior.cu)."""
    mean0, var0 = gp.posterior(xs)
    gp.int8, gp.int8_impl = True, impl
    gp._packed_i8 = None
    try:
        gp.posterior(xs[:65536])                                   # slices L^-1, allocates the workspace
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mean1, var1 = gp.posterior(xs)
        e1.record()
        torch.cuda.synchronize()
    finally:
        gp.int8 = False
        gp._ws_i8 = None
        gp._packed_i8 = None
    ms = e0.elapsed_time(e1)
    v = xs.shape[0] / (ms * 1e-3)
    notes = {"fused": "opt-in (ExactGPPosterior(int8=True) / MGB_POSTERIOR_PATH=int8): error-free 8-bit slicing of K* and L^-1 (carry pass keeps the digits in int8); "
                      "the 28 slice products, the float64 recombination and the sum of squares in ONE hand-written kernel "
                      "(tcgen05.mma kind::i8, seven int32 planes of a 128 x 64 tile in tensor memory, operands pre-tiled "
                      "in the tensor core's shared-memory image and staged by cp.async.bulk, exact triangular k range); "
                      "parity bar 1e-10",
             "library": "int8='library' / MGB_INT8_IMPL=library: the same slicing, 28 slice products as 7 x 5 int8 GEMMs "
                        "(library kernel: CUTLASS collective builder, tcgen05 kind::i8) over column blocks of the "
                        "triangular factor, float64 recombination pass over the int32 planes in HBM"}
    return {"value": v, "unit": "candidates/s", "impl": impl, "sample_candidates": int(xs.shape[0]), "ms": ms,
            "ratio_to_fp64_path": v / fp64_value,
            "max_var_diff_rel_to_max_var": float((var1 - var0).abs().max() / var0.abs().max()),
            "max_mean_diff_rel_to_max_mean": float((mean1 - mean0).abs().max() / mean0.abs().max()),
            "note": notes[impl]}


def posterior_roofline(gp, xc, n, lib):
    """Average duration of posterior_reduce_kernel launches (the FLOP-dominant kernel), CUDA events on the
    launching stream, against the FP64 tensor (DMMA) peak micro-benchmarked in this same run."""
    from materngabo_b200 import _lib

    chunk = gp.chunk
    reps = min(6, max(1, xc.shape[0] // chunk))
    if xc.shape[0] < chunk:
        chunk = xc.shape[0]
        reps = 1
    kb = torch.zeros(lib.mgb_posterior_tiled_bytes(chunk, n) // 8, dtype=F64, device=xc.device)
    part = torch.empty(max(1, lib.mgb_posterior_reduce_workspace_bytes(chunk, n) // 8), dtype=F64, device=xc.device)
    kss = torch.full((chunk,), gp.kss_value, dtype=F64, device=xc.device)
    mean = torch.empty(chunk, dtype=F64, device=xc.device)
    var = torch.empty(chunk, dtype=F64, device=xc.device)
    d = gp.spec.desc(gp.coef.shape[-1])
    st = _lib.stream_ptr(xc.device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(reps + 1)]
    for i in range(reps + 1):      # first pass warms up
        x = xc[(i % reps) * chunk:(i % reps) * chunk + chunk]
        ev[i][0].record()
        _lib.check(lib.mgb_series_gram_tiled(C.byref(d), _lib.ptr(x), chunk, x.stride(0), _lib.ptr(gp.train_prepared), n,
                                             gp.train_prepared.stride(0), _lib.ptr(gp.coef), _lib.ptr(kb), st), "gram")
        ev[i][1].record()
        _lib.check(lib.mgb_posterior_reduce(_lib.ptr(kb), chunk, n, _lib.ptr(gp._packed), _lib.ptr(kss),
                                            gp.mean_const, _lib.ptr(mean), _lib.ptr(var), _lib.ptr(part),
                                            part.numel() * 8, st), "reduce")
        ev[i][2].record()
    torch.cuda.synchronize()
    gram_ms = statistics.mean(ev[i][0].elapsed_time(ev[i][1]) for i in range(1, reps + 1))
    red_ms = statistics.mean(ev[i][1].elapsed_time(ev[i][2]) for i in range(1, reps + 1))
    tf, pk_ms = C.c_double(), C.c_double()
    _lib.check(lib.mgb_fp64_peak(1, 3, C.byref(tf), C.byref(pk_ms)), "mgb_fp64_peak")
    tf_dfma = C.c_double()
    _lib.check(lib.mgb_fp64_peak(0, 3, C.byref(tf_dfma), None), "mgb_fp64_peak")
    alg_flops = chunk * (float(n) * n + 4.0 * n)      # triangular product + mean + norm, per launch
    achieved = alg_flops / (red_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "posterior_reduce_kernel (FP64 DMMA m8n8k4)", "achieved": achieved,
            "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
            "traffic": NCU_REDUCE_TRAFFIC_BYTES if (chunk, n) == (37888, 5000) else None,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture "
                              "profiles/r01b_ncu_posterior_reduce.txt (chunk 37888, n 5000); the kernel is FP64-tensor-pipe bound, this traffic is 7 %% of the HBM peak; algorithmic input bytes per "
                              "launch = Gram chunk 8*chunk*n = %.2e + packed factor %.2e"
                              % (8.0 * chunk * n, 8.0 * n * n / 2),
            "peak_source": "FP64 DMMA register-resident micro-benchmark measured in this run "
                           "(MEASURED_PEAKS.json has no FP64 figure; nominal %.1f; DFMA micro-benchmark %.1f)"
                           % (NOMINAL_FP64_TFLOPS, tf_dfma.value),
            "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
            "launch_ms": red_ms, "algorithmic_flops_per_launch": alg_flops, "candidates_per_launch": chunk,
            "gram_kernel": {"launch_ms": gram_ms, "achieved_GBps": 8.0 * chunk * n / (gram_ms * 1e-3) / 1e9,
                            "hbm_peak_GBps": measured_peaks().get("hbm_gbs"), "bytes_per_entry": 8}}


def posterior_e2e(args, gp, rank, world, local, m_total, m_local, n, xc, dist):
    """Same metric through the persistent C-ABI handle (include/mgb.h: mgb_posterior_create_from_device /
    mgb_posterior_eval_host): pinned HOST candidates in, HOST mean / var out, H2D + D2H inside the timed region.  The
    fitted state is the model - it is resident on every rank's device after the NCCL broadcast of step(), exactly as
    gpytorch keeps its prediction caches on the device between acq(X) calls - so each call moves candidates and results
    only."""
    device = torch.device("cuda", local)
    if gp.int8 and gp.int8_impl != "fused":
        return posterior_e2e_int8(args, gp, rank, world, local, m_total, m_local, n, xc, dist)
    xc_h = xc.cpu().pin_memory()
    mean_h = torch.empty(m_local, dtype=F64).pin_memory()
    var_h = torch.empty(m_local, dtype=F64).pin_memory()
    handle = gp.handle(gp.chunk)          # INT8 mode (fused kernel) follows gp.int8: mgb_posterior_use_int8
    via = "INT8 tensor-core contraction, csrc/int8_fused.cu" if handle.int8 else "FP64 DMMA kernel"
    steps = max(1, min(args.steps, 2))
    handle.eval_host(xc_h[: min(m_local, 4 * gp.chunk)].contiguous())       # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        handle.eval_host(xc_h, mean_h, var_h)
    dt = time.perf_counter() - t0
    handle.destroy()
    t = torch.tensor([dt], dtype=F64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    return {"value": m_total / (dt / steps), "unit": "candidates/s", "h2d_bytes_per_step": 8 * m_local * 11,
            "d2h_bytes_per_step": 8 * 2 * m_local, "steps": steps,
            "api": "mgb_posterior_eval_host on a persistent handle (C-ABI, pinned host buffers, per rank; fitted state "
                   "resident on the device; %s)" % via,
            "sample_mean0": float(mean_h[0]), "timer": "host wall clock around the synchronous C-ABI call, max over ranks"}


def posterior_e2e_int8(args, gp, rank, world, local, m_total, m_local, n, xc, dist):
    """Same metric through ExactGPPosterior.posterior_host on the INT8 path: pinned HOST candidates in, HOST mean / var
    out; the fitted state (sliced factor, alpha, training points) is already on every rank's device after step()."""
    device = torch.device("cuda", local)
    xc_h = xc.cpu().pin_memory()
    mean_h = torch.empty(m_local, dtype=F64).pin_memory()
    var_h = torch.empty(m_local, dtype=F64).pin_memory()
    steps = max(1, min(args.steps, 2))
    gp.posterior_host(xc_h[: min(m_local, 1 << 20)])          # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        gp.posterior_host(xc_h, mean_h, var_h)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=F64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    return {"value": m_total / (dt / steps), "unit": "candidates/s", "h2d_bytes_per_step": 8 * m_local * 11,
            "d2h_bytes_per_step": 8 * 2 * m_local, "steps": steps,
            "api": "ExactGPPosterior.posterior_host (INT8 path, pinned host buffers, per rank; fitted state resident on the device)",
            "sample_mean0": float(mean_h[0]), "timer": "host wall clock around the synchronous call, max over ranks"}


# ---------------------------------------------------------------------------------------------------
# Gram workloads (single GPU or row-sharded, weak scaling: rows per GPU fixed)
# ---------------------------------------------------------------------------------------------------
def gram_setup(name, device, m, n):
    from materngabo_b200 import kernels_hyperbolic, kernels_so, kernels_spd, kernels_sphere, kernels_torus

    gen = torch.Generator(device=device).manual_seed(0)
    if name in ("gram-s2", "gram-s10", "config1-s2"):
        dim = 10 if name == "gram-s10" else 2
        k = kernels_sphere.SphereRiemannianMaternKernel(dim=dim, nu=2.5)
        k.lengthscale = 0.5
        x1, x2 = unit_rows(m, dim + 1, gen, device), unit_rows(n, dim + 1, gen, device)
        flops = 2 * (dim + 1) + 2 + 6 * 8 + 6
    elif name == "gram-so3":
        k = kernels_so.SORiemannianMaternKernel(dim=3, nu=2.5)
        k.lengthscale = 0.5

        def rot(cnt):
            q, r = torch.linalg.qr(torch.randn(cnt, 3, 3, generator=gen, device=device, dtype=F64))
            q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
            q[:, :, 0] = q[:, :, 0] * torch.sign(torch.linalg.det(q)).unsqueeze(-1)
            return q.reshape(cnt, 9).contiguous()
        x1, x2 = rot(m), rot(n)
        flops = 62
    elif name == "gram-so5":
        # SO(5) character sum at the canonical eigen-angles (csrc/son_gram.cu): latency / FP64 bound,
        # no roofline convention in the survey - reported as entries/s with the oracle-port CPU baseline beside it
        k = kernels_so.SORiemannianMaternKernel(dim=5, nu=2.5)
        k.lengthscale = 0.5

        def rot5(cnt):
            q, r = torch.linalg.qr(torch.randn(cnt, 5, 5, generator=gen, device=device, dtype=F64))
            q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
            q[:, :, 0] = q[:, :, 0] * torch.sign(torch.linalg.det(q)).unsqueeze(-1)
            return q.reshape(cnt, 25).contiguous()
        x1, x2 = rot5(m), rot5(n)
        flops = 0
    elif name == "gram-t10":
        k = kernels_torus.TorusProductOfManifoldsRiemannianMaternKernel(dim=10, nu=2.5)
        for kk in k.torus_kernel.kernels:
            kk.lengthscale = 0.5
        x1 = k._to_circles(torch.rand(m, 10, generator=gen, device=device, dtype=F64) * 6.283185307179586)
        x2 = k._to_circles(torch.rand(n, 10, generator=gen, device=device, dtype=F64) * 6.283185307179586)
        # 10 factors x (6 + 4 n_eff) flops, n_eff = retained cosine terms per factor
        spec_, coef_ = k.series()
        n_eff = int(coef_.shape[-1])
        flops = 10 * (6 + 4 * n_eff)
        # executed FMA slots: 2-wide inner product + clamp/product bookkeeping (3) + one FMA per term in the monomial
        # basis (two in the Chebyshev basis), 2 flops each
        EXECUTED["gram-t10"] = 10 * (2 + 3 + (1 if spec_.basis == 0 else 2) * n_eff) * 2
    elif name == "gram-h2":
        k = kernels_hyperbolic.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=64)
        k.lengthscale = 1.0

        def lor(cnt):
            z = torch.randn(cnt, 2, generator=gen, device=device, dtype=F64)
            return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1).contiguous()
        x1, x2 = lor(m), lor(n)
        flops = 4096 * 32
    elif name == "gram-spd2":
        k = kernels_spd.SpdRiemannianIntegratedMaternKernel(2, nu=2.5, nb_points_integral_b=64, nb_points_integral_t=64)
        k.lengthscale = 1.0

        def spd(cnt):
            a = torch.randn(cnt, 2, 2, generator=gen, device=device, dtype=F64)
            s = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2, device=device, dtype=F64)
            return torch.stack([s[:, 0, 0], s[:, 1, 1], s[:, 0, 1] * 2 ** 0.5], -1).contiguous()
        x1, x2 = spd(m), spd(n)
        flops = 4096 * 32
    elif name == "gram-h2-table":
        # OPT-IN tabulated profile (materngabo_b200/tabulated.py): NOT the per-entry quadrature of config 4 - the
        # quadrature is evaluated once per table node and each entry interpolates.  Reported under its own name.
        from materngabo_b200.tabulated import TabulatedRadialKernel
        base = kernels_hyperbolic.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=64).to(device)
        base.lengthscale = 1.0

        def lor(cnt):
            z = torch.randn(cnt, 2, generator=gen, device=device, dtype=F64)
            return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1).contiguous()
        x1, x2 = lor(m), lor(n)
        k = TabulatedRadialKernel.for_inputs(base, x1, x2)
        flops = 0
    elif name == "gram-spd3":
        # The reference has no intrinsic SPD(3) kernel (kernels_spd.py:59-60 raises for dim != 2); SPD(3) goes through
        # the product kernel R^3 x SO(3) on decomposed inputs (eigenvalues | flattened rotation), SURVEY.md 8d cfg 4.
        k = kernels_spd.SpdProductOfRSOManifoldsRiemannianMaternKernel(dim=3, nu=2.5, spd_input=False)
        k.Euclidean_kernel.lengthscale = 1.0
        k.so_kernel.lengthscale = 0.5

        def dec(cnt):
            q, r = torch.linalg.qr(torch.randn(cnt, 3, 3, generator=gen, device=device, dtype=F64))
            q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
            q[:, :, 0] *= torch.sign(torch.linalg.det(q)).unsqueeze(-1)
            ev = 0.5 + 2.0 * torch.rand(cnt, 3, generator=gen, device=device, dtype=F64)
            return torch.cat([ev, q.reshape(cnt, 9)], -1).contiguous()
        x1, x2 = dec(m), dec(n)
        flops = 100 * 32 + 62        # 100 t nodes of the Euclidean factor (one exp each) + the SO(3) series
    else:
        raise SystemExit("unknown workload %s" % name)
    return k, x1, x2, flops


EXECUTED = {}


GRAM_DEFAULTS = {"config1-s2": (1000, 1000), "gram-s2": (1_000_000, 2000), "gram-s10": (1_000_000, 2000), "gram-so3": (1_000_000, 2000),
                 "gram-t10": (1_000_000, 2000), "gram-h2": (2000, 2000), "gram-spd2": (2000, 2000), "gram-spd3": (2000, 2000),
                 "gram-h2-table": (1_000_000, 2000), "gram-so5": (65536, 2048)}


# FP64 instructions (DFMA + DMUL + DADD, thread level) the quadrature Gram kernels EXECUTE per entry at 64 x 64 nodes,
# from the per-opcode warp-instruction counts of the committed ncu captures (x 32 lanes / entries of the profiled
# launch).  SURVEY.md 8d's "32 flops per node" is a convention the factorised kernels undercut, so the honest
# utilisation figure is executed instructions / s against the FP64 pipe's issue rate (measured DFMA TFLOP/s / 2).
EXECUTED_FP64_INSTR = {
    "gram-h2": (11440.0, "profiles/r01d_ncu_gram_h2_fwd4.txt: (200.43 M DFMA + 148.55 M DMUL + 8.54 M DADD) warp-instructions x 32 / 1.0e6 entries"),
    "gram-spd2": (31450.0, "profiles/r01d_ncu_gram_spd2_pairs.txt: (708.05 M DFMA + 226.96 M DMUL + 47.80 M DADD) warp-instructions x 32 / 1.0e6 entries"),
}


class _GramArgs:
    def __init__(self, workload, m, n, steps, warmup, no_cpu_baseline):
        self.workload, self.m, self.n, self.steps, self.warmup, self.no_cpu_baseline = workload, m, n, steps, warmup, no_cpu_baseline


def run_gram(args, rank, world, local, min_seconds=0.0, cpu_seconds=8.0):
    import torch.distributed as dist
    from materngabo_b200 import _lib

    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    lib = _lib.load()
    m, n = GRAM_DEFAULTS[args.workload]
    m, n = args.m or m, args.n or n
    kernel, x1, x2, flops = gram_setup(args.workload, device, m, n)

    def step():
        with torch.no_grad():
            return kernel.forward(x1, x2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step()
        del out
    barrier()
    steps = args.steps
    if min_seconds > 0:
        # sustained timing: enough steps for >= min_seconds of timed work (>= 10 clock samples at 200 ms), so that the
        # board reaches its load power / clocks instead of a 15 ms burst from idle
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(3):
            out = step()
            del out
        p1.record()
        torch.cuda.synchronize()
        est = max(p0.elapsed_time(p1) / 3.0, 1e-3)
        steps = max(args.steps, int(min_seconds * 1e3 / est) + 1)
        if 8.0 * m * n < 4 * 126e6:
            steps = min(steps, 20000)          # small problems: the untimed L2 flush dominates the wall clock
    sampler = ClockSampler(local).start() if rank == 0 else None
    launches0 = lib.mgb_launch_count()
    small = 8.0 * m * n < 4 * 126e6            # output would stay L2 resident between steps: flush L2 in between
    out = None
    protocol = None
    graphable = args.workload in ("config1-s2", "gram-s2", "gram-s10", "gram-so3")   # quadrature kernels read their node
    if small and not graphable:                                                       # grids back on the host: eager
        flush = torch.empty(64 * 1024 * 1024, dtype=F64, device=device)     # 512 MB > 126 MB L2
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            del out
            flush.fill_(1.0)               # untimed: evicts x1, x2 and the previous output
            a.record()
            out = step()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
    elif small:
        # Launch-bound size: the step is replayed from a CUDA graph so that the event pair brackets device time only
        # (eagerly, the Python time to enqueue the launch would sit between the two events).
        flush = torch.empty(64 * 1024 * 1024, dtype=F64, device=device)     # 512 MB > 126 MB L2
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            out = step()
        torch.cuda.current_stream(device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = step()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.fill_(1.0)               # untimed: evicts x1, x2 and the previous output
            a.record()
            graph.replay()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches0 -= steps            # replays do not pass through the library's launch counter
        # What the protocol itself costs: the same flush / event / graph-replay / event sequence around a ONE-ELEMENT
        # kernel.  On B200 that floor is ~6 us (gpurun_out/r2_config1_probe.log), i.e. most of a 10 us reading is the
        # replay launch path between the two events, not the Gram kernel.
        tiny = torch.zeros(32, dtype=F64, device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            tiny.add_(1.0)
        torch.cuda.current_stream(device).wait_stream(side)
        fgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(fgraph):
            tiny.add_(1.0)
        fevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(steps, 400))]
        for a, b in fevs:
            flush.fill_(1.0)
            a.record()
            fgraph.replay()
            b.record()
        barrier()
        floor_us = statistics.median(a.elapsed_time(b) for a, b in fevs) * 1e3
        median_us = statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3
        protocol = {"event_to_event_us_median": median_us, "floor_us_one_element_kernel": floor_us,
                    "above_floor_us": median_us - floor_us,
                    "entries_per_s_above_floor": float(m) * n / max((median_us - floor_us) * 1e-6, 1e-9),
                    "note": "value / ms_per_step are the plain event-to-event figures; the floor is the same flush + event + "
                            "CUDA-graph replay + event sequence around a one-element kernel"}
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            del out        # the caching allocator hands the same block back: no cudaMalloc inside the timed region
            out = step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
    launches = lib.mgb_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, float(launches)], dtype=F64, device=device)
    if world > 1:
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(tsum[1])
    ms_per_step = ms / steps
    entries = float(m) * n * world
    value = entries / (ms_per_step * 1e-3)
    if rank != 0:
        return None
    peaks = measured_peaks()
    quad = args.workload in ("gram-h2", "gram-spd2", "gram-spd3", "gram-t10", "gram-so5")
    if quad:
        tf = C.c_double()
        _lib.check(lib.mgb_fp64_peak(0, 3, C.byref(tf), None), "mgb_fp64_peak")
        rate = float(m) * n / (ms_per_step * 1e-3)                          # entries/s of one GPU
        ach = (flops or 0) * rate / 1e12 if flops else None
        executed = EXECUTED.get(args.workload)                              # flops (FMA = 2) where the mix is all-FMA
        executed_frac = executed * rate / 1e12 / tf.value if executed else None
        executed_src = "FMA slots of the kernel's inner loops x 2 (see gram_setup)" if executed else None
        if args.workload in EXECUTED_FP64_INSTR and (m, n) == GRAM_DEFAULTS[args.workload]:
            instr, executed_src = EXECUTED_FP64_INSTR[args.workload]
            executed = instr                                                 # FP64 instructions per entry
            executed_frac = instr * rate / (tf.value * 1e12 / 2.0)          # against the pipe's instruction issue rate
        roof = {"bound": "fp64", "achieved": ach, "peak": tf.value, "unit": "TFLOP/s",
                "frac": executed_frac if executed_frac is not None else ((ach / tf.value) if ach else None),
                "frac_definition": "executed FP64 operations / s over the measured DFMA peak" if executed_frac is not None
                else "SURVEY.md 8d algorithmic flops / s over the measured DFMA peak",
                "frac_survey_convention": (ach / tf.value) if ach else None, "traffic": None,
                "algorithmic_flops_per_entry": flops,
                "executed_per_entry": executed, "executed_unit": "FP64 instructions" if args.workload in EXECUTED_FP64_INSTR else "flops (FMA = 2)",
                "executed_source": executed_src, "executed_frac": executed_frac,
                "note": ("SURVEY.md 8d: 10 x (6 + 4 n_eff) flops per entry, n_eff = %d retained cosine terms per circle "
                         "factor (of the reference's 201)" % ((flops // 10 - 6) // 4)) if args.workload == "gram-t10" else
                        "frac = executed FP64 instructions / s over the pipe's issue rate; frac_survey_convention uses "
                        "SURVEY.md 8d's 32 flops per quadrature node (one exp = 28), which the factorised kernels undercut "
                        "(it can exceed 1 and is kept only for comparison with the survey's roof)",
                "peak_source": "DFMA micro-benchmark in this run"}
    else:
        ach = 8.0 * m * n / (ms_per_step * 1e-3) / 1e9
        hbm = peaks.get("hbm_gbs", 6650.0)
        tf = C.c_double()
        _lib.check(lib.mgb_fp64_peak(1, 3, C.byref(tf), None), "mgb_fp64_peak")
        hbm_roof = hbm * 1e9 / 8.0                                  # entries/s at 8 B per entry
        fp64_roof = tf.value * 1e12 / flops if flops else None      # SURVEY.md 8d algorithmic flops per entry
        lower = min(hbm_roof, fp64_roof) if fp64_roof else hbm_roof
        w = int(x1.shape[1])
        slots = 4 * ((w + 3) // 4) + (int(kernel.series()[1].shape[-1]) - 1 if hasattr(kernel, "series") else 9)
        if args.workload == "gram-so3" and 8.0 * m * n >= 4e6 * 8:
            slots = 4 + 1 + 9                       # quaternion form: 4-wide inner product, square, 10-term Horner
        exec_roof = tf.value * 1e12 / (2.0 * slots)
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "bytes_per_entry": 8,
                "fp64_executed": {"fma_slots_per_entry": slots, "roof_entries_per_s": exec_roof, "frac": value / world / exec_roof,
                                  "note": "executed FMA slots on the shared FP64 data path: DMMA inner product padded to a "
                                          "multiple of 4 + one DFMA per Horner term (DESIGN.md section 4)"},
                "survey_8d": {"algorithmic_flops_per_entry": flops, "hbm_roof_entries_per_s": hbm_roof,
                              "fp64_roof_entries_per_s": fp64_roof, "fp64_peak_tflops_measured": tf.value,
                              "frac_of_lower_roof": value / lower}}
    return {"metric": "kernel_entries_per_sec", "value": value, "unit": "entries/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "timed_seconds": ms * 1e-3,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s K(X*, X) %d x %d per GPU" % (args.workload, m, n), "rows": m, "cols": n,
                       "l2_policy": (("L2 flushed (512 MB fill, untimed) before every timed step; each step is "
                                      + ("a CUDA-graph replay" if graphable else "an eager call")
                                      + " timed by its own CUDA event pair") if small
                                     else "output %.2f GB per step >> L2" % (8e-9 * m * n))},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "e2e": None,
            "checksum": float(out[:8, :8].sum()), "timing_protocol": protocol,
            "cpu_baseline": None if (args.no_cpu_baseline or world > 1) else cpu_gram_baseline(args.workload, x1, x2, cpu_seconds)}


GRAM_LEG = ("config1-s2", "gram-s2", "gram-s10", "gram-so3", "gram-t10", "gram-h2", "gram-spd2")


def gram_leg(args, rank, world, local):
    """The "kernel entries/sec (fp64)" half of BASELINE.json's metric inside the default (driver-run) line: every Gram
    workload of configs[0], [2], [3] (+ S^2 / S^10 at the 1 M x 2 k shape), each timed for >= 2 s with its own clock
    record, roofline (HBM-write and FP64, executed-operation fractions) and oracle-port CPU baseline."""
    out = {}
    for wl in GRAM_LEG:
        ga = _GramArgs(wl, None, None, 5, 3, args.no_cpu_baseline)
        try:
            r = run_gram(ga, rank, world, local, min_seconds=2.0, cpu_seconds=2.0)
        except Exception as exc:                      # one failing workload must not take the headline line with it
            out[wl] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            continue
        out[wl] = {k: r[k] for k in ("value", "unit", "steps", "ms_per_step", "timed_seconds", "config", "clocks", "roofline",
                                     "cpu_baseline", "gpu_launches", "checksum", "timing_protocol")}
        torch.cuda.empty_cache()
    return out


def run_bo_latency(args, local):
    """BASELINE.json configs[1]: GaBO on S^3 with 100 training points - microseconds per call of what the BO loop
    calls (SURVEY.md 8d cfg 2: latency-bound, no roofline): one MLL step (Gram + Cholesky + backward to
    raw_lengthscale), the posterior fit, EI for b = 100 / 500 candidates in botorch's b x 1 x 4 layout, and one
    trust-region inner step (EI value + gradient at one candidate)."""
    from materngabo_b200 import _lib, kernels_sphere
    from materngabo_b200.gpytorch_compat import ScaleKernel
    from materngabo_b200.posterior import ExactGPPosterior, ExpectedImprovement

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    gen = torch.Generator().manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(100, 4, generator=gen), dim=-1).to(dev)
    y = torch.sin(3 * x[:, 0])
    base = kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5)
    sk = ScaleKernel(base).to(dev)
    eye = 1e-2 * torch.eye(100, device=dev)

    def timeit(fn, reps):
        for _ in range(max(args.warmup, 3) * 5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    def mll_step():
        for prm in sk.parameters():
            prm.grad = None
        k = sk(x)
        L = torch.linalg.cholesky(k + eye)
        a = torch.cholesky_solve(y.unsqueeze(-1), L).squeeze(-1)
        (0.5 * (y * a).sum() + torch.log(torch.diagonal(L)).sum()).backward()

    reps = 50 * args.steps
    sampler = ClockSampler(local).start()
    launches0 = lib.mgb_launch_count()
    lat = {}
    base.raw_lengthscale.requires_grad_(True)
    lat["mll_step_fwd_bwd_us"] = timeit(mll_step, reps)
    base.raw_lengthscale.requires_grad_(False)
    with torch.no_grad():
        lat["gram_100x100_fwd_us"] = timeit(lambda: base.forward(x, x), reps)
    gp = ExactGPPosterior(base, x, y, outputscale=1.0, noise=1e-2).fit()
    lat["fit_n100_us"] = timeit(lambda: gp.fit(), reps)
    acq = ExpectedImprovement(gp, best_f=float(y.min()), maximize=False)
    for b in (100, 500):
        xc = torch.nn.functional.normalize(torch.randn(b, 1, 4, generator=gen), dim=-1).to(dev)
        with torch.no_grad():
            lat["ei_b%d_us" % b] = timeit(lambda: acq(xc), reps)
    x1 = torch.nn.functional.normalize(torch.randn(1, 1, 4, generator=gen), dim=-1).to(dev)

    def tr_step():
        xx = x1.clone().requires_grad_(True)
        return torch.autograd.grad(-acq(xx).sum(), [xx])[0]

    lat["ei_value_and_grad_b1_us"] = timeit(tr_step, reps)

    # the same two launch-bound calls replayed from CUDA graphs (materngabo_b200/graphs.py)
    from materngabo_b200.graphs import GraphedValueAndGrad, exact_mll
    base.raw_lengthscale.requires_grad_(True)
    wrt = [base.raw_lengthscale, sk.raw_outputscale]

    def mll_fused_eager():
        loss, _ = exact_mll(sk, x, y, 1e-2, 0.0)          # Gram + Cholesky + solves + gradients in ONE launch (csrc/mll.cu)
        return torch.autograd.grad(loss, wrt)

    lat["mll_step_fused_eager_us"] = timeit(mll_fused_eager, reps)
    for name, fused in (("mll_step_graph_unfused_us", False), ("mll_step_graph_us", True)):
        gm = GraphedValueAndGrad(lambda: exact_mll(sk, x, y, 1e-2, 0.0, fused=fused), inputs=(), wrt=wrt)

        def mll_graph():
            value, grads = gm()
            return float(value)             # the L-BFGS driver reads the loss back every evaluation

        lat[name] = timeit(mll_graph, reps)
    base.raw_lengthscale.requires_grad_(False)
    v1 = torch.randn(1, 1, 4, generator=gen).to(dev)
    gt = GraphedValueAndGrad(lambda xx, vv: -acq(xx).sum(), inputs=(x1, v1), wrt=[0], hvp_index=1)

    def tr_graph():
        value, grads, hvp = gt(x1, v1)
        return float(value)

    lat["ei_value_grad_hvp_b1_graph_us"] = timeit(tr_graph, reps)

    # the same three quantities from ONE fused launch (csrc/acquisition.cu), one candidate and all 5 restarts at once
    def tr_fused():
        ei, g, h = acq.value_grad_hvp(x1, v1)
        return float(ei)                     # pymanopt's solver reads the cost back on the host

    lat["ei_value_grad_hvp_b1_fused_us"] = timeit(tr_fused, reps)
    x5 = torch.nn.functional.normalize(torch.randn(5, 1, 4, generator=gen), dim=-1).to(dev)
    v5 = torch.randn(5, 1, 4, generator=gen).to(dev)

    def tr_fused5():
        ei, g, h = acq.value_grad_hvp(x5, v5)
        return ei.cpu()

    lat["ei_value_grad_hvp_b5_fused_us"] = timeit(tr_fused5, reps)
    # the same marginal-likelihood step on the other manifolds of the reference's BO scripts, N = 100, CUDA-graph replay
    # (sync-free hyper-parameter chains: mgb_matern_t_grid / circle_coef_sync_free)
    from materngabo_b200 import kernels_hyperbolic, kernels_spd, kernels_torus

    def _hyp(dim):
        z = 0.7 * torch.randn(100, dim, generator=gen)
        return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)

    def _spd():
        a = torch.randn(100, 2, 2, generator=gen)
        sm = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
        return torch.stack([sm[:, 0, 0], sm[:, 1, 1], sm[:, 0, 1] * 2 ** 0.5], -1)

    for tag, kern, pts in (("torus_T3", kernels_torus.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5), torch.rand(100, 3, generator=gen) * 6.283185307179586),
                           ("hyperbolic_H2", kernels_hyperbolic.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=2, nu=2.5, nb_points_integral=128), _hyp(2)),
                           ("hyperbolic_H3", kernels_hyperbolic.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=3, nu=2.5), _hyp(3)),
                           ("spd2", kernels_spd.SpdRiemannianIntegratedMaternKernel(2, nu=2.5), _spd())):
        try:
            skm = ScaleKernel(kern).to(dev)
            xm = pts.to(dev)
            ym = torch.sin(3 * xm[:, 0])
            prm = [p_ for p_ in skm.parameters() if p_.requires_grad]
            gmm = GraphedValueAndGrad(lambda: exact_mll(skm, xm, ym, 5e-2, 0.0), inputs=(), wrt=prm)

            def _step():
                value, grads = gmm()
                return float(value)

            lat["mll_step_graph_%s_us" % tag] = timeit(_step, reps)
        except Exception as exc:
            lat["mll_step_graph_%s_us" % tag] = "failed: %s" % type(exc).__name__
    launches = lib.mgb_launch_count() - launches0
    clocks = sampler.stop()
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import restated as R
        use_all_host_threads()
        xc_, yc_ = x.cpu(), y.cpu()
        raw = torch.zeros(1, 1, requires_grad=True)

        def cpu_mll():
            raw.grad = None
            ls = torch.nn.functional.softplus(raw)
            k = R.sphere_matern(xc_, xc_, 3, 2.5, ls) * 0.6931471805599453
            L = torch.linalg.cholesky(k + 1e-2 * torch.eye(100))
            a = torch.cholesky_solve(yc_.unsqueeze(-1), L).squeeze(-1)
            (0.5 * (yc_ * a).sum() + torch.log(torch.diagonal(L)).sum()).backward()

        for _ in range(3):
            cpu_mll()
        t0 = time.perf_counter()
        for _ in range(20):
            cpu_mll()
        cpu = {"value": (time.perf_counter() - t0) / 20 * 1e6, "unit": "us per MLL step", "cores": torch.get_num_threads(),
               "kind": "port", "sample": "oracle port (reference torch operator sequence): 20 MLL steps, N = 100, S^3"}
    return {"metric": "mll_step_latency", "value": lat["mll_step_graph_us"], "unit": "us", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": lat["mll_step_graph_us"] * 1e-3,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: GaBO on S^3 (Matern nu=2.5), 100 training points: MLL step (value = CUDA-graph "
                                   "replay of the fused single-CTA kernel + coefficient chain; eager and unfused variants in latencies_us), "
                                   "fit, EI posterior b=100/500 (b x 1 x 4), trust-region step b=1; host wall clock per call "
                                   "incl. Python and launch overhead, %d calls each" % reps,
                       "l2_policy": "latency-bound: working set (80 KB) is cache resident by construction"},
            "latencies_us": lat, "gpu_launches": launches, "clocks": clocks, "roofline": None, "e2e": None,
            "cpu_baseline": cpu}


def cpu_gram_baseline(workload, x1, x2, seconds=8.0):
    """Oracle port (the reference's torch CPU algorithm, all host threads) on a bounded block of the same inputs:
    entries/s.  Block sizes are chosen so that one evaluation takes seconds (the reference materialises
    O(P x M x N) temporaries for the quadrature kernels and O(M x N x D) for the distances)."""
    from oracle import restated as R
    use_all_host_threads()
    blocks = {"gram-s2": (4000, 2000), "gram-s10": (4000, 2000), "config1-s2": (1000, 1000), "gram-so3": (800, 500), "gram-so5": (24, 24),
              "gram-t10": (100, 200), "gram-h2": (200, 300), "gram-spd2": (150, 100), "gram-spd3": (600, 400)}
    if workload not in blocks:
        return None
    bm, bn = blocks[workload]
    a, b = x1[:bm].cpu(), x2[:bn].cpu()
    fns = {"gram-s2": lambda: R.sphere_matern(a, b, 2, 2.5, 0.5), "config1-s2": lambda: R.sphere_matern(a, b, 2, 2.5, 0.5),
           "gram-s10": lambda: R.sphere_matern(a, b, 10, 2.5, 0.5), "gram-so3": lambda: R.so3_matern(a, b, 2.5, 0.5),
           "gram-so5": lambda: R.son_matern(a, b, 5, 2.5, 0.5),
           "gram-t10": lambda: R.torus_matern(a, b, 10, 2.5, 0.5), "gram-h2": lambda: R.hyperbolic_integrated_matern(a, b, 2, 2.5, 1.0, 64),
           "gram-spd2": lambda: R.spd2_integrated_matern(a, b, 2.5, 1.0, 64, 64),
           "gram-spd3": lambda: R.spd_product_matern(a, b, 3, 2.5, 1.0, 0.5, "so")}
    fn = fns[workload]
    with torch.no_grad():
        fn()                                   # warm-up
        t0 = time.perf_counter()
        reps = 0
        while reps < 20 and (time.perf_counter() - t0 < seconds or reps < 2):
            fn()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
    return {"value": a.shape[0] * b.shape[0] / dt, "unit": "entries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle port (reference's torch CPU algorithm) on a %d x %d block of the same inputs, %d repetitions, "
                      "%.3f s each" % (a.shape[0], b.shape[0], reps, dt)}


def run_tr_latency(args, local):
    """Trust-region inner step (EI value + gradient (+ Hessian-vector product)) at ONE candidate, N = 100 training
    points, on every manifold of the reference's BO scripts: fused acquisition kernel against torch autograd through the
    differentiable posterior path (what the pymanopt closures of the reference do).  Microseconds per call incl. the
    host read-back of the cost."""
    from materngabo_b200 import _lib, kernels_hyperbolic, kernels_so, kernels_spd, kernels_sphere, kernels_torus
    from materngabo_b200.posterior import ExactGPPosterior, ExpectedImprovement

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    gen = torch.Generator().manual_seed(0)
    n = args.n or 100

    def unit(m, d):
        return torch.nn.functional.normalize(torch.randn(m, d, generator=gen), dim=-1)

    def rot(m):
        q, r = torch.linalg.qr(torch.randn(m, 3, 3, generator=gen))
        q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
        q[:, :, 0] *= torch.sign(torch.linalg.det(q)).unsqueeze(-1)
        return q.reshape(m, 9)

    def lor(m, d):
        z = torch.randn(m, d, generator=gen)
        return torch.cat([torch.sqrt(1 + (z ** 2).sum(-1, keepdim=True)), z], -1)

    def spd(m):
        a = torch.randn(m, 2, 2, generator=gen)
        sm = a @ a.transpose(-1, -2) + 0.5 * torch.eye(2)
        return torch.stack([sm[:, 0, 0], sm[:, 1, 1], sm[:, 0, 1] * 2 ** 0.5], -1)

    def tor(m, d):
        a = torch.rand(m, d, generator=gen) * 6.283185307179586
        return torch.stack([torch.cos(a), torch.sin(a)], -1).reshape(m, 2 * d)

    cases = [("sphere S^3", kernels_sphere.SphereRiemannianMaternKernel(dim=3, nu=2.5), unit(n + 1, 4), 0.5, 1e-2, True),
             ("SO(3)", kernels_so.SORiemannianMaternKernel(dim=3, nu=2.5), rot(n + 1), 0.5, 1e-2, True),
             ("torus T^3", kernels_torus.TorusProductOfManifoldsRiemannianMaternKernel(dim=3, nu=2.5), tor(n + 1, 3), 0.5, 1e-2, True),
             ("hyperbolic H^3", kernels_hyperbolic.HyperbolicRiemannianMillsonIntegratedMaternKernel(dim=3, nu=2.5), lor(n + 1, 3), 1.0, 1e-2, True),
             ("SPD(2)", kernels_spd.SpdRiemannianIntegratedMaternKernel(2, nu=2.5), spd(n + 1), 1.0, 5e-2, False)]

    def timeit(fn, reps):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    reps = 30 * args.steps
    sampler = ClockSampler(local).start()
    launches0 = lib.mgb_launch_count()
    lat = {}
    for name, k, pts, ls, noise, has_hvp in cases:
        k = k.to(dev)
        if hasattr(k, "torus_kernel"):
            for kk in k.torus_kernel.kernels:
                kk.lengthscale = ls
        else:
            k.lengthscale = ls
        for prm in k.parameters():
            prm.requires_grad_(False)
        xt, xc = pts[:n].to(dev), pts[n:].to(dev)
        y = torch.sin(2.0 * xt[:, 0]) + 0.3 * xt[:, 1]
        gp = ExactGPPosterior(k, xt, y, outputscale=1.0, noise=noise).fit()
        acq = ExpectedImprovement(gp, best_f=float(y.min()), maximize=False)
        v = torch.randn(1, xc.shape[-1], generator=gen).to(dev)

        def fused():
            ei, g, h = acq.value_grad_hvp(xc, v if has_hvp else None)
            return float(ei)

        def autograd():
            xa = xc.clone().requires_grad_(True)
            val = -acq(xa.unsqueeze(-2)).sum()
            (g,) = torch.autograd.grad(val, [xa], create_graph=has_hvp)
            if has_hvp:
                torch.autograd.grad((g * v).sum(), [xa])
            return float(val)

        lat[name] = {"fused_us": timeit(fused, reps), "autograd_us": timeit(autograd, max(10, reps // 3)),
                     "quantities": "value + gradient + Hessian-vector product" if has_hvp else "value + gradient"}
    launches = lib.mgb_launch_count() - launches0
    clocks = sampler.stop()
    return {"metric": "trust_region_step_latency", "value": lat["sphere S^3"]["fused_us"], "unit": "us", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": lat["sphere S^3"]["fused_us"] * 1e-3,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "trust-region inner step at one candidate, %d training points, Matern nu=2.5 on every manifold "
                                   "of the reference's BO scripts; fused acquisition kernels vs torch autograd through the "
                                   "differentiable posterior (value = S^3 fused); host wall clock per call" % n,
                       "l2_policy": "latency-bound: working set is cache resident by construction"},
            "latencies_us": lat, "gpu_launches": launches, "clocks": clocks, "roofline": None, "e2e": None,
            "cpu_baseline": None}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------
# CPU legs (oracle port = the reference's torch algorithm on the host cores)
# ---------------------------------------------------------------------------------------------------
def cpu_posterior_state(n):
    from oracle import restated as R

    gen = torch.Generator().manual_seed(0)
    xt = torch.randn(n, 11, generator=gen, dtype=F64)
    xt = xt / xt.norm(dim=-1, keepdim=True)
    y = torch.sin(3.0 * xt[:, 0]) + 0.5 * xt[:, 1] * xt[:, 2]
    kfn = lambda a, b: R.sphere_matern(a, b, 10, 2.5, 0.5)  # noqa: E731
    return R, kfn, xt, y


_CPU_FIT = {}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core."""
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)


def cpu_posterior_rate(n, sample, chunk=512):
    """candidates/s of the oracle port (reference algorithm, torch CPU ops, all host threads).  The fit
    (Cholesky of n x n) is done once outside the timed region, like rank 0's fit in our arm."""
    use_all_host_threads()
    R, kfn, xt, y = cpu_posterior_state(n)
    if n not in _CPU_FIT:
        _CPU_FIT[n] = R.gp_fit(kfn, xt, y, outputscale=1.0, noise=1e-2, mean_const=0.0)
    L, alpha = _CPU_FIT[n]
    gen = torch.Generator().manual_seed(1)
    xc = torch.randn(sample, 11, generator=gen, dtype=F64)
    xc = xc / xc.norm(dim=-1, keepdim=True)
    t0 = time.perf_counter()
    mean, var = R.gp_predict(kfn, xt, L, alpha, xc, outputscale=1.0, mean_const=0.0, chunk=chunk)
    dt = time.perf_counter() - t0
    return sample / dt, dt, float(mean[0])


def cpu_sample_size(n, seconds, lo=512, hi=65536):
    """Candidates the oracle port processes in about ``seconds`` on this host (pilot run of 512 candidates)."""
    rate, _, _ = cpu_posterior_rate(n, 512)
    return int(min(hi, max(lo, rate * seconds))) // 512 * 512


def run_reference(args, rank, world):
    if rank != 0:
        return None
    torch.set_default_dtype(F64)
    n = args.n or 5000
    use_all_host_threads()
    cores = torch.get_num_threads()
    # bounded sample: about 8 s of host work per step so that --steps K --warmup W ends within a few minutes
    sample = cpu_sample_size(n, 8.0)
    for _ in range(min(args.warmup, 1)):
        cpu_posterior_rate(n, 512)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, dt, _ = cpu_posterior_rate(n, sample)
        rates.append(r)
    total = time.perf_counter() - t0
    value = statistics.mean(rates)
    desc = ("oracle port (oracle/restated.py: the reference's torch operator sequence) on %d host threads; each step = "
            "posterior (Gram rows + triangular solve) of a %d-candidate sample of the S^10 workload against %d training "
            "points; fit outside the timed region" % (cores, sample, n))
    return {"impl": "reference", "metric": "posterior_cands_per_sec", "value": value, "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "S^10 Riemannian Matern nu=2.5 kappa=0.5 T=10 exact-GP posterior mean/var, "
                                   "%d training points; bounded sample of %d candidates per step" % (n, sample),
                       "train": n, "sample_candidates": sample},
            "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank, world, local = dist_env()
    torch.set_default_dtype(F64)
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the product path has no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        import datetime
        # a step is seconds at most: a rank that waits minutes for a peer is a bug, not a slow collective
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    try:
        if args.workload == "posterior-s10":
            out = run_posterior(args, rank, world, local)
        elif args.workload == "config2-bo-s3":
            out = run_bo_latency(args, local) if rank == 0 else None
        elif args.workload == "tr-step":
            out = run_tr_latency(args, local) if rank == 0 else None
        else:
            out = run_gram(args, rank, world, local)
        if rank == 0:
            if not args.no_cpu_baseline and args.workload == "posterior-s10" and world == 1:   # N = 1 only
                n = args.n or 5000
                sample = cpu_sample_size(n, 12.0)          # about 10-30 s of host work
                rate, dt, _ = cpu_posterior_rate(n, sample)
                out["cpu_baseline"] = {"value": rate, "unit": "candidates/s", "cores": torch.get_num_threads(),
                                       "kind": "port",
                                       "sample": "oracle port (reference's torch CPU algorithm): posterior of %d "
                                                 "candidates x %d training points, %.1f s (fit excluded)" % (sample, n, dt)}
            elif "cpu_baseline" not in out:
                out["cpu_baseline"] = None
            print(json.dumps(out), flush=True)
    except BaseException:
        if world > 1:
            # the peers are (or will be) waiting in a collective for this rank: report and take the job down now instead
            # of meeting them in a barrier that can never complete
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)
        raise
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
