"""not gpu: the reference arm of bench.py (the only arm that runs without a GPU) prints ONE JSON line with the keys the
driver reads; our arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--train", "200"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "posterior_cands_per_sec" and d["unit"] == "candidates/s"
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--train", "200"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_json_line_on_the_gpu():
    """Small instance of the default workload through bench.py: one JSON line with value, roofline, e2e (host buffers,
    copies counted), launches of our own kernels and clocks."""
    r = _run(["--steps", "1", "--warmup", "3", "--cands", "300000", "--train", "1000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "posterior_cands_per_sec" and d["n_gpus"] == 1 and d["value"] > 0 and d["dtype"] == "f64"
    assert d["gpu_launches"] > 0 and d["scaling"] in ("weak", "strong") and d["vs_baseline"] is None
    roof = d["roofline"]
    assert roof["bound"] == "tensor" and 0 < roof["frac"] <= 1.05 and roof["achieved"] > 0 and roof["peak"] > 0
    e2e = d["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] >= 8 * 300000 * 11 and e2e["d2h_bytes_per_step"] == 16 * 300000
    assert abs(e2e["value"] - d["value"]) > 0              # a separate measurement, not a copy of the device number
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert "sm_mhz" in d["clocks"] and "workload" in d["config"] and "l2_policy" in d["config"]
    i8 = d["int8_path"]
    assert i8 is None or (i8["value"] > 0 and i8["max_var_diff_rel_to_max_var"] < 1e-10)


@pytest.mark.gpu
def test_default_path_picks_the_int8_contraction_at_large_sizes_and_carries_the_fp64_measurement():
    """--posterior-path auto (the default) with >= 16384 candidates x >= 1024 training points: the headline runs on the
    INT8 tensor-core contraction, the same workload on the FP64 kernel and the difference between the two result sets
    over every candidate are part of the line."""
    r = _run(["--steps", "1", "--warmup", "3", "--cands", "200000", "--train", "2048", "--no-cpu-baseline"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["posterior_path"] == "int8" and d["posterior_path_requested"] == "auto" and d["dtype"] == "f64+i8"
    f = d["fp64_path"]
    assert f["value"] > 0 and f["compared_candidates"] == 200000
    assert f["max_var_diff_rel_to_max_var"] < 1e-10 and f["max_mean_diff_rel_to_max_mean"] < 1e-10
    roof = d["roofline"]
    assert roof["bound"] == "tensor" and roof["fp64_kernel"]["frac"] > 0 and roof["fp64_equivalent"]["frac"] > 0
    assert "csrc/int8_fused.cu" in d["e2e"]["api"] and d["e2e"]["value"] > 0
    r = _run(["--steps", "1", "--warmup", "3", "--cands", "200000", "--train", "2048", "--no-cpu-baseline", "--posterior-path", "fp64",
              "--no-e2e"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["posterior_path"] == "fp64" and d["dtype"] == "f64" and d["fp64_path"] is None


@pytest.mark.gpu
@pytest.mark.parametrize("workload,rows,cols", [("config1-s2", 500, 500), ("gram-so3", 4096, 512), ("gram-t10", 4096, 256),
                                                ("gram-h2", 256, 256), ("gram-spd2", 256, 256), ("gram-spd3", 256, 256),
                                                ("gram-h2-table", 4096, 512), ("gram-so5", 512, 256)])
def test_gram_workloads_run_small(workload, rows, cols):
    """Every Gram workload of bench.py on a small block (also the L2-flushed / graph-replayed timing branch)."""
    r = _run(["--workload", workload, "--cands", str(rows), "--train", str(cols), "--steps", "2", "--warmup", "3", "--no-cpu-baseline"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "kernel_entries_per_sec" and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["config"]["rows"] == rows and d["config"]["cols"] == cols and "l2_policy" in d["config"]
