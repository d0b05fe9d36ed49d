"""Tabulates the JSON lines written by tools/bench_all.sh.
usage: python tools/bench_summary.py <tag> >> profiles/<round>_bench_summary.md"""
import glob
import json
import os
import sys


def last_json(path):
    for ln in reversed(open(path).read().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return None


def main(tag):
    rows = []
    for path in sorted(glob.glob("gpurun_out/bench_%s_*.json" % tag)):
        d = last_json(path)
        if d is None:
            rows.append((os.path.basename(path), "no JSON line", "", "", "", ""))
            continue
        roof = d.get("roofline") or {}
        cpu = d.get("cpu_baseline") or {}
        e2e = d.get("e2e") or {}
        rows.append((os.path.basename(path)[len("bench_%s_" % tag):-5], d["config"]["workload"][:78], "%.4g" % d["value"], d["unit"],
                     ("%.3f (%s)" % (roof["frac"], roof["bound"])) if roof.get("frac") is not None else "-",
                     ("%.4g" % e2e["value"]) if e2e.get("value") else "-",
                     ("%.4g %s, %s cores" % (cpu["value"], cpu.get("unit", ""), cpu.get("cores"))) if cpu.get("value") else "-"))
    print("| workload | config | value | unit | roofline frac (bound) | e2e | CPU baseline (oracle port) |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        print("| " + " | ".join(r) + " |")
    for name in ("config2-bo-s3", "tr-step"):
        path = "gpurun_out/bench_%s_%s.json" % (tag, name)
        if os.path.exists(path):
            d = last_json(path)
            if d and "latencies_us" in d:
                print("\n%s latencies (us per call):\n" % name)
                print("```")
                print(json.dumps(d["latencies_us"], indent=1))
                print("```")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r")
