"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py gpurun_out/launches.csv "command line that was profiled" > profiles/rNN_launches_..._summary.txt"""
import csv
import sys
from collections import defaultdict


def main(path, command):
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        rows.append((r[ik], v * scale))
    tot = defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        k = k.split("(")[0][:70]
        tot[k][0] += 1
        tot[k][1] += ms
    total = sum(ms for _, ms in rows)
    print("ncu --metrics gpu__time_duration.sum --clock-control none: %s" % command)
    print("kernel | launches | total ms | share of the captured launches\n---|---|---|---")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%s | %d | %.3f | %.2f %%" % (k, n, ms, 100.0 * ms / total))
    print("all captured launches: %d, %.1f ms" % (len(rows), total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
