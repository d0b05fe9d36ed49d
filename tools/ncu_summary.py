"""Turns an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_xxx.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_elapsed.avg.per_second"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path):
    rows = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=== %s" % r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("  %-88s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = []
        for h, v in zip(hdr, r):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  top stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]))
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv"]))))
    data = [r for r in src[2:] if len(r) > 5 and r[0].startswith("0x")]
    if data:
        tot = sum(int(r[2]) for r in data) or 1
        c, ex = Counter(), Counter()
        for r in data:
            toks = r[1].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            c[op] += int(r[2])
            ex[op] += int(r[5])
        print("  stall samples by opcode (source page, all launches in the report):")
        for op, v in c.most_common(8):
            print("    %-22s %5.1f%%   warp-instructions executed %d" % (op, 100.0 * v / tot, ex[op]))


if __name__ == "__main__":
    main(sys.argv[1])
