"""Print the headline numbers of a bench.py JSON line (default workload incl. its Gram leg)."""
import json
import sys

d = [json.loads(ln) for ln in open(sys.argv[1]) if ln.startswith("{")][-1]
print("value", d["value"], "e2e", (d.get("e2e") or {}).get("value"), "roofline.frac", (d.get("roofline") or {}).get("frac"))
print("gram_cool_down", d.get("gram_cool_down"))
for k, v in (d.get("gram") or {}).items():
    if "error" in v:
        print(k, "ERROR", v["error"])
        continue
    r = v["roofline"]
    print(k, "%.3e %s" % (v["value"], v["unit"]), "ms/step %.4f" % v["ms_per_step"], "steps", v["steps"], "timed %.2fs" % v["timed_seconds"],
          "clocks", v["clocks"], "frac", r.get("frac"), "fp64_exec", (r.get("fp64_executed") or {}).get("frac"),
          "exec_frac", r.get("executed_frac"), "survey", r.get("frac_survey_convention"), "cpu", (v["cpu_baseline"] or {}).get("value"))
