"""one-line digest of bench.py's JSON line (stdin): evals/s, ms/step, per-kernel ms and roofline fraction"""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        sys.stdout.write(line)
        continue
    d = json.loads(line)
    k = d.get("roofline", {}).get("kernels", {})
    print("%.2f M/s  %.4f ms/step  %s  e2e %.2f M/s" % (
        d["value"] / 1e6, d["ms_per_step"],
        {n: (round(v["ms_per_launch"], 4), round(v["frac_of_fp64_peak"], 3)) for n, v in k.items()},
        d.get("e2e", {}).get("value", 0) / 1e6))
