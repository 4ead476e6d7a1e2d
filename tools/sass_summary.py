"""SASS evidence per kernel (no GPU needed): opcode histogram of every kernel matching a pattern in the built
library, plus the lines that prove the data-movement path (UBLKCP / SYNCS = cp.async.bulk + mbarrier, LDGSTS =
cp.async, DMMA = FP64 tensor pipe).   python tools/sass_summary.py <pattern> > profiles/r2_sass_<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cosmoslik_b200", "libcosmoslik_b200.so")
pat = sys.argv[1] if len(sys.argv) > 1 else "k_trigemm_chi2_bulk"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        if pat in cur:
            funcs[cur] = []
        continue
    if cur in funcs and re.search(r"/\*[0-9a-f]{4,6}\*/", ln):
        funcs[cur].append(ln)
for name, lines in funcs.items():
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    ops = collections.Counter()
    for ln in lines:
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
        if m:
            ops[m.group(1)] += 1
    print("# %s" % dem)
    print("# %d instructions; opcode histogram:" % sum(ops.values()))
    print("#   " + ", ".join("%s %d" % kv for kv in ops.most_common(24)))
    keys = ("UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "FENCE", "MEMBAR", "UCGABAR", "BAR.")
    shown = collections.Counter()
    for ln in lines:
        for k in keys:
            if k in ln and shown[k] < 4:
                shown[k] += 1
                print("    " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", ln).strip())
    print("#   totals: " + ", ".join("%s %d" % (k, sum(1 for ln in lines if k in ln)) for k in keys + ("DMMA", "LDS", "DFMA")))
    print()
