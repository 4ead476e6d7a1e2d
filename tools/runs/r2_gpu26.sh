#!/bin/bash
# the slab substitution kernel at n = 7497 (118 row blocks): bitwise against the two-CTA kernel, and its speed
mkdir -p gpurun_out
timeout 240 python - <<'PY' 2>&1 | tail -8
import sys, json, numpy as np, torch
sys.path.insert(0, 'tools')
import r2_chi2_sweep as S
prog, prob = S.program("unbinned", chi2_mode="trsm")
print("n", prog.n, "n_pad", prog.n_pad, "chi2_mode", prog.meta.get("chi2_mode"))
for W in (256, 8192):
    ms2, c2 = S.time_chi2(prog, W, iters=3, trsm_shape=2)
    ms4, c4 = S.time_chi2(prog, W, iters=3, trsm_shape=4)
    ms4b, c4b = S.time_chi2(prog, W, iters=3)
    n = prog.n
    print(W, "v2 %.3f ms" % ms2, "slab %.3f ms (%.3f of peak)" % (ms4, n * n * W / ms4 / 1e9 / 35.43), "bitwise", bool(torch.equal(c2, c4)), bool(torch.equal(c4, c4b)))
PY
