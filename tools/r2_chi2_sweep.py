"""A/B sweep of the chi^2 kernels (run on the GPU box): cp.async pipelines vs bulk copies + mbarrier, tile shapes,
stages, launch sizes.  Every variant must give bitwise the chi^2 of the round-1 kernel.  Prints one JSON line per
variant; `python tools/r2_chi2_sweep.py [planck|spt|unbinned]`."""
import ctypes as C
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmoslik_b200 as cb                      # noqa: E402
from cosmoslik_b200 import _lib as L, synthetic  # noqa: E402

lib = L.lib()
ns = types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl, egfs=cb.models.egfs)


def program(kind, chi2_mode="auto"):
    prob = {"planck": synthetic.planck_lite_problem, "spt": synthetic.spt_multifreq_problem,
            "unbinned": synthetic.unbinned_problem}[kind]()
    slik = cb.Slik(synthetic.build_script(ns, prob)())
    return slik.program(chi2_mode=chi2_mode), prob


def set_opts(**kw):
    for k in ("pipe", "trigemm_rows", "trigemm_stages", "trigemm_group", "trsm_shape", "slab_warps"):
        L.check(lib.csl_set_option(k.encode(), int(kw.get(k, -1))))


def time_chi2(prog, W, iters=20, **opts):
    set_opts(**opts)
    npad = prog.n_pad
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    R0 = torch.randn((npad, W), dtype=torch.float64, device="cuda", generator=g)
    R0[prog.n:] = 0
    R = R0.clone()
    chi2 = torch.empty(W, dtype=torch.float64, device="cuda")
    part = torch.zeros((npad // 32, W), dtype=torch.float64, device="cuda")
    st = L.stream_ptr()
    call = lambda: L.check(lib.csl_chi2(C.byref(prog.struct), W, R.data_ptr(), W, chi2.data_ptr(), part.data_ptr(), st))
    inplace = not prog.meta.get("chi2_mode")
    for _ in range(3):
        if inplace:
            R.copy_(R0)
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if inplace:
            R.copy_(R0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), chi2.clone()


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "planck"
    peak = json.load(open(os.path.join(ROOT, "profiles", "measured_fp64_peak.json")))["fp64_tflops_sustained"]
    sizes = {"planck": [4096, 8192, 16384, 32768, 65536], "spt": [8192, 16384, 65536], "unbinned": [8192]}[kind]
    modes = ("auto", "trsm") if kind != "unbinned" else ("auto",)
    if len(sys.argv) > 2:
        modes = (sys.argv[2],)
    for mode in modes:
        prog, prob = program(kind, chi2_mode=mode)
        n = prog.n
        for W in sizes:
            ref = None
            variants = [dict(pipe=0)]
            if prog.meta.get("chi2_mode"):
                variants += [dict(pipe=0, trigemm_rows=32), dict(pipe=0, trigemm_rows=64),
                             dict(pipe=1), dict(pipe=1, trigemm_rows=64, trigemm_stages=3),
                             dict(pipe=1, trigemm_rows=64, trigemm_stages=4),
                             dict(pipe=1, trigemm_rows=32, trigemm_stages=3), dict(pipe=1, trigemm_rows=32, trigemm_stages=4)]
            else:
                variants = [dict(trsm_shape=0), dict(trsm_shape=2), dict(trsm_shape=4, slab_warps=14)] + [dict(trsm_shape=4)] * int(os.environ.get('SLAB_REPEATS', '2'))
            for v in variants:
                try:
                    ms, chi2 = time_chi2(prog, W, **v)
                except Exception as e:
                    print(json.dumps(dict(kind=kind, mode=mode, W=W, variant=v, error=str(e)))); continue
                if ref is None:
                    ref = chi2
                same = bool(torch.equal(chi2, ref))
                rel = float(((chi2 - ref).abs() / ref.abs()).max())
                print(json.dumps(dict(kind=kind, mode=mode, n=n, W=W, variant=v, ms=round(ms, 5),
                                      tflops=round(n * n * W / ms / 1e9, 3), frac=round(n * n * W / ms / 1e9 / peak, 4),
                                      bitwise=same, maxrel=rel)), flush=True)
    set_opts()


if __name__ == "__main__":
    main()
