"""A/B sweep of the DMMA binning kernels on the GPU box: k_bin_residual (round 1) vs k_bin_residual_ws with ell
splits 1 / 2 / 4, at several launch sizes.  `python tools/r2_bin_sweep.py [spt|spt_lowl|planck_dmma]`."""
import ctypes as C
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cosmoslik_b200 as cb                      # noqa: E402
from cosmoslik_b200 import _lib as L, synthetic  # noqa: E402

lib = L.lib()
ns = types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl, egfs=cb.models.egfs)


def set_opts(**kw):
    for k in ("pipe", "bin_variant", "bin_split"):
        L.check(lib.csl_set_option(k.encode(), int(kw.get(k, -1))))


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "spt"
    peak = json.load(open(os.path.join(ROOT, "profiles", "measured_fp64_peak.json")))["fp64_tflops_sustained"]
    if kind == "spt_lowl":
        import b200_scripts as B
        data = np.load(os.path.join(ROOT, "tests", "golden", "spt_lowl_data.npz"))
        slik = cb.Slik(B.spt_script(data, "s12")())
        prog = slik.program()
        x0 = np.array([1.0, 20.0, 5.0, 5.0]); sc = np.array([0.026, 3, 3, 3]) * 0.2
        sizes = [4096, 16384, 65536]
    else:
        prob = synthetic.spt_multifreq_problem(0) if kind == "spt" else synthetic.planck_lite_problem(0)
        slik = cb.Slik(synthetic.build_script(ns, prob)())
        prog = slik.program(**(dict(binning="dmma") if kind == "planck_dmma" else {}))
        names = sorted(prob["truth"])
        x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
        sizes = [8192, 16384, 32768, 65536]
    fl = prog.flops
    alg = fl["bin_nnz"] + fl["build_ops"]
    st = L.stream_ptr()
    P = C.byref(prog.struct)
    for W in sizes:
        X = torch.from_numpy(x0 + sc * np.random.RandomState(1).standard_normal((W, len(x0)))).cuda()
        ws = prog.workspace(W)
        L.check(lib.csl_eval_prepare(P, W, X.data_ptr(), ws["coef"].data_ptr(), W, ws["prior"].data_ptr(), st))
        ref = None
        for v in (dict(bin_variant=0, bin_split=1), dict(bin_variant=0, bin_split=2), dict(bin_variant=0, bin_split=4),
                  dict(bin_variant=0), dict(bin_variant=2, bin_split=1), dict(bin_variant=2)):
            set_opts(**v)
            call = lambda: L.check(lib.csl_bin_residual(P, W, ws["coef"].data_ptr(), W, ws["R"].data_ptr(), W, st))
            try:
                for _ in range(3):
                    call()
                torch.cuda.synchronize()
                ts = []
                for _ in range(10):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); call(); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
            except Exception as e:
                print(json.dumps(dict(kind=kind, W=W, variant=v, error=str(e))), flush=True); continue
            Rn = ws["R"][:prog.n].clone()
            if ref is None:
                ref = Rn
            ms = float(np.median(ts))
            print(json.dumps(dict(kind=kind, W=W, variant=v, ms=round(ms, 5), tflops=round(alg * W / ms / 1e9, 3),
                                  frac=round(alg * W / ms / 1e9 / peak, 4), bitwise=bool(torch.equal(Rn, ref)),
                                  maxabs=float((Rn - ref).abs().max()), scale=float(ref.abs().max()))), flush=True)
    set_opts()


if __name__ == "__main__":
    main()
