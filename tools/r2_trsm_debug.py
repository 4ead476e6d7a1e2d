"""where does the bulk-copy blocked solve go wrong?  repeated runs, first differing row block / column pattern"""
import ctypes as C, json, os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmoslik_b200 as cb
from cosmoslik_b200 import _lib as L, synthetic
lib = L.lib()
ns = types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl, egfs=cb.models.egfs)
prob = synthetic.planck_lite_problem()
prog = cb.Slik(synthetic.build_script(ns, prob)()).program(chi2_mode="trsm")
def opts(**kw):
    for k in ("pipe", "trsm_shape"):
        L.check(lib.csl_set_option(k.encode(), int(kw.get(k, -1))))
def run(W, **kw):
    opts(**kw)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    R = torch.randn((prog.n_pad, W), dtype=torch.float64, device="cuda", generator=g); R[prog.n:] = 0
    chi2 = torch.empty(W, dtype=torch.float64, device="cuda")
    L.check(lib.csl_chi2(C.byref(prog.struct), W, R.data_ptr(), W, chi2.data_ptr(), None, L.stream_ptr()))
    torch.cuda.synchronize()
    return R, chi2
for W in (32768, 65536):
    Rref, cref = run(W, pipe=0)
    for shape in [1] * 24 + [17] * 24 + [3] * 8 + [33] * 6:
        R, c = run(W, pipe=1, trsm_shape=shape)
        bad = (R != Rref)
        nbad = int(bad.sum())
        rec = dict(W=W, shape=shape, chi2_equal=bool(torch.equal(c, cref)), nbad=nbad)
        if nbad:
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            first = int(rows[0])
            fc = bad[first].nonzero().flatten()
            rec.update(first_row=first, first_block=first // 64, nrows_bad=len(rows), ncols_bad=len(cols),
                       col_tiles=sorted(set((cols // 128).tolist()))[:12], n_col_tiles=len(set((cols // 128).tolist())),
                       first_row_cols=fc[:16].tolist(), first_row_ncols=len(fc),
                       first_row_col_in_tile=sorted(set((fc % 128).tolist()))[:40],
                       maxrel=float(((R - Rref).abs().max() / Rref.abs().max())))
        print(json.dumps(rec), flush=True)
opts()
