"""Generate tests/golden/mh_rejected.npz by running the UNMODIFIED reference sampler with
``yield_rejected=True`` (metropolis_hastings.py:160) and injected streams.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is mounted:
    python -m oracle.make_golden_rejected
Two scripts: the 2-parameter quickstart (no priors) and the spt_lowl 's12' script (hard-prior rejections give
weight-0 rows with lnl = +inf), both with max_weight = 4 so that flush rows follow rejected rows.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402
from oracle.make_golden import OUT, pack_rows, run_injected_mh, synthetic_tt  # noqa: E402


def main():
    R = refshim.load()
    core = R.core
    rng = np.random.RandomState(20260202)
    out = {}

    class quick(core.SlikPlugin):   # doc/quickstart.md:26-43
        def __init__(self):
            super().__init__()
            self.a = core.param(start=0, scale=1)
            self.b = core.param(start=0, scale=1)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    n = 400
    z, u = rng.standard_normal((n, 2)), rng.uniform(size=n)
    names, rows, _ = run_injected_mh(R, quick, n, z, u, yield_rejected=True, max_weight=4)
    lnl, w, x = pack_rows(rows)
    out.update(quick_z=z, quick_u=u, quick_lnl=lnl, quick_weight=w, quick_x=x)
    print("quick:", len(rows), "rows,", int((w == 0).sum()), "with weight 0")

    tt = synthetic_tt()

    class spt(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.spt = R.spt_lowl.spt_lowl("s12")
            self.cmb = {"TT": tt}

        def __call__(self):
            return self.spt(self.cmb)

    n = 300
    z, u = rng.standard_normal((n, 4)), rng.uniform(size=n)
    names, rows, _ = run_injected_mh(R, spt, n, z, u, yield_rejected=True, max_weight=4)
    lnl, w, x = pack_rows(rows)
    out.update(spt_z=z, spt_u=u, spt_lnl=lnl, spt_weight=w, spt_x=x, spt_names=np.array(names))
    print("spt:", len(rows), "rows,", int((w == 0).sum()), "with weight 0,", int(np.isinf(lnl).sum()), "with lnl = inf")
    np.savez_compressed(os.path.join(OUT, "mh_rejected.npz"), **out)


if __name__ == "__main__":
    main()
