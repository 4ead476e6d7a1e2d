"""Generate tests/golden/mh_extras.npz + ref_mh_extras.chain by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (reference tree mounted):
    python -m oracle.make_golden_extras
A separate script so that the fixtures of oracle/make_golden.py keep their bytes.

What is produced: the reference's metropolis_hastings with ``output_extra_params`` (derived quantities
the script stores on itself in ``__call__``; the sampler writes ``s.extra[k]`` next to every sample,
metropolis_hastings.py:103, 183-186, 221, 228) on a two-parameter script with INJECTED streams:
the yielded rows, the extras of every yielded sample, and the chain file the reference wrote.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402
from oracle.make_golden import OUT  # noqa: E402

EXTRAS = ["r2", "ratio", "fixed", "a"]


def main():
    R = refshim.load()
    core, mh = R.core, R.mh
    from oracle.runtime import proposal_factor
    rng = np.random.RandomState(20260202)
    n = 1200
    z = rng.standard_normal((n, 2))
    u = rng.uniform(size=n)

    class derived(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = core.param(start=0, scale=1)
            self.b = core.param(start=0.5, scale=1, min=-4)

        def __call__(self):
            self.r2 = self.a ** 2 + self.b ** 2
            self.ratio = (self.a - 1.5) / (2 + self.b ** 2)
            self.fixed = 3.25
            return self.r2 / 2

    zs, us = iter(z), iter(u)

    class injected(mh.metropolis_hastings):
        def draw_x(self, cur_x):
            F = proposal_factor(self.cov_est, len(self.x0), self.proposal_scale)
            return np.asarray(cur_x) + np.dot(next(zs), F)

    chainfile = os.path.join(OUT, "ref_mh_extras.chain")
    old_uniform = mh.uniform
    mh.uniform = lambda: next(us)
    try:
        s = derived()
        s.sampler = injected(s, num_samples=n, reseed=False, output_file=chainfile, output_extra_params=EXTRAS,
                             mpi_comm_freq=50, max_weight=5)
        slik = core.Slik(s)
        rows, extras = [], []
        for r in slik.sample():
            rows.append((r.lnl, r.weight, np.array(r.x, dtype=float)))
            extras.append([float(r.extra[k]) for k in EXTRAS])
    finally:
        mh.uniform = old_uniform
    names = list(slik.get_sampled().keys())
    loaded = R.chains.load_chain(chainfile)[0]
    out = dict(z=z, u=u, names=np.array(names), extra_names=np.array(EXTRAS),
               lnl=np.array([r[0] for r in rows]), weight=np.array([r[1] for r in rows], dtype=float),
               x=np.array([r[2] for r in rows]), extras=np.array(extras))
    for k in ["lnl", "weight"] + names + EXTRAS:
        out["file_" + k] = np.asarray(loaded[k], dtype=float)
    np.savez_compressed(os.path.join(OUT, "mh_extras.npz"), **out)
    print("mh_extras rows", len(rows), "file rows", len(loaded["lnl"]), "file keys", list(loaded.keys()))


if __name__ == "__main__":
    main()
