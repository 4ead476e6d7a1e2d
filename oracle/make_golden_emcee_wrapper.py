"""Generate tests/golden/emcee_wrapper.npz + ref_emcee.chain by running the UNMODIFIED reference wrapper
``cosmoslik_plugins/samplers/emcee.py`` (through the reference's ``Slik.sample``) over the stand-in of the absent
``emcee`` package (``oracle/emcee_standin.py``: read its header for what that does and does not pin -- the wrapper's
own bookkeeping yes, the stretch move no).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is mounted:
    python -m oracle.make_golden_emcee_wrapper

The script is the two-parameter one of oracle/make_golden_extras.py (a hard prior ``b >= -4``, derived quantities
stored on the script in ``__call__``), so the file also carries ``output_extra_params`` columns.  Stored: the initial
positions the wrapper drew (numpy's global generator, seeded here), the injected streams, every yielded sample in
order (the reference's ``sample(-l, x, weight)`` fills the positional fields (x, lnl, weight) with (lnl, x, weight):
stored as they come), the number of lnprob evaluations, and the pickled chain file.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import emcee_standin, refshim  # noqa: E402
from oracle.make_golden import OUT  # noqa: E402

K, NSTEPS, FREQ = 24, 9, 3
EXTRAS = ["r2", "ratio"]


def run(out_file, k=K, nsteps=NSTEPS, freq=FREQ, stream_seed=20260401, scatter_seed=77):
    """the reference wrapper on the derived-quantities script -> dict of what it drew, was fed and yielded"""
    R = refshim.load()
    core = R.core
    emcee_standin.install()
    wrapper = importlib.import_module("cosmoslik_plugins.samplers.emcee")

    class derived(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = core.param(start=0, scale=1)
            self.b = core.param(start=0.5, scale=1, min=-4)
            self.sampler = wrapper.emcee(self, num_samples=k * nsteps, nwalkers=k, output_file=out_file,
                                         output_freq=freq, output_extra_params=EXTRAS)

        def __call__(self):
            self.r2 = self.a ** 2 + self.b ** 2
            self.ratio = (self.a - 1.5) / (2 + self.b ** 2)
            self.fixed = 3.25
            return self.r2 / 2

    slik = core.Slik(derived())
    names = list(slik.get_sampled())
    rs = np.random.RandomState(stream_seed)
    streams = [(rs.uniform(size=(2, k // 2)), rs.randint(k // 2, size=(2, k // 2)), rs.uniform(size=(2, k // 2)))
               for _ in range(nsteps)]
    emcee_standin.STREAMS = lambda i: streams[i]
    emcee_standin.EVALS[0] = 0
    np.random.seed(scatter_seed)                     # the wrapper draws its initial scatter from the global generator
    state = np.random.get_state()
    pos0 = np.random.multivariate_normal([v.start for v in slik.sampler.sampled.values()], slik.sampler.cov_est, size=k)
    np.random.set_state(state)                       # ... so the wrapper draws exactly pos0
    first, second, third = [], [], []
    for s in slik.sample():
        first.append(np.atleast_1d(np.asarray(s.x, dtype=float)))     # what lands in the fields, as the reference fills them
        second.append(np.atleast_1d(np.asarray(s.lnl, dtype=float)))
        third.append(float(s.weight))
    return dict(names=np.array(names), pos0=pos0, cov_est=np.asarray(slik.sampler.cov_est, dtype=float),
                u_z=np.array([s[0] for s in streams]), j=np.array([s[1] for s in streams]),
                u_acc=np.array([s[2] for s in streams]), field_x=np.array(first), field_lnl=np.array(second),
                field_weight=np.array(third), evals=np.array(emcee_standin.EVALS[0]),
                nwalkers=np.array(k), nsteps=np.array(nsteps), output_freq=np.array(freq))


def main():
    out_file = os.path.join(OUT, "ref_emcee.chain")
    g = run(out_file)
    np.savez_compressed(os.path.join(OUT, "emcee_wrapper.npz"), **g)
    print("yielded %d samples from %d walkers x %d steps; %d lnprob evaluations; file %s (%d bytes)"
          % (len(g["field_weight"]), K, NSTEPS, int(g["evals"]), out_file, os.path.getsize(out_file)))


if __name__ == "__main__":
    main()
