"""Sample i.i.d. from the uniform priors (reference: samplers/priors.py:11-47).

Same constructor as the reference (a metropolis_hastings with proposal_update off).  The draws come
from numpy's global generator in the reference's order (sample by sample, parameters in sorted-name
order: ``[uniform(*prior[k]) for k in sampled]``), so ``numpy.random.seed`` reproduces the reference's
points bit for bit; the posterior of all points is evaluated in one batch on the device and every
sample has weight 1.
"""
import numpy as np

from ..core import sample
from ..likelihoods.priors import priors as likelihoods_priors
from .metropolis_hastings import metropolis_hastings
from .utils import extras_of, program_of


class priors(metropolis_hastings):

    def __init__(self, params, output_extra_params=None, **kwargs):
        kwargs.pop("proposal_update", None)
        super().__init__(params, output_extra_params=output_extra_params, proposal_update=False, **kwargs)
        self.uniform_priors = {}
        boxes = likelihoods_priors(params).uniform_priors
        for k in self.sampled:
            pr = [(n, lo, hi) for (n, lo, hi) in boxes if n == k]
            if len(pr) > 2:
                raise Exception("Can't sample prior for parameters with multiple priors at once.")
            elif len(pr) == 0:
                raise Exception("Prior for parameter '%s' is improper." % k)
            self.uniform_priors[k] = pr[0][1:]

    def draw_x(self, n=1):
        lo = np.array([self.uniform_priors[k][0] for k in self.sampled], dtype=float)
        hi = np.array([self.uniform_priors[k][1] for k in self.sampled], dtype=float)
        return np.random.uniform(lo, hi, size=(n, len(lo)))

    def sample(self, lnl):
        prog = program_of(lnl)
        names = list(self.output_extra_params)
        extras = extras_of(lnl, names)
        n = int(self.num_samples / float(max(1, int(self.nchains))))
        block = 65536
        for i0 in range(0, n, block):
            X = self.draw_x(min(block, n - i0))
            vals = prog.lnl_host(X)
            E = extras.values_host(X) if extras is not None else None
            for i, (x, v) in enumerate(zip(X, vals)):
                s = sample(x.copy(), v, 1, dict(zip(names, E[i])) if E is not None else None)
                s.chain = 0
                yield s
