"""A stand-in for the third-party package ``emcee`` (2.x API; not in the reference tree, not in its requirements,
not installable offline) -- just enough of ``EnsembleSampler`` for the reference's WRAPPER
(``cosmoslik_plugins/samplers/emcee.py``) to run its own code here.

TEST INFRASTRUCTURE ONLY (used by ``oracle/make_golden_emcee_wrapper.py`` and ``tests/test_oracle_vs_reference.py``).
Nothing in the product imports it.

What this buys: everything the reference itself wrote around the move -- the initial scatter
(``multivariate_normal`` from the covariance estimate, emcee.py:49-53), the sign convention of ``lnprob`` (:56-58),
the loop of ``run_mcmc(pos, 1)`` calls, the weights of walkers that did not move, which rows are yielded and with
which lnl (that of the NEXT position beside the OLD one, :66-74), the last-step rule, the per-walker file records of
``output_freq`` rows with their extra columns (:36-45, :77-84) -- is executed from the reference's own file and
stored as golden data.  What it cannot buy: the stretch move inside ``run_mcmc`` is NOT emcee's code; it is
``oracle/samplers.py::ensemble_step`` (Goodman & Weare 2010 as emcee 2.x ``_propose_stretch`` states it) fed with
injected random streams.  The move stays PARITY UNPINNED; the wrapper does not.

emcee 2.x behaviour relied on by the wrapper and stated here: ``lnprobfn(x)`` returns ``(lnprob, blob)``;
``run_mcmc(pos0, N)`` with no ``lnprob0`` evaluates every walker at ``pos0`` first, then runs N iterations, and returns
``(pos, lnprob, rstate, blobs)`` with the blobs of the final positions.
"""
import sys
import types

import numpy as np

from . import samplers as osamp

STREAMS = None          # callable step -> (u_z, j, u_acc), each [2, nwalkers / 2]; set by the caller before sampling
EVALS = [0]             # lnprob evaluations made (the wrapper's per-call re-evaluation is counted too)


class EnsembleSampler(object):

    def __init__(self, nwalkers, dim, lnprobfn, a=2.0, pool=None, **kw):
        self.k, self.dim, self.fn, self.a = int(nwalkers), int(dim), lnprobfn, float(a)
        self.iterations = 0

    def _eval(self, P):
        """lnprob of the rows of P; blobs remembered by position"""
        out = np.empty(len(P))
        for i, x in enumerate(np.atleast_2d(P)):
            lp, blob = self.fn(x)
            EVALS[0] += 1
            self._blob[x.tobytes()] = blob
            out[i] = lp
        return out

    def run_mcmc(self, pos0, N, **kw):
        self._blob = {}
        pos = np.array(pos0, dtype=float)
        lnprob = self._eval(pos)
        for _ in range(int(N)):
            u_z, j, u_acc = STREAMS(self.iterations)
            pos, lnprob, _ = osamp.ensemble_step(pos, lnprob, self._eval, u_z, j, u_acc, self.a)
            self.iterations += 1
        return pos, lnprob, None, [self._blob[x.tobytes()] for x in pos]


def install():
    mod = types.ModuleType("emcee")
    mod.EnsembleSampler = EnsembleSampler
    sys.modules["emcee"] = mod
    return sys.modules[__name__]
