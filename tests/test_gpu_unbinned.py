"""BASELINE config 5 shape at reduced size: UNBINNED multi-spectrum likelihood (identity windows, one
column per data point), dense covariance whose factor exceeds what the explicit-inverse path takes in
extended precision, MH chains with the Gelman-Rubin reduction.  Exercises the blocked triangular solve
with many row blocks and the segmented binning with one-column bins."""
import numpy as np
import pytest

from namespaces import ORACLE, PRODUCT
from cosmoslik_b200 import Slik, samplers
from oracle import runtime as ort, samplers as osamp

pytestmark = pytest.mark.gpu


def unbinned_problem(lmax=450, seed=0):
    rs = np.random.RandomState(seed)
    ell = np.arange(lmax, dtype=float)
    tt = 3000.0 * np.exp(-(ell / 300.0) ** 1.1) * (1 + 0.3 * np.cos(ell / 40.0)) + 30
    te = 60.0 * np.exp(-(ell / 300.0)) * np.sin(ell / 40.0)
    ee = 20.0 * np.exp(-(ell / 350.0)) * (1 - 0.5 * np.cos(ell / 40.0)) + 0.1
    T, E = "T", "E"
    use = {(T, T): (2, lmax), (T, E): (2, lmax), (E, E): (2, lmax)}
    nl = lmax - 2
    windows = {k: np.eye(nl) for k in use}
    n = 3 * nl
    truth = dict(A=1.0, n_s=0.0, A_ps=40.0, cal=1.0)
    m = np.concatenate([truth["A"] * ee[2:lmax], truth["A"] * te[2:lmax],
                        truth["A"] * tt[2:lmax] + truth["A_ps"] * (ell[2:lmax] / 3000.0) ** 2])  # sorted keys: EE, TE, TT
    sig = 0.05 * np.abs(m) + 1.0
    band = np.eye(n) + 0.2 * (np.eye(n, k=1) + np.eye(n, k=-1)) + 0.05 * (np.eye(n, k=nl) + np.eye(n, k=-nl))
    u = rs.standard_normal((n, 3)) * 0.1
    cov = (band + u @ u.T) * np.outer(sig, sig)        # banded + low rank, dense storage
    d = m + np.linalg.cholesky(cov) @ rs.standard_normal(n)
    signal = {(E, E): d[:nl], (T, E): d[nl:2 * nl], (T, T): d[2 * nl:]}
    return dict(use=use, windows=windows, signal=signal, cov=cov, tt=tt, te=te, ee=ee, lmax=lmax, n=n)


def script_for(ns, prob, sampler=None):
    lmax = prob["lmax"]
    ell = np.arange(lmax)

    class fg(ns.egfs):
        def get_egfs(self, spectra, lmax, spec, A_ps, **_):
            if spec[0] == "T" and spec[1] == "T":
                return {"ps": A_ps * (np.arange(lmax) / 3000.0) ** 2}
            return {}

    class script(ns.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.A = ns.param(start=1.0, scale=0.003)
            self.n_s = ns.param(start=0.0, scale=0.003)
            self.A_ps = ns.param(start=40.0, scale=4.0, min=0)
            self.cal = ns.param(start=1.0, scale=0.002, gaussian_prior=(1, 0.01))
            self.fg = fg()
            self.like = ns.mspec_lnl(prob["use"], prob["windows"], prob["signal"], prob["cov"])
            self.sampler = sampler(self) if sampler else None

        def __call__(self):
            tilt = (np.maximum(ell, 1.0) / 100.0) ** self.n_s
            cmb = {"cl_TT": self.A * (tilt * prob["tt"]), "cl_TE": self.A * (tilt * prob["te"]),
                   "cl_EE": self.A * (tilt * prob["ee"])}
            self.like.cal = {"T": self.cal, "E": self.cal}
            return self.like(cmb, self.fg(A_ps=self.A_ps))

    return script


@pytest.mark.parametrize("mode", ["trsm", "auto"])
def test_unbinned_lnl_matches_the_oracle(mode):
    prob = unbinned_problem()
    assert prob["n"] == 1344
    oslik = ort.Slik(script_for(ORACLE, prob)())
    slik = Slik(script_for(PRODUCT, prob)())
    prog = slik.program(chi2_mode=mode)
    assert prog.meta["n_pad"] == 1344 and prog.meta["nsegclass"] >= 1 and prog.meta["nclass"] == 0
    names = list(slik.get_sampled())
    x0 = np.array([oslik.get_start()[k] for k in names])
    sc = np.array([oslik.get_sampled()[k].scale for k in names])
    X = x0 + sc * np.random.RandomState(1).standard_normal((300, len(names))) * 2
    X[3, names.index("A_ps")] = -1
    ref = np.array([oslik.evaluate(*x)[0] for x in X])
    got = slik.evaluate_batch(X)
    f = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), f)
    assert np.max(np.abs(got[f] - ref[f]) / np.abs(ref[f])) < 1e-10


def test_unbinned_mh_chains_with_gelman_rubin():
    prob = unbinned_problem(lmax=200)
    nch, nsteps = 64, 400
    slik = Slik(script_for(PRODUCT, prob, lambda s: samplers.metropolis_hastings(
        s, num_samples=nch * nsteps, nchains=nch, seed=9, proposal_update_start=100, mpi_comm_freq=50))())
    st = slik.sampler.run(slik.evaluate)
    assert len(slik.sampler.adaptations) >= 4
    rows, nrows = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
    assert nrows.min() > 20
    burn = 10
    mom = [osamp.chain_moments(rows[c, burn:nrows[c], 2:], rows[c, burn:nrows[c], 1]) for c in range(nch)]
    ref = osamp.gelman_rubin([m[0] for m in mom], [m[1] for m in mom], [m[2] for m in mom])
    np.testing.assert_allclose(slik.sampler.gelman_rubin(burn_rows=burn), ref, rtol=1e-10)
    # every emitted row's lnl is the oracle's lnl at that position
    oslik = ort.Slik(script_for(ORACLE, prob)())
    for c in (0, 17, 63):
        for r in rows[c, :nrows[c]][::25]:
            ref_l = oslik.evaluate(*r[2:])[0]
            assert abs(r[0] - ref_l) <= 1e-10 * abs(ref_l)
