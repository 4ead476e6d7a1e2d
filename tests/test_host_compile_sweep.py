"""Randomised sweep of the host compilation (no GPU): problem shapes that cross every tile / group
boundary of the device tables -- 1..3 maps, bin counts around the 64-row tile and the 8-row group,
ell ranges that are not multiples of the 16-column chunk, dense / banded / top-hat / partly empty
windows, every binning option -- interpreted by tests/table_interp.py and compared with the oracle."""
import numpy as np
import pytest

import table_interp as TI
from cosmoslik_b200 import Slik
from cosmoslik_b200.program import CompiledProgram
from oracle import runtime as ort
from test_host_compile import mspec_scripts


def make_problem(rs, nspec, nbins, lmin, lmax, kind):
    maps = ["T090", "T150", "T220"][:nspec]
    keys = [(a, b) for i, a in enumerate(maps) for b in maps[i:]]
    nl = lmax - lmin
    ell = np.arange(lmax)
    cmb = {"cl_TT": 4000.0 * np.exp(-(ell / 900.0) ** 1.1) + 20.0}
    use, windows, signal = {}, {}, {}
    for k in keys:
        use[k] = (lmin, lmax)
        if kind == "tophat":
            edges = np.linspace(0, nl, nbins + 1).astype(int)
            w = np.zeros((nbins, nl))
            for b in range(nbins):
                if edges[b + 1] > edges[b]:
                    w[b, edges[b]:edges[b + 1]] = 1.0 / (edges[b + 1] - edges[b])
        else:
            centers = np.linspace(0, nl, nbins + 2)[1:-1]
            w = np.exp(-0.5 * ((np.arange(nl)[None, :] - centers[:, None]) / max(nl / nbins / 2.0, 1.0)) ** 2)
            w /= w.sum(1, keepdims=True)
            if kind == "banded":
                w[w < 1e-6] = 0.0
            else:
                w += 1e-10
        if kind != "dense" and nbins > 2:
            w[rs.randint(nbins)] = 0.0                   # a window row with no support at all
        windows[k] = w
        signal[k] = w @ cmb["cl_TT"][lmin:lmax] * (1 + 0.02 * rs.standard_normal(nbins)) + 30
    n = nbins * len(keys)
    d = np.concatenate([signal[k] for k in sorted(keys)])
    A = rs.standard_normal((n, n)) * 0.05
    cov = np.diag((0.03 * d) ** 2) + (A @ A.T) * np.outer(0.03 * d, 0.03 * d)
    return dict(maps=maps, keys=keys, use=use, windows=windows, signal=signal, cov=cov, cmb=cmb)


CASES = [  # nspec, nbins, lmin, lmax, window kind, binning option
    (1, 1, 2, 19, "dense", "auto"),
    (1, 7, 30, 333, "banded", "auto"),
    (1, 8, 30, 334, "tophat", "auto"),
    (1, 9, 50, 700, "tophat", "segsum"),
    (2, 33, 50, 601, "dense", "dmma"),
    (2, 63, 2, 515, "banded", "dmma"),
    (1, 64, 2, 1000, "tophat", "dmma"),
    (1, 65, 10, 1001, "dense", "auto"),
    (2, 70, 10, 999, "tophat", "auto"),
    (3, 20, 50, 700, "banded", "segsum"),
    (3, 5, 100, 164, "dense", "segsum"),
    (2, 129, 2, 2500, "tophat", "auto"),
]


@pytest.mark.parametrize("nspec,nbins,lmin,lmax,kind,binning", CASES)
def test_tables_reproduce_the_oracle(nspec, nbins, lmin, lmax, kind, binning):
    rs = np.random.RandomState(nspec * 1000 + nbins)
    prob = make_problem(rs, nspec, nbins, lmin, lmax, kind)
    o_cls, b_cls = mspec_scripts(prob)
    oslik = ort.Slik(o_cls())
    slik = Slik(b_cls())
    cp = CompiledProgram(list(slik.get_sampled()), slik.params.priors, slik.trace(), binning=binning)
    nkeys = nspec * (nspec + 1) // 2
    assert cp.meta["n"] == nbins * nkeys and cp.meta["n_pad"] % 64 == 0 and cp.meta["n_pad"] >= cp.meta["n"]
    if binning == "dmma":
        assert cp.meta["nsegclass"] == 0 and cp.meta["nclass"] >= 1
    if binning == "segsum":
        assert cp.meta["nclass"] == 0 and cp.meta["nsegclass"] >= 1
    x0 = np.array([oslik.get_start()[k] for k in cp.names])
    for i in range(3):
        x = x0 * (1 + 0.05 * rs.standard_normal(len(x0)))
        ref = oslik.evaluate(*x)[0]
        got = TI.lnl(cp, x)
        assert np.isfinite(ref) and abs(got - ref) <= 1e-10 * abs(ref), (got, ref)
