"""The shapes of tests/test_host_compile_sweep.py through the CUDA path: bin counts around the 64-row tile
and the 8-row group, ell ranges shorter than one 16-column chunk or not a multiple of it, dense / banded /
top-hat windows with an empty row, every binning and chi^2 option, ragged walker counts -- lnL against the
oracle to 1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from cosmoslik_b200 import Slik
from oracle import runtime as ort
from test_host_compile import mspec_scripts
from test_host_compile_sweep import CASES, make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nspec,nbins,lmin,lmax,kind,binning", CASES)
def test_device_lnl_on_edge_shapes(nspec, nbins, lmin, lmax, kind, binning):
    rs = np.random.RandomState(nspec * 1000 + nbins)
    prob = make_problem(rs, nspec, nbins, lmin, lmax, kind)
    o_cls, b_cls = mspec_scripts(prob)
    oslik = ort.Slik(o_cls())
    slik = Slik(b_cls())
    names = list(slik.get_sampled())
    x0 = np.array([oslik.get_start()[k] for k in names])
    W = 131                                                   # one full 128-walker tile + a ragged one
    X = x0 * (1 + 0.05 * rs.standard_normal((W, len(x0))))
    X[7, names.index("Aps")] = -1.0                           # hard prior
    ref = np.array([oslik.evaluate(*x)[0] for x in X[:24]])
    for chi2_mode in ("trsm", "inverse"):
        prog = slik.program(binning=binning, chi2_mode=chi2_mode)
        got = prog.lnl_host(X)
        assert np.isinf(got[7]) and np.isinf(ref[7])
        f = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got[:24]), f)
        assert np.max(np.abs(got[:24][f] - ref[f]) / np.abs(ref[f])) < 1e-10
        # the ragged tail is evaluated like the rest: a walker's value does not depend on its position
        again = prog.lnl_host(X[::-1].copy())[::-1]
        np.testing.assert_array_equal(again, got)
