"""Model-dependent covariance (SPTSZ_lowl_2017-style) on the GPU: the beam modes
go through the binning kernel as extra window rows, chi^2 through the per-walker Cholesky kernel."""
import numpy as np
import pytest

import sptsz_problem as SP
from cosmoslik_b200 import Slik, samplers
from oracle import runtime as ort

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nmode,scale", [(6, 1.0), (8, 5.0), (0, 1.0)])
def test_lnl_matches_the_oracle(nmode, scale):
    o_cls, b_cls = SP.scripts(nmode=nmode, scale=scale)
    oslik = ort.Slik(o_cls())
    slik = Slik(b_cls())
    names = list(slik.get_sampled())
    assert names == list(oslik.get_sampled())
    x0 = np.array([oslik.get_start()[k] for k in names])
    sc = np.array([oslik.get_sampled()[k].scale for k in names])
    X = x0 + sc * np.random.RandomState(0).standard_normal((133, len(names)))
    X[0] = x0
    X[4, names.index("egfs.Aps")] = -0.5
    ref = np.array([oslik.evaluate(*x)[0] for x in X])
    got = slik.evaluate_batch(X)
    f = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), f) and (~f).sum() >= 1
    assert np.max(np.abs(got[f] - ref[f]) / np.abs(ref[f])) < 1e-10


def test_mh_chain_on_the_model_dependent_covariance():
    """generic MH path over the new likelihood: every emitted row's lnl is the oracle's at that point"""
    nch, nsteps = 32, 60
    o_cls, b_cls = SP.scripts(nmode=6, sampler=lambda s: samplers.metropolis_hastings(
        s, num_samples=nch * nsteps, nchains=nch, seed=3, proposal_update=False))
    slik = Slik(b_cls())
    st = slik.sampler.run(slik.evaluate)
    rows, nrows = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
    assert nrows.sum() > nch
    oslik = ort.Slik(o_cls())
    for c in (0, 13, 31):
        for r in rows[c, :nrows[c]][::7]:
            ref = oslik.evaluate(*r[2:])[0]
            assert abs(r[0] - ref) <= 1e-10 * abs(ref)


@pytest.mark.parametrize("tag", ["m6_ab0", "m6_ab1", "m8_ab0", "m8_ab1"])
def test_lnl_matches_the_reference_itself(tag):
    """PINNED: the device path (binning kernel with the beam modes as extra window rows, aberration factor built per
    walker in its Cl stage, per-walker Cholesky) against the values the reference's OWN SPTSZ_lowl.__call__ /
    get_cl_model_and_beam_cov / DlnClDlnl returned for the reference's own data files
    (tests/golden/sptsz.npz, oracle/make_golden_sptsz.py)"""
    script, X, want = SP.reference_golden_script(tag)
    slik = Slik(script())
    got = slik.evaluate_batch(np.ascontiguousarray(X))
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-10


def test_aberration_lnl_matches_the_oracle_and_is_batch_independent():
    o_cls, b_cls = SP.scripts(nmode=6, ab_on=True)
    oslik = ort.Slik(o_cls())
    slik = Slik(b_cls())
    names = list(slik.get_sampled())
    x0 = np.array([oslik.get_start()[k] for k in names])
    sc = np.array([oslik.get_sampled()[k].scale for k in names])
    X = x0 + sc * np.random.RandomState(5).standard_normal((300, len(names)))
    ref = np.array([oslik.evaluate(*x)[0] for x in X[:60]])
    got = slik.evaluate_batch(X)
    f = np.isfinite(ref)
    assert np.max(np.abs(got[:60][f] - ref[f]) / np.abs(ref[f])) < 1e-10
    np.testing.assert_array_equal(got, np.concatenate([slik.evaluate_batch(X[:77]), slik.evaluate_batch(X[77:])]))
