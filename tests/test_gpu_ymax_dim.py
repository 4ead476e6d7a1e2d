"""Maximum parameter count (CSL_MAX_DIM = 64 sampled parameters): the warp-per-walker sampler kernels
hold a row as two registers per lane, the fused Gaussian kernel keeps 64-dimensional state in registers;
MH (fused and per-step) and the stretch move against the oracle with injected streams, and the first
count past the limit is refused on the host."""
import numpy as np
import pytest

import b200_scripts as B
import oracle_scripts as S
from cosmoslik_b200 import Slik, samplers
from cosmoslik_b200 import _lib as L
from oracle import runtime as ort, samplers as osamp

pytestmark = pytest.mark.gpu
ND = 64


def _cov():
    A = np.random.RandomState(64).standard_normal((ND, ND))
    return A @ A.T / ND + np.eye(ND)


@pytest.mark.parametrize("mode", ["delta", "z"])
def test_mh_at_the_maximum_dimension(mode):
    """per-step path (the whole-run fused kernel serves ndim <= 32 only and is not selected here): injected
    offsets -> chains bit for bit; injected normals -> the 64-term z.F product runs on the device (positions to
    rounding, decisions and weights exactly)"""
    assert L.MAX_DIM == ND
    Cv = _cov()
    rs = np.random.RandomState(1)
    nch, nsteps = 5, 120
    z, u = rs.standard_normal((nsteps, nch, ND)), rs.uniform(size=(nsteps, nch))
    F = samplers.proposal_factor(np.eye(ND), ND, 2.4)
    delta = z @ F
    streams = dict(delta=delta, u=u) if mode == "delta" else dict(z=z, u=u)
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, streams=streams, proposal_update=False))())
    assert slik.program().gauss_direct is None
    got = {}
    for s in slik.sample():
        got.setdefault(s.chain, []).append((s.lnl, s.weight, s.x))
    oslik = ort.Slik(S.gauss_script(Cv)())
    lnl_batch = lambda X: np.array([oslik.evaluate(*x)[0] for x in X])
    rows, _, _ = osamp.mh_lockstep(lnl_batch, np.zeros((nch, ND)), nsteps, np.eye(ND), delta=delta, u=u,
                                   proposal_update=False)
    assert sum(len(r) for r in rows) > nch
    for c in range(nch):
        assert len(got.get(c, [])) == len(rows[c]), "chain %d: a decision flipped" % c
        np.testing.assert_array_equal([r[1] for r in got[c]], [r[1] for r in rows[c]])
        if mode == "delta":
            np.testing.assert_array_equal([r[2] for r in got[c]], [r[2] for r in rows[c]])
        else:
            np.testing.assert_allclose([r[2] for r in got[c]], [r[2] for r in rows[c]], rtol=0, atol=1e-12)
        np.testing.assert_allclose([r[0] for r in got[c]], [r[0] for r in rows[c]], rtol=1e-10)


def test_stretch_move_at_the_maximum_dimension():
    Cv = _cov()
    k, nsteps = 160, 4
    rs = np.random.RandomState(2)
    streams = [(rs.uniform(size=(2, k // 2)), rs.randint(k // 2, size=(2, k // 2)), rs.uniform(size=(2, k // 2)))
               for _ in range(nsteps)]
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, streams=lambda i: streams[i], seed=3), start=0.1)())
    pos0 = slik.sampler.initial_positions()
    slik.sampler.init_pos = pos0
    st = slik.sampler.run(slik.evaluate)
    oslik = ort.Slik(S.gauss_script(Cv)())
    lnprob = lambda q: np.array([-oslik.evaluate(*x)[0] for x in np.atleast_2d(q)])
    _, opos, olnp = osamp.emcee_wrapper_rows(pos0, lnprob, nsteps, lambda i: streams[i])
    np.testing.assert_array_equal(st["pos"].cpu().numpy(), opos)
    np.testing.assert_allclose(st["lnl"].cpu().numpy(), -olnp, rtol=1e-10)
    assert 0 < st["accepted"] < k * nsteps


def test_one_parameter_too_many_is_refused():
    Cv = np.eye(ND + 1)
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(s, num_samples=10))())
    with pytest.raises(ValueError, match="1..%d" % ND):
        slik.program()


def test_adaptive_mh_at_the_maximum_dimension():
    """pooled proposal update at 64 parameters: the device moments kernel (1 + 64 + 64^2 sums) and the host
    SVD factor against oracle.mh_lockstep"""
    Cv = _cov()
    rs = np.random.RandomState(5)
    nch, nsteps = 16, 300
    z, u = rs.standard_normal((nsteps, nch, ND)), rs.uniform(size=(nsteps, nch))
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, streams=dict(z=z, u=u), proposal_update_start=100,
        mpi_comm_freq=100))())
    got = {}
    for s in slik.sample():
        got.setdefault(s.chain, []).append((s.lnl, s.weight, s.x))
    oslik = ort.Slik(S.gauss_script(Cv)())
    lnl_batch = lambda X: np.array([oslik.evaluate(*x)[0] for x in X])
    rows, _, adapt = osamp.mh_lockstep(lnl_batch, np.zeros((nch, ND)), nsteps, np.eye(ND), z=z, u=u,
                                       proposal_update_start=100, exchange=100)
    assert len(adapt) == len(slik.sampler.adaptations) == 1 and adapt[0][0] == slik.sampler.adaptations[0][0] == 200
    np.testing.assert_allclose(slik.sampler.adaptations[0][1], adapt[0][1], rtol=1e-9, atol=1e-12)
    for c in range(nch):
        assert len(got.get(c, [])) == len(rows[c]), "chain %d: a decision flipped" % c
        np.testing.assert_array_equal([r[1] for r in got[c]], [r[1] for r in rows[c]])
        np.testing.assert_allclose([r[2] for r in got[c]], [r[2] for r in rows[c]], rtol=0, atol=1e-9)
