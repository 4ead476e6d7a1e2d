"""``output_extra_params`` on the GPU: derived quantities a script stores on itself in ``__call__`` are
compiled to the coefficient bytecode and evaluated by the priors/bytecode kernel for every emitted row
(reference: metropolis_hastings.py:103, 183-186, 221, 228).  Checked against a run of the UNMODIFIED
reference with the same injected streams (tests/golden/mh_extras.npz, oracle/make_golden_extras.py)."""
import numpy as np
import pytest

from conftest import golden
import b200_scripts as B
from cosmoslik_b200 import Slik, load_chain, samplers

pytestmark = pytest.mark.gpu


def test_mh_extra_columns_match_the_reference(tmp_path):
    g = golden("mh_extras.npz")
    names = [str(k) for k in g["extra_names"]]
    n = len(g["u"])
    out = str(tmp_path / "extras.chain")
    streams = dict(z=g["z"][:, None, :], u=g["u"][:, None])
    slik = Slik(B.derived_script(lambda s: samplers.metropolis_hastings(
        s, num_samples=n, streams=streams, output_file=out, output_extra_params=names, mpi_comm_freq=50,
        max_weight=5))())
    rows = list(slik.sample())
    assert len(rows) == len(g["weight"]), "different number of emitted rows: a decision flipped"
    np.testing.assert_array_equal([s.weight for s in rows], g["weight"])
    np.testing.assert_allclose([s.x for s in rows], g["x"], rtol=0, atol=1e-13)
    np.testing.assert_allclose([s.lnl for s in rows], g["lnl"], rtol=1e-10, atol=1e-13)
    got = np.array([[s.extra[k] for k in names] for s in rows])
    np.testing.assert_allclose(got, g["extras"], rtol=1e-12, atol=1e-12)
    assert np.all(got[:, names.index("fixed")] == 3.25)
    # the chain file: header carries the extra names, every record the extra columns, and the
    # single-process block structure (last row of each block dropped) is the reference's
    c = load_chain(out)[0]
    assert list(c.keys()) == ["lnl", "weight", "a", "b", "r2", "ratio", "fixed"]
    assert len(c["lnl"]) == len(g["file_lnl"])
    np.testing.assert_array_equal(c["weight"], g["file_weight"])
    for k in ("a", "b", "r2", "ratio", "fixed"):
        np.testing.assert_allclose(c[k], g["file_" + k], rtol=1e-12, atol=1e-12)


def test_extras_program_on_a_batch():
    """Slik.extras_program evaluates the derived quantities for any batch of points (ragged sizes,
    the empty batch) and the priors sampler attaches them to its samples"""
    slik = Slik(B.derived_script(lambda s: samplers.metropolis_hastings(s, num_samples=10))())
    ex = slik.extras_program(["ratio", "r2"])
    X = np.random.RandomState(3).standard_normal((333, 2))
    got = ex.values_host(X)
    ref = np.stack([(X[:, 0] - 1.5) / (2 + X[:, 1] ** 2), X[:, 0] ** 2 + X[:, 1] ** 2], axis=1)
    np.testing.assert_allclose(got, ref, rtol=1e-14, atol=0)
    assert ex.values_host(np.empty((0, 2))).shape == (0, 2)
    assert slik.extras_program(["ratio", "r2"]) is ex
    with pytest.raises(KeyError):
        slik.extras_program(["nope"])


def test_emcee_file_extras_are_those_of_the_next_position(tmp_path):
    """the emcee wrapper writes extras to the chain FILE only, and -- like the lnl -- those of the
    position the walker moved to (emcee.py:66-78)"""
    import oracle_scripts as S
    from oracle import runtime as ort, samplers as osamp
    k, nsteps, names = 24, 13, ["ratio", "r2"]
    rs = np.random.RandomState(11)
    streams = [(rs.uniform(size=(2, k // 2)), rs.randint(k // 2, size=(2, k // 2)), rs.uniform(size=(2, k // 2)))
               for _ in range(nsteps)]
    out = str(tmp_path / "emcee_extras.chain")
    slik = Slik(B.derived_script(lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, streams=lambda i: streams[i], seed=3, output_file=out,
        output_freq=4, output_extra_params=names))())
    pos0 = slik.sampler.initial_positions()
    slik.sampler.init_pos = pos0
    yielded = list(slik.sample())
    assert all(s.extra is None for s in yielded)                 # emcee.py:74: no extras on the samples
    oslik = ort.Slik(S.derived_script()())

    def lnprob(q):
        return np.array([-oslik.evaluate(*x)[0] for x in np.atleast_2d(q)])

    def blob(x):
        p = oslik.evaluate(*x)[1]
        return [p[n] for n in names]

    rows, opos, _ = osamp.emcee_wrapper_rows(pos0, lnprob, nsteps, lambda i: streams[i], blob_fn=blob)
    np.testing.assert_array_equal(slik.sampler.stats["pos"].cpu().numpy(), opos)
    chains = load_chain(out)
    assert len(chains) == k and list(chains[0].keys()) == ["lnl", "weight", "a", "b", "ratio", "r2"]
    for w in range(k):
        np.testing.assert_array_equal(chains[w]["weight"], [r[1] for r in rows[w]])
        np.testing.assert_array_equal(chains[w]["a"], [r[2][0] for r in rows[w]])
        for i, n in enumerate(names):
            np.testing.assert_allclose(chains[w][n], [r[3][i] for r in rows[w]], rtol=1e-13, atol=0)


def test_emcee_chain_file_matches_the_reference_wrapper(tmp_path):
    """the chain file of samplers.emcee against the file the UNMODIFIED reference wrapper wrote for the same script,
    initial positions and injected streams (tests/golden/ref_emcee.chain: oracle/make_golden_emcee_wrapper.py ran
    cosmoslik_plugins/samplers/emcee.py over a stand-in of the absent ``emcee`` package): header, walkers, records,
    weights and positions exact, lnl and the derived columns (those of the NEXT position) to 1e-13; both files are
    read by this package's loader"""
    import os
    from conftest import GOLDEN, golden
    g = golden("emcee_wrapper.npz")
    k, nsteps, freq = int(g["nwalkers"]), int(g["nsteps"]), int(g["output_freq"])
    out = str(tmp_path / "emcee_ref.chain")
    slik = Slik(B.derived_script(lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, streams=lambda i: (g["u_z"][i], g["j"][i], g["u_acc"][i]), seed=3,
        output_file=out, output_freq=freq, output_extra_params=["r2", "ratio"]))())
    assert list(slik.get_sampled()) == list(g["names"])
    np.testing.assert_allclose(slik.sampler.cov_est, g["cov_est"], rtol=0, atol=0)
    slik.sampler.init_pos = np.array(g["pos0"])
    yielded = list(slik.sample())
    assert len(yielded) == len(g["field_weight"])
    mine, ref = load_chain(out), load_chain(os.path.join(GOLDEN, "ref_emcee.chain"))
    assert len(mine) == len(ref) == k
    for w in range(k):
        assert list(mine[w].keys()) == list(ref[w].keys()) == ["lnl", "weight", "a", "b", "r2", "ratio"]
        for col in ("weight", "a", "b"):
            np.testing.assert_array_equal(mine[w][col], ref[w][col])
        for col in ("lnl", "r2", "ratio"):
            np.testing.assert_allclose(mine[w][col], ref[w][col], rtol=1e-13, atol=0)


def test_evaluate_returns_the_evaluated_tree():
    """cosmoslik.py:141: the second element of Slik.evaluate carries what the script stored on itself"""
    slik = Slik(B.derived_script(lambda s: samplers.metropolis_hastings(s, num_samples=10))())
    assert slik.derived_names() == ["r2", "ratio"]
    lnl, p = slik.evaluate(0.3, -0.4)
    assert abs(lnl - 0.125) < 1e-15 and p["a"] == 0.3 and p["b"] == -0.4
    assert abs(p["r2"] - 0.25) < 1e-15 and abs(p["ratio"] - (0.3 - 1.5) / (2 + 0.16)) < 1e-15
    lnl, p = slik.evaluate(a=0.3, b=-5.0)                    # below b's min: the script is not called (:135-137)
    assert lnl == np.inf and "r2" not in p


def test_polychord_callbacks_on_a_batch():
    """samplers/polychord.py:38-52 batched: cube -> parameters -> (-lnl, extras) for all live points at once"""
    from cosmoslik_b200 import SlikPlugin, param

    class script(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = param(start=0, scale=1, min=-2, max=3)
            self.b = param(start=1, scale=1, range=(0.5, 1.5))
            self.sampler = samplers.polychord(self, output_file="chains/pc", output_extra_params=["s"])

        def __call__(self):
            self.s = self.a + 2 * self.b
            return (self.a ** 2 + self.b ** 2) / 2

    slik = Slik(script())
    cube = np.random.RandomState(8).uniform(size=(257, 2))
    theta = slik.sampler.prior_transform(cube)
    assert theta[:, 0].min() >= -2 and theta[:, 0].max() <= 3 and theta[:, 1].min() >= 0.5 and theta[:, 1].max() <= 1.5
    logl, ex = slik.sampler.loglikelihood(slik.evaluate, theta)
    np.testing.assert_allclose(logl, -(theta[:, 0] ** 2 + theta[:, 1] ** 2) / 2, rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(ex[:, 0], theta[:, 0] + 2 * theta[:, 1], rtol=1e-15, atol=0)


def test_typed_extra_columns(tmp_path):
    """output_extra_params=(name, dtype) (metropolis_hastings.py:46-56): the derived value is cast to the column's
    scalar dtype in the yielded samples and in the chain file; array / object columns are refused loudly"""
    out = str(tmp_path / "typed.chain")
    n = 300
    slik = Slik(B.derived_script(lambda s: samplers.metropolis_hastings(
        s, num_samples=n, seed=4, output_file=out, output_extra_params=["ratio", ("r2", "f4"), ("fixed", "i8")]))())
    rows = list(slik.sample())
    assert all(isinstance(s.extra["fixed"], np.int64) and s.extra["fixed"] == 3 for s in rows)
    assert all(isinstance(s.extra["r2"], np.float32) for s in rows)
    r2 = np.array([s.x[0] ** 2 + s.x[1] ** 2 for s in rows])
    np.testing.assert_array_equal([s.extra["r2"] for s in rows], r2.astype(np.float32))
    c = load_chain(out)[0]
    assert c["fixed"].dtype == np.int64 and c["r2"].dtype == np.float32 and c["ratio"].dtype == np.float64
    with pytest.raises(NotImplementedError):
        B.derived_script(lambda s: samplers.metropolis_hastings(s, num_samples=10, output_extra_params=[("r2", "object")]))()
    with pytest.raises(NotImplementedError):
        B.derived_script(lambda s: samplers.metropolis_hastings(s, num_samples=10, output_extra_params=[("r2", "(2,2)d")]))()
