"""The CPU oracle side by side with the UNMODIFIED reference, live, on inputs the committed golden
vectors do not contain (fresh seeds, other sampler options).  Runs only where the reference tree is
mounted (the build container); skipped on the GPU box, where tests/test_oracle_golden.py checks the
oracle against the committed outputs of the same reference code instead."""
import numpy as np
import pytest

from oracle import likes, refshim, runtime, samplers
import oracle_scripts as S

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def R():
    return refshim.load()


def _ref_spt_script(R, which, tt):
    core = R.core

    class script(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.spt = R.spt_lowl.spt_lowl(which)
            self.cmb = {"TT": tt}
            self.sampler = None

        def __call__(self):
            return self.spt(self.cmb)

    return script


@pytest.mark.parametrize("which", ["s12", "k11"])
def test_spt_lowl_lnl_at_fresh_points(R, spt_data, which):
    tt = S.synthetic_tt() * 1.03 + 7.0                      # not the golden spectrum
    ref = R.core.Slik(_ref_spt_script(R, which, tt)())
    ora = runtime.Slik(S.spt_script(spt_data, which, tt)())
    names = list(ref.get_sampled().keys())
    assert [str(n) for n in names] == list(ora.get_sampled())
    start = np.array([ref.get_start()[k] for k in names], dtype=float)
    scale = np.array([ref.get_sampled()[k].scale for k in names], dtype=float)
    X = start + 2.0 * scale * np.random.RandomState(777).standard_normal((40, len(names)))
    X[3, -1] = -0.5                                          # an amplitude below its min -> hard prior
    n_inf = 0
    for x in X:
        a, b = ref.evaluate(*x)[0], ora.evaluate(*x)[0]
        if np.isinf(a):
            assert np.isinf(b)
            n_inf += 1
        else:
            assert abs(a - b) <= 1e-13 * abs(a)
    assert n_inf >= 1


def test_clust_poisson_egfs_at_fresh_parameters(R):
    ref, ora = R.clust_poisson_egfs.clust_poisson_egfs(), likes.ClustPoissonEgfs()
    for Aps, Acib, ncib in np.random.RandomState(5).uniform([0.1, 0.1, 0.1], [50, 30, 2.0], size=(10, 3)):
        np.testing.assert_array_equal(ora(Aps=Aps, Acib=Acib, ncib=ncib)(spectra="cl_TT", lmax=2000),
                                      ref(Aps=Aps, Acib=Acib, ncib=ncib)(spectra="cl_TT", lmax=2000))


def test_priors_plugin_on_random_trees(R):
    """likelihoods/priors.py:6-30 on trees mixing every way of declaring a prior"""
    rs = np.random.RandomState(9)
    decls = [dict(start=0, scale=1), dict(start=0, min=-1), dict(start=0, max=2), dict(start=0, min=-1, max=1),
             dict(start=0, range=(-3, 3)), dict(start=0, uniform_prior=(-2, 2), gaussian_prior=(0.3, 1.5)),
             dict(start=0, gaussian_prior=(-1.0, 0.25))]
    core = R.core
    for _ in range(20):
        pick = [decls[i] for i in rs.randint(len(decls), size=5)]
        rt = core.SlikDict(**{"p%d" % i: core.param(**d) for i, d in enumerate(pick)})
        ot = runtime.Tree(**{"p%d" % i: runtime.Param(**d) for i, d in enumerate(pick)})
        rp, op = R.priors.priors(rt), runtime.Priors(ot)
        assert [tuple(u) for u in rp.uniform_priors] == [tuple(u) for u in op.uniform_priors]
        assert [tuple(g) for g in rp.gaussian_priors] == [tuple(g) for g in op.gaussian_priors]
        for _ in range(8):
            vals = rs.uniform(-3.5, 3.5, size=5)
            r2, o2 = rt.deepcopy(), ot.clone()
            for i, v in enumerate(vals):
                r2["p%d" % i] = v
                o2["p%d" % i] = v
            a, b = rp(r2), op(o2)
            assert (np.isinf(a) and np.isinf(b)) or a == b


@pytest.mark.parametrize("opts", [dict(), dict(max_weight=3), dict(temp=2.5), dict(proposal_scale=1.1, max_weight=25)])
def test_mh_rows_with_fresh_streams_and_options(R, opts):
    """metropolis_hastings._mcmc (:144-164) with injected streams: rows bit for bit"""
    from oracle.make_golden import run_injected_mh
    core = R.core

    class quick(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = core.param(start=0.3, scale=0.7, min=-2.5)
            self.b = core.param(start=-0.2, scale=1.3, gaussian_prior=(0.0, 2.0))

        def __call__(self):
            return (self.a ** 2 + self.b ** 2 + self.a * self.b) / 2

    class oquick(runtime.Plugin):
        def __init__(self):
            super().__init__()
            self.a = runtime.Param(start=0.3, scale=0.7, min=-2.5)
            self.b = runtime.Param(start=-0.2, scale=1.3, gaussian_prior=(0.0, 2.0))

        def __call__(self):
            return (self.a ** 2 + self.b ** 2 + self.a * self.b) / 2

    rs = np.random.RandomState(31337)
    n = 1500
    z, u = rs.standard_normal((n, 2)), rs.uniform(size=n)
    names, rrows, _ = run_injected_mh(R, quick, n, z, u, **opts)
    oslik = runtime.Slik(oquick())
    x0 = [oslik.get_sampled()[k].start for k in oslik.get_sampled()]
    F = runtime.proposal_factor(runtime.initialize_covariance(oslik.get_sampled()), 2, opts.get("proposal_scale", 2.4))
    zs, us = iter(z), iter(u)
    kw = {k: v for k, v in opts.items() if k in ("max_weight", "temp")}
    orows = list(samplers.mh_mcmc(oslik.evaluate, x0, n, lambda cx: np.asarray(cx) + np.dot(next(zs), F),
                                  lambda: next(us), **kw))
    assert [str(k) for k in names] == list(oslik.get_sampled())
    assert len(orows) == len(rrows) > 50
    np.testing.assert_array_equal([r.weight for r in orows], [r[1] for r in rrows])
    np.testing.assert_array_equal([r.x for r in orows], [r[2] for r in rrows])
    np.testing.assert_array_equal([r.lnl for r in orows], [r[0] for r in rrows])


def test_covariance_sources(R, tmp_path):
    """samplers/utils.py:4-23 + chains.py:891-927: dict of scales, (names, array), covariance file,
    later sources overriding earlier ones"""
    core = R.core
    covfile = str(tmp_path / "prop.covmat")
    with open(covfile, "w") as f:
        f.write("# b c\n1.5 0.2\n0.2 0.8\n")
    decl = dict(a=dict(start=0, scale=2.0), b=dict(start=1, scale=0.5), c=dict(start=1), d=dict(start=0, scale=0.1))
    rs_ = core.SlikDict(**{k: core.param(**v) for k, v in decl.items()}).find_sampled()
    os_ = runtime.Tree(**{k: runtime.Param(**v) for k, v in decl.items()}).find_sampled()
    sources = [(["c", "a"], np.array([[4.0, 0.3], [0.3, 9.0]])), covfile]
    np.testing.assert_array_equal(R.sutils.initialize_covariance(rs_, sources), runtime.initialize_covariance(os_, sources))
    np.testing.assert_array_equal(R.sutils.initialize_covariance(rs_, covfile), runtime.initialize_covariance(os_, covfile))
    dat = np.random.RandomState(2).standard_normal((300, 4)) @ np.random.RandomState(3).standard_normal((4, 4))
    w = np.random.RandomState(4).randint(0, 9, size=300).astype(float)
    np.testing.assert_allclose(runtime.get_covariance(dat, w), R.chains.get_covariance(dat, w), rtol=1e-14)


def test_product_chain_statistics_against_the_reference_container(R):
    """the host container the samplers fill (cosmoslik_b200.chains.Chain) against cosmoslik.chains.Chain
    on the same weighted samples"""
    from cosmoslik_b200 import Chain
    rs = np.random.RandomState(12)
    n = 500
    d = {"weight": rs.randint(1, 7, size=n).astype(float), "lnl": rs.uniform(0, 9, size=n),
         "a": rs.standard_normal(n), "b": rs.gamma(2.0, size=n), "c.d": rs.standard_normal(n) * 3 + 1}
    mine, ref = Chain(d), R.chains.Chain(d)
    ps = ["a", "b", "c.d"]
    np.testing.assert_allclose(mine.mean(ps), ref.mean(ps), rtol=1e-14)
    np.testing.assert_allclose(mine.std(ps), ref.std(ps), rtol=1e-13)
    np.testing.assert_allclose(mine.cov(ps), ref.cov(ps), rtol=1e-13)
    np.testing.assert_allclose(mine.corr(ps), ref.corr(ps), rtol=1e-13)
    assert mine.acceptance() == ref.acceptance() and mine.length() == ref.length()
    assert mine.length(unique=False) == ref.length(unique=False)
    for k in ps:
        np.testing.assert_array_equal(mine.burnin(120)[k], ref.burnin(120)[k])
        np.testing.assert_array_equal(mine.thin(7)[k], ref.thin(7)[k])
        np.testing.assert_array_equal(mine.thin(7)["weight"], ref.thin(7)["weight"])
    bm, br = mine.best_fit(), ref.best_fit()
    assert {k: bm[k] for k in ps} == {k: br[k] for k in ps}
    g_m = mine.add_gauss_prior(["a", "b"], [0.1, 2.0], np.array([[1.0, 0.2], [0.2, 2.0]]))
    g_r = ref.add_gauss_prior(["a", "b"], [0.1, 2.0], np.array([[1.0, 0.2], [0.2, 2.0]]))
    np.testing.assert_allclose(g_m["weight"], g_r["weight"], rtol=1e-13)


@pytest.mark.parametrize("freq,n", [(50, 1200), (37, 900), (100, 400)])
def test_chain_file_block_structure_live(R, tmp_path, freq, n):
    """the product's file writer (samplers.metropolis_hastings._FileBlocks, single-process branch) fed with
    the rows the reference yields must reproduce, record by record, the file the reference wrote in the
    same run (metropolis_hastings.py:219-244), for several mpi_comm_freq"""
    import pickle
    from oracle.make_golden import run_injected_mh
    from cosmoslik_b200.samplers.metropolis_hastings import _FileBlocks
    core = R.core

    class quick(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = core.param(start=0, scale=1)
            self.b = core.param(start=0, scale=1)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    rs = np.random.RandomState(freq)
    ref_file, my_file = str(tmp_path / "ref.chain"), str(tmp_path / "mine.chain")
    names, rows, _ = run_injected_mh(R, quick, n, rs.standard_normal((n, 2)), rs.uniform(size=n), tmpfile=ref_file,
                                     mpi_comm_freq=freq)
    fb = _FileBlocks(my_file, ["lnl", "weight"] + [str(k) for k in names], freq, single=True)
    arr = np.array([np.hstack([r[0], r[1], r[2]]) for r in rows])
    for i in range(0, len(arr), 41):
        fb.add(0, arr[i:i + 41])
    fb.finish()

    def records(path):
        out = []
        with open(path, "rb") as f:
            while True:
                try:
                    out.append(pickle.load(f))
                except EOFError:
                    return out

    mine, ref = records(my_file), records(ref_file)
    assert list(mine[0]) == list(ref[0]) and len(mine) == len(ref)
    for (i1, r1), (i2, r2) in zip(mine[1:], ref[1:]):
        assert i1 == i2 and r1.dtype == r2.dtype and len(r1) == len(r2)
        np.testing.assert_array_equal(r1.view(float), r2.view(float))


def test_product_covariance_sources_against_the_reference(R, tmp_path):
    """the product's host helpers (samplers.initialize_covariance, chains.combine_covs) on every source
    format the reference accepts (samplers/utils.py:4-23; chains.py:891-927): scales, (names, array),
    covariance file, a Chain, and lists of them with later entries overriding earlier ones"""
    import cosmoslik_b200 as cb
    core = R.core
    covfile = str(tmp_path / "prop.covmat")
    with open(covfile, "w") as f:
        f.write("# b c\n1.5 0.2\n0.2 0.8\n")
    rs = np.random.RandomState(21)
    cd = {"weight": rs.randint(1, 5, size=80).astype(float), "lnl": rs.uniform(size=80),
          "a": rs.standard_normal(80), "d": rs.standard_normal(80) * 0.3}
    decl = dict(a=dict(start=0, scale=2.0), b=dict(start=1, scale=0.5), c=dict(start=1), d=dict(start=0, scale=0.1))
    r_s = core.SlikDict(**{k: core.param(**v) for k, v in decl.items()}).find_sampled()
    b_s = cb.SlikDict(**{k: cb.param(**v) for k, v in decl.items()}).find_sampled()
    assert list(r_s.keys()) == list(b_s.keys())
    tup = (["c", "a"], np.array([[4.0, 0.3], [0.3, 9.0]]))
    for r_src, b_src in ((covfile, covfile), ([tup, covfile], [tup, covfile]), ([covfile, tup], [covfile, tup]),
                         ([covfile, R.chains.Chain(cd)], [covfile, cb.Chain(cd)])):
        np.testing.assert_allclose(cb.samplers.initialize_covariance(b_s, b_src),
                                   R.sutils.initialize_covariance(r_s, r_src), rtol=1e-14, atol=0)
    with pytest.raises(ValueError):
        cb.samplers.initialize_covariance(b_s)                 # 'c' has neither a scale nor a covariance
    with pytest.raises(ValueError):
        R.sutils.initialize_covariance(r_s)


def test_spt_lowl_loader_and_trimming_match_the_live_reference():
    """likelihoods.spt_lowl(datadir=...): read_newdat on the reference's own tarballs and the lmin / lmax / drop
    trimming (spt_lowl.py:65-96) against the unmodified plugin, for both data sets and several cuts"""
    import os
    from cosmoslik_b200 import likelihoods
    R = refshim.load()
    datadir = os.path.join(os.path.dirname(R.spt_lowl.__file__), "data")
    for which in ("s12", "k11"):
        for kw in (dict(), dict(lmin=650), dict(lmax=2000), dict(lmin=800, lmax=2500), dict(drop=[0, 5, 46]),
                   dict(lmin=700, drop=[2, 3])):
            ref = R.spt_lowl.spt_lowl(which, **kw)
            mine = likelihoods.spt_lowl(which, datadir=datadir, **kw)
            np.testing.assert_array_equal(mine.spec, ref.spec)
            np.testing.assert_array_equal(mine.sigma, ref.sigma)
            np.testing.assert_array_equal(mine.windows, np.array(ref.windows))
            assert (mine.windowrange.start, mine.windowrange.stop, mine.lmax) == (ref.windowrange.start, ref.windowrange.stop, ref.lmax)
            np.testing.assert_allclose(mine.ells, ref.ells, rtol=1e-14)
    # the default foreground templates read from the same directory
    ref = R.spt_lowl.spt_lowl("s12")
    mine = likelihoods.spt_lowl("s12", datadir=datadir)
    for t in ("template_sz", "template_cl", "template_ps"):
        n = min(len(getattr(mine.egfs, t)), len(getattr(ref.egfs, t)))
        np.testing.assert_array_equal(getattr(mine.egfs, t)[:n], getattr(ref.egfs, t)[:n])


def test_mspec_lnl_live_over_the_container_stand_in(R):
    """the reference's own mspec_lnl.py (over oracle/mspec_standin.py) against the oracle restatement AND the
    product's compiled tables on a FRESH problem and fresh points -- the live version of tests/golden/mspec.npz"""
    import table_interp as TI
    from namespaces import ORACLE, PRODUCT
    from cosmoslik_b200 import Slik, synthetic
    from cosmoslik_b200.program import CompiledProgram
    from oracle.make_golden_mspec import points, reference_namespace
    prob = synthetic.spt_multifreq_problem(17, nbins=7, lmax=451)
    R2, ns = reference_namespace()
    rslik = R2.core.Slik(synthetic.build_script(ns, prob)())
    names, X = points(prob, 6, 5)
    assert list(rslik.get_sampled()) == names
    want = np.array([rslik.evaluate(*x)[0] for x in X])
    oslik = runtime.Slik(synthetic.build_script(ORACLE, prob)())
    got = np.array([oslik.evaluate(*x)[0] for x in X])
    fin = np.isfinite(want)
    assert fin.sum() >= 4 and not fin.all()
    np.testing.assert_array_equal(np.isfinite(got), fin)
    np.testing.assert_allclose(got[fin], want[fin], rtol=1e-12)
    pslik = Slik(synthetic.build_script(PRODUCT, prob)())
    cp = CompiledProgram(list(pslik.get_sampled()), pslik.params.priors, pslik.trace())
    tab = np.array([TI.lnl(cp, x) for x in X])
    np.testing.assert_array_equal(np.isfinite(tab), fin)
    np.testing.assert_allclose(tab[fin], want[fin], rtol=1e-10)


def test_emcee_wrapper_live_over_the_move_stand_in(R, tmp_path):
    """the reference's own samplers/emcee.py (over oracle/emcee_standin.py) with FRESH streams, walker count, step
    count and output_freq, against the oracle's restatement of the wrapper -- the live version of
    tests/golden/emcee_wrapper.npz"""
    import pickle
    import oracle_scripts as S
    from oracle.make_golden_emcee_wrapper import run
    out = str(tmp_path / "live_emcee.chain")
    g = run(out, k=10, nsteps=7, freq=2, stream_seed=5, scatter_seed=6)
    slik = runtime.Slik(S.derived_script()())
    lnprob = lambda q: np.array([-slik.evaluate(*x)[0] for x in np.atleast_2d(q)])
    blob = lambda x: [slik.evaluate(*x)[1][n] for n in ("r2", "ratio")]
    rows, _, _ = samplers.emcee_wrapper_rows(g["pos0"], lnprob, 7, lambda i: (g["u_z"][i], g["j"][i], g["u_acc"][i]),
                                             blob_fn=blob)
    per = {}
    with open(out, "rb") as f:
        pickle.load(f)
        while True:
            try:
                w, rec = pickle.load(f)
            except EOFError:
                break
            assert len(rec) <= 2
            per.setdefault(int(w), []).extend(list(r) for r in rec)
    for w in range(10):
        mine = np.array([[r[0], r[1]] + list(r[2]) + list(r[3]) for r in rows[w]])
        np.testing.assert_allclose(mine, np.array(per[w], dtype=float), rtol=1e-13)
    assert int(g["evals"]) == 2 * 10 * 7
