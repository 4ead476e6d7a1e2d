"""Host-side tracing + compilation (no GPU): the tables handed to the kernels, interpreted in
numpy by tests/table_interp.py, must reproduce the oracle / the reference's golden outputs."""
import numpy as np
import pytest

from conftest import golden
import b200_scripts as B
import oracle_scripts as S
import table_interp as TI
from cosmoslik_b200 import Slik, SlikDict, param, SlikPlugin, likelihoods, models
from cosmoslik_b200.program import CompiledProgram
from cosmoslik_b200 import _lib as L
from oracle import runtime as ort, likes as olikes


def compiled(script_cls):
    s = script_cls()
    slik = Slik(s)
    return slik, CompiledProgram(list(slik.get_sampled()), slik.params.priors, slik.trace())


@pytest.mark.parametrize("which", ["s12", "k11"])
def test_spt_lowl_tables_reproduce_reference(spt_data, which):
    g = golden("spt_lowl_lnl.npz")
    slik, cp = compiled(B.spt_script(spt_data, which, g["tt"]))
    assert cp.names == list(g[which + "_names"])
    got = np.array([TI.lnl(cp, x) for x in g[which + "_x"][:24]])
    ref = g[which + "_lnl"][:24]
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    f = np.isfinite(ref)
    np.testing.assert_allclose(got[f], ref[f], rtol=1e-11)
    assert cp.meta["n"] == 47 and cp.meta["n_pad"] == 64 and cp.meta["ntile"] == 1
    assert cp.flops["bin_dense"] == 2 * 47 * 3251


def test_quickstart_is_pure_bytecode():
    slik, cp = compiled(B.quick_script())
    assert cp.meta["resid_mode"] == 0 and cp.meta["nscalar"] == 1
    for x in ([0.3, -1.2], [0, 0], [5, 7]):
        assert TI.lnl(cp, x) == (x[0] ** 2 + x[1] ** 2) / 2


def test_gaussian_plugin_tables():
    A = np.random.RandomState(0).standard_normal((30, 30))
    C = A @ A.T / 30 + np.eye(30)
    slik, cp = compiled(B.gauss_script(C))
    assert cp.gauss_direct is not None and list(cp.gauss_direct["perm"]) == list(range(30))
    oslik = ort.Slik(S.gauss_script(C)())
    xs = np.random.RandomState(1).standard_normal((8, 30))
    for x in xs:
        assert abs(TI.lnl(cp, x) - oslik.evaluate(*x)[0]) < 1e-12 * abs(oslik.evaluate(*x)[0])
    # diagonal-block inverses really invert the factor's diagonal blocks
    npad = cp.meta["n_pad"]
    Lm = cp.tables["Lmat"].reshape(npad, npad)
    inv = cp.tables["Linv_diag"].reshape(-1, L.BM, L.BM)
    for b in range(npad // L.BM):
        s = slice(b * L.BM, (b + 1) * L.BM)
        np.testing.assert_allclose(inv[b] @ Lm[s, s], np.eye(L.BM), atol=1e-12)


def make_mspec_problem(ns, seed=3, nspec=3, nbins=20, lmin=50, lmax=700, banded=False):
    """A small multi-frequency bandpower problem on both the oracle and the product plugins."""
    rs = np.random.RandomState(seed)
    maps = ["T090", "T150", "T220"][:nspec]
    keys = [(a, b) for i, a in enumerate(maps) for b in maps[i:]]
    nl = lmax - lmin
    ell = np.arange(lmax)
    cmb = {"cl_TT": 4000.0 * np.exp(-(ell / 900.0) ** 1.1) + 20.0}
    use, windows, signal = {}, {}, {}
    for k in keys:
        use[k] = (lmin, lmax)
        centers = np.linspace(0, nl, nbins + 2)[1:-1]
        w = np.exp(-0.5 * ((np.arange(nl)[None, :] - centers[:, None]) / (nl / nbins / 2.0)) ** 2)
        w /= w.sum(1, keepdims=True)
        if banded:
            w[w < 1e-6] = 0.0
        else:
            w += 1e-10
        windows[k] = w
        signal[k] = w @ cmb["cl_TT"][lmin:lmax] * (1 + 0.02 * rs.standard_normal(nbins)) + 30
    n = nbins * len(keys)
    d = np.concatenate([signal[k] for k in sorted(keys)])
    Acov = rs.standard_normal((n, n)) * 0.05
    cov = np.diag((0.03 * d) ** 2) + (Acov @ Acov.T) * np.outer(0.03 * d, 0.03 * d)
    return dict(maps=maps, keys=keys, use=use, windows=windows, signal=signal, cov=cov, cmb=cmb)


def mspec_scripts(prob):
    """(oracle script class, product script class) for the same problem"""
    starts = dict(Aps=12.0, Acib=6.0, ncib=0.8, cal150=1.0, cal220=1.0, Acmb=1.0)

    def build(ns_param, ns_Plugin, like_cls, egfs_cls):
        class script(ns_Plugin):
            def __init__(self):
                super().__init__()
                self.Aps = ns_param(start=starts["Aps"], scale=2, min=0)
                self.Acib = ns_param(start=starts["Acib"], scale=1, min=0, max=50)
                self.ncib = ns_param(start=starts["ncib"], scale=0.1, range=(0.1, 2.5))
                self.cal150 = ns_param(start=1.0, scale=0.01, gaussian_prior=(1, 0.02))
                self.cal220 = ns_param(start=1.0, scale=0.02, gaussian_prior=(1, 0.05))
                self.Acmb = ns_param(start=1.0, scale=0.01)
                self.egfs = egfs_cls()
                self.like = like_cls(prob["use"], prob["windows"], prob["signal"], prob["cov"])
                self.sampler = None

            def __call__(self):
                self.like.cal = {"T150": self.cal150, "T220": self.cal220}
                cmb = {"cl_TT": self.Acmb * prob["cmb"]["cl_TT"]}
                return self.like(cmb, self.egfs(Aps=self.Aps, Acib=self.Acib, ncib=self.ncib))
        return script

    o = build(ort.Param, ort.Plugin, olikes.MspecLike, olikes.ClustPoissonEgfs)
    b = build(param, SlikPlugin, likelihoods.mspec_lnl, models.clust_poisson_egfs)
    return o, b


@pytest.mark.parametrize("banded", [False, True])
def test_mspec_with_power_law_foregrounds(banded):
    prob = make_mspec_problem(None, banded=banded)
    o_cls, b_cls = mspec_scripts(prob)
    oslik = ort.Slik(o_cls())
    slik, cp = compiled(b_cls)
    assert list(oslik.get_sampled()) == cp.names == ["Acib", "Acmb", "Aps", "cal150", "cal220", "ncib"]
    rs = np.random.RandomState(4)
    x0 = np.array([oslik.get_start()[k] for k in cp.names])
    for i in range(10):
        x = x0 * (1 + 0.05 * rs.standard_normal(len(x0)))
        if i == 3: x[cp.names.index("Aps")] = -1.0
        if i == 4: x[cp.names.index("ncib")] = 2.6
        ref = oslik.evaluate(*x)[0]
        got = TI.lnl(cp, x)
        if np.isinf(ref):
            assert np.isinf(got)
        else:
            assert abs(got - ref) <= 1e-11 * abs(ref)
    if banded:
        assert cp.flops["bin_exec"] < 0.7 * 2 * 64 * 656 * 6   # zero blocks are skipped
    assert cp.meta["nspec"] == 6 and cp.meta["n"] == 120 and cp.meta["n_pad"] == 128


def test_find_sampled_order_and_errors():
    t = SlikDict(b=param(start=1, scale=1), a=SlikDict(z=param(start=2, scale=1), y=3), c=4)
    assert list(t.find_sampled()) == ["a.z", "b"]
    t["a.z"] = 5
    assert t.a.z == 5 and t["a.z"] == 5 and "a.q" not in t and "a.y" in t
    c = t.deepcopy()
    c["a.y"] = 9
    assert t.a.y == 3 and c.a.y == 9
    with pytest.raises(ValueError):
        t[3] = 1
    slik = Slik(B.quick_script(sampler=lambda s: None)())
    with pytest.raises(ValueError, match="Expected 2 parameters"):
        slik._bind((1,), {})
    with pytest.raises(ValueError, match="Missing"):
        slik._bind((), {"a": 1})
    with pytest.raises(ValueError, match="either"):
        slik._bind((1, 2), {"a": 1})


def test_untraceable_script_fails_loudly():
    class bad(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = param(start=0, scale=1)
            self.sampler = None

        def __call__(self):
            return np.exp(self.a)   # numpy ufunc on a symbolic parameter

    with pytest.raises(TypeError, match="cannot be batched"):
        Slik(bad()).trace()


@pytest.mark.parametrize("by_modes", [False, True])
def test_model_dependent_covariance_tables(by_modes):
    """SPTSZ-style likelihood: C = sigma + V V^T with the beam modes as extra window rows"""
    import sptsz_problem as SP
    o_cls, b_cls = SP.scripts(nmode=5)
    oslik = ort.Slik(o_cls())
    slik = Slik(b_cls(by_modes))
    cp = CompiledProgram(list(slik.get_sampled()), slik.params.priors, slik.trace())
    assert cp.meta["resid_mode"] == 3 and cp.meta["n_base"] == 47 and cp.meta["nmode"] == 5
    assert cp.meta["n"] == 6 * 47 and cp.names == list(oslik.get_sampled())
    x0 = np.array([oslik.get_start()[k] for k in cp.names])
    for f in (1.0, 1.03, 0.95):
        ref = oslik.evaluate(*(x0 * f))[0]
        assert abs(TI.lnl(cp, x0 * f) - ref) <= 1e-11 * abs(ref)


@pytest.mark.parametrize("tag", ["m6_ab0", "m6_ab1", "m8_ab1"])
def test_sptsz_tables_against_the_reference_golden(tag):
    """the compiled tables (aberration table included) interpreted in numpy against what the reference's own
    SPTSZ_lowl.__call__ returned for the same arrays (tests/golden/sptsz.npz)"""
    import sptsz_problem as SP
    script, X, want = SP.reference_golden_script(tag)
    slik = Slik(script())
    assert list(slik.get_sampled()) == ["cal", "egfs.Acl", "egfs.Aps", "egfs.Asz"]
    cp = CompiledProgram(list(slik.get_sampled()), slik.params.priors, slik.trace())
    assert cp.meta["has_aberr"] == int(tag.endswith("ab1")) and cp.meta["nsegclass"] == 0
    for x, ref in list(zip(X, want))[::6]:
        assert abs(TI.lnl(cp, x) - ref) <= 1e-10 * abs(ref)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): exactly one JSON line on
    stdout with the contract's keys; the oracle port on all host cores, no GPU needed."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "evals/s" and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_output_extra_params_compile_to_bytecode():
    """derived quantities (metropolis_hastings.py:228, ``s.extra[k]``) are read off the traced tree and
    compiled to coefficient rows; the bytecode reproduces the reference's values at its own samples"""
    from cosmoslik_b200.program import compile_extras
    g = golden("mh_extras.npz")
    names = [str(k) for k in g["extra_names"]]
    slik = Slik(B.derived_script()())
    cp, rows = compile_extras(list(slik.get_sampled()), slik.extra_exprs(names))
    assert cp.meta["nbox"] == 0 and cp.meta["ngauss"] == 0 and cp.meta["resid_mode"] == 0
    assert len(rows) == len(names) == 4
    for x, ref in zip(g["x"][:200], g["extras"][:200]):
        coef = TI.run_bytecode(cp, x)
        np.testing.assert_allclose([coef[r] for r in rows], ref, rtol=4e-16, atol=0)
    with pytest.raises(KeyError):
        slik.extra_exprs(["nonexistent"])

    class vec(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = param(start=0, scale=1)
            self.sampler = None

        def __call__(self):
            self.v = np.arange(3.0)
            return self.a ** 2 / 2

    with pytest.raises(TypeError):
        Slik(vec()).extra_exprs(["v"])


def test_sampler_extra_params_bookkeeping():
    """constructor side of output_extra_params: name -> dtype map like the reference (:103); scalar numeric
    dtypes are cast, array / object columns are refused"""
    from cosmoslik_b200 import samplers
    s = B.derived_script(lambda p: samplers.metropolis_hastings(p, num_samples=10, output_extra_params=["r2", ("ratio", "float64")]))()
    assert list(s.sampler.output_extra_params.items()) == [("r2", "float64"), ("ratio", "float64")]
    s = B.derived_script(lambda p: samplers.metropolis_hastings(p, num_samples=10, output_extra_params=[("n", "int64")]))()
    assert list(s.sampler.output_extra_params.items()) == [("n", "int64")]
    for bad in ("object", "(3,)f8", "U8"):
        with pytest.raises(NotImplementedError):
            B.derived_script(lambda p: samplers.metropolis_hastings(p, num_samples=10, output_extra_params=[("n", bad)]))()


def test_mspec_cleaning_and_resgal():
    """mspec_lnl.py:38-42: `cleaning` (linear combinations of the maps applied to the data AND their covariance)
    and `resgal` (residual Galactic power subtracted from the data).  Product (host preprocessing -> device tables,
    interpreted in numpy) against the oracle's separately written restatement; plus the identities that pin the
    semantics: the identity combination changes nothing, and a one-map combination with normalize=False IS the
    calibration the reference applies with the same call (:44-45)."""
    from cosmoslik_b200.likelihoods.bandpower import lincombo
    base = make_mspec_problem(None, nbins=12)
    rs = np.random.RandomState(11)
    cleaning = {"Tclean": {"T150": 1.0, "T220": -0.25}, "T090": {"T090": 1.0}}
    new_keys = [("T090", "T090"), ("T090", "Tclean"), ("Tclean", "Tclean")]
    resgal = {k: 2.0 + rs.uniform(0, 1, 12) for k in new_keys}
    w = base["windows"][("T090", "T090")]
    prob = dict(base, use={k: (50, 700) for k in new_keys}, windows={k: w for k in new_keys})

    def scripts(extra):
        def build(ns_param, ns_Plugin, like_cls, egfs_cls):
            class script(ns_Plugin):
                def __init__(self):
                    super().__init__()
                    self.Aps = ns_param(start=12.0, scale=2, min=0)
                    self.Acib = ns_param(start=6.0, scale=1, min=0)
                    self.ncib = ns_param(start=0.8, scale=0.1, range=(0.1, 2.5))
                    self.calc = ns_param(start=1.0, scale=0.01, gaussian_prior=(1, 0.02))
                    self.egfs = egfs_cls()
                    self.like = like_cls(prob["use"], prob["windows"], base["signal"], base["cov"], **extra)
                    self.sampler = None

                def __call__(self):
                    self.like.cal = {"Tclean": self.calc}
                    return self.like({"cl_TT": prob["cmb"]["cl_TT"]}, self.egfs(Aps=self.Aps, Acib=self.Acib, ncib=self.ncib))
            return script
        return (build(ort.Param, ort.Plugin, olikes.MspecLike, olikes.ClustPoissonEgfs),
                build(param, SlikPlugin, likelihoods.mspec_lnl, models.clust_poisson_egfs))

    for extra in (dict(cleaning=cleaning), dict(cleaning=cleaning, resgal=resgal)):
        o_cls, b_cls = scripts(extra)
        oslik = ort.Slik(o_cls())
        slik, cp = compiled(b_cls)
        assert cp.meta["n"] == 36
        x0 = np.array([oslik.get_start()[k] for k in cp.names])
        vals = []
        for i in range(6):
            x = x0 * (1 + 0.05 * rs.standard_normal(len(x0)))
            ref = oslik.evaluate(*x)[0]
            assert abs(TI.lnl(cp, x) - ref) <= 1e-10 * abs(ref)
            vals.append(ref)
    # identities
    same, cov_same = lincombo(base["signal"], base["cov"], {m: {m: 1.0} for m in base["maps"]})
    assert all(np.array_equal(same[k], base["signal"][k]) for k in base["keys"]) and np.allclose(cov_same, base["cov"], rtol=1e-14)
    cal = {"T090": 1.0, "T150": 1.013, "T220": 0.96}
    scaled, _ = lincombo(base["signal"], None, {m: {m: cal[m]} for m in base["maps"]}, normalize=False)
    for (a, b) in base["keys"]:
        np.testing.assert_allclose(scaled[(a, b)], base["signal"][(a, b)] * cal[a] * cal[b], rtol=1e-15)


@pytest.mark.parametrize("name", ["spt", "subset", "clean", "planck"])
def test_mspec_tables_against_the_reference_golden(name):
    """the product's traced + compiled multi-spectrum programs (device tables interpreted in numpy) against what
    the UNMODIFIED reference's mspec_lnl.py returns for the same scripts (tests/golden/mspec.npz, generated over the
    container stand-in: oracle/make_golden_mspec.py): per-map calibration, a `use` subset, cleaning + resgal,
    TT / TE / EE with calibration on the data"""
    import mspec_cases
    from namespaces import PRODUCT
    from cosmoslik_b200 import synthetic
    prob, ns, X, want = mspec_cases.case(name, PRODUCT)
    slik, cp = compiled(synthetic.build_script(ns, prob))
    n = 10 if name != "planck" else len(X)
    got = np.array([TI.lnl(cp, x) for x in X[:n]])
    fin = np.isfinite(want[:n])
    np.testing.assert_array_equal(np.isfinite(got), fin)
    np.testing.assert_allclose(got[fin], want[:n][fin], rtol=1e-10)
