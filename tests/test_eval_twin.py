"""The DEVICE source of the per-walker evaluation pieces (cosmoslik_b200/csrc/eval_device.cuh: priors,
coefficient bytecode, lnl combine) compiled for the host by g++ (tests/host_twin/eval_twin.cpp) and run on
the CPU over the tables a CompiledProgram holds: against the reference's golden lnL for the pure-bytecode
quickstart script, against the oracle's priors (boxes, Gaussians, the min/max triple, NaN), and against the
numpy table interpreter for the coefficients of the bandpower scripts."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import golden
import b200_scripts as B
import oracle_scripts as S
import table_interp as TI
from cosmoslik_b200 import Slik, _lib as L
from cosmoslik_b200.program import CompiledProgram, compile_extras
from oracle import runtime as ort
from test_host_compile import make_mspec_problem, mspec_scripts

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("twin") / "eval_twin.so")
    subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "host_twin", "eval_twin.cpp")])
    return C.CDLL(so)


def host_struct(cp):
    """csl_program_t whose table pointers are the HOST numpy arrays of a CompiledProgram"""
    P = L.Program()
    for k, v in cp.meta.items():
        setattr(P, k, int(v))
    for k in cp._PTR_FIELDS:
        setattr(P, k, cp.tables[k].ctypes.data if k in cp.tables else None)
    return P


def run(twin, cp, X, chi2=None):
    X = np.ascontiguousarray(X, dtype=float)
    W = len(X)
    Wp = (W + 127) // 128 * 128
    P = host_struct(cp)
    coef = np.full((max(cp.meta["ncoef"], 1), Wp), np.nan)
    prior = np.full(Wp, np.nan)
    lnl = np.empty(W)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    twin.twin_prepare(C.byref(P), W, Wp, vp(X), vp(coef), Wp, vp(prior))
    twin.twin_combine(C.byref(P), W, vp(coef), Wp, vp(prior), vp(chi2) if chi2 is not None else None, vp(lnl))
    return coef, prior, lnl


def compiled(script_cls):
    from cosmoslik_b200.core import _tree_lookup
    slik = Slik(script_cls())
    tree, out = slik._trace()
    return CompiledProgram(list(slik.get_sampled()), slik.params.priors, out, derived=_tree_lookup(tree))


def test_quickstart_lnl_through_the_device_source(twin):
    """config 1 is pure bytecode: prepare + combine IS the posterior; the reference's golden chain supplies
    points and values (x**2 there is libm pow -> 1 ulp)"""
    g = golden("mh_quick.npz")
    cp = compiled(B.quick_script())
    coef, prior, lnl = run(twin, cp, g["x"])
    np.testing.assert_allclose(lnl, g["lnl"], rtol=4e-16, atol=0)
    assert np.all(prior[:len(g["x"])] == 0.0)
    # tile padding replays the last walker (k_prepare), never NaN
    assert not np.isnan(coef).any() and np.all(coef[:, len(g["x"]):] == coef[:, [len(g["x"]) - 1]])


def test_priors_through_the_device_source(twin, spt_data):
    """likelihoods/priors.py:24-30 on the spt_lowl 's12' script (Gaussian prior on cal, min = 0 on the
    amplitudes) and on a tree with a min/max triple; NaN fails a box"""
    g = golden("spt_lowl_lnl.npz")
    cp = compiled(B.spt_script(spt_data, "s12", g["tt"]))
    oslik = ort.Slik(S.spt_script(spt_data, "s12", g["tt"])())
    X = np.array(g["s12_x"])
    X[2, 2] = np.nan          # an amplitude with a min=0 box: NaN fails `lo <= x <= hi` -> +inf (priors.py:26-28)
    X[8, 0] = np.nan          # the calibration has only a Gaussian prior: the penalty itself is NaN, as in the reference
    _, prior, _ = run(twin, cp, X)
    for x, got in zip(X, prior):
        if np.isnan(x[2]):
            assert got == np.inf
            continue
        if np.isnan(x[0]):
            assert np.isnan(got) or got == np.inf
            continue
        params = oslik.params.clone()
        for k, v in zip(oslik.get_sampled(), x):
            params[k] = v
        ref = params.priors(params)
        assert got == ref or (np.isinf(got) and np.isinf(ref))
    assert np.isinf(prior[:len(X)]).sum() >= 3


@pytest.mark.parametrize("banded", [False, True])
def test_coefficients_through_the_device_source(twin, banded):
    """multi-frequency script: calibration products, amplitudes, tilt indices -- every coefficient row the
    binning kernels read, bit for bit against the numpy interpreter of the same tables"""
    prob = make_mspec_problem(None, banded=banded)
    _, b_cls = mspec_scripts(prob)
    cp = compiled(b_cls)
    rs = np.random.RandomState(6)
    x0 = np.array([12.0 if n == "Aps" else 6.0 if n == "Acib" else 0.8 if n == "ncib" else 1.0 for n in cp.names])
    X = x0 * (1 + 0.05 * rs.standard_normal((131, len(x0))))
    X[5, cp.names.index("Aps")] = -1.0
    coef, prior, _ = run(twin, cp, X)
    for w, x in enumerate(X):
        np.testing.assert_array_equal(coef[:cp.meta["ncoef"], w], TI.run_bytecode(cp, x)[:cp.meta["ncoef"]])
        assert prior[w] == TI.prior(cp, x) or (np.isinf(prior[w]) and np.isinf(TI.prior(cp, x)))
    assert np.isinf(prior[5])


def test_derived_quantities_through_the_device_source(twin):
    """output_extra_params: the golden extras of the unmodified reference from the device interpreter"""
    g = golden("mh_extras.npz")
    names = [str(k) for k in g["extra_names"]]
    slik = Slik(B.derived_script()())
    cp, rows = compile_extras(list(slik.get_sampled()), slik.extra_exprs(names))
    coef, _, _ = run(twin, cp, g["x"])
    np.testing.assert_allclose(coef[rows, :len(g["x"])].T, g["extras"], rtol=4e-16, atol=0)


def test_priors_on_derived_quantities_match_the_live_reference(twin):
    """cosmoslik.py:139-141: priors are evaluated again after the script's __call__, so add_gaussian_prior /
    add_uniform_prior may name quantities the script derives there.  The device source (negative prior indices ->
    coefficient rows) against the unmodified reference on the same script, when /root/reference is mounted, and
    against the numpy table interpreter always."""
    from cosmoslik_b200 import SlikPlugin, param

    def build(ns_plugin, ns_param):
        class derived(ns_plugin):
            def __init__(self):
                super().__init__()
                self.a = ns_param(start=0.3, scale=1)
                self.b = ns_param(start=0.5, scale=1, min=-4)
                self.sampler = None

            def __call__(self):
                self.r2 = self.a ** 2 + self.b ** 2
                self.ratio = (self.a - 1.5) / (2 + self.b ** 2)
                self.fixed = 3.25
                return self.r2 / 2
        return derived

    def with_priors(script, priors_cls):
        script.priors = priors_cls(script)
        script.priors.add_gaussian_prior("r2", 1.0, 0.7)
        script.priors.add_uniform_prior("ratio", -0.6, 0.1)
        script.priors.add_gaussian_prior("a", 0.1, 2.0)
        script.priors.add_gaussian_prior("fixed", 3.0, 0.5)
        script.priors.add_gaussian_prior("never_set", 0.0, 1.0)     # skipped by the reference (param_ok), and here
        return script

    from cosmoslik_b200.likelihoods.priors import priors as my_priors
    from cosmoslik_b200.core import _tree_lookup
    slik = Slik(with_priors(build(SlikPlugin, param)(), my_priors))
    tree, out = slik._trace()
    with pytest.warns(UserWarning, match="never_set"):
        cp = CompiledProgram(list(slik.get_sampled()), slik.params.priors, out, derived=_tree_lookup(tree))
    assert cp.derived_priors == 3 and (cp.tables["gauss_idx"][:cp.meta["ngauss"]] < 0).sum() == 2
    rs = np.random.RandomState(4)
    X = rs.uniform(-2, 2, (200, 2))
    X[7] = [0.5, -5.0]                   # the box of a sampled parameter
    coef, prior, lnl = run(twin, cp, X)
    want = np.array([TI.lnl(cp, x) for x in X])
    np.testing.assert_array_equal(np.isinf(lnl), np.isinf(want))
    f = np.isfinite(want)
    np.testing.assert_allclose(lnl[f], want[f], rtol=1e-14)
    assert np.isinf(lnl).sum() > 20 and f.sum() > 20
    if os.path.isdir("/root/reference"):
        from oracle import refshim
        R = refshim.load()
        script = build(R.core.SlikPlugin, R.core.param)()
        script.sampler = R.mh.metropolis_hastings(script, num_samples=1)
        rslik = R.core.Slik(with_priors(script, R.priors.priors if hasattr(R, "priors") else
                                        __import__("cosmoslik_plugins.likelihoods.priors", fromlist=["priors"]).priors))
        ref = np.array([rslik.evaluate(*x)[0] for x in X])
        np.testing.assert_array_equal(np.isinf(lnl), np.isinf(ref))
        np.testing.assert_allclose(lnl[f], ref[f], rtol=1e-13)
