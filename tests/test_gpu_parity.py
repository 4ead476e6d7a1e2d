"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the golden outputs of
the unmodified reference.  Tolerances: lnL 1e-10 relative (BASELINE.json north_star); positions,
weights and accept/reject decisions exact."""
import numpy as np
import pytest

from conftest import golden
import b200_scripts as B
import oracle_scripts as S
from namespaces import ORACLE, PRODUCT
from cosmoslik_b200 import Slik, load_chain, samplers, synthetic
from cosmoslik_b200 import _lib as L
from oracle import philox as ophilox, runtime as ort, samplers as osamp

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def torch():
    import torch
    torch.cuda.set_device(0)
    return torch


def relerr(got, ref):
    f = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), f)
    return np.max(np.abs(got[f] - ref[f]) / np.abs(ref[f])) if f.any() else 0.0


# ---------------------------------------------------------------- kernels in isolation
def test_dmma_gemm_matches_fp64_matmul(torch):
    lib = L.lib()
    rs = np.random.RandomState(0)
    M, N, K = 128, 256, 208
    A = torch.from_numpy(rs.standard_normal((M, K))).cuda()
    Bm = torch.from_numpy(rs.standard_normal((K, N))).cuda()
    Cm = torch.empty((M, N), dtype=torch.float64, device="cuda")
    L.check(lib.csl_dgemm_dmma(M, N, K, A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), L.stream_ptr()))
    ref = A.cpu().numpy() @ Bm.cpu().numpy()
    np.testing.assert_allclose(Cm.cpu().numpy(), ref, rtol=0, atol=1e-12)


def test_philox_streams_match_the_oracle(torch):
    lib = L.lib()
    W, nd, seed, step, c0 = 1000, 7, 0x1234567890ABCDEF, 41, 17
    z = torch.empty((W, nd), dtype=torch.float64, device="cuda")
    u = torch.empty(W, dtype=torch.float64, device="cuda")
    L.check(lib.csl_philox_mh(W, nd, seed, step, c0, z.data_ptr(), u.data_ptr(), L.stream_ptr()))
    zo, uo = ophilox.mh_streams(seed, step, c0 + np.arange(W), nd)
    np.testing.assert_array_equal(u.cpu().numpy(), uo)                 # integer arithmetic: exact
    np.testing.assert_allclose(z.cpu().numpy(), zo, rtol=0, atol=1e-14)  # log/sqrt/sincos: a few ulp
    Ns, Nc = 777, 5000
    uz = torch.empty(Ns, dtype=torch.float64, device="cuda")
    ua = torch.empty(Ns, dtype=torch.float64, device="cuda")
    j = torch.empty(Ns, dtype=torch.int64, device="cuda")
    L.check(lib.csl_philox_stretch(Ns, Nc, seed, step, 1, c0, uz.data_ptr(), j.data_ptr(), ua.data_ptr(), L.stream_ptr()))
    uzo, jo, uao = ophilox.stretch_streams(seed, step, 1, c0 + np.arange(Ns), Nc)
    np.testing.assert_array_equal(uz.cpu().numpy(), uzo)
    np.testing.assert_array_equal(ua.cpu().numpy(), uao)
    np.testing.assert_array_equal(j.cpu().numpy(), jo)


# ---------------------------------------------------------------- posterior evaluation
@pytest.mark.parametrize("which", ["s12", "k11"])
def test_spt_lowl_lnl_matches_the_reference(torch, spt_data, which):
    g = golden("spt_lowl_lnl.npz")
    slik = Slik(B.spt_script(spt_data, which, g["tt"])())
    assert list(slik.get_sampled()) == list(g[which + "_names"])
    X, ref = g[which + "_x"], g[which + "_lnl"]
    got = slik.evaluate_batch(X)                     # host buffers -> csl_lnl_host
    assert relerr(got, ref) < RTOL
    got_dev = slik.evaluate_batch(torch.from_numpy(X).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got_dev, got)      # same kernels, same bits
    one, params = slik.evaluate(*X[0])               # scalar API = batch of one
    assert one == got[0] and params["spt.egfs.Acl"] == X[0][list(g[which + "_names"]).index("spt.egfs.Acl")]
    assert slik.evaluate(**dict(zip(g[which + "_names"], X[1])))[0] == got[1]
    assert ref[0] == {"s12": 26542.27655449845, "k11": 16488.100618743654}[which]


def test_quickstart_and_gaussian_lnl(torch):
    slik = Slik(B.quick_script()())
    X = np.random.RandomState(0).standard_normal((300, 2)) * 3
    np.testing.assert_array_equal(slik.evaluate_batch(X), (X[:, 0] ** 2 + X[:, 1] ** 2) / 2)
    A = np.random.RandomState(0).standard_normal((30, 30))
    Cv = A @ A.T / 30 + np.eye(30)
    slik = Slik(B.gauss_script(Cv)())
    oslik = ort.Slik(S.gauss_script(Cv)())
    X = np.random.RandomState(1).standard_normal((257, 30))
    ref = np.array([oslik.evaluate(*x)[0] for x in X])
    assert relerr(slik.evaluate_batch(X), ref) < RTOL


@pytest.mark.parametrize("kind", ["spt_dense", "planck_banded"])
def test_multispectrum_lnl_matches_the_oracle(torch, kind):
    prob = synthetic.spt_multifreq_problem(2, nbins=20, lmax=1001) if kind == "spt_dense" \
        else synthetic.planck_lite_problem(2)
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    names = list(slik.get_sampled())
    assert names == list(oslik.get_sampled())
    rs = np.random.RandomState(5)
    x0 = np.array([prob["truth"][k] for k in names])
    sc = np.array([prob["scales"][k] for k in names])
    X = x0 + sc * rs.standard_normal((200, len(names))) * 2
    X[0] = x0
    X[7, names.index("n_cib")] = 2.5      # outside its range -> +inf
    ref = np.array([oslik.evaluate(*x)[0] for x in X])
    assert np.isinf(ref).sum() >= 1
    got = slik.evaluate_batch(X)
    assert relerr(got, ref) < RTOL


@pytest.mark.parametrize("name", ["spt", "subset", "clean", "planck"])
def test_multispectrum_lnl_matches_the_reference_golden(torch, name):
    """the CUDA path against what the UNMODIFIED reference's mspec_lnl.py returns for the same scripts
    (tests/golden/mspec.npz: oracle/make_golden_mspec.py ran them through the reference's Slik.evaluate over the
    stand-in of its absent container package): lnL within 1e-10, the same hard-prior rejections"""
    import mspec_cases
    prob, ns, X, want = mspec_cases.case(name, PRODUCT)
    slik = Slik(synthetic.build_script(ns, prob)())
    got = slik.evaluate_batch(X)
    fin = np.isfinite(want)
    assert not fin.all()
    np.testing.assert_array_equal(np.isfinite(got), fin)
    np.testing.assert_allclose(got[fin], want[fin], rtol=RTOL)


@pytest.mark.parametrize("mode", ["trsm", "inverse"])
def test_both_chi2_kernels_match_the_oracle(torch, spt_data, mode):
    """blocked forward substitution (k_trsm_chi2) and the explicit-inverse tile product (k_trigemm_chi2)"""
    prob = synthetic.planck_lite_problem(3)
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    prog = slik.program(chi2_mode=mode)
    assert prog.meta["chi2_mode"] == (1 if mode == "inverse" else 0)
    assert slik.program(chi2_mode=mode) is prog
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    X = x0 + sc * np.random.RandomState(8).standard_normal((130, len(names))) * 3
    ref = np.array([oslik.evaluate(*x)[0] for x in X])
    assert relerr(slik.evaluate_batch(X), ref) < RTOL
    # the real SPT covariance (47 bins) through the same kernel
    g = golden("spt_lowl_lnl.npz")
    s2 = Slik(B.spt_script(spt_data, "s12", g["tt"])())
    s2.program(chi2_mode=mode)
    assert relerr(s2.evaluate_batch(g["s12_x"]), g["s12_lnl"]) < RTOL


@pytest.mark.parametrize("binning", ["dmma", "segsum"])
def test_both_binning_kernels_match_the_oracle(torch, binning):
    """top-hat (block-sparse) windows through the DMMA tile product with zero-block skipping and through
    the per-bin segmented sum"""
    prob = synthetic.planck_lite_problem(4)
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    prog = slik.program(binning=binning)
    assert (prog.meta["nsegclass"] > 0) == (binning == "segsum") and (prog.meta["nclass"] > 0) == (binning == "dmma")
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    X = x0 + sc * np.random.RandomState(9).standard_normal((200, len(names))) * 3
    ref = np.array([oslik.evaluate(*x)[0] for x in X])
    assert relerr(slik.evaluate_batch(X), ref) < RTOL


def test_tile_shapes_give_the_same_bits(torch, monkeypatch):
    """The chi^2 tile product has a 64-row bulk shape and a 32-row shape for small launches, picked by launch
    size: forcing either must not change a single bit (so neither does the number of GPUs a batch is sharded
    over), and both match the oracle."""
    prob = synthetic.spt_multifreq_problem(2, nbins=40, lmax=1201)       # dense windows, 240 rows, tiles of 40
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    prog = slik.program(chi2_mode="inverse")
    assert prog.meta["nclass"] > 0 and prog.meta["chi2_mode"] == 1
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    X = x0 + sc * np.random.RandomState(13).standard_normal((700, len(names))) * 2
    out = {}
    for rows in ("64", "32"):
        monkeypatch.setenv("CSL_TRIGEMM_ROWS", rows)
        out[rows] = slik.evaluate_batch(X)
    np.testing.assert_array_equal(out["64"], out["32"])
    monkeypatch.delenv("CSL_TRIGEMM_ROWS")
    np.testing.assert_array_equal(slik.evaluate_batch(X), out["64"])       # whatever the heuristic picks
    ref = np.array([oslik.evaluate(*x)[0] for x in X[:60]])
    assert relerr(out["32"][:60], ref) < RTOL


def test_segmented_binning_index_bounds_and_position_independence(torch):
    """The Taylor degree of the power-law recurrence is fixed per bin group from the host's bound on the
    index: a prior box is a hard bound, a parameter without one gets the soft bound 4 and walkers beyond it
    are re-run in the exact-exp mode.  Either way lnl matches the oracle and a walker's value does not depend
    on its position in the batch (bitwise)."""
    prob = synthetic.planck_lite_problem(6)

    def unboxed(ns):
        scr = synthetic.build_script(ns, prob)()
        for k in ("n_cib", "n_tilt"):
            scr[k].__dict__.pop("range", None)
        return scr

    oslik, slik = ort.Slik(unboxed(ORACLE)), Slik(unboxed(PRODUCT))
    prog = slik.program(binning="segsum")
    bounds = prog.host["spec_tab"][:, L.S_IDXBOUND:L.S_IDXBOUND + 2].view(np.float32)
    assert np.all((bounds == 0) | (bounds >= 4.0))                 # soft (positive) bounds only
    boxed = Slik(synthetic.build_script(PRODUCT, prob)()).program(binning="segsum")
    hb = boxed.host["spec_tab"][:, L.S_IDXBOUND:L.S_IDXBOUND + 2].view(np.float32)
    assert np.all(hb <= 0) and np.any(np.isclose(hb, -2.0)) and np.any(np.isclose(hb, -0.5))   # hard = negative
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    rs = np.random.RandomState(12)
    X = x0 + sc * rs.standard_normal((1024, len(names)))
    wild = rs.choice(1024, 37, replace=False)                      # indices far outside the soft bound
    X[wild, names.index("n_cib")] = rs.uniform(4.5, 6.0, wild.size) * rs.choice([-1, 1], wild.size)
    X[wild[:9], names.index("n_tilt")] = rs.uniform(4.1, 5.0, 9)
    got = slik.evaluate_batch(X)
    pick = np.concatenate([wild, rs.choice(1024, 40, replace=False)])
    ref = np.array([oslik.evaluate(*X[i])[0] for i in pick])
    assert relerr(got[pick], ref) < RTOL
    perm = rs.permutation(1024)
    np.testing.assert_array_equal(slik.evaluate_batch(X[perm]), got[perm])
    np.testing.assert_array_equal(slik.evaluate_batch(X[:300]), got[:300])        # another tile shape, same bits


def test_segmented_binning_wide_windows_take_several_staging_passes(torch):
    """Numerically dense windows forced through the segmented sum: a group of 8 bins has ~6000 columns, far
    more than one staging pass holds, so bins are cut by pass boundaries and keep their running sums and
    power-law state in registers across passes; the bins all start at column 0, so none inherits the state."""
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    prog = slik.program(binning="segsum")
    assert prog.meta["nsegclass"] > 0 and prog.meta["nclass"] == 0
    gt = prog.host["grp_tab"]
    assert gt[:, L.G_PRE + L.SEG_GROUP].max() > 3072 // 4            # more columns than any pass can stage
    assert not gt[:, L.G_CONT:L.G_CONT + L.SEG_GROUP].any()
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    X = x0 + sc * np.random.RandomState(14).standard_normal((300, len(names))) * 2
    ref = np.array([oslik.evaluate(*x)[0] for x in X[:50]])
    got = slik.evaluate_batch(X)
    assert relerr(got[:50], ref) < RTOL
    # and the DMMA tile product on the same windows agrees with it to rounding
    dm = Slik(synthetic.build_script(PRODUCT, prob)())
    dm.program(binning="dmma")
    assert relerr(dm.evaluate_batch(X), got) < 1e-12


def test_segmented_binning_with_ragged_walker_counts(torch):
    """walker counts that leave the last 256-walker CTA of k_bin_segsum half empty"""
    prob = synthetic.planck_lite_problem(5)
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    X = x0 + sc * np.random.RandomState(10).standard_normal((700, len(names)))
    ref = np.array([oslik.evaluate(*x)[0] for x in X[:40]] + [oslik.evaluate(*x)[0] for x in X[-40:]])
    for W in (700, 300, 129):
        got = slik.evaluate_batch(X[:W])
        assert relerr(got[:40], ref[:40]) < RTOL
        if W == 700:
            assert relerr(got[-40:], ref[40:]) < RTOL


def test_script_file_runs_through_the_launcher(torch, tmp_path):
    """python -m cosmoslik_b200 script.py --num_samples N  (load_script + argparse + drained sampler)"""
    import subprocess, sys as _sys
    from conftest import ROOT
    script = tmp_path / "myscript.py"
    out = tmp_path / "my.chain"
    script.write_text("""
from cosmoslik_b200 import *

class myscript(SlikPlugin):
    def __init__(self, ndim, num_samples=1000):
        super().__init__()
        ndim = self.ndim = int(ndim)
        for i in range(ndim):
            self["param%i" % i] = param(start=0, scale=1)
        self.sampler = samplers.metropolis_hastings(self, num_samples=num_samples, output_file="OUTFILE", nchains=8)

    def __call__(self):
        return sum([self["param%i" % i] ** 2 for i in range(self.ndim)]) / 2
""".replace("OUTFILE", str(out)))
    r = subprocess.run([_sys.executable, "-m", "cosmoslik_b200", str(script), "3", "--num_samples", "2400"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    chains = load_chain(str(out))
    assert len(chains) == 8 and set(chains[0].keys()) == {"lnl", "weight", "param0", "param1", "param2"}
    assert sum(c["weight"].sum() for c in chains) > 0.9 * 2400


def test_ragged_and_tiny_batches(torch, spt_data):
    """W not a multiple of the 128-walker tile, W = 1, and a batch where every walker is outside the prior"""
    g = golden("spt_lowl_lnl.npz")
    slik = Slik(B.spt_script(spt_data, "s12", g["tt"])())
    X, ref = g["s12_x"], g["s12_lnl"]
    for W in (1, 5, 96, 95):
        assert relerr(slik.evaluate_batch(X[:W]), ref[:W]) < RTOL
    bad = X[:3].copy()
    bad[:, 2] = -1.0
    assert np.all(np.isinf(slik.evaluate_batch(bad)))
    with pytest.raises(ValueError):
        slik.program().lnl(torch.zeros((4, 3), dtype=torch.float64, device="cuda"))


# ---------------------------------------------------------------- Metropolis-Hastings
def run_mh(script_factory, golden_name, **kw):
    g = golden(golden_name)
    n = len(g["u"])
    streams = dict(z=g["z"][:, None, :], u=g["u"][:, None])
    slik = Slik(script_factory(lambda s: samplers.metropolis_hastings(s, num_samples=n, streams=streams, **kw))())
    rows = [(s.lnl, s.weight, s.x) for s in slik.sample()]
    return g, slik, rows


def check_rows(rows, g, lnl_exact=False):
    lnl = np.array([r[0] for r in rows]); w = np.array([r[1] for r in rows]); x = np.array([r[2] for r in rows])
    assert x.shape == g["x"].shape, "different number of emitted rows: a decision flipped"
    np.testing.assert_array_equal(w, g["weight"])
    if lnl_exact:
        np.testing.assert_array_equal(lnl, g["lnl"])
    else:
        np.testing.assert_allclose(lnl, g["lnl"], rtol=RTOL)
    return x


def test_empty_batch_and_bad_shapes(torch, spt_data):
    g = golden("spt_lowl_lnl.npz")
    slik = Slik(B.spt_script(spt_data, "s12", g["tt"])())
    nd = len(slik.get_sampled())
    assert slik.evaluate_batch(np.empty((0, nd))).shape == (0,)
    assert tuple(slik.evaluate_batch(torch.empty((0, nd), dtype=torch.float64, device="cuda")).shape) == (0,)
    with pytest.raises(ValueError):
        slik.evaluate_batch(np.zeros((3, nd + 1)))
    with pytest.raises(ValueError):
        slik.evaluate_batch(torch.zeros((3, nd), dtype=torch.float32, device="cuda"))
    # NaN parameters: a NaN fails every prior box (priors.py:26-28) -> +inf, like the reference
    x = np.array(g["s12_x"][:4])
    x[1, 0] = np.nan
    out = slik.evaluate_batch(x)
    ref = np.array(g["s12_lnl"][:4])
    assert np.isfinite(out[[0, 2, 3]]).all() and relerr(out[[0, 2, 3]], ref[[0, 2, 3]]) < RTOL
    assert np.isinf(out[1]) or np.isnan(out[1])


def test_mh_quickstart_matches_the_reference_chain(torch, tmp_path):
    """config 1 script, generic per-step path; z @ F runs on the device so positions agree to
    rounding of a 2-term dot product, decisions and weights exactly"""
    out = str(tmp_path / "quick.chain")
    g, slik, rows = run_mh(B.quick_script, "mh_quick.npz", output_file=out)
    x = check_rows(rows, g)
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-13)
    # the chain file has the reference's single-process block structure (last row of each block dropped)
    c = load_chain(out)[0]
    assert len(c["lnl"]) == len(g["file_lnl"])
    np.testing.assert_array_equal(c["weight"], g["file_weight"])
    np.testing.assert_allclose(c["a"], g["file_a"], rtol=0, atol=1e-13)


def test_mh_exact_with_injected_offsets(torch):
    """with the proposal OFFSETS injected (one exact add on both sides) the chain contents are
    bit-identical to the reference's"""
    g = golden("mh_quick.npz")
    F = samplers.proposal_factor(np.eye(2), 2, 2.4)
    delta = (g["z"] @ F)[:, None, :]
    n = len(g["u"])
    slik = Slik(B.quick_script(lambda s: samplers.metropolis_hastings(
        s, num_samples=n, streams=dict(delta=delta, u=g["u"][:, None])))())
    rows = [(s.lnl, s.weight, s.x) for s in slik.sample()]
    x = check_rows(rows, g)
    np.testing.assert_array_equal(x, g["x"])
    # lnl: the reference squares numpy scalars through libm pow(x, 2.0), which is not always the
    # correctly rounded x*x the device computes -> agreement to 1 ulp, not bit for bit
    np.testing.assert_allclose([r[0] for r in rows], g["lnl"], rtol=4e-16, atol=0)


@pytest.mark.parametrize("fused", [True, False])
def test_mh_gauss30_matches_the_reference_chain(torch, fused):
    """config 2 likelihood (30 correlated parameters): whole-run fused kernel and generic path"""
    g = golden("mh_gauss30.npz")
    F = samplers.proposal_factor(np.eye(30), 30, 2.4)
    delta = (g["z"] @ F)[:, None, :]
    n = len(g["u"])
    slik = Slik(B.gauss_script(g["cov"], lambda s: samplers.metropolis_hastings(
        s, num_samples=n, streams=dict(delta=delta, u=g["u"][:, None]), proposal_update=False))())
    if not fused:
        slik.program().gauss_direct = None
    rows = [(s.lnl, s.weight, s.x) for s in slik.sample()]
    x = check_rows(rows, g)
    np.testing.assert_array_equal(x, g["x"])


def test_mh_spt_lowl_matches_the_reference_chain(torch, spt_data):
    """bandpower likelihood with hard-prior rejections, max_weight flushes"""
    g = golden("mh_spt.npz")
    cov = np.diag([0.026, 3, 3, 3]) ** 2
    F = samplers.proposal_factor(cov, 4, 2.4)
    delta = (g["z"] @ F)[:, None, :]
    n = len(g["u"])
    slik = Slik(B.spt_script(spt_data, "s12", None, lambda s: samplers.metropolis_hastings(
        s, num_samples=n, max_weight=4, streams=dict(delta=delta, u=g["u"][:, None])))())
    rows = [(s.lnl, s.weight, s.x) for s in slik.sample()]
    x = check_rows(rows, g)
    np.testing.assert_array_equal(x, g["x"])


def test_mh_lockstep_adaptive_matches_the_oracle(torch):
    """many chains, pooled proposal adaptation (device moments + host SVD) vs oracle.mh_lockstep"""
    rs = np.random.RandomState(3)
    nd, nch, nsteps = 6, 40, 900
    A = rs.standard_normal((nd, nd)); Cv = A @ A.T / nd + np.eye(nd)
    z = rs.standard_normal((nsteps, nch, nd)); u = rs.uniform(size=(nsteps, nch))
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, streams=dict(z=z, u=u), proposal_update_start=300,
        mpi_comm_freq=100))())
    got = {}
    for s in slik.sample():
        got.setdefault(s.chain, []).append((s.lnl, s.weight, s.x))
    oslik = ort.Slik(S.gauss_script(Cv)())
    lnl_batch = lambda X: np.array([oslik.evaluate(*x)[0] for x in X])
    rows, final, adapt = osamp.mh_lockstep(lnl_batch, np.zeros((nch, nd)), nsteps, np.eye(nd), z=z, u=u,
                                           proposal_update_start=300, exchange=100)
    assert len(adapt) == len(slik.sampler.adaptations) == 5
    for (t1, c1), (t2, c2) in zip(adapt, slik.sampler.adaptations):
        assert t1 == t2
        np.testing.assert_allclose(c2, c1, rtol=1e-9)
    for c in range(nch):
        assert len(got.get(c, [])) == len(rows[c]), "chain %d: a decision flipped" % c
        np.testing.assert_array_equal([r[1] for r in got[c]], [r[1] for r in rows[c]])
        np.testing.assert_allclose([r[2] for r in got[c]], [r[2] for r in rows[c]], rtol=0, atol=1e-9)
        np.testing.assert_allclose([r[0] for r in got[c]], [r[0] for r in rows[c]], rtol=1e-8)


def test_mh_device_streams_reproduce_the_oracle(torch):
    """throughput mode: Philox streams generated on the device give the chain the oracle gives
    when fed oracle/philox.py's statement of the same streams"""
    nd, nch, nsteps, seed = 5, 33, 150, 987654321
    Cv = np.eye(nd) + 0.3
    z = np.empty((nsteps, nch, nd)); u = np.empty((nsteps, nch))
    for t in range(nsteps):
        z[t], u[t] = ophilox.mh_streams(seed, t, np.arange(nch), nd)
    for fused in (True, False):
        slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
            s, num_samples=nsteps * nch, nchains=nch, seed=seed, proposal_update=False))())
        if not fused:
            slik.program().gauss_direct = None
        got = {}
        for s in slik.sample():
            got.setdefault(s.chain, []).append((s.lnl, s.weight, s.x))
        oslik = ort.Slik(S.gauss_script(Cv)())
        rows, _, _ = osamp.mh_lockstep(lambda X: np.array([oslik.evaluate(*x)[0] for x in X]), np.zeros((nch, nd)),
                                       nsteps, np.eye(nd), z=z, u=u, proposal_update=False)
        for c in range(nch):
            assert len(got.get(c, [])) == len(rows[c])
            np.testing.assert_array_equal([r[1] for r in got[c]], [r[1] for r in rows[c]])
            np.testing.assert_allclose([r[2] for r in got[c]], [r[2] for r in rows[c]], rtol=0, atol=1e-11)


def test_gelman_rubin_on_device_matches_the_oracle(torch):
    nd, nch, nsteps = 4, 16, 800
    Cv = np.eye(nd) + 0.2
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=77, proposal_update=False))())
    st = slik.sampler.run(slik.evaluate)
    rows, nrows = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
    burn = 20
    mom = [osamp.chain_moments(rows[c, burn:nrows[c], 2:], rows[c, burn:nrows[c], 1]) for c in range(nch)]
    ref = osamp.gelman_rubin([m[0] for m in mom], [m[1] for m in mom], [m[2] for m in mom])
    got = slik.sampler.gelman_rubin(burn_rows=burn)
    np.testing.assert_allclose(got, ref, rtol=1e-11)
    assert np.all(got < 1.2)


def test_prior_sampler_matches_the_reference(torch):
    from cosmoslik_b200 import SlikPlugin, param
    g = golden("prior_sampler.npz")

    class boxed(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = param(start=0, scale=1, range=(-2, 3))
            self.b = param(start=0, scale=1, uniform_prior=(0.5, 1.5), gaussian_prior=(1.0, 0.2))
            self.sampler = samplers.priors(self, num_samples=200, reseed=False)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    np.random.seed(int(g["seed"]))
    rows = [(s.lnl, s.weight, s.x) for s in Slik(boxed()).sample()]
    np.testing.assert_array_equal([r[2] for r in rows], g["x"])          # same points, bit for bit
    np.testing.assert_array_equal([r[1] for r in rows], g["weight"])
    np.testing.assert_allclose([r[0] for r in rows], g["lnl"], rtol=4e-16)

    class improper(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = param(start=0, scale=1)
            self.sampler = samplers.priors(self, num_samples=2)

        def __call__(self):
            return self.a ** 2

    with pytest.raises(Exception, match="improper"):
        improper()


def test_device_summary_matches_chain_statistics(torch):
    """weighted mean / covariance of all chains reduced on the device == Chains.join().mean() / .cov()"""
    from cosmoslik_b200 import Chain, Chains
    from cosmoslik_b200.chains import device_summary
    nd, nch, nsteps = 4, 12, 500
    Cv = np.eye(nd) + 0.3
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=21, proposal_update=False))())
    st = slik.sampler.run(slik.evaluate)
    rows, nrows = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
    names = list(slik.get_sampled())
    burn = 15
    chains = Chains([Chain(dict(zip(["lnl", "weight"] + names, rows[c, burn:nrows[c]].T))) for c in range(nch)])
    joined = chains.join()
    got = device_summary(st, burn_rows=burn)
    np.testing.assert_allclose(got["mean"], joined.mean(names), rtol=1e-12)
    np.testing.assert_allclose(got["cov"], joined.cov(names), rtol=1e-10)
    assert got["weight"] == joined["weight"].sum()
    np.testing.assert_allclose(got["std"], joined.std(names), rtol=1e-10)
    # Chain.add_gauss_prior on the device rows: one parameter with a standard deviation, then two with a covariance
    from cosmoslik_b200.chains import device_add_gauss_prior
    full = Chains([Chain(dict(zip(["lnl", "weight"] + names, rows[c, :nrows[c]].T))) for c in range(nch)])
    Cp = np.array([[0.5, 0.1], [0.1, 0.8]])
    want = Chains([c.add_gauss_prior(names[1], 0.2, 0.7).add_gauss_prior([names[0], names[3]], [0.1, -0.3], Cp)
                   for c in full])
    device_add_gauss_prior(st, names[1], 0.2, 0.7, names)
    device_add_gauss_prior(st, [names[0], names[3]], [0.1, -0.3], Cp, names)
    rows2 = st["rows"].cpu().numpy()
    for c in range(nch):
        np.testing.assert_allclose(rows2[c, :nrows[c], 0], want[c]["lnl"], rtol=1e-13)
        np.testing.assert_allclose(rows2[c, :nrows[c], 1], want[c]["weight"], rtol=1e-12)
        np.testing.assert_array_equal(rows2[c, :nrows[c], 2:], rows[c, :nrows[c], 2:])
    got2 = device_summary(st, burn_rows=0)
    np.testing.assert_allclose(got2["mean"], want.join().mean(names), rtol=1e-11)
    np.testing.assert_allclose(got2["cov"], want.join().cov(names), rtol=1e-9)


# ---------------------------------------------------------------- stretch move
def test_stretch_move_matches_the_oracle(torch, tmp_path):
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    k, nsteps = 96, 6
    rs = np.random.RandomState(0)
    streams = [(rs.uniform(size=(2, k // 2)), rs.randint(k // 2, size=(2, k // 2)), rs.uniform(size=(2, k // 2)))
               for _ in range(nsteps)]
    out = str(tmp_path / "emcee.chain")
    slik = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, streams=lambda i: streams[i], seed=3, output_file=out,
        output_freq=4))())
    pos0 = slik.sampler.initial_positions()
    slik.sampler.init_pos = pos0
    got = {}
    for s in slik.sample():
        got.setdefault(s.chain, []).append((s.lnl, s.weight, s.x))
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    lnprob = lambda q: np.array([-oslik.evaluate(*x)[0] for x in np.atleast_2d(q)])
    rows, opos, olnp = osamp.emcee_wrapper_rows(pos0, lnprob, nsteps, lambda i: streams[i])
    st = slik.sampler.stats
    np.testing.assert_array_equal(st["pos"].cpu().numpy(), opos)          # positions: bit exact
    assert relerr(st["lnl"].cpu().numpy(), -olnp) < RTOL
    assert 0 < st["accepted"] < k * nsteps
    for w in range(k):
        assert len(got[w]) == len(rows[w])
        np.testing.assert_array_equal([r[1] for r in got[w]], [r[1] for r in rows[w]])
        np.testing.assert_array_equal([r[2] for r in got[w]], [r[2] for r in rows[w]])
        np.testing.assert_allclose([r[0] for r in got[w]], [r[0] for r in rows[w]], rtol=RTOL)
    # chain file: emcee-style list header, per-walker records of <= output_freq rows
    chains = load_chain(out)
    assert len(chains) == k
    np.testing.assert_array_equal(chains[5]["weight"], [r[1] for r in rows[5]])


def test_stretch_device_streams_and_sampling(torch):
    """device-generated streams equal oracle/philox.py's; and the ensemble samples a Gaussian"""
    nd, k, nsteps, seed = 4, 64, 5, 4242
    Cv = np.eye(nd) * 0.5 + 0.5
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.emcee(s, num_samples=k * nsteps, nwalkers=k, seed=seed),
                               start=0.1)())
    pos0 = slik.sampler.initial_positions()
    slik.sampler.init_pos = pos0
    st = slik.sampler.run(slik.evaluate)
    oslik = ort.Slik(S.gauss_script(Cv)())
    lnprob = lambda q: np.array([-oslik.evaluate(*x)[0] for x in np.atleast_2d(q)])

    def streams(i):
        a = [ophilox.stretch_streams(seed, i, h, np.arange(k // 2) + h * (k // 2), k // 2) for h in (0, 1)]
        return (np.array([a[0][0], a[1][0]]), np.array([a[0][1], a[1][1]]), np.array([a[0][2], a[1][2]]))

    _, opos, olnp = osamp.emcee_wrapper_rows(pos0, lnprob, nsteps, streams)
    np.testing.assert_array_equal(st["pos"].cpu().numpy(), opos)
    # statistical check on a longer run
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.emcee(s, num_samples=512 * 400, nwalkers=512, seed=1),
                               start=0.1)())
    st = slik.sampler.run(slik.evaluate)
    p = st["pos"].cpu().numpy()
    assert np.all(np.abs(p.mean(0)) < 0.15)
    np.testing.assert_allclose(np.cov(p.T), Cv, atol=0.2)
    assert 0.3 < st["accepted"] / st["evals"] < 0.9


@pytest.mark.parametrize("chi2_mode,k", [("inverse", 300), ("trsm", 256)])
def test_native_fused_step_loop_matches_the_oracle_and_the_staged_kernels(torch, monkeypatch, chi2_mode, k):
    """csl_stretch_run (fused propose+prepare and combine+accept kernels, device Philox streams) on a binned
    bandpower likelihood: positions bit-exact against the oracle driven by oracle/philox.py's streams, and
    bitwise equal to the per-stage Python loop and to the unfused C loop (CSL_NO_FUSE=1); a walker count that
    leaves the 128-walker tiles ragged; both chi^2 kernels."""
    prob = synthetic.planck_lite_problem(7)
    nsteps, seed = 3, 99

    def run(python_loop=False):
        sl = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
            s, num_samples=k * nsteps, nwalkers=k, seed=seed))())
        sl.program(chi2_mode=chi2_mode)
        sl.sampler.init_pos = pos0
        if python_loop:
            sl.sampler.python_loop = True
        st = sl.sampler.run(sl.evaluate)
        return st["pos"].cpu().numpy(), st["lnl"].cpu().numpy(), st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()

    names = sorted(prob["truth"])
    x0 = np.array([prob["truth"][n] for n in names]); sc = np.array([prob["scales"][n] for n in names])
    pos0 = x0 + sc * np.random.RandomState(5).standard_normal((k, len(names)))
    native = run()
    staged = run(python_loop=True)
    monkeypatch.setenv("CSL_NO_FUSE", "1")
    unfused = run()
    monkeypatch.delenv("CSL_NO_FUSE")
    for a, b, c in zip(native, staged, unfused):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    lnprob = lambda q: np.array([-oslik.evaluate(*x)[0] for x in np.atleast_2d(q)])

    def streams(i):
        a = [ophilox.stretch_streams(seed, i, h, np.arange(k // 2) + h * (k // 2), k // 2) for h in (0, 1)]
        return (np.array([a[0][0], a[1][0]]), np.array([a[0][1], a[1][1]]), np.array([a[0][2], a[1][2]]))

    _, opos, olnp = osamp.emcee_wrapper_rows(pos0, lnprob, nsteps, streams)
    np.testing.assert_array_equal(native[0], opos)
    assert relerr(native[1], -olnp) < RTOL


# ---------------------------------------------------------------- full BASELINE sizes (properties)
def test_spt_multifreq_full_size_properties(torch):
    """BASELINE configs[2] at its full size (6 cross-spectra x 50 bins, ell 50..3300, dense windows, 16 384
    walkers) through the DMMA binning kernel: a random sample of walkers against the oracle, invariance under
    walker permutation (bitwise), and lnl = prior alone on noise-free data at the truth."""
    prob = synthetic.spt_multifreq_problem(0)
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    prog = slik.program()
    assert prog.meta["nclass"] > 0 and prog.meta["nsegclass"] == 0 and prog.meta["n"] == 300
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    rs = np.random.RandomState(3)
    W = 16384
    X = x0 + sc * rs.standard_normal((W, len(names)))
    got = slik.evaluate_batch(torch.from_numpy(X).cuda()).cpu().numpy()
    pick = rs.choice(W, 24, replace=False)
    ref = np.array([oslik.evaluate(*X[i])[0] for i in pick])
    assert relerr(got[pick], ref) < RTOL
    perm = rs.permutation(W)
    np.testing.assert_array_equal(slik.evaluate_batch(torch.from_numpy(X[perm]).cuda()).cpu().numpy(), got[perm])
    np.testing.assert_array_equal(slik.evaluate_batch(X[:1000]), got[:1000])          # a smaller launch, same bits


def test_planck_lite_full_size_properties(torch):
    """config 4 at 65 536 walkers: a random sample of walkers against the oracle, quadratic scaling
    of chi^2 along a ray from the best fit, and invariance under walker permutation."""
    prob = synthetic.planck_lite_problem(0)
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    names = list(slik.get_sampled())
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    rs = np.random.RandomState(0)
    W = 65536
    X = x0 + sc * rs.standard_normal((W, len(names)))
    Xd = torch.from_numpy(X).cuda()
    got = slik.evaluate_batch(Xd).cpu().numpy()
    pick = rs.choice(W, 48, replace=False)
    ref = np.array([oslik.evaluate(*X[i])[0] for i in pick])
    assert relerr(got[pick], ref) < RTOL
    perm = rs.permutation(W)
    np.testing.assert_array_equal(slik.evaluate_batch(torch.from_numpy(X[perm]).cuda()).cpu().numpy(), got[perm])
    # noise-free data at the truth point: residual 0 -> lnl is the prior term alone (= 0 at the prior centres)
    clean = dict(prob)
    m = synthetic.model_vector(prob, prob["truth"])
    sig, i = {}, 0
    for key in sorted(prob["use"]):
        nb = prob["windows"][key].shape[0]; sig[key] = m[i:i + nb]; i += nb
    clean["signal"] = sig
    s2 = Slik(synthetic.build_script(PRODUCT, clean)())
    l0 = s2.evaluate_batch(x0[None, :])[0]
    assert abs(l0) < 1e-12 * float(m @ m)
