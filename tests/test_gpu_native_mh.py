"""The native MH step loop (csl_mh_run / CUDA graph per block), the on-device proposal adaptation, yield_rejected
and the streamed row blocks -- against the staged kernels, the reference's golden chains and the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden
import b200_scripts as B
import oracle_scripts as S
from namespaces import PRODUCT
from cosmoslik_b200 import Slik, load_chain, samplers, synthetic
from cosmoslik_b200 import _lib as L
from oracle import runtime as ort

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def staged_mh(torch, slik, nch, nsteps, seed, max_weight=10.0, temp=1.0):
    """the per-stage kernels of round 1 (csl_mh_propose -> csl_lnl -> csl_mh_accept), driven from Python"""
    lib = L.lib()
    prog = slik.program()
    names = list(slik.get_sampled())
    nd = len(names)
    dev = prog.device
    kw = dict(dtype=torch.float64, device=dev)
    x0 = [slik.get_sampled()[k].start for k in names]
    cur_x = torch.tensor(x0, **kw).repeat(nch, 1).contiguous()
    cur_lnl = prog.lnl(cur_x)
    cur_w = torch.zeros(nch, **kw)
    test_x = torch.empty_like(cur_x)
    rows = torch.zeros((nch, nsteps, 2 + nd), **kw)
    nrows = torch.zeros(nch, dtype=torch.int32, device=dev)
    acc = torch.zeros(nch, dtype=torch.uint8, device=dev)
    F = torch.from_numpy(samplers.proposal_factor(slik.sampler.cov_est, nd, 2.4)).to(dev)
    st = L.stream_ptr()
    for t in range(nsteps):
        L.check(lib.csl_mh_propose(nch, nd, cur_x.data_ptr(), test_x.data_ptr(), 2, None, F.data_ptr(), seed, t, 0, st))
        test_lnl = prog.lnl(test_x)
        L.check(lib.csl_mh_accept(nch, nd, cur_x.data_ptr(), cur_lnl.data_ptr(), cur_w.data_ptr(), test_x.data_ptr(),
                                  test_lnl.data_ptr(), None, seed, t, 0, temp, max_weight, rows.data_ptr(),
                                  nrows.data_ptr(), nsteps, acc.data_ptr(), None, st))
    return cur_x.cpu().numpy(), cur_lnl.cpu().numpy(), cur_w.cpu().numpy(), rows.cpu().numpy(), nrows.cpu().numpy()


@pytest.mark.parametrize("graph", [True, False])
def test_native_loop_is_bitwise_the_staged_kernels(torch, spt_data, graph):
    """bandpower likelihood, device Philox streams, hard-prior rejections and max_weight flushes: the fused
    proposal+prepare / combine+accept kernels enqueued from C (plain loop, and one CUDA graph per 50-step block
    plus a tail) leave exactly the state and rows the per-stage kernels leave"""
    nch, nsteps, seed = 300, 130, 4242
    make = lambda: Slik(B.spt_script(spt_data, "s12", None, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=seed, max_weight=4, mpi_comm_freq=50, proposal_update=False))())
    slik = make()
    if not graph:
        slik.sampler.no_graph = True
    st = slik.sampler.run(slik.evaluate)
    assert ("graph" in st["loop"]) == graph
    if graph:
        assert st["graph_nodes"] >= 50 * 4
    x, lnl, w, rows, nrows = staged_mh(torch, make(), nch, nsteps, seed, max_weight=4.0)
    np.testing.assert_array_equal(st["cur_x"].cpu().numpy(), x)
    np.testing.assert_array_equal(st["cur_lnl"].cpu().numpy(), lnl)
    np.testing.assert_array_equal(st["cur_weight"].cpu().numpy(), w)
    got_n = st["nrows"].cpu().numpy()
    np.testing.assert_array_equal(got_n, nrows)
    got = st["rows"].cpu().numpy()
    for c in range(nch):
        np.testing.assert_array_equal(got[c, :nrows[c]], rows[c, :nrows[c]])
    assert nrows.sum() > nch and np.isinf(lnl).sum() == 0


@pytest.mark.parametrize("tag", ["quick", "spt"])
def test_yield_rejected_matches_the_reference(torch, spt_data, tag):
    """metropolis_hastings.py:160 through the native loop, against the unmodified reference's rows"""
    g = golden("mh_rejected.npz")
    sub = {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}
    if tag == "quick":
        cov, script = np.eye(2), lambda smp: B.quick_script(smp)
    else:
        cov, script = np.diag([0.026, 3, 3, 3]) ** 2, lambda smp: B.spt_script(spt_data, "s12", None, smp)
    nd = cov.shape[0]
    delta = (sub["z"] @ samplers.proposal_factor(cov, nd, 2.4))[:, None, :]
    n = len(sub["u"])
    slik = Slik(script(lambda s: samplers.metropolis_hastings(
        s, num_samples=n, max_weight=4, yield_rejected=True, streams=dict(delta=delta, u=sub["u"][:, None])))())
    rows = [(s.lnl, s.weight, s.x) for s in slik.sample()]
    assert len(rows) == len(sub["weight"]), "a decision flipped"
    np.testing.assert_array_equal([r[1] for r in rows], sub["weight"])
    np.testing.assert_array_equal(np.array([r[2] for r in rows]), sub["x"])
    lnl = np.array([r[0] for r in rows])
    fin = np.isfinite(sub["lnl"])
    np.testing.assert_array_equal(np.isfinite(lnl), fin)
    np.testing.assert_allclose(lnl[fin], sub["lnl"][fin], rtol=1e-10)


def test_device_adaptation_matches_the_host_path(torch):
    """the fused one-pass moments + device Cholesky factor against the two-pass moments + host SVD of round 1:
    same pooled covariance (the reference's get_new_cov) and a factor with F^T F = cov / ndim * scale^2"""
    lib = L.lib()
    rs = np.random.RandomState(5)
    nd, nch, cap = 7, 333, 40
    dev = torch.device("cuda")
    mean = rs.uniform(-50, 50, nd)
    A = rs.standard_normal((nd, nd))
    rows = np.zeros((nch, cap, 2 + nd))
    nrows = rs.randint(0, cap + 1, nch).astype(np.int32)
    rows[:, :, 1] = rs.randint(0, 11, (nch, cap))
    rows[:, :, 2:] = mean + rs.standard_normal((nch, cap, nd)) @ A.T
    d_rows, d_n = torch.from_numpy(rows).to(dev), torch.from_numpy(nrows).to(dev)
    m = 1 + nd + nd * nd
    part = torch.empty((592, m), dtype=torch.float64, device=dev)
    mom = torch.zeros(m, dtype=torch.float64, device=dev)
    ref = torch.from_numpy(mean + 3.0).to(dev)
    cov = torch.empty((nd, nd), dtype=torch.float64, device=dev)
    F = torch.empty((nd, nd), dtype=torch.float64, device=dev)
    mu = torch.empty(nd, dtype=torch.float64, device=dev)
    st = L.stream_ptr()
    L.check(lib.csl_mh_moments_fused(nch, nd, d_rows.data_ptr(), d_n.data_ptr(), cap, ref.data_ptr(), part.data_ptr(),
                                     mom.data_ptr(), st))
    L.check(lib.csl_mh_adapt_factor(nd, mom.data_ptr(), 2.4, ref.data_ptr(), cov.data_ptr(), F.data_ptr(), mu.data_ptr(), st))
    xs = [rows[c, nrows[c] // 2:nrows[c], 2:] for c in range(nch)]
    ws = [rows[c, nrows[c] // 2:nrows[c], 1] for c in range(nch)]
    want = ort.get_covariance(np.concatenate(xs), np.concatenate(ws))
    np.testing.assert_allclose(cov.cpu().numpy(), want, rtol=1e-10, atol=1e-12)
    Fh = F.cpu().numpy()
    np.testing.assert_allclose(Fh.T @ Fh, want / nd * 2.4 ** 2, rtol=1e-10, atol=1e-12)
    wm = np.concatenate(ws)
    np.testing.assert_allclose(mu.cpu().numpy(), (np.concatenate(xs) * wm[:, None]).sum(0) / wm.sum(), rtol=1e-12)
    # a direction nobody moved in: zero column instead of NaN
    rows[:, :, 2 + 3] = 1.25
    d_rows = torch.from_numpy(rows).to(dev)
    L.check(lib.csl_mh_moments_fused(nch, nd, d_rows.data_ptr(), d_n.data_ptr(), cap, ref.data_ptr(), part.data_ptr(),
                                     mom.data_ptr(), st))
    L.check(lib.csl_mh_adapt_factor(nd, mom.data_ptr(), 2.4, ref.data_ptr(), cov.data_ptr(), F.data_ptr(), None, st))
    Fh = F.cpu().numpy()
    assert np.isfinite(Fh).all() and np.abs(Fh[:, 3]).max() < 1e-4 * np.abs(Fh).max()


def test_adaptive_run_on_the_device_recovers_the_target(torch):
    """device Philox streams + device adaptation end to end: no host round trip in the loop, the pooled
    covariance converges to the target's"""
    nd, nch, nsteps = 6, 512, 1500
    A = np.random.RandomState(9).standard_normal((nd, nd)); Cv = A @ A.T / nd + np.eye(nd)
    for generic in (False, True):
        slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
            s, num_samples=nsteps * nch, nchains=nch, seed=31, proposal_update_start=300, mpi_comm_freq=100))())
        if generic:
            slik.program().gauss_direct = None
        st = slik.sampler.run(slik.evaluate)
        ad = slik.sampler.adaptations
        assert len(ad) == 11 and [t for t, _ in ad] == list(range(400, 1500, 100))
        np.testing.assert_allclose(ad[-1][1], Cv, rtol=0.15, atol=0.05)
        np.testing.assert_array_equal(slik.sampler.cov_est, ad[-1][1])
        rows, nrows = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
        x = np.concatenate([rows[c, nrows[c] // 2:nrows[c], 2:] for c in range(nch)])
        w = np.concatenate([rows[c, nrows[c] // 2:nrows[c], 1] for c in range(nch)])
        np.testing.assert_allclose(ort.get_covariance(x, w), Cv, rtol=0.1, atol=0.05)


def test_emcee_streamed_file_equals_the_device_history(torch, tmp_path):
    """output_file: rows leave the device in blocks of output_freq steps (ring buffer + pinned copies); the file
    must hold exactly the rows a run that keeps its whole history on the device has"""
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    k, nsteps = 160, 23
    make = lambda **kw: Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, seed=11, **kw))())
    hist = make()
    st = hist.sampler.run(hist.evaluate)
    rows, nrows = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
    out = str(tmp_path / "ens.chain")
    streamed = make(output_file=out, output_freq=5)
    st2 = streamed.sampler.run(streamed.evaluate)
    assert st2["rows"] is None and st2["row_storage_bytes"] < rows.nbytes
    np.testing.assert_array_equal(st2["pos"].cpu().numpy(), st["pos"].cpu().numpy())
    chains = load_chain(out)
    assert len(chains) == k
    names = list(hist.get_sampled())
    for c in range(k):
        got = np.column_stack([chains[c]["lnl"], chains[c]["weight"]] + [chains[c][n] for n in names])
        np.testing.assert_array_equal(got, rows[c, :nrows[c]])
    # the generator protocol sees the same rows
    gen = make(output_freq=7)
    per = {}
    for s in gen.sample():
        per.setdefault(s.chain, []).append(np.hstack([s.lnl, s.weight, s.x]))
    for c in range(k):
        np.testing.assert_array_equal(np.array(per[c]), rows[c, :nrows[c]])


def test_emcee_row_storage_guard(torch):
    """a run whose whole row history cannot fit asks for streaming instead of failing inside an allocation"""
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    slik = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=64 * 10 ** 9, nwalkers=64, seed=1))())
    with pytest.raises(MemoryError, match="output_file"):
        slik.sampler.run(slik.evaluate)


def test_lnl_host_two_streams_equals_the_single_batch(torch, spt_data):
    """csl_lnl_host evaluates the two half-batches on two streams (copies under kernels): same bits as one batch"""
    g = golden("spt_lowl_lnl.npz")
    slik = Slik(B.spt_script(spt_data, "s12", g["tt"])())
    rs = np.random.RandomState(2)
    X = g["s12_x"][rs.randint(0, len(g["s12_x"]), 1500)] * (1 + 1e-3 * rs.standard_normal((1500, 4)))
    a = slik.evaluate_batch(X)
    b = slik.evaluate_batch(torch.from_numpy(X).cuda()).cpu().numpy()
    np.testing.assert_array_equal(a, b)
    c = np.concatenate([slik.evaluate_batch(X[:100]), slik.evaluate_batch(X[100:])])    # 100 < 4 tiles: single stream
    np.testing.assert_array_equal(a, c)


def test_device_thin_and_reweighted_match_the_chain_methods(torch):
    """chains.py:235-243 (thin) and :178-203 (reweighted) applied to every chain's rows on the device"""
    from cosmoslik_b200 import Chain, Chains
    from cosmoslik_b200.chains import device_reweighted, device_thin
    from cosmoslik_b200.symbolic import exp
    nd, nch, nsteps = 3, 37, 600
    Cv = np.eye(nd) + 0.3
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=12, proposal_update=False))())
    st = slik.sampler.run(slik.evaluate)
    names = list(slik.get_sampled())
    rows, nrows = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
    full = Chains([Chain(dict(zip(["lnl", "weight"] + names, rows[c, :nrows[c]].T))) for c in range(nch)])
    # reweighted by a function of the parameters (traced once, evaluated by the bytecode kernel)
    f_host = lambda **p: np.exp(-(p[names[0]] - 0.3) ** 2 / 2 / 0.5 ** 2)
    f_sym = lambda **p: exp(-(p[names[0]] - 0.3) ** 2 / 2 / 0.5 ** 2)
    want = Chains([c.reweighted(lambda weight, lnl, **p: f_host(**p)) for c in full])
    device_reweighted(st, names, f_sym)
    r2 = st["rows"].cpu().numpy()
    for c in range(nch):
        np.testing.assert_allclose(r2[c, :nrows[c], 1], want[c]["weight"], rtol=1e-13)
        np.testing.assert_array_equal(r2[c, :nrows[c], 2:], rows[c, :nrows[c], 2:])
    # thin: integer weights again (the MH weights), every chain on its own
    st["rows"].copy_(torch.from_numpy(rows).cuda())
    for delta in (3, 7.5):
        st["rows"].copy_(torch.from_numpy(rows).cuda()); st["nrows"].copy_(torch.from_numpy(nrows).cuda())
        device_thin(st, delta)
        r3, n3 = st["rows"].cpu().numpy(), st["nrows"].cpu().numpy()
        for c in range(nch):
            t = full[c].thin(delta)
            assert n3[c] == t.length()
            np.testing.assert_array_equal(r3[c, :n3[c], 1], t["weight"])
            np.testing.assert_array_equal(r3[c, :n3[c], 0], t["lnl"])
            np.testing.assert_array_equal(r3[c, :n3[c], 2], t[names[0]])
