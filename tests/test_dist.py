"""Walker sharding: world_size-2 runs through torch.distributed.run (gloo).  The CPU test covers the
host-side helpers; the GPU tests run both ranks on cuda:0 and check that the sharded samplers give the
single-process result -- bit for bit for the stretch move (streams are keyed by global walker id)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def launch(mode, *args, nproc=2, port=29731):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           mode] + list(args)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_dist_helpers_world_size_2():
    launch("helpers", port=29741)


def test_dist_helpers_world_size_3():
    launch("helpers", nproc=3, port=29742)


@pytest.mark.gpu
def test_sharded_stretch_move_is_bitwise_the_single_process_run(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from namespaces import PRODUCT
    from cosmoslik_b200 import Slik, samplers, synthetic
    out = str(tmp_path / "stretch_rank%d.npz")
    launch("stretch", out, port=29743)
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    k, nsteps = 192, 5
    slik = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, seed=11))())
    st = slik.sampler.run(slik.evaluate)
    pos, lnl = st["pos"].cpu().numpy(), st["lnl"].cpu().numpy()
    seen, acc = [], 0
    for r in range(2):
        z = np.load(out % r)
        ids = z["ids"]
        np.testing.assert_array_equal(z["pos"], pos[ids])
        np.testing.assert_array_equal(z["lnl"], lnl[ids])
        np.testing.assert_array_equal(z["nrows"], st["nrows"].cpu().numpy()[ids])
        seen += list(ids); acc += int(z["accepted"])
    assert sorted(seen) == list(range(k)) and acc == st["accepted"]


@pytest.mark.gpu
def test_sharded_adaptive_mh_matches_the_single_process_run(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import b200_scripts as B
    from cosmoslik_b200 import Slik, samplers
    out = str(tmp_path / "mh_rank%d.npz")
    launch("mh", out, port=29744)
    nd, nch, nsteps = 5, 24, 700
    A = np.random.RandomState(3).standard_normal((nd, nd)); Cv = A @ A.T / nd + np.eye(nd)
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=5, proposal_update_start=200, mpi_comm_freq=100))())
    st = slik.sampler.run(slik.evaluate)
    x = st["cur_x"].cpu().numpy()
    for r in range(2):
        z = np.load(out % r)
        assert int(z["nadapt"]) == len(slik.sampler.adaptations) == 4
        # the pooled covariance is summed in a different order across ranks: equal to rounding
        np.testing.assert_allclose(z["cov"], slik.sampler.cov_est, rtol=1e-9)
        np.testing.assert_allclose(z["cur_x"], x[int(z["lo"]):int(z["hi"])], rtol=0, atol=1e-7)


@pytest.mark.gpu
def test_all_ranks_write_one_chain_file(tmp_path):
    """metropolis_hastings.py:181-186, 209-212 ("everything still gets dumped into one file"): with two ranks,
    load_chain(output_file) must return ALL walkers / chains, with the rows of the single-process run"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import b200_scripts as B
    from namespaces import PRODUCT
    from cosmoslik_b200 import Slik, load_chain, samplers, synthetic
    # ---- stretch move
    out = str(tmp_path / "ens.chain")
    launch("stretch_file", out, port=29745)
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    k, nsteps = 194, 9
    single = str(tmp_path / "ens1.chain")
    slik = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, seed=11, output_file=single, output_freq=4))())
    slik.sampler.run(slik.evaluate)
    got, want = load_chain(out), load_chain(single)
    assert len(got) == len(want) == k
    for a, b in zip(got, want):
        assert list(a.keys()) == list(b.keys())
        for key in a:
            np.testing.assert_array_equal(a[key], b[key])
    yielded = [int(np.load(out + ".rank%d.npz" % r)["n"]) for r in range(2)]
    assert sum(yielded) == sum(c.length() for c in want) and min(yielded) > 0
    assert not os.path.exists(out + ".rank1")
    # ---- Metropolis-Hastings
    out = str(tmp_path / "mh.chain")
    launch("mh_file", out, port=29746)
    nd, nch, nsteps = 4, 11, 230
    Cv = np.eye(nd) + 0.25
    single = str(tmp_path / "mh1.chain")
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=6, proposal_update=False, mpi_comm_freq=50, output_file=single))())
    slik.sampler.run(slik.evaluate)
    got, want = load_chain(out), load_chain(single)
    assert len(got) == len(want) == nch
    for a, b in zip(got, want):
        for key in a:
            np.testing.assert_array_equal(a[key], b[key])
