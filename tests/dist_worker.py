"""Worker for the multi-process tests (launched with torch.distributed.run, backend gloo).

mode 'helpers' (CPU): checks cosmoslik_b200.dist on CPU tensors.
mode 'stretch' / 'mh' (needs a GPU; both ranks share cuda:0, collectives through gloo): runs the sharded
sampler and saves each rank's result so the parent test can compare it with the single-process run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cosmoslik_b200 import dist  # noqa: E402


def helpers():
    rank, size = dist.get_rank(), dist.get_size()
    assert size == int(os.environ["WORLD_SIZE"]) and dist.is_master() == (rank == 0)
    for n in (10, 11, 7, 2, 1):
        covered = []
        for r in range(size):
            lo, hi = dist.shard_bounds(n, r, size)
            covered += list(range(lo, hi))
        assert covered == list(range(n))
    t = torch.full((5,), float(rank + 1), dtype=torch.float64)
    dist.all_reduce_sum(t)
    assert torch.all(t == sum(range(1, size + 1)))
    # even and ragged all-gather of row blocks
    for total in (8, 7):
        lo, hi = dist.shard_bounds(total, rank, size)
        local = torch.arange(lo, hi, dtype=torch.float64)[:, None].repeat(1, 3)
        full = dist.all_gather_rows(local, total)
        assert full.shape == (total, 3) and torch.all(full[:, 0] == torch.arange(total, dtype=torch.float64))
    dist.barrier()


def stretch(out):
    import b200_scripts as B  # noqa: F401
    from namespaces import PRODUCT
    from cosmoslik_b200 import Slik, samplers, synthetic
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    k, nsteps = 192, 5
    slik = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, seed=11))())
    st = slik.sampler.run(slik.evaluate)
    np.savez(out % dist.get_rank(), pos=st["pos"].cpu().numpy(), lnl=st["lnl"].cpu().numpy(), ids=st["walker_ids"],
             nrows=st["nrows"].cpu().numpy(), weight=st["weight"].cpu().numpy(), accepted=st["accepted"])


def mh(out):
    import b200_scripts as B
    from cosmoslik_b200 import Slik, samplers
    nd, nch, nsteps = 5, 24, 700
    A = np.random.RandomState(3).standard_normal((nd, nd)); Cv = A @ A.T / nd + np.eye(nd)
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=5, proposal_update_start=200, mpi_comm_freq=100))())
    st = slik.sampler.run(slik.evaluate)
    lo, hi = st["chains"]
    np.savez(out % dist.get_rank(), cur_x=st["cur_x"].cpu().numpy(), lo=lo, hi=hi,
             nrows=st["nrows"].cpu().numpy(), cov=slik.sampler.cov_est, nadapt=len(slik.sampler.adaptations))


def stretch_file(out):
    """both ranks run with the SAME output_file: rank 0 must write every rank's walkers into it"""
    from namespaces import PRODUCT
    from cosmoslik_b200 import Slik, samplers, synthetic
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    k, nsteps = 194, 9                      # 97 walkers per half: ragged shares over 2 ranks
    slik = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, seed=11, output_file=out, output_freq=4))())
    n = sum(1 for _ in slik.sample())       # the generator yields this rank's walkers only
    np.savez(out + ".rank%d.npz" % dist.get_rank(), n=n, ids=slik.sampler.stats["walker_ids"])


def mh_file(out):
    import b200_scripts as B
    from cosmoslik_b200 import Slik, samplers
    nd, nch, nsteps = 4, 11, 230
    Cv = np.eye(nd) + 0.25
    slik = Slik(B.gauss_script(Cv, lambda s: samplers.metropolis_hastings(
        s, num_samples=nsteps * nch, nchains=nch, seed=6, proposal_update=False, mpi_comm_freq=50, output_file=out))())
    slik.sampler.run(slik.evaluate)


if __name__ == "__main__":
    mode = sys.argv[1]
    d.init_process_group("gloo")
    if mode == "helpers":
        helpers()
    else:
        torch.cuda.set_device(0)
        {"stretch": stretch, "mh": mh, "stretch_file": stretch_file, "mh_file": mh_file}[mode](sys.argv[2])
    d.destroy_process_group()
