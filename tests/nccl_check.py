"""Multi-GPU check (run with torchrun on >= 2 GPUs, NCCL): the sharded stretch move -- with the
symmetric-memory replica updated by the accept kernel, and with plain NCCL all_gather -- must be
bitwise the single-process run.  Prints one line per mode; exit code 0 on success."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cosmoslik_b200 import Slik, dist, samplers, synthetic   # noqa: E402
from namespaces import PRODUCT                               # noqa: E402


def run(k, nsteps):
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)
    slik = Slik(synthetic.build_script(PRODUCT, prob, sampler=lambda s: samplers.emcee(
        s, num_samples=k * nsteps, nwalkers=k, seed=11))())
    st = slik.sampler.run(slik.evaluate)
    return st["pos"].cpu().numpy(), st["lnl"].cpu().numpy(), st["walker_ids"], st["accepted"]


def main():
    rank, size = dist.init_from_env()
    k, nsteps = 1024 + 2 * size, 6        # ragged shares when size does not divide k/2
    results = {}
    for mode in ("symm", "nccl"):
        os.environ["COSMOSLIK_B200_NO_SYMM"] = "0" if mode == "symm" else "1"
        results[mode] = run(k, nsteps)
    # single-process reference on every rank: pretend there is no process group
    saved = (dist.get_rank, dist.get_size, dist._dist)
    dist.get_rank, dist.get_size, dist._dist = (lambda: 0), (lambda: 1), (lambda: None)
    ref_pos, ref_lnl, ref_ids, ref_acc = run(k, nsteps)
    dist.get_rank, dist.get_size, dist._dist = saved
    ok = True
    for mode, (pos, lnl, ids, acc) in results.items():
        same = np.array_equal(pos, ref_pos[ids]) and np.array_equal(lnl, ref_lnl[ids])
        tot = torch.tensor([acc], device="cuda"); torch.distributed.all_reduce(tot)
        same = same and int(tot.item()) == ref_acc
        print("rank %d mode %s: %s" % (rank, mode, "bitwise equal to the single-process run" if same else "MISMATCH"))
        ok = ok and same
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
