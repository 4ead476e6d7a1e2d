#!/usr/bin/env python
"""Drive the HBM-bound kernels of the path (proposal, priors + coefficient bytecode, accept, row
bookkeeping) on the headline workload so that ncu can capture them:

  ncu --set full --clock-control none --import-source on \
      -k regex:'k_stretch_propose_prepare|k_stretch_combine_accept|k_emcee_rows|k_mh_propose|k_mh_accept|k_prepare' \
      --launch-skip 12 -c 12 -o gpurun_out/hbm_kernels python tools/profile_hbm_kernels.py

Without ncu it prints CUDA-event durations of the same kernels (per-stage timers of the staged loops)
and the algorithmic bytes per launch, i.e. the achieved HBM GB/s of each.
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import cosmoslik_b200 as cb
    from cosmoslik_b200 import synthetic
    torch.cuda.set_device(0)
    W = int(os.environ.get("CSL_PROFILE_WALKERS", "65536"))
    prob = synthetic.planck_lite_problem(0)
    ns = types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl,
                               egfs=cb.models.egfs)
    names = sorted(prob["truth"])
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    nd = len(names)

    # stretch move: native fused loop (k_stretch_propose_prepare, k_stretch_combine_accept, k_emcee_rows)
    slik = cb.Slik(synthetic.build_script(ns, prob, sampler=lambda s: cb.samplers.emcee(
        s, num_samples=W, nwalkers=W, seed=1234))())
    prog = slik.program()
    smp = slik.sampler
    smp.init_pos = x0 + sc * np.random.RandomState(99).standard_normal((W, nd))
    short = os.environ.get("CSL_PROFILE_SHORT") is not None      # under ncu: as few launches as possible
    smp.row_capacity = 16
    smp.run(prog, nsteps=1 if short else 3)
    smp.run(prog, nsteps=2 if short else 3, resume=True)
    # staged loop with a CUDA-event pair around every stage
    timers = {}
    smp.timers = timers
    smp.run(prog, nsteps=1 if short else 5, resume=True)
    torch.cuda.synchronize()
    del smp["timers"]
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}

    # MH over the same likelihood: k_mh_propose, k_prepare, k_mh_accept (one chain per walker)
    try:
        slik2 = cb.Slik(synthetic.build_script(ns, prob, sampler=lambda s: cb.samplers.metropolis_hastings(
            s, num_samples=(2 if short else 6) * W, nchains=W, seed=5, proposal_update=False))())
        slik2._program, slik2._program_options = prog, slik._program_options
        slik2.sampler.run(prog)
        torch.cuda.synchronize()
    except Exception as e:          # the stretch-move numbers above are still worth printing
        sys.stderr.write("MH leg failed: %r\n" % (e,))

    half = W // 2
    ncoef = int(prog.meta.get("ncoef", 0))
    alg = {  # algorithmic bytes per launch: rows read + rows written + scalars
        "propose": half * (3 * nd * 8 + 16),              # own row + partner row in, proposal out, u_z / z
        "prepare": half * (nd * 8 + (ncoef + 1) * 8),     # x in, coefficients + prior out
        "combine": half * (2 + 1) * 8,
        "accept": half * (3 * nd * 8 + 32),               # x, q in, x out, lnl / lnl_q / u in, lnl out
    }
    out = {"walkers": W, "ndim": nd, "ncoef": ncoef, "stage_ms": stage_ms,
           "achieved_GBps": {k: alg[k] / (stage_ms[k] * 1e-3) / 1e9 for k in alg if k in stage_ms},
           "algorithmic_bytes_per_launch": alg}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
