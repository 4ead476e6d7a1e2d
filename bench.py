#!/usr/bin/env python
"""Headline benchmark: posterior evaluations / second for the batched bandpower lnL.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]

Workload (BASELINE.json configs[3], the one the metric is quoted on): Planck-lite-sized TT/TE/EE
bandpower likelihood, 613 bins, dense FP64 covariance, 9 sampled parameters, stretch-move ensemble.
One "step" = one full stretch-move step (both half-ensemble updates, each = proposal + priors +
foreground build + DMMA binning + DMMA triangular solve + accept) over all walkers = one posterior
evaluation per walker.  Scaling is weak: --walkers walkers PER GPU (default 65536), sharded row-wise;
the complementary half is all-gathered over NCCL before each half-update.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import multiprocessing
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "posterior evals/sec (whole box) for batched bandpower lnL"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="planck_lite", choices=["planck_lite", "spt_multifreq", "gauss30_mh", "unbinned_mh"])
    ap.add_argument("--walkers", type=int, default=65536, help="walkers per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--binning", default="auto", choices=["auto", "dmma", "segsum"],
                    help="window-binning kernel: DMMA tile product, per-bin segmented sum, or by window structure")
    ap.add_argument("--chi2", default="auto", choices=["auto", "trsm", "inverse"])
    return ap.parse_args()


def make_problem(name):
    from cosmoslik_b200 import synthetic
    return synthetic.planck_lite_problem(0) if name == "planck_lite" else synthetic.spt_multifreq_problem(0)


def config_of(args, prob, nwalkers_total):
    win = ("top-hat plik-lite bins (block-sparse window matrix; binning kernel: %s)" % args.binning
           if args.workload == "planck_lite" else "Gaussian bumps + 1e-10 tails (numerically dense windows)")
    return {"workload": "BASELINE configs[3] planck_lite TT/TE/EE 613 bins" if args.workload == "planck_lite"
            else "BASELINE configs[2] spt_multifreq 6 spectra x 50 bins",
            "n_data": int(prob["n"]), "ndim": len(prob["truth"]), "sampler": "stretch move a=2",
            "walkers_total": int(nwalkers_total), "walkers_per_gpu": int(args.walkers),
            "windows": win, "l2": "working set (R matrix %d MB per half-step) exceeds L2; no explicit flush"
            % (prob["n"] * args.walkers // 2 * 8 // 2 ** 20), "parallelism": "walker-sharded x%d" % args.gpus}


# ------------------------------------------------------------------------------------- oracle arm
def _oracle_slik(workload):
    from cosmoslik_b200 import synthetic
    from oracle import likes as olikes, runtime as ort
    ns = types.SimpleNamespace(SlikPlugin=ort.Plugin, param=ort.Param, mspec_lnl=olikes.MspecLike, egfs=olikes.Egfs)
    prob = make_problem(workload)
    return prob, ort.Slik(synthetic.build_script(ns, prob)())


_worker_state = {}


def _limit_blas_threads():
    """one BLAS thread per process, as BASELINE.md's plan prescribes (OMP_NUM_THREADS=1)"""
    try:
        from threadpoolctl import threadpool_limits
        _worker_state["blas_limit"] = threadpool_limits(1)
    except Exception:
        pass


def _worker_init(workload):
    _limit_blas_threads()
    _worker_state["prob"], _worker_state["slik"] = _oracle_slik(workload)


def _worker_eval(X):
    slik = _worker_state["slik"]
    return [slik.evaluate(*x)[0] for x in X]


def _sample_points(prob, n, seed=1):
    names = sorted(prob["truth"])
    x0 = np.array([prob["truth"][k] for k in names])
    sc = np.array([prob["scales"][k] for k in names])
    return x0 + sc * np.random.RandomState(seed).standard_normal((n, len(names)))


def cpu_baseline_single(workload, budget_s):
    """the oracle port (numpy/scipy, scalar per-walker calls like the reference) on ONE host core"""
    _limit_blas_threads()
    prob, slik = _oracle_slik(workload)
    X = _sample_points(prob, 4096)
    slik.evaluate(*X[0])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and n < len(X):
        slik.evaluate(*X[n]); n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d oracle Slik.evaluate calls of the same script in %.1f s (1 core, OMP_NUM_THREADS=1)" % (n, dt)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the reference is pure Python and
    cannot travel to the GPU box) on all host cores, a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    prob = make_problem(args.workload)
    per_core = 24
    X = _sample_points(prob, per_core * cores)
    chunks = np.array_split(X, cores)
    ctx = multiprocessing.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(args.workload,)) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_worker_eval, chunks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_eval, chunks)
        dt = time.perf_counter() - t0
    value = len(X) * args.steps / dt
    sample = "%d evaluations per step (%d per core) through the oracle's Slik.evaluate, Pool(%d)" % (len(X), per_core, cores)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_of(args, prob, args.walkers * args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region (NVML, 20 ms period; nvidia-smi fallback)."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake", 0x80), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.max_mhz, self.nv, self.h = None, None, None
        try:                                  # NVML is initialised BEFORE the timed region
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.nv = nv
        except Exception:
            self.nv = None

    def _run(self):
        nv, h = self.nv, self.h
        if nv is not None:
            while not self.stop:
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    try:
                        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    self.samples.append((mhz, mask))
                except Exception:
                    pass
                time.sleep(0.05)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.max_mhz = float(f[1])
                self.samples.append((float(f[0]), int(f[2], 16)))
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.thread.start(); return self

    def __exit__(self, *a):
        self.stop = True; self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        mask = 0
        for s in self.samples:
            mask |= s[1]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.REASONS if mask & bit], "samples": len(sm)}


# ------------------------------------------------------------------------------------- B200 arm
def fp64_peak():
    """FP64 tensor roofline denominator: measured cuBLAS DGEMM on this pool's B200 (tools/measure_peaks.py,
    committed under profiles/); MEASURED_PEAKS.json has no FP64 figure."""
    path = os.path.join(ROOT, "profiles", "measured_fp64_peak.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d["fp64_tflops_sustained"], "measured cuBLAS DGEMM sustained (profiles/measured_fp64_peak.json)"
    return 37.0, "fallback: B200 datasheet FP64 37 TFLOP/s (no measurement yet)"


def run_b200(args):
    import torch
    import cosmoslik_b200 as cb
    from cosmoslik_b200 import dist, synthetic
    rank, size = dist.init_from_env()
    if size == 1:
        torch.cuda.set_device(0)
    dev = torch.cuda.current_device()
    prob = make_problem(args.workload)
    ns = types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl,
                               egfs=cb.models.egfs)
    k_total = args.walkers * size
    sampler = lambda s: cb.samplers.emcee(s, num_samples=k_total, nwalkers=k_total, seed=1234)
    slik = cb.Slik(synthetic.build_script(ns, prob, sampler=sampler)())
    prog = slik.program(chi2_mode=args.chi2, binning=args.binning)
    smp = slik.sampler
    names = list(slik.get_sampled())
    # tight initial ball around the truth (a burnt-in ensemble), identical on every rank
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    smp.init_pos = x0 + sc * np.random.RandomState(99).standard_normal((k_total, len(names)))

    # ---- device-resident timing: K full stretch steps, CUDA events, max over ranks
    smp.row_capacity = args.steps            # same row buffer for the warm-up and the timed run
    smp.run(prog, nsteps=max(args.warmup, 3))      # warm-up: start positions evaluated, replica built
    smp.python_loop = True                          # ... and the instrumented Python loop's code paths
    smp.run(prog, nsteps=3, resume=True)
    del smp["python_loop"]
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev) as clk:
        e0.record()
        smp.run(prog, nsteps=args.steps, resume=True)     # exactly K steps, continuing the warmed-up ensemble
        e1.record()
        torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if size > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(ms.item())
    value = k_total * args.steps / (ms * 1e-3)
    acc_frac = smp.stats["accepted"] / float(smp.stats["evals"])
    host_ms = smp.stats.get("host_ms_per_step")
    # ---- second pass of K more steps with a CUDA-event pair around every stage (this disables the C-side
    # step loop: each stage is launched from Python between events), for the per-kernel durations
    timers = {}
    smp.timers = timers
    dist.barrier(); torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    smp.run(prog, nsteps=args.steps, resume=True)
    f1.record(); torch.cuda.synchronize()
    ms_instrumented = f0.elapsed_time(f1) / args.steps
    del smp["timers"]

    # ---- per-kernel durations from the event pairs recorded inside the timed region
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}
    stage_n = {k: len(v) for k, v in timers.items()}
    half = args.walkers // 2
    fl = prog.flops
    peak, peak_src = fp64_peak()
    # algorithmic FP64 work per launch (DESIGN.md "Kernels"): chi2 = n^2 per walker (one triangular solve,
    # SURVEY.md 8d); DMMA binning = 2*nnz(windows) + foreground build (2 flops per template FMA, a power law
    # counted as 20 FMAs) per walker, "executed" adds padding / zero blocks actually issued; segmented-sum
    # binning = 2 x the FP64 instructions it issues per walker (program.flops["seg_flops"]).
    per_kernel = {}
    for name, kern, alg, exe in (("chi2", "k_trigemm_chi2" if prog.meta.get("chi2_mode") else "k_trsm_chi2", fl["trsm_alg"], fl["trsm"]),
                                 ("bin_residual", "k_bin_segsum", fl["seg_flops"], fl["seg_flops"])
                                 if prog.meta.get("nsegclass") and not prog.meta.get("nclass") else
                                 ("bin_residual", "k_bin_residual", fl["bin_nnz"] + fl["build_ops"],
                                  fl["bin_exec"] + fl["build_ops"])):
        if name in stage_ms:
            t = stage_ms[name] * 1e-3
            per_kernel[kern] = {"ms_per_launch": stage_ms[name], "launches": stage_n[name],
                                "algorithmic_tflops": alg * half / t / 1e12, "executed_tflops": exe * half / t / 1e12,
                                "frac_of_fp64_peak": alg * half / t / 1e12 / peak}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"])
    traffic = None     # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f).get(dom)
        if tj and tj["walkers_per_launch"] == half:
            traffic = tj["dram_bytes_per_launch"]
    roofline = {"bound": "tensor", "kernel": dom, "achieved": per_kernel[dom]["algorithmic_tflops"], "peak": peak,
                "unit": "TFLOP/s", "frac": per_kernel[dom]["frac_of_fp64_peak"], "traffic": traffic,
                "peak_source": peak_src, "kernels": per_kernel,
                "other_stages_ms": {k: v for k, v in stage_ms.items() if k not in ("chi2", "bin_residual")}}

    # ---- end to end through the plugin API with HOST buffers (Slik.evaluate_batch -> csl_lnl_host)
    Xh = smp.stats["pos"].cpu().pin_memory()           # this rank's walkers, as a host array
    Oh = torch.empty(args.walkers, dtype=torch.float64).pin_memory()
    xh, oh = Xh.numpy(), Oh.numpy()
    for _ in range(3):
        prog.lnl_host(xh, oh)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        prog.lnl_host(xh, oh)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if size > 1:
        torch.distributed.all_reduce(t_e2e, op=torch.distributed.ReduceOp.MAX)
    e2e = {"value": k_total * args.steps / float(t_e2e.item()), "unit": UNIT,
           "h2d_bytes_per_step": int(xh.nbytes), "d2h_bytes_per_step": int(oh.nbytes),
           "api": "Slik.evaluate_batch(numpy [W, ndim]) -> numpy [W] (csl_lnl_host, pinned host buffers)"}

    if rank != 0:
        return
    cpu = None
    if size == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single(args.workload, args.cpu_seconds)
    # my kernels per stretch step in the native loop (csl_stretch_run): per half-update the fused
    # propose+prepare, one binning launch per spectrum class, chi2, the fused combine+accept (+ peer barrier
    # when sharded); + the row bookkeeping.  CSL_NO_FUSE=1 runs the staged kernels instead (propose, prepare,
    # binning, chi2, partial sums, combine, accept).
    nbin = int(prog.meta.get("nclass", 0)) + int(prog.meta.get("nsegclass", 0))
    if os.environ.get("CSL_NO_FUSE") is None:
        launches_per_step = 2 * (1 + nbin + 1 + 1 + (1 if size > 1 else 0)) + 1
    else:
        launches_per_step = 2 * (2 + nbin + (2 if prog.meta.get("chi2_mode") else 1) + 1 + 1 + (1 if size > 1 else 0)) + 1
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": size, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_of(args, prob, k_total), exchange=smp.stats.get("exchange")), "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk.summary(),
            "acceptance": acc_frac, "host_enqueue_ms_per_step": host_ms,
            "ms_per_step_instrumented_pass": ms_instrumented,
            "flops_per_eval": {"trsm_algorithmic": fl["trsm_alg"], "trsm_executed": fl["trsm"],
                               "binning_dense_windows": fl["bin_dense"], "binning_nonzero_windows": fl["bin_nnz"],
                               "binning_executed_dmma": fl["bin_exec"], "foreground_build": fl["build_ops"],
                               "segsum_fp64_instructions_x2": fl.get("seg_flops", 0)}}
    emit(line)


def run_gauss30(args):
    """BASELINE configs[1] (secondary line, not the headline): 30-parameter correlated Gaussian,
    adaptive MH, --walkers chains per GPU (default 4096 when left at 65536), one bench step = one block of
    100 MH steps = one launch of the fused kernel csl_mh_gauss_run (Philox streams on the device)."""
    import torch
    import cosmoslik_b200 as cb
    from cosmoslik_b200 import dist
    rank, size = dist.init_from_env()
    if size == 1:
        torch.cuda.set_device(0)
    nch = (4096 if args.walkers == 65536 else args.walkers) * size
    nd, block = 30, 100
    A = np.random.RandomState(0).standard_normal((nd, nd))
    C = A @ A.T / nd + np.eye(nd)

    class script(cb.SlikPlugin):
        def __init__(self, nsteps):
            super().__init__()
            for i in range(nd):
                self["p%02d" % i] = cb.param(start=0, scale=1)
            self.like = cb.likelihoods.gaussian(np.zeros(nd), C)
            self.sampler = cb.samplers.metropolis_hastings(self, num_samples=nsteps * nch, nchains=nch, seed=3,
                                                           mpi_comm_freq=block, proposal_update_start=200)

        def __call__(self):
            return self.like([self["p%02d" % i] for i in range(nd)])

    def run(nblocks):
        slik = cb.Slik(script(nblocks * block))
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = slik.sampler.run(slik.evaluate)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1), st, slik

    run(max(args.warmup, 3))
    clk = ClockSampler(torch.cuda.current_device())
    with clk:
        ms, st, slik = run(args.steps)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if size > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    if rank != 0:
        return
    nrows = int(st["nrows"].sum().item())
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    bytes_alg = nrows * (2 + nd) * 8 + 2 * (nch // size) * (nd + 2) * 8        # rows written + state in/out per launch set
    line = {"metric": METRIC.replace("bandpower lnL", "Gaussian lnL (secondary workload)"), "unit": UNIT,
            "value": nch * args.steps * block / (ms * 1e-3), "n_gpus": size, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] 30-parameter correlated Gaussian, adaptive MH",
                       "chains_total": nch, "mh_steps_per_bench_step": block, "streams": "Philox4x32-10 on device",
                       "proposal_update": "pooled, every 100 steps after 200", "adaptations": len(slik.sampler.adaptations)},
            "roofline": {"bound": "hbm", "kernel": "k_mh_gauss_run", "achieved": bytes_alg / (ms * 1e-3) / 1e9,
                         "peak": hbm, "unit": "GB/s", "frac": bytes_alg / (ms * 1e-3) / 1e9 / hbm, "traffic": None,
                         "note": "state lives in registers for the whole block; only emitted rows touch HBM, so this "
                                 "kernel is latency/issue bound, not HBM bound"},
            "gpu_launches": args.steps + 2 * len(slik.sampler.adaptations) * 2 + 4, "clocks": clk.summary(),
            "acceptance": nrows / float(nch * args.steps * block)}
    emit(line)


def run_unbinned(args):
    """BASELINE configs[4] (secondary line): unbinned TT/TE/EE, 3 x 2499 = 7497 data points, dense 7497^2
    covariance, --walkers MH chains per GPU (default 8192), pooled proposal adaptation + Gelman-Rubin.
    One bench step = one MH step of every chain (propose, priors, foregrounds, chi^2, accept)."""
    import torch
    import cosmoslik_b200 as cb
    from cosmoslik_b200 import dist, synthetic
    rank, size = dist.init_from_env()
    if size == 1:
        torch.cuda.set_device(0)
    per_gpu = 8192 if args.walkers == 65536 else args.walkers
    nch = per_gpu * size
    t0 = time.perf_counter()
    prob = synthetic.unbinned_problem()
    ns = types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl,
                               egfs=cb.models.egfs)

    def make(nsteps):
        return cb.Slik(synthetic.build_script(ns, prob, sampler=lambda s: cb.samplers.metropolis_hastings(
            s, num_samples=nsteps * nch, nchains=nch, seed=5, mpi_comm_freq=20, proposal_update_start=20))())

    slik = make(max(args.warmup, 3))
    prog = slik.program(chi2_mode=args.chi2)
    setup_s = time.perf_counter() - t0
    slik.sampler.run(prog)
    slik2 = make(args.steps)
    slik2._program, slik2._program_options = prog, slik._program_options
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(torch.cuda.current_device())
    with clk:
        e0.record()
        st = slik2.sampler.run(prog)
        rhat = slik2.sampler.gelman_rubin(burn_rows=0)
        e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if size > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    timers = {}
    for _ in range(3):
        prog.lnl(st["cur_x"], timers=timers)
    torch.cuda.synchronize()
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}
    if rank != 0:
        return
    peak, peak_src = fp64_peak()
    fl = prog.flops
    kern = "k_trigemm_chi2" if prog.meta.get("chi2_mode") else "k_trsm_chi2"
    ach = fl["trsm_alg"] * per_gpu / (stage_ms["chi2"] * 1e-3) / 1e12
    line = {"metric": METRIC.replace("bandpower lnL", "unbinned multi-spectrum lnL (secondary workload)"), "unit": UNIT,
            "value": nch * args.steps / (ms * 1e-3), "n_gpus": size, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4] unbinned TT/TE/EE ell<=2500, n=%d, dense Cholesky" % prob["n"],
                       "chains_total": nch, "sampler": "adaptive MH, exchange every 20 steps", "setup_seconds": setup_s,
                       "gelman_rubin_max": float(np.max(rhat)), "adaptations": len(slik2.sampler.adaptations)},
            "roofline": {"bound": "tensor", "kernel": kern, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                         "frac": ach / peak, "traffic": None, "peak_source": peak_src, "stages_ms": stage_ms,
                         "executed_tflops": fl["trsm"] * per_gpu / (stage_ms["chi2"] * 1e-3) / 1e12},
            "gpu_launches": 8 * args.steps, "clocks": clk.summary()}
    emit(line)


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print banners there (NCCL's version line under
    # NCCL_DEBUG=VERSION, for one) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        return _main(args)
    finally:
        if args.impl == "b200":
            try:
                from cosmoslik_b200 import dist
                dist.shutdown()
            except Exception:
                pass
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
        if _LINES:
            sys.stdout.write("\n".join(_LINES) + "\n")
            sys.stdout.flush()


_LINES = []


def emit(line):
    _LINES.append(json.dumps(line))


def _main(args):
    if args.workload == "gauss30_mh" and args.impl == "b200":
        return run_gauss30(args)
    if args.workload == "unbinned_mh" and args.impl == "b200":
        return run_unbinned(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
