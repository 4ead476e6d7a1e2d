#!/usr/bin/env python
"""Headline benchmark: posterior evaluations / second for the batched bandpower lnL.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]

Headline workload (BASELINE.json configs[3], the one the metric is quoted on): Planck-lite-sized TT/TE/EE
bandpower likelihood, 613 bins, dense FP64 covariance, 9 sampled parameters, stretch-move ensemble of
65 536 walkers IN TOTAL, sharded over the N GPUs ("strong" scaling: the literal config).  One "step" = one full
stretch-move step (both half-ensemble updates, each = proposal + priors + foreground build + binning + FP64
DMMA quadratic form + accept) = one posterior evaluation per walker.  The same line carries, under "weak", the
record with 65 536 walkers PER GPU, and under "secondary" short runs of the other BASELINE configs (quickstart
single chain, 30-parameter Gaussian adaptive MH, SPT-like multi-frequency stretch move, unbinned multi-spectrum
adaptive MH with Gelman-Rubin, real SPT 's12' bandpowers with MH), each with its own roofline / clocks / e2e /
cpu_baseline.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import hashlib
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
LITERAL_WALKERS = 65536
PREROLL = int(os.environ.get("CSL_BENCH_PREROLL", "40"))   # untimed steps run right before every timed region (>= --warmup)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="planck_lite",
                    choices=["planck_lite", "spt_multifreq", "gauss30_mh", "unbinned_mh", "quickstart", "spt_lowl_mh"])
    ap.add_argument("--walkers", type=int, default=LITERAL_WALKERS,
                    help="walkers IN TOTAL of the headline record (and per GPU of the weak record)")
    ap.add_argument("--secondary", default="all", help="'all', 'none' or a comma-separated list of secondary workloads")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling record at N > 1")
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="budget of the headline cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--binning", default="auto", choices=["auto", "dmma", "segsum"],
                    help="window-binning kernel: DMMA tile product, per-bin segmented sum, or by window structure")
    ap.add_argument("--chi2", default="auto", choices=["auto", "trsm", "inverse"])
    return ap.parse_args()


def make_problem(name):
    from cosmoslik_b200 import synthetic
    return synthetic.planck_lite_problem(0) if name == "planck_lite" else synthetic.spt_multifreq_problem(0)


def config_of(args, prob, total=None):
    """the workload's description: a function of the command line only, so both arms print the same dict"""
    total = int(args.walkers if total is None else total)
    win = ("top-hat plik-lite bins (block-sparse window matrix; binning kernel: %s)" % args.binning
           if args.workload == "planck_lite" else "Gaussian bumps + 1e-10 tails (numerically dense windows)")
    per = total // max(args.gpus, 1)
    return {"workload": "BASELINE configs[3] planck_lite TT/TE/EE 613 bins" if args.workload == "planck_lite"
            else "BASELINE configs[2] spt_multifreq 6 spectra x 50 bins",
            "n_data": int(prob["n"]), "ndim": len(prob["truth"]), "sampler": "stretch move a=2",
            "walkers_total": total, "walkers_per_gpu": per,
            "windows": win, "l2": "R matrix of a half-step is %d MB per GPU; no explicit flush (the working set of a "
            "step -- R, partial sums, coefficients, rows -- is rewritten every half-step)"
            % (prob["n"] * (per // 2) * 8 // 2 ** 20), "parallelism": "walker-sharded x%d" % args.gpus}


# ------------------------------------------------------------------------------------- oracle arm
def _oracle_ns():
    from oracle import likes as olikes, runtime as ort
    return types.SimpleNamespace(SlikPlugin=ort.Plugin, param=ort.Param, mspec_lnl=olikes.MspecLike, egfs=olikes.Egfs)


def _oracle_slik(workload):
    from cosmoslik_b200 import synthetic
    from oracle import runtime as ort
    prob = make_problem(workload)
    return prob, ort.Slik(synthetic.build_script(_oracle_ns(), prob)())


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


def time_oracle_evals(slik, X, budget_s, what):
    """bounded sample of scalar oracle evaluations on ONE host core"""
    _limit_blas_threads()
    slik.evaluate(*X[0])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and n < len(X):
        slik.evaluate(*X[n]); n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d oracle Slik.evaluate calls of %s in %.1f s (1 core, one BLAS thread)" % (n, what, dt)}


def cpu_baseline_single(workload, budget_s):
    """the oracle port (numpy/scipy, scalar per-walker calls like the reference) on ONE host core"""
    prob, slik = _oracle_slik(workload)
    return time_oracle_evals(slik, _sample_points(prob, 4096), budget_s, "the same script")


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
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_of(args, prob),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region (NVML, 8 ms period (CSL_CLOCK_PERIOD_MS); nvidia-smi fallback).
    Construct it BEFORE the barrier that opens the timed region: nvmlInit is slow and contended when several
    ranks start at once."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("hw_power_brake", 0x80), ("sw_power_cap", 0x4))
    _nvml = None

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.period = float(os.environ.get("CSL_CLOCK_PERIOD_MS", "8")) * 1e-3
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.max_mhz, self.nv, self.h = None, None, None
        try:
            if ClockSampler._nvml is None:
                import pynvml as nv
                nv.nvmlInit()
                ClockSampler._nvml = nv
            nv = ClockSampler._nvml
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
                time.sleep(self.period)
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


# ------------------------------------------------------------------------------------- helpers of the B200 arm
def fp64_peak():
    """FP64 tensor roofline denominator: measured cuBLAS DGEMM on this pool's B200 (tools/measure_peaks.py,
    committed under profiles/); MEASURED_PEAKS.json has no FP64 figure."""
    path = os.path.join(ROOT, "profiles", "measured_fp64_peak.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        src = "measured cuBLAS DGEMM sustained (profiles/measured_fp64_peak.json), the FP64 analogue of MEASURED_PEAKS.json's cuBLAS bf16 figure"
        if d.get("dmma_pipe_tflops"):
            src += "; the bare DMMA issue rate measures %.2f TFLOP/s (tools/micro/pipes.cu), so a kernel can exceed 1.0 of the library peak by up to %.3f" % (
                d["dmma_pipe_tflops"], d["dmma_pipe_tflops"] / d["fp64_tflops_sustained"])
        return d["fp64_tflops_sustained"], src
    return 37.0, "fallback: B200 datasheet FP64 37 TFLOP/s (no measurement yet)"


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f).get("hbm_gbs", 6650.0)), "of measured (MEASURED_PEAKS.json)"
    return 6650.0, "of fallback (B200_PROFILING.md)"


def kernel_source_digest():
    """digest of the kernel sources: an ncu capture recorded under profiles/ is quoted only while the
    sources it was taken from are unchanged"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "cosmoslik_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh")):
            with open(os.path.join(d, fn), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic(kernel, walkers_per_launch):
    """DRAM bytes per launch of `kernel` from this round's `ncu --set full` capture (profiles/r2_traffic.json),
    or None when there is no capture of exactly these sources and this launch size."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        tj = json.load(f)
    if tj.get("source_digest") != kernel_source_digest():
        return None
    rec = tj.get(kernel)
    if rec and rec.get("walkers_per_launch") == walkers_per_launch:
        return rec.get("dram_bytes_per_launch")
    return None


def product_ns():
    import cosmoslik_b200 as cb
    return types.SimpleNamespace(SlikPlugin=cb.SlikPlugin, param=cb.param, mspec_lnl=cb.likelihoods.mspec_lnl,
                                 egfs=cb.models.egfs)


def all_max(torch, size, v):
    t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
    if size > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def all_list(torch, size, v):
    t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
    if size == 1:
        return [float(v)]
    out = torch.empty(size, dtype=torch.float64, device="cuda")
    torch.distributed.all_gather_into_tensor(out, t)
    return [float(x) for x in out.cpu()]


def time_e2e(torch, dist, size, prog, X_dev, steps):
    """the metric end to end through the plugin API with HOST buffers: Slik.evaluate_batch(numpy) -> csl_lnl_host
    (pinned host arrays; H2D of the positions, the kernels and D2H of lnl all inside the timed region)"""
    W = int(X_dev.shape[0])
    Xh = X_dev.cpu().pin_memory()
    Oh = torch.empty(W, dtype=torch.float64).pin_memory()
    xh, oh = Xh.numpy(), Oh.numpy()
    for _ in range(3):
        prog.lnl_host(xh, oh)
    # three repetitions of `steps` calls, the MEDIAN reported (every call synchronises, so a repetition is exposed to
    # whatever else the host is doing: 43 .. 50 M/s between otherwise identical runs of one repetition)
    reps = []
    for _ in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            prog.lnl_host(xh, oh)
        torch.cuda.synchronize()
        reps.append(all_max(torch, size, time.perf_counter() - t0))
    dt = sorted(reps)[1]
    return {"value": W * size * steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(xh.nbytes),
            "d2h_bytes_per_step": int(oh.nbytes), "repetitions": [W * size * steps / r for r in reps],
            "api": "Slik.evaluate_batch(numpy [W, ndim]) -> numpy [W] (csl_lnl_host: two half-batches on two "
                   "streams, pinned host buffers)"}


# ------------------------------------------------------------------------------------- stretch-move records
def measure_stretch(args, prob, k_total, instrument=True, e2e=True):
    """One record: K stretch-move steps of `k_total` walkers sharded over the ranks, device resident."""
    import torch
    import cosmoslik_b200 as cb
    from cosmoslik_b200 import dist, synthetic
    rank, size = dist.get_rank(), dist.get_size()
    dev = torch.cuda.current_device()
    sampler = lambda s: cb.samplers.emcee(s, num_samples=k_total, nwalkers=k_total, seed=1234)
    slik = cb.Slik(synthetic.build_script(product_ns(), prob, sampler=sampler)())
    prog = slik.program(chi2_mode=args.chi2, binning=args.binning)
    smp = slik.sampler
    names = list(slik.get_sampled())
    # tight initial ball around the truth (a burnt-in ensemble), identical on every rank
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    smp.init_pos = x0 + sc * np.random.RandomState(99).standard_normal((k_total, len(names)))

    clk = ClockSampler(dev)                         # NVML initialised BEFORE the warm-up: the GPU does not idle (and
                                                    # no rank is late) between the warm-up and the timed region
    smp.row_capacity = max(args.steps, PREROLL)     # same row buffer for the warm-up and the timed run
    if instrument:
        smp.python_loop = True                      # the instrumented Python loop's code paths ...
        smp.run(prog, nsteps=3)
        del smp["python_loop"]
    # untimed pre-roll of the native loop right before the timed run: the W warm-up steps asked for, topped up to
    # PREROLL steps (a 20-step timed region is 27 ms; with only 5 steps = 7 ms of work before it the first timed
    # steps still ran on a GPU that was leaving idle: 1.33 .. 1.81 ms per step between otherwise identical runs)
    smp.run(prog, nsteps=max(args.warmup, 3, PREROLL), resume=instrument)
    replica = smp["_replica"][1] if "_replica" in smp and smp["_replica"] is not None else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    # Short regions (K steps x 10 launches that fit the launch queue) are PRE-QUEUED: a host-released gate kernel
    # holds the stream while the host enqueues the barrier, the first event, all K steps and the last event, and
    # is opened by the host afterwards.  A sampler run of thousands of steps looks like that to the device anyway
    # (the host enqueues a step ~8x faster than the device runs it).  Insurance against a host stall in the first
    # steps of a 5-26 ms region (8 processes + their NVML samplers on 16 cores), not a speed-up: gated and ungated
    # 20-step regions agree within 0.6 % at 1, 2 and 8 GPUs (profiles/r2_gate_ab.json).
    from cosmoslik_b200 import _lib as L
    nlaunch = args.steps * (2 * (2 + int(L.lib().csl_bin_residual_launches(prog.struct)) + 1) + 2)
    use_gate = (not os.environ.get("CSL_BENCH_NO_GATE")) and nlaunch <= 600
    gate = torch.zeros(1, dtype=torch.int32).pin_memory()
    def open_region():
        # every rank's timed region opens at the same point of the device timeline: a device-side barrier on the
        # stream right before the first event (a host barrier alone leaves the ranks' launch skew in the step time)
        if use_gate:
            L.check(L.lib().csl_host_gate(gate.data_ptr(), 2.0, L.stream_ptr()), "csl_host_gate")   # gives up after 2 s (a profiler that serialises launches never lets the host reach the release)
        if replica is not None:
            replica.barrier()
        e0.record()
    def close_region():
        e1.record()
        if use_gate:
            gate[0] = 1          # everything is queued: let the device go

    # the events sit on the stream immediately before the first and after the last launch of the K steps: the
    # sampler's host-side set-up before its first launch and its closing device->host read of the acceptance counter
    # are host work around the steps, not part of them (at 20 steps of 0.28 ms they were a sixth of the region)
    smp.hooks = dict(start=open_region, end=close_region)
    with clk:
        smp.run(prog, nsteps=args.steps, resume=True)     # exactly K steps, continuing the warmed-up ensemble
        torch.cuda.synchronize()
    del smp["hooks"]
    dist.barrier()
    my_ms = e0.elapsed_time(e1)
    per_rank = all_list(torch, size, my_ms)
    ms = max(per_rank)
    rec = {"walkers_total": k_total, "walkers_per_gpu": k_total // size, "value": k_total * args.steps / (ms * 1e-3),
           "ms_per_step": ms / args.steps, "ms_per_rank": [m / args.steps for m in per_rank],
           "untimed_preroll_steps": max(args.warmup, 3, PREROLL),
           "launches": "pre-queued behind a host-released gate (csl_host_gate)" if use_gate else "enqueued while the region runs",
           "acceptance": smp.stats["accepted"] / float(smp.stats["evals"]),
           "host_enqueue_ms_per_step": smp.stats.get("host_ms_per_step"), "exchange": smp.stats.get("exchange"),
           "clocks": clk.summary()}
    nbin = int(L.lib().csl_bin_residual_launches(prog.struct))       # launches of the binning stage
    fused = os.environ.get("CSL_NO_FUSE") is None and prog.meta["resid_mode"] == 1
    # my kernels per stretch step in the native loop: per half-update the fused propose+prepare(+wait), one
    # binning launch per spectrum class, chi2, the fused combine+accept(+rows+publish+signal)
    if fused:           # the row bookkeeping rides in the accept kernels (CSL_NO_FOLD_ROWS=1: its own launch per step)
        per_step = 2 * (1 + nbin + 1 + 1) + (1 if os.environ.get("CSL_NO_FOLD_ROWS") else 0)
    else:
        per_step = 2 * (2 + nbin + 1 + 1 + 1 + (1 if size > 1 else 0)) + 1
    rec["gpu_launches"] = per_step * args.steps + (1 if (fused and size > 1) else 0)
    half = (k_total // size) // 2
    fl = prog.flops
    if instrument:
        # second pass of K more steps with a CUDA-event pair around every stage (this disables the C-side step
        # loop: each stage is launched from Python between events), for the per-kernel durations
        timers = {}
        smp.timers = timers
        dist.barrier(); torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        smp.run(prog, nsteps=args.steps, resume=True)
        f1.record(); torch.cuda.synchronize()
        rec["ms_per_step_instrumented_pass"] = f0.elapsed_time(f1) / args.steps
        del smp["timers"]
        stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}
        stage_n = {k: len(v) for k, v in timers.items()}
        peak, peak_src = fp64_peak()
        per_kernel = {}
        if "chi2" in stage_ms:
            t = stage_ms["chi2"] * 1e-3
            per_kernel["k_trigemm_chi2" if prog.meta.get("chi2_mode") else "k_trsm_chi2"] = {
                "ms_per_launch": stage_ms["chi2"], "launches": stage_n["chi2"],
                "algorithmic_tflops": fl["trsm_alg"] * half / t / 1e12, "executed_tflops": fl["trsm"] * half / t / 1e12,
                "frac_of_fp64_peak": fl["trsm_alg"] * half / t / 1e12 / peak}
        if "bin_residual" in stage_ms:
            t = stage_ms["bin_residual"] * 1e-3
            alg = fl["bin_nnz"] + fl["build_ops"]            # 2 nnz(W) + the foreground build: the work the maths needs
            if prog.meta.get("nsegclass") and not prog.meta.get("nclass"):
                per_kernel["k_bin_segsum"] = {
                    "ms_per_launch": stage_ms["bin_residual"], "launches": stage_n["bin_residual"],
                    "algorithmic_tflops": alg * half / t / 1e12, "frac_of_fp64_peak": alg * half / t / 1e12 / peak,
                    "fp64_pipe_utilisation": fl["seg_flops"] * half / t / 1e12 / peak,
                    "note": "fp64_pipe_utilisation = 2 x the FP64 instructions the kernel issues (Taylor recurrence "
                            "of the power laws included) / peak: a pipe figure, not an algorithmic one.  The algorithmic "
                            "figure counts the reference formulation (2 nnz(W) + 2 NT + 40 NP flops per ell, an exp = 40): "
                            "it exceeds 1.0 because the kernel advances the power laws by a 7-instruction recurrence"}
            else:
                per_kernel["k_bin_residual"] = {
                    "ms_per_launch": stage_ms["bin_residual"], "launches": stage_n["bin_residual"],
                    "algorithmic_tflops": alg * half / t / 1e12,
                    "executed_tflops": (fl["bin_exec"] + fl["build_ops"]) * half / t / 1e12,
                    "frac_of_fp64_peak": alg * half / t / 1e12 / peak}
        if per_kernel:
            dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"])
            rec["roofline"] = {"bound": "tensor", "kernel": dom, "achieved": per_kernel[dom]["algorithmic_tflops"],
                               "peak": peak, "unit": "TFLOP/s", "frac": per_kernel[dom]["frac_of_fp64_peak"],
                               "traffic": measured_traffic(dom, half), "peak_source": peak_src, "kernels": per_kernel,
                               "walkers_per_launch": half,
                               "other_stages_ms": {k: v for k, v in stage_ms.items() if k not in ("chi2", "bin_residual")}}
    if e2e:
        rec["e2e"] = time_e2e(torch, dist, size, prog, smp.stats["pos"], args.steps)
    rec["flops_per_eval"] = {"trsm_algorithmic": fl["trsm_alg"], "trsm_executed": fl["trsm"],
                             "binning_dense_windows": fl["bin_dense"], "binning_nonzero_windows": fl["bin_nnz"],
                             "binning_executed_dmma": fl["bin_exec"], "foreground_build": fl["build_ops"],
                             "segsum_fp64_instructions_x2": fl.get("seg_flops", 0)}
    return rec


def exchange_selfcheck(k=None, nsteps=6):
    """N > 1: a small sharded ensemble -- through the fused NVLink exchange (symmetric-memory replica updated by
    the accept kernel, signal / wait folded into the kernels) and through NCCL all_gather -- against a
    single-rank replay of the WHOLE ensemble on every rank.  Returns 'bitwise' or a description of the mismatch."""
    import torch
    import cosmoslik_b200 as cb
    from cosmoslik_b200 import dist, synthetic
    size = dist.get_size()
    if size == 1:
        return None
    k = 1024 + 2 * size if k is None else k         # ragged shares when size does not divide k/2
    prob = synthetic.spt_multifreq_problem(1, nbins=12, lmax=801)

    def run():
        slik = cb.Slik(synthetic.build_script(product_ns(), prob, sampler=lambda s: cb.samplers.emcee(
            s, num_samples=k * nsteps, nwalkers=k, seed=11))())
        st = slik.sampler.run(slik.evaluate)
        return st["pos"].cpu().numpy(), st["lnl"].cpu().numpy(), st["walker_ids"], st["accepted"], st["exchange"]

    results = {}
    for mode in ("symm", "nccl"):
        os.environ["COSMOSLIK_B200_NO_SYMM"] = "0" if mode == "symm" else "1"
        results[mode] = run()
    os.environ["COSMOSLIK_B200_NO_SYMM"] = "0"
    saved = (dist.get_rank, dist.get_size, dist._dist)
    dist.get_rank, dist.get_size, dist._dist = (lambda: 0), (lambda: 1), (lambda: None)
    try:
        ref_pos, ref_lnl, _, ref_acc, _ = run()
    finally:
        dist.get_rank, dist.get_size, dist._dist = saved
    bad = []
    for mode, (pos, lnl, ids, acc, exch) in results.items():
        same = np.array_equal(pos, ref_pos[ids]) and np.array_equal(lnl, ref_lnl[ids])
        tot = torch.tensor([acc], device="cuda"); torch.distributed.all_reduce(tot)
        same = same and int(tot.item()) == ref_acc
        flag = torch.tensor([1 if same else 0], device="cuda")
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if not int(flag.item()):
            bad.append("%s (%s)" % (mode, exch))
    return "bitwise" if not bad else "MISMATCH in " + ", ".join(bad)


def run_b200(args):
    import torch
    from cosmoslik_b200 import dist
    rank, size = dist.init_from_env()
    if size == 1:
        torch.cuda.set_device(0)
    prob = make_problem(args.workload)
    literal = measure_stretch(args, prob, args.walkers)                 # the literal config: walkers IN TOTAL
    weak = None
    if size > 1 and not args.no_weak:
        weak = measure_stretch(args, prob, args.walkers * size)         # the same walkers PER GPU
    parity = exchange_selfcheck() if size > 1 else None
    secondary = run_secondary(args, torch, dist) if args.workload == "planck_lite" else []
    if rank != 0:
        return
    cpu = None
    if size == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single(args.workload, args.cpu_seconds)
    line = {"metric": METRIC, "value": literal["value"], "unit": UNIT, "n_gpus": size, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": literal["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args, prob), "e2e": literal.get("e2e"), "gpu_launches": literal["gpu_launches"],
            "roofline": literal.get("roofline"), "cpu_baseline": cpu, "clocks": literal["clocks"],
            "ms_per_rank": literal["ms_per_rank"], "exchange": literal["exchange"], "exchange_parity": parity,
            "untimed_preroll_steps": literal["untimed_preroll_steps"], "launches": literal["launches"],
            "acceptance": literal["acceptance"], "host_enqueue_ms_per_step": literal["host_enqueue_ms_per_step"],
            "ms_per_step_instrumented_pass": literal.get("ms_per_step_instrumented_pass"),
            "flops_per_eval": literal["flops_per_eval"],
            "weak": None if weak is None else dict(weak, scaling="weak", config=config_of(args, prob, args.walkers * size)),
            "secondary": secondary}
    emit(line)


# ------------------------------------------------------------------------------------- secondary workloads
def _secondary_line(name, metric, value, ms_per_step, steps, config, roofline, clocks, e2e, cpu, extra=None):
    d = {"name": name, "metric": metric, "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps,
         "higher_is_better": True, "dtype": "f64", "data": "synthetic", "config": config, "roofline": roofline,
         "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu}
    d.update(extra or {})
    return d


def run_secondary(args, torch, dist):
    want = args.secondary
    if want == "none":
        return []
    names = ["quickstart", "gauss30_mh", "spt_multifreq", "unbinned_mh", "spt_lowl_mh"]
    if want != "all":
        names = [n for n in names if n in want.split(",")]
    out = []
    for n in names:
        t0 = time.perf_counter()
        try:
            rec = {"quickstart": sec_quickstart, "gauss30_mh": sec_gauss30, "spt_multifreq": sec_spt_multifreq,
                   "unbinned_mh": sec_unbinned, "spt_lowl_mh": sec_spt_lowl}[n](args, torch, dist)
        except Exception as e:                    # a secondary workload never takes the headline down with it
            rec = {"name": n, "error": "%s: %s" % (type(e).__name__, e)}
        if rec is not None:
            rec["wall_s"] = time.perf_counter() - t0
            out.append(rec)
        dist.barrier()
    return out


def _mh_timed(torch, dist, make_slik, nsteps_warm, nsteps, prog=None):
    """warm-up run, then a timed run of the MH sampler; returns (ms max over ranks, stats, slik, clocks)"""
    size = dist.get_size()
    sl = make_slik(nsteps_warm)
    if prog is not None:
        sl._program, sl._program_options = prog
    sl.sampler.run(sl.evaluate)
    sl2 = make_slik(nsteps)
    sl2._program, sl2._program_options = sl._program, sl._program_options
    clk = ClockSampler(torch.cuda.current_device())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    # the events sit right before the first and after the last block of steps: the sampler's set-up (a row buffer of
    # gigabytes to allocate, the initial evaluation, the host factor of the initial covariance) is host work before
    # the steps, and on one box it put 120 ms into a 30 ms region (gauss30_mh: 7.7 ms per block against 1.6)
    sl2.sampler.hooks = dict(start=e0.record, end=e1.record)
    with clk:
        st = sl2.sampler.run(sl2.evaluate)
        torch.cuda.synchronize()
    del sl2.sampler["hooks"]
    ms = all_max(torch, size, e0.elapsed_time(e1))
    return ms, st, sl2, clk.summary()


def sec_quickstart(args, torch, dist):
    """BASELINE configs[0]: the quickstart script (doc/quickstart.md:26-43), ONE chain, 1e5 MH steps.  The
    reference runs it on one CPU core; here the single chain runs through the native step loop (a CUDA graph
    per 100 steps) -- one warp of one SM: a latency measurement, not a throughput one."""
    import cosmoslik_b200 as cb
    if dist.get_size() > 1 and dist.get_rank() != 0:
        return None
    nsteps = 100000 if args.steps >= 20 else 5000 * max(args.steps, 1)

    class quick(cb.SlikPlugin):
        def __init__(self, n):
            super().__init__()
            self.a = cb.param(start=0, scale=1)
            self.b = cb.param(start=0, scale=1)
            self.sampler = cb.samplers.metropolis_hastings(self, num_samples=n, seed=2)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    saved = (dist.get_rank, dist.get_size, dist._dist)         # a single chain lives on rank 0 alone
    dist.get_rank, dist.get_size, dist._dist = (lambda: 0), (lambda: 1), (lambda: None)
    try:
        sl = cb.Slik(quick(2000)); sl.sampler.run(sl.evaluate)
        sl = cb.Slik(quick(nsteps))
        clk = ClockSampler(torch.cuda.current_device())
        torch.cuda.synchronize()
        with clk:
            t0 = time.perf_counter()
            st = sl.sampler.run(sl.evaluate)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        X = st["cur_x"].repeat(4096, 1).contiguous()
        e2e = time_e2e(torch, types.SimpleNamespace(barrier=lambda: None), 1, sl.program(), X, 5)
    finally:
        dist.get_rank, dist.get_size, dist._dist = saved
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import runtime as ort, samplers as osamp
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_scripts as O
        oslik = ort.Slik(O.quick_script()())
        n = 20000
        rs = np.random.RandomState(0)
        cov = np.eye(2) / 2 * 2.4 ** 2                      # cov_est / ndim * proposal_scale^2 (:141-142)
        t0 = time.perf_counter()
        nrow = sum(1 for _ in osamp.mh_mcmc(lambda *x: oslik.evaluate(*x), [0.0, 0.0], n,
                                            lambda cur: rs.multivariate_normal(cur, cov), rs.uniform))
        dtc = time.perf_counter() - t0
        cpu = {"value": n / dtc, "unit": "MH steps/s", "cores": 1, "kind": "port",
               "sample": "%d steps (%d rows) of the oracle's _mcmc restatement, numpy multivariate_normal (SVD) per "
                         "draw as the reference, in %.1f s" % (n, nrow, dtc)}
    hbm, src = hbm_peak()
    return _secondary_line(
        "quickstart", "MH steps/sec, single chain (BASELINE configs[0])", nsteps / dt, 1e3 * dt / nsteps, nsteps,
        {"workload": "BASELINE configs[0] quickstart 2-parameter Gaussian, samplers.metropolis_hastings, 1 chain",
         "loop": st.get("loop")},
        {"bound": "hbm", "kernel": "k_mh_propose_prepare + k_mh_combine_accept", "achieved": 1e-9 * (24 * 2 + 32) * nsteps / dt,
         "peak": hbm, "unit": "GB/s", "frac": 1e-9 * (24 * 2 + 32) * nsteps / dt / hbm, "traffic": None,
         "note": "one chain = one thread: launch-latency bound by construction (%s)" % src},
        clk.summary(), e2e, cpu, {"unit": "MH steps/s", "accepted": st.get("accepted")})


def sec_gauss30(args, torch, dist):
    """BASELINE configs[1]: 30-parameter correlated Gaussian, adaptive MH, 4096 chains per GPU; one bench step =
    one block of 100 MH steps = one launch of csl_mh_gauss_run; pooled proposal update on the device."""
    import cosmoslik_b200 as cb
    size = dist.get_size()
    per_gpu, nd, block = 4096, 30, 100
    nch = per_gpu * size
    A = np.random.RandomState(0).standard_normal((nd, nd))
    Cm = A @ A.T / nd + np.eye(nd)

    def make(nblocks):
        class script(cb.SlikPlugin):
            def __init__(self):
                super().__init__()
                for i in range(nd):
                    self["p%02d" % i] = cb.param(start=0, scale=1)
                self.like = cb.likelihoods.gaussian(np.zeros(nd), Cm)
                self.sampler = cb.samplers.metropolis_hastings(self, num_samples=nblocks * block * nch, nchains=nch, seed=3,
                                                               mpi_comm_freq=block, proposal_update_start=200)

            def __call__(self):
                return self.like([self["p%02d" % i] for i in range(nd)])
        return cb.Slik(script())

    nblk = max(args.steps, 4)
    ms, st, sl, clocks = _mh_timed(torch, dist, make, 4, nblk)
    value = nch * nblk * block / (ms * 1e-3)
    nrows = int(st["nrows"].sum().item())
    e2e = time_e2e(torch, dist, size, sl.program(), st["cur_x"], 10)
    hbm, src = hbm_peak()
    bytes_alg = nrows * (2 + nd) * 8 + 2 * per_gpu * (nd + 2) * 8 * nblk
    cpu = None
    if dist.get_rank() == 0 and size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_scripts as O
        from oracle import runtime as ort
        oslik = ort.Slik(O.gauss_script(Cm)())
        cpu = time_oracle_evals(oslik, np.random.RandomState(1).standard_normal((20000, nd)), 2.0, "the 30-parameter Gaussian script")
    return _secondary_line(
        "gauss30_mh", METRIC.replace("bandpower lnL", "Gaussian lnL (BASELINE configs[1])"), value, ms / nblk, nblk,
        {"workload": "BASELINE configs[1] 30-parameter correlated Gaussian, adaptive MH", "chains_total": nch,
         "chains_per_gpu": per_gpu, "mh_steps_per_bench_step": block, "streams": "Philox4x32-10 on device",
         "proposal_update": "pooled over all GPUs, every 100 steps after 200: one fused moment pass + ONE all-reduce "
                            "+ device Cholesky factor (no host round trip)",
         "adaptations": len(sl.sampler._adapt), "loop": st.get("loop")},
        {"bound": "hbm", "kernel": "k_mh_gauss_run", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
         "frac": bytes_alg / (ms * 1e-3) / 1e9 / hbm, "traffic": None,
         "note": "state lives in registers for the whole block; only emitted rows touch HBM, so this kernel is "
                 "latency / issue bound, not HBM bound (%s)" % src},
        clocks, e2e, cpu, {"acceptance": nrows / float(nch * nblk * block), "scaling": "weak"})


def sec_spt_multifreq(args, torch, dist):
    """BASELINE configs[2]: 6 cross-spectra x 50 bins, numerically dense windows (DMMA binning), 16 384
    stretch-move walkers in total."""
    from cosmoslik_b200 import synthetic
    a2 = argparse.Namespace(**vars(args))
    a2.workload, a2.binning, a2.chi2 = "spt_multifreq", "auto", "auto"
    prob = synthetic.spt_multifreq_problem(0)
    rec = measure_stretch(a2, prob, 16384)
    cpu = None
    if dist.get_rank() == 0 and dist.get_size() == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single("spt_multifreq", 3.0)
    return _secondary_line(
        "spt_multifreq", METRIC.replace("bandpower lnL", "bandpower lnL (BASELINE configs[2])"), rec["value"],
        rec["ms_per_step"], args.steps, config_of(a2, prob, 16384), rec.get("roofline"), rec["clocks"], rec.get("e2e"), cpu,
        {"gpu_launches": rec["gpu_launches"], "acceptance": rec["acceptance"], "scaling": "strong",
         "ms_per_rank": rec["ms_per_rank"], "flops_per_eval": rec["flops_per_eval"]})


def sec_unbinned(args, torch, dist):
    """BASELINE configs[4]: unbinned TT/TE/EE, 3 x 2499 = 7497 data points, dense 7497^2 covariance, 8192 MH
    chains per GPU, pooled proposal adaptation + Gelman-Rubin.  One bench step = one MH step of every chain."""
    import cosmoslik_b200 as cb
    from cosmoslik_b200 import synthetic
    size = dist.get_size()
    per_gpu = 8192
    nch = per_gpu * size
    t0 = time.perf_counter()
    prob = synthetic.unbinned_problem()

    def make(nsteps):
        return cb.Slik(synthetic.build_script(product_ns(), prob, sampler=lambda s: cb.samplers.metropolis_hastings(
            s, num_samples=nsteps * nch, nchains=nch, seed=5, mpi_comm_freq=5, proposal_update_start=5))())

    first = make(3)
    prog = first.program(chi2_mode=args.chi2)
    setup_s = time.perf_counter() - t0
    nsteps = max(args.steps, 16)
    ms, st, sl, clocks = _mh_timed(torch, dist, make, 6, nsteps, prog=(prog, first._program_options))
    rhat = sl.sampler.gelman_rubin(burn_rows=0)
    timers = {}
    for _ in range(3):
        prog.lnl(st["cur_x"], timers=timers)
    torch.cuda.synchronize()
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}
    peak, peak_src = fp64_peak()
    fl = prog.flops
    kern = "k_trigemm_chi2" if prog.meta.get("chi2_mode") else "k_trsm_chi2"
    ach = fl["trsm_alg"] * per_gpu / (stage_ms["chi2"] * 1e-3) / 1e12
    e2e = time_e2e(torch, dist, size, prog, st["cur_x"], 3)
    cpu = None
    if dist.get_rank() == 0 and size == 1 and not args.no_cpu_baseline:
        from oracle import runtime as ort
        oslik = ort.Slik(synthetic.build_script(_oracle_ns(), prob)())
        cpu = time_oracle_evals(oslik, _sample_points(prob, 512), 3.0, "the unbinned script")
    return _secondary_line(
        "unbinned_mh", METRIC.replace("bandpower lnL", "unbinned multi-spectrum lnL (BASELINE configs[4])"),
        nch * nsteps / (ms * 1e-3), ms / nsteps, nsteps,
        {"workload": "BASELINE configs[4] unbinned TT/TE/EE ell<=2500, n=%d, dense Cholesky" % prob["n"],
         "chains_total": nch, "chains_per_gpu": per_gpu, "sampler": "adaptive MH, pooled update every 5 steps after 5",
         "setup_seconds": setup_s, "gelman_rubin_max": float(np.max(rhat)), "adaptations": len(sl.sampler._adapt),
         "loop": st.get("loop")},
        {"bound": "tensor", "kernel": kern, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
         "traffic": None, "peak_source": peak_src, "stages_ms": stage_ms,
         "executed_tflops": fl["trsm"] * per_gpu / (stage_ms["chi2"] * 1e-3) / 1e12},
        clocks, e2e, cpu, {"scaling": "weak", "host_enqueue_ms_per_step": st.get("host_ms_per_step")})


def sec_spt_lowl(args, torch, dist):
    """The reference's own likelihood on its own data: spt_lowl('s12') (47 bandpowers, 47 x 3251 dense windows, the
    bundled SPT covariance; arrays from tests/golden/spt_lowl_data.npz) with a fixed synthetic TT spectrum,
    adaptive MH, 4096 chains per GPU through the native step loop (a CUDA graph per 100-step block)."""
    import cosmoslik_b200 as cb
    size = dist.get_size()
    per_gpu, block = 4096, 100
    nch = per_gpu * size
    data = np.load(os.path.join(ROOT, "tests", "golden", "spt_lowl_data.npz"))
    ell = np.arange(6000)
    tt = 5000.0 * np.exp(-(ell / 1200.0) ** 1.2) + 50.0

    def make(nsteps):
        class script(cb.SlikPlugin):
            def __init__(self):
                super().__init__()
                self.spt = cb.likelihoods.spt_lowl("s12", data=data)
                self.cmb = {"TT": tt}
                self.sampler = cb.samplers.metropolis_hastings(self, num_samples=nsteps * nch, nchains=nch, seed=8,
                                                               mpi_comm_freq=block, proposal_update_start=block)

            def __call__(self):
                return self.spt(self.cmb)
        return cb.Slik(script())

    nsteps = block * max(2, min(args.steps // 5, 4))
    ms, st, sl, clocks = _mh_timed(torch, dist, make, block, nsteps)
    prog = sl.program()
    timers = {}
    for _ in range(5):
        prog.lnl(st["cur_x"], timers=timers)
    torch.cuda.synchronize()
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}
    peak, peak_src = fp64_peak()
    fl = prog.flops
    alg = fl["bin_nnz"] + fl["build_ops"]
    ach = alg * per_gpu / (stage_ms["bin_residual"] * 1e-3) / 1e12
    e2e = time_e2e(torch, dist, size, prog, st["cur_x"], 20)
    cpu = None
    if dist.get_rank() == 0 and size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_scripts as O
        from oracle import runtime as ort
        oslik = ort.Slik(O.spt_script(data, "s12")())
        x0 = np.array([1.0, 20.0, 5.0, 5.0]); sc = np.array([0.026, 3.0, 3.0, 3.0]) * 0.3
        X = x0 + sc * np.random.RandomState(4).standard_normal((8192, 4))
        cpu = time_oracle_evals(oslik, X, 3.0, "the spt_lowl('s12') script")
    return _secondary_line(
        "spt_lowl_mh", METRIC.replace("bandpower lnL", "spt_lowl('s12') lnL, adaptive MH"),
        nch * nsteps / (ms * 1e-3), ms / nsteps, nsteps,
        {"workload": "reference likelihoods.spt_lowl('s12') on its bundled bandpowers / windows / covariance, synthetic TT",
         "chains_total": nch, "chains_per_gpu": per_gpu, "n_data": 47, "ndim": 4,
         "sampler": "adaptive MH, pooled update every 100 steps", "adaptations": len(sl.sampler._adapt),
         "loop": st.get("loop"), "graph_nodes_per_block": st.get("graph_nodes")},
        {"bound": "tensor", "kernel": "k_bin_residual", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
         "traffic": None, "peak_source": peak_src, "stages_ms": stage_ms,
         "note": "4096 chains are 32 walker tiles x 1 window tile: a fraction of one wave, so this line is launch / "
                 "latency bound; the kernel's roofline figure at scale is the spt_multifreq entry"},
        clocks, e2e, cpu, {"scaling": "weak", "host_enqueue_ms_per_step": st.get("host_ms_per_step"),
                           "acceptance": (st.get("accepted") or 0) / float(per_gpu * nsteps)})


def run_single_secondary(args):
    import torch
    from cosmoslik_b200 import dist
    rank, size = dist.init_from_env()
    if size == 1:
        torch.cuda.set_device(0)
    fn = {"quickstart": sec_quickstart, "gauss30_mh": sec_gauss30, "spt_multifreq": sec_spt_multifreq,
          "unbinned_mh": sec_unbinned, "spt_lowl_mh": sec_spt_lowl}[args.workload]
    rec = fn(args, torch, dist)
    if rank == 0 and rec is not None:
        rec.update({"n_gpus": size, "warmup": max(args.warmup, 3), "vs_baseline": None})
        rec.setdefault("scaling", "weak")
        emit(rec)


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
    if args.impl == "reference":
        if args.workload not in ("planck_lite", "spt_multifreq"):
            args.workload = "planck_lite"
        return run_reference(args)
    if args.workload in ("planck_lite", "spt_multifreq"):
        return run_b200(args)
    return run_single_secondary(args)


if __name__ == "__main__":
    main()
