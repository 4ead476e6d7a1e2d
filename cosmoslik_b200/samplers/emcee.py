"""Affine-invariant ensemble sampler (stretch move), batched on the device.

Constructor keywords and ``sample(lnl)`` follow the reference wrapper
(samplers/emcee.py:13-88).  The reference delegates the move itself to the
third-party package ``emcee`` (2.x API); here it is implemented directly
(Goodman & Weare 2010; emcee 2.x ``_propose_stretch``): the ensemble is split
in two halves, each half is updated against the other,
    z = ((a-1) u + 1)^2 / a,   q = c_j - z (c_j - s),
    accept iff (ndim-1) ln z + lnp(q) - lnp(s) > ln u',
with ``csl_stretch_propose`` / ``DeviceProgram.lnl`` / ``csl_stretch_accept``.
The wrapper's bookkeeping -- weights of walkers that did not move, rows
``(lnl_next, weight, x_old)``, per-walker file records of ``output_freq`` rows
-- is samplers/emcee.py:66-88 (``csl_emcee_rows``).

With several GPUs the walkers of BOTH halves are row-sharded; before each
half-update the complementary half's positions are all-gathered over NCCL
(this replaces the reference's ``mpi_map`` pool, mpi.py:40-64).  Partner
indices and uniforms are keyed by GLOBAL walker id, so the chain is identical
for any number of GPUs.

Extra keywords: ``a`` (stretch scale, emcee default 2), ``seed``, ``streams``
(callable ``step -> (u_z, j, u_acc)`` each [2, nwalkers/2], injected for parity
runs) and ``init_pos`` ([nwalkers, ndim], replaces the initial Gaussian scatter
of emcee.py:49-53).

Waived reference behaviour: emcee.py:74 yields ``sample(-l, x, weight)`` -- the
positional order of ``sample(x, lnl, weight)`` swapped, so ``run_chain`` labels
columns wrongly there.  Here the yielded objects carry x / lnl in their proper
fields; the row CONTENT (the lnl of the next position next to the old position)
is reproduced.
"""
import ctypes as C
import time
from collections import OrderedDict
from math import ceil

import numpy as np

from .. import _lib as L
from .. import dist
from ..chains import ChainWriter
from ..core import SlikSampler, arguments, sample
from .utils import append_extras_of_next, extras_of, initialize_covariance, program_of


class emcee(SlikSampler):
    pool_chains = True    # run_chain joins all walkers into one Chain, as the reference does

    def __init__(self, params, num_samples, nwalkers=100, cov_est=None, output_extra_params=None,
                 output_file=None, output_freq=10, a=2.0, seed=None, streams=None, init_pos=None):
        if output_extra_params is None:
            output_extra_params = []
        super().__init__(**arguments(exclude=["params"]))
        # name -> dtype name (emcee.py:28); float columns only, evaluated on the device (Slik.extras_program)
        self.output_extra_params = OrderedDict(k if isinstance(k, tuple) else (k, np.dtype("float").name)
                                               for k in self.output_extra_params)
        for k_, dt in self.output_extra_params.items():
            if np.dtype(dt) != np.dtype("float"):
                raise NotImplementedError("output_extra_params: column '%s' has dtype %s; the batched sampler "
                                          "writes float64 columns only" % (k_, dt))
        self.sampled = params.find_sampled()
        self.x0 = [params[k].start for k in self.sampled]
        self.cov_est = initialize_covariance(self.sampled, self.cov_est)
        self.stats = {}

    def initial_positions(self):
        """Scatter of the walkers around the start point (emcee.py:49-53)."""
        if self.init_pos is not None:
            return np.ascontiguousarray(self.init_pos, dtype=float)
        rs = np.random.RandomState(None if self.seed is None else int(self.seed) % (2 ** 32))
        return rs.multivariate_normal([v.start for v in self.sampled.values()], self.cov_est, size=int(self.nwalkers))

    def sample(self, lnl):
        for blocks in self._run(program_of(lnl), emit=True, extras=self._file_extras(lnl)):
            for wid, rows in blocks:
                for r in rows:
                    s = sample(r[2:].copy(), r[0], r[1], None)
                    s.chain = wid
                    yield s

    @staticmethod
    def _collect_rows(rows, nrows, prev_n, gid, rl):
        """device row buffer -> [(global walker id, new rows)] since the previous collection"""
        W = rows.shape[0]
        n_now = nrows.cpu().numpy().astype(np.int64)
        a_, b_ = (int(prev_n.min()), int(n_now.max())) if W else (0, 0)
        slab = rows[:, a_:b_].cpu().numpy() if b_ > a_ else np.zeros((W, 0, rl))
        return [(int(gid[c]), slab[c, prev_n[c] - a_:n_now[c] - a_]) for c in range(W)], n_now

    def run(self, lnl_or_program, nsteps=None, resume=False):
        """Batched entry point: advance the ensemble `nsteps` steps and return the state dict.
        ``resume=True`` continues from the state the previous call left (positions, lnl, weights)
        without re-evaluating the start positions."""
        if hasattr(lnl_or_program, "struct"):
            prog, extras = lnl_or_program, None
            if self.output_extra_params and self.output_file is not None:
                raise ValueError("output_extra_params: pass the bound Slik.evaluate, not the program, so that "
                                 "the derived quantities can be compiled")
        else:
            prog, extras = program_of(lnl_or_program), self._file_extras(lnl_or_program)
        for _ in self._run(prog, emit=False, nsteps=nsteps, resume=resume, extras=extras):
            pass
        return self.stats

    def _file_extras(self, lnl):
        """The reference's wrapper puts extra columns in the chain FILE only (emcee.py:74 yields
        ``sample(-l, x, weight)`` without them, :78 writes them)."""
        if self.output_file is None:
            return None
        return extras_of(lnl, list(self.output_extra_params))

    def _run(self, prog, emit, nsteps=None, resume=False, extras=None):
        torch, lib = prog.torch, prog.lib
        dev = prog.device
        nd = len(self.x0)
        k = int(self.nwalkers)
        if k % 2 or k < 2:
            raise ValueError("nwalkers must be even")
        h = k // 2
        rank, size = dist.get_rank(), dist.get_size()
        lo0, hi0 = dist.shard_bounds(h, rank, size)          # my slice of the first half
        lo1, hi1 = dist.shard_bounds(k - h, rank, size)      # my slice of the second half
        n0, n1 = hi0 - lo0, hi1 - lo1
        W = n0 + n1
        gid = np.concatenate([np.arange(lo0, hi0), h + np.arange(lo1, hi1)])   # global walker ids
        if nsteps is None:
            nsteps = int(ceil(self.num_samples / float(k)))
        seed = int(self.seed) if self.seed is not None else int(np.random.SeedSequence().entropy & (2 ** 63 - 1))
        kw = dict(dtype=torch.float64, device=dev)
        rl = 2 + nd
        cap = max(nsteps, int(getattr(self, "row_capacity", 0)), 1)
        prev = self.stats if (resume and self.stats and tuple(self.stats["pos"].shape) == (W, nd)) else None
        if prev is None:
            pos_all = self.initial_positions()
            if pos_all.shape != (k, nd):
                raise ValueError("initial positions must be [nwalkers, ndim]")
            pos = torch.from_numpy(pos_all[gid]).to(dev).contiguous()   # local: first-half then second-half slice
            lnl_pos = prog.lnl(pos)
        else:
            pos, lnl_pos = prev["pos"], prev["lnl"]
        halves = [(0, n0, lo0, h), (n0, W, h + lo1, k - h)]          # (row0, row1, first global id, half size)
        # several GPUs: a replica of the whole ensemble in NVLink symmetric memory, updated by the accept
        # kernel itself (fused exchange); otherwise NCCL all_gather of the complementary half
        replica = None
        if size > 1:                        # the rendezvous is expensive: one replica per sampler and shape
            cached = getattr(self, "_replica", None)
            if cached is not None and cached[0] == (k, nd, str(dev)):
                replica = cached[1]
            else:
                replica = dist.EnsembleReplica.create(k, nd, dev)
                dict.__setitem__(self, "_replica", ((k, nd, str(dev)), replica))
        if replica is not None and prev is None:
            replica.tensor.copy_(torch.from_numpy(pos_all).to(dev))
            torch.cuda.synchronize()
            dist.barrier()
        q = torch.empty_like(pos)
        zz = torch.empty(W, **kw)
        weight = prev["weight"] if prev is not None else torch.ones(W, **kw)
        buf = getattr(self, "_buffers", None)           # reuse the big row buffer across runs of equal shape
        if buf is None or buf[0] != (W, cap, rl, str(dev)):
            buf = ((W, cap, rl, str(dev)), torch.empty((W, cap, rl), **kw))
            dict.__setitem__(self, "_buffers", buf)
        rows = buf[1]
        nrows = torch.zeros(W, dtype=torch.int32, device=dev)
        accepted = torch.zeros(W, dtype=torch.uint8, device=dev)
        pos_old = pos.clone()                        # refreshed by csl_emcee_rows at the end of every step
        writer = None
        if self.output_file is not None:
            path = self.output_file if rank == 0 else "%s.rank%d" % (self.output_file, rank)
            writer = ChainWriter(path, ["lnl", "weight"] + list(self.sampled) + list(self.output_extra_params), "list")
        prev_n = np.zeros(W, dtype=np.int64)
        st = L.stream_ptr()
        timers = getattr(self, "timers", None)     # bench.py: CUDA-event pairs per stage
        nacc = torch.zeros((), dtype=torch.int64, device=dev)
        F = int(self.output_freq)
        t_host0 = time.perf_counter()
        step0 = int(prev["steps_done"]) if prev is not None else 0       # Philox step counter continues on resume

        class timed(object):
            """CUDA-event pair around a stage, appended to timers[name] (bench.py); no-op without timers"""
            def __init__(self, name):
                self.name = name

            def __enter__(self):
                if timers is not None:
                    self.a = torch.cuda.Event(enable_timing=True)
                    self.a.record()

            def __exit__(self, *exc):
                if timers is not None:
                    b_ev = torch.cuda.Event(enable_timing=True)
                    b_ev.record()
                    timers.setdefault(self.name, []).append((self.a, b_ev))

        def to_dev(arr, dtype):
            return torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(dev)

        def file_rows(blocks):
            """rows as the chain file holds them: with the extra columns, if any, appended"""
            if extras is None:
                return blocks
            return append_extras_of_next(blocks, pos.cpu().numpy(), extras, nd)

        # Fast path: whole steps are enqueued from C (csl_stretch_run) -- no Python between the ~11 kernel
        # launches of a step.  Needs device streams and either one GPU or the symmetric-memory replica.
        native = self.streams is None and timers is None and (size == 1 or replica is not None) \
            and not getattr(self, "python_loop", False)
        ens = None
        if native:
            ws = prog.workspace(max(n0, n1, 1))
            ens = L.Ensemble(ndim=nd, n0=n0, n1=n1, h0=h, h1=k - h, gid0=lo0, gid1=h + lo1, cap=cap, a=float(self.a),
                             seed=seed, pos=pos.data_ptr(), lnl=lnl_pos.data_ptr(), q=q.data_ptr(),
                             lnl_q=ws["lnl"].data_ptr(), zz=zz.data_ptr(), weight=weight.data_ptr(),
                             pos_old=pos_old.data_ptr(), rows=rows.data_ptr(), nrows=nrows.data_ptr(),
                             accepted=accepted.data_ptr(), naccepted=nacc.data_ptr(), ws_coef=ws["coef"].data_ptr(),
                             ws_prior=ws["prior"].data_ptr(), ws_R=L.ptr(ws["R"]), ws_chi2=L.ptr(ws["chi2"]),
                             ws_part=L.ptr(ws["part"]))
            if replica is not None:
                ens.replica, ens.peer_ptrs = replica.tensor.data_ptr(), replica.peers_dev.data_ptr()
                ens.flag_ptrs, ens.multicast_ptr = replica.flag_ptrs.data_ptr(), replica.multicast
                ens.epoch, ens.npeers, ens.rank = replica._epoch, replica.npeers, replica.handle.rank

        try:
            i = 0
            while native and i < nsteps:
                # one call per output block (or for the whole run when nobody consumes rows)
                nb = min(F - (i % F), nsteps - i) if (emit or writer is not None) else nsteps - i
                L.check(lib.csl_stretch_run(C.byref(prog.struct), C.byref(ens), nb, step0 + i, step0 + nsteps - 1, st),
                        "csl_stretch_run")
                if replica is not None:
                    replica._epoch = int(ens.epoch)
                i += nb
                if (emit or writer is not None) and (i % F == 0 or i == nsteps):
                    blocks, prev_n = self._collect_rows(rows, nrows, prev_n, gid, rl)
                    if writer is not None:
                        for wid, r in file_rows(blocks):
                            for c0 in range(0, len(r), F):
                                writer.write(wid, r[c0:c0 + F])
                        writer.flush()
                    if emit:
                        yield blocks
            for i in (range(nsteps) if not native else ()):
                step = step0 + i
                inj = self.streams(step) if self.streams is not None else None
                for hidx, (r0, r1, g0, hsize) in enumerate(halves):
                    o0, o1, _, csize = halves[1 - hidx]
                    with timed("all_gather"):
                        if replica is not None:
                            cbase = 0 if hidx == 1 else h                 # global row of the complementary half
                            comp = replica.tensor[cbase:cbase + csize]
                        else:
                            comp = dist.all_gather_rows(pos[o0:o1], csize)    # the complementary half, all ranks' rows
                    ns = r1 - r0
                    if ns == 0:
                        if replica is not None:
                            replica.barrier()
                        continue
                    uz = jj = ua = None
                    if inj is not None:
                        first = g0 - (0 if hidx == 0 else h)              # index of my first walker within its half
                        uz = to_dev(inj[0][hidx][first:first + ns], np.float64)
                        jj = to_dev(inj[1][hidx][first:first + ns], np.int64)
                        ua = to_dev(inj[2][hidx][first:first + ns], np.float64)
                    s_view, l_view, q_view = pos[r0:r1], lnl_pos[r0:r1], q[r0:r1]
                    with timed("propose"):
                        L.check(lib.csl_stretch_propose(ns, csize, nd, s_view.data_ptr(), comp.data_ptr(),
                                                        float(self.a), L.ptr(uz), L.ptr(jj), seed, step, hidx, g0,
                                                        q_view.data_ptr(), zz[r0:r1].data_ptr(), st),
                                "csl_stretch_propose")
                    lnl_q = prog.lnl(q_view, timers=timers)
                    if replica is None:
                        with timed("accept"):
                            L.check(lib.csl_stretch_accept(ns, csize, nd, s_view.data_ptr(), l_view.data_ptr(),
                                                           q_view.data_ptr(), lnl_q.data_ptr(), zz[r0:r1].data_ptr(),
                                                           L.ptr(ua), seed, step, hidx, g0,
                                                           accepted[r0:r1].data_ptr(), nacc.data_ptr(), st),
                                    "csl_stretch_accept")
                    else:
                        with timed("accept_publish"):
                            L.check(lib.csl_stretch_accept_publish(
                                ns, csize, nd, s_view.data_ptr(), l_view.data_ptr(), q_view.data_ptr(),
                                lnl_q.data_ptr(), zz[r0:r1].data_ptr(), L.ptr(ua), seed, step, hidx, g0,
                                accepted[r0:r1].data_ptr(), nacc.data_ptr(), replica.peers_dev.data_ptr(),
                                replica.npeers, replica.multicast, g0, st), "csl_stretch_accept_publish")
                        with timed("replica_barrier"):
                            replica.barrier()      # every rank's stores have landed before anyone proposes again
                L.check(lib.csl_emcee_rows(W, nd, pos_old.data_ptr(), pos.data_ptr(), lnl_pos.data_ptr(),
                                           weight.data_ptr(), 1 if i == nsteps - 1 else 0, rows.data_ptr(),
                                           nrows.data_ptr(), cap, st), "csl_emcee_rows")
                if (emit or writer is not None) and ((i + 1) % F == 0 or i == nsteps - 1):
                    blocks, prev_n = self._collect_rows(rows, nrows, prev_n, gid, rl)
                    if writer is not None:
                        for wid, r in file_rows(blocks):
                            for c0 in range(0, len(r), F):      # records of at most output_freq rows (emcee.py:80-84)
                                writer.write(wid, r[c0:c0 + F])
                        writer.flush()
                    if emit:
                        yield blocks
        finally:
            if writer is not None:
                writer.close()
        host_ms = 1e3 * (time.perf_counter() - t_host0) / max(nsteps, 1)   # CPU time to ENQUEUE one step
        exchange = ("none" if size == 1 else "nccl all_gather" if replica is None else
                    "accept kernel stores to every rank's replica over NVLink (%s) + device barrier"
                    % ("multimem.st multicast" if replica.multicast else "peer pointers"))
        self.stats = dict(exchange=exchange, host_ms_per_step=host_ms, pos=pos, lnl=lnl_pos, weight=weight, rows=rows,
                          nrows=nrows, nsteps=nsteps, steps_done=step0 + nsteps, walker_ids=gid, seed=seed,
                          accepted=int(nacc.item()), evals=nsteps * W)
