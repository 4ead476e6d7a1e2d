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
from .rowstream import RowBlockStream, blocks_of
from .utils import append_extras_of_next, extras_of, initialize_covariance, program_of


class emcee(SlikSampler):
    pool_chains = True    # run_chain joins all walkers into one Chain, as the reference does

    def __init__(self, params, num_samples, nwalkers=100, cov_est=None, output_extra_params=None,
                 output_file=None, output_freq=10, a=2.0, seed=None, streams=None, init_pos=None):
        if output_extra_params is None:
            output_extra_params = []
        super().__init__(**arguments(exclude=["params"]))
        # name -> dtype name (emcee.py:28); scalar numeric columns, evaluated on the device in float64
        # (Slik.extras_program) and cast to the column's dtype in the file
        self.output_extra_params = OrderedDict(k if isinstance(k, tuple) else (k, np.dtype("float").name)
                                               for k in self.output_extra_params)
        for k_, dt in self.output_extra_params.items():
            d = np.dtype(dt)
            if d.kind not in "fiub" or d.shape != ():
                raise NotImplementedError("output_extra_params: column '%s' has dtype %s; the batched sampler derives "
                                          "scalar numeric columns only" % (k_, dt))
        self.sampled = params.find_sampled()
        self.x0 = [params[k].start for k in self.sampled]
        self.cov_est = initialize_covariance(self.sampled, self.cov_est)
        self.stats = {}

    def initial_positions(self, seed=None):
        """Scatter of the walkers around the start point (emcee.py:49-53)."""
        if self.init_pos is not None:
            return np.ascontiguousarray(self.init_pos, dtype=float)
        seed = self.seed if seed is None else seed
        rs = np.random.RandomState(None if seed is None else int(seed) % (2 ** 32))
        return rs.multivariate_normal([v.start for v in self.sampled.values()], self.cov_est, size=int(self.nwalkers))

    def sample(self, lnl):
        for blocks in self._run(program_of(lnl), emit=True, extras=self._file_extras(lnl)):
            for wid, rows in blocks:
                for r in rows:
                    s = sample(r[2:].copy(), r[0], r[1], None)
                    s.chain = wid
                    yield s

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

        def local_ids(r):
            a0, b0 = dist.shard_bounds(h, r, size)               # slice of the first half
            a1, b1 = dist.shard_bounds(k - h, r, size)           # slice of the second half
            return a0, b0, a1, b1, np.concatenate([np.arange(a0, b0), h + np.arange(a1, b1)])

        lo0, hi0, lo1, hi1, gid = local_ids(rank)                # global walker ids of the local rows
        n0, n1 = hi0 - lo0, hi1 - lo1
        W = n0 + n1
        if nsteps is None:
            nsteps = int(ceil(self.num_samples / float(k)))
        # an unseeded run draws its key on rank 0: every rank must use the same key AND the same start scatter
        seed = int(self.seed) if self.seed is not None else \
            dist.broadcast_int(int(np.random.SeedSequence().entropy & (2 ** 63 - 1)))
        kw = dict(dtype=torch.float64, device=dev)
        rl = 2 + nd
        F = max(int(self.output_freq), 1)
        consume = emit or self.output_file is not None
        # Rows: with a consumer (generator or chain file) the device holds ONE block of output_freq steps per
        # walker and hands it to the host asynchronously (rowstream); without one (run(): batched use, tests,
        # device statistics) the whole history of the call stays on the device.
        ring = consume
        cap = F if ring else max(nsteps, int(getattr(self, "row_capacity", 0)), 1)
        need = W * cap * rl * 8
        if need > 0.8 * torch.cuda.mem_get_info(dev)[0]:
            raise MemoryError("row buffer for %d walkers x %d steps needs %.1f GB: run with an output_file / as a "
                              "generator (rows are then streamed in blocks of output_freq steps), or in shorter calls"
                              % (W, cap, need / 1e9))
        prev = self.stats if (resume and self.stats and tuple(self.stats["pos"].shape) == (W, nd)) else None
        if prev is None:
            pos_all = self.initial_positions(seed)
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
        # ONE chain file for all ranks, written by rank 0 (the reference's master writes one file)
        file_all = self.output_file is not None and size > 1
        writer = None
        if self.output_file is not None and rank == 0:
            writer = ChainWriter(self.output_file, ["lnl", "weight"] + list(self.sampled) + list(self.output_extra_params), "list",
                                 dtypes=["f8"] * (2 + nd) + [np.dtype(d).str for d in self.output_extra_params.values()])
        all_ids = np.concatenate([local_ids(r)[4] for r in range(size)]) if file_all else gid
        per_rank = [len(local_ids(r)[4]) for r in range(size)]
        stream = RowBlockStream(torch, dev, W, cap, rl, all_ranks=file_all, total=k, counts=per_rank) if consume else None
        local_stream = RowBlockStream(torch, dev, W, cap, rl) if (consume and file_all and emit and rank != 0) else None
        status = replica.sync if replica is not None else None
        st = L.stream_ptr()
        timers = getattr(self, "timers", None)     # bench.py: CUDA-event pairs per stage
        nacc = torch.zeros((), dtype=torch.int64, device=dev)
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

        pending = [None]

        def submit_block():
            """end of an output block (stream order): hand its rows over, start an empty block"""
            pos_snap = None
            if extras is not None:                 # the file's derived columns belong to where a walker moved TO
                pos_snap = dist.all_gather_rows(pos.clone(), k, per_rank) if file_all else pos.clone()
            ticket = (stream.submit(rows, nrows, status),
                      local_stream.submit(rows, nrows) if local_stream is not None else None, pos_snap)
            nrows.zero_()
            return ticket

        def hand_over(ticket):
            """host side of a block: file records on rank 0 (every rank's walkers), local rows to the consumer"""
            slot, lslot, pos_snap = ticket
            got = stream.collect(slot)
            mine = None
            if got is not None:
                blocks = blocks_of(got[0], got[1], all_ids)
                if writer is not None:
                    recs = blocks if extras is None else append_extras_of_next(blocks, pos_snap.cpu().numpy(), extras, nd)
                    for wid, r in recs:
                        for c0 in range(0, len(r), F):      # records of at most output_freq rows (emcee.py:80-84)
                            writer.write(wid, r[c0:c0 + F])
                    writer.flush()
                if file_all:
                    at = int(sum(per_rank[:rank]))
                    mine = blocks[at:at + W]
                else:
                    mine = blocks
            if lslot is not None:
                got = local_stream.collect(lslot)
                mine = blocks_of(got[0], got[1], gid)
            return mine

        def end_of_block():
            """submit this block, then consume the PREVIOUS one on the host while the device runs on"""
            out = None
            ticket = submit_block()
            if pending[0] is not None:
                out = hand_over(pending[0])
            pending[0] = ticket
            return out

        hooks = getattr(self, "hooks", None) or {}      # bench.py: callables run right before the first / after the last launch
        if "start" in hooks:
            hooks["start"]()
        # Fast path: whole steps are enqueued from C (csl_stretch_run) -- no Python between the kernel
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
                ens.sync = replica.sync.data_ptr()

        try:
            i = 0
            while native and i < nsteps:
                # one call per output block (or for the whole run when nobody consumes rows)
                nb = min(F - (i % F), nsteps - i) if consume else nsteps - i
                L.check(lib.csl_stretch_run(C.byref(prog.struct), C.byref(ens), nb, step0 + i, step0 + nsteps - 1, st),
                        "csl_stretch_run")
                if replica is not None:
                    replica._epoch = int(ens.epoch)
                i += nb
                if consume and (i % F == 0 or i == nsteps):
                    out = end_of_block()
                    if emit and out is not None:
                        yield out
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
                if consume and ((i + 1) % F == 0 or i == nsteps - 1):
                    out = end_of_block()
                    if emit and out is not None:
                        yield out
            if "end" in hooks:
                hooks["end"]()
            if pending[0] is not None:
                out = hand_over(pending[0])
                pending[0] = None
                if emit and out is not None:
                    yield out
        finally:
            if writer is not None:
                writer.close()
        host_ms = 1e3 * (time.perf_counter() - t_host0) / max(nsteps, 1)   # CPU time to ENQUEUE one step
        exchange = ("none" if size == 1 else "nccl all_gather" if replica is None else
                    "accept kernel stores to every rank's replica over NVLink (%s); release signal in its last CTA, "
                    "acquire wait at the head of the next proposal kernel"
                    % ("multimem.st multicast" if replica.multicast else "peer pointers"))
        n_acc = int(nacc.item())                  # synchronises: the status word below is final
        if replica is not None:
            replica.check()
        self.stats = dict(exchange=exchange, host_ms_per_step=host_ms, pos=pos, lnl=lnl_pos, weight=weight,
                          rows=None if ring else rows, nrows=None if ring else nrows, nsteps=nsteps,
                          steps_done=step0 + nsteps, walker_ids=gid, seed=seed, accepted=n_acc, evals=nsteps * W,
                          row_storage_bytes=int(rows.numel() * 8 + (stream.W * stream.width * 16 if stream else 0)))
