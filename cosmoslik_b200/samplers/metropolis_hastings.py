"""Batched adaptive Metropolis-Hastings.

Same constructor keywords and ``sample(lnl)`` generator protocol as the
reference sampler (samplers/metropolis_hastings.py:30-125).  ``nchains`` chains
advance in lock step as rows of device tensors; every step is
  csl_mh_propose -> DeviceProgram.lnl (priors, foregrounds, binning, chi^2) -> csl_mh_accept,
or, for a dense Gaussian posterior, the whole block of steps is one launch of
``csl_mh_gauss_run``.  Per chain the accept rule, the weights and the emitted
rows are those of ``_mcmc`` (:144-164).

Where the reference runs N worker processes and a master that pools their
samples (:188-241), chains here are rows; every ``mpi_comm_freq`` steps the
proposal covariance is replaced by the pooled weighted covariance of the last
half of every chain's rows (``get_new_cov``, :249-253), computed on the device
and, with several GPUs, summed by one NCCL all-reduce.  As in the reference's
MPI mode each of the N chains takes ``num_samples / N`` steps (:150).

The step loop is native: ``csl_mh_run`` enqueues propose+priors+bytecode ->
residual -> chi^2 -> combine+accept+rows for every step of a block from C, and a
block of ``mpi_comm_freq`` steps is captured ONCE as a CUDA graph and replayed
per block (the step counter lives in device memory), so the host issues one
call per block.  The proposal update never leaves the device: one fused pass
gives ``[sum w, sum w dx, sum w dx dx^T]`` about the start point, ONE all-reduce
pools the GPUs, and ``csl_mh_adapt_factor`` writes the new covariance and its
factor in place.  (Runs with injected NORMAL streams keep the host SVD factor
numpy's ``multivariate_normal`` builds, so their proposals equal the reference's.)

Rows reach the host in blocks through ``rowstream.RowBlockStream`` (pinned memory,
one asynchronous copy per block on a side stream); with several GPUs rank 0
receives every rank's block and writes ONE chain file (:181-186, 209-212).

Extra keywords (not in the reference): ``nchains``, ``seed`` (Philox key for the
on-device streams) and ``streams`` (dict of injected host arrays for parity
runs: 'z' [steps, nchains, ndim] standard normals or 'delta' offsets, and 'u'
[steps, nchains] uniforms).
"""
import ctypes as C
import os
import time
from collections import OrderedDict

import numpy as np

from .. import _lib as L
from .. import dist
from ..chains import ChainWriter
from ..core import SlikSampler, arguments, sample
from .rowstream import RowBlockStream, blocks_of
from .utils import append_extras, extras_of, initialize_covariance, program_of, proposal_factor


class metropolis_hastings(SlikSampler):

    def __init__(self, params, output_file=None, output_extra_params=None, num_samples=100, print_level=0,
                 cov_est=None, proposal_scale=2.4, proposal_update=True, proposal_update_start=1000,
                 mpi_comm_freq=100, max_weight=10, debug_output=False, temp=1, reseed=True,
                 yield_rejected=False, nchains=1, seed=None, streams=None):
        if output_extra_params is None:
            output_extra_params = []
        super().__init__(**arguments(exclude=["params"]))
        # name -> dtype name, as the reference (:103).  The values are scalar expressions of the sampled parameters
        # evaluated on the device in float64 (Slik.extras_program) and cast to the column's dtype; any scalar
        # numeric dtype works, array- or object-valued columns (:46-52) cannot be derived that way
        self.output_extra_params = OrderedDict(k if isinstance(k, tuple) else (k, np.dtype("float").name)
                                               for k in self.output_extra_params)
        for k, dt in self.output_extra_params.items():
            d = np.dtype(dt)
            if d.kind not in "fiub" or d.shape != ():
                raise NotImplementedError("output_extra_params: column '%s' has dtype %s; the batched sampler derives "
                                          "scalar numeric columns only" % (k, dt))
        self.sampled = params.find_sampled()
        self.x0 = [params[k].start for k in self.sampled]
        self.cov_est = initialize_covariance(self.sampled, self.cov_est)
        self._adapt = []            # (step, covariance: numpy array or device tensor) per adaptation
        self.stats = {}

    @property
    def adaptations(self):
        """(step, cov_est) history; covariances computed on the device are copied to the host on first access"""
        hist = self._adapt
        for i, (t, c) in enumerate(hist):
            if not isinstance(c, np.ndarray):
                hist[i] = (t, c.cpu().numpy())
        return hist

    # ------------------------------------------------------------------
    def sample(self, lnl):
        """Generator of ``sample(x, lnl, weight)`` objects (attribute ``chain`` = chain id),
        chain by chain within each block of ``mpi_comm_freq`` steps."""
        nd, extra_names = len(self.x0), list(self.output_extra_params)
        for blocks in self._run(program_of(lnl), emit=True, extras=extras_of(lnl, extra_names)):
            for cid, rows in blocks:
                for r in rows:
                    extra = None
                    if extra_names:
                        extra = {k: np.dtype(dt).type(v) for (k, dt), v in zip(self.output_extra_params.items(), r[2 + nd:])}
                    s = sample(r[2:2 + nd].copy(), r[0], r[1], extra)
                    s.chain = cid
                    yield s

    def run(self, lnl_or_program, keep_rows=True):
        """Batched entry point: run to completion and return a dict with the final state
        (device tensors) and the device row buffer ``rows [nchains_local, cap, 2+ndim]`` with per-chain
        counts ``nrows``."""
        if hasattr(lnl_or_program, "struct"):
            prog, extras = lnl_or_program, None
            if self.output_extra_params:
                raise ValueError("output_extra_params: pass the bound Slik.evaluate, not the program, so that "
                                 "the derived quantities can be compiled")
        else:
            prog, extras = program_of(lnl_or_program), extras_of(lnl_or_program, list(self.output_extra_params))
        for _ in self._run(prog, emit=False, extras=extras):
            pass
        return self.stats

    def gelman_rubin(self, burn_rows=0):
        """Potential scale reduction R-hat per sampled parameter over ALL chains of the last run
        (new functionality; the reference has no convergence statistic).  Per-chain weighted
        count / mean / variance are reduced on the device (``csl_chain_moments``); with several
        GPUs the per-chain moments are all-gathered; the O(nchains x ndim) combination is
        chains.rhat_from_moments."""
        from ..chains import rhat_from_moments
        st = self.stats
        if not st:
            raise RuntimeError("gelman_rubin: run the sampler first")
        rows, nrows = st["rows"], st["nrows"]
        W, cap, rl = rows.shape
        nd = rl - 2
        torch = __import__("torch")
        out = torch.empty((W, 1 + 2 * nd), dtype=torch.float64, device=rows.device)
        L.check(L.lib().csl_chain_moments(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, int(burn_rows),
                                          out.data_ptr(), L.stream_ptr()), "csl_chain_moments")
        allm = dist.all_gather_rows(out, int(self.nchains)).cpu().numpy()
        return rhat_from_moments(allm[:, 0], allm[:, 1:1 + nd], allm[:, 1 + nd:])

    # ------------------------------------------------------------------
    def _run(self, prog, emit, extras=None):
        torch, lib = prog.torch, prog.lib
        dev = prog.device
        nd = len(self.x0)
        rank, size = dist.get_rank(), dist.get_size()
        ntot = int(self.nchains)
        if ntot < size:
            raise ValueError("nchains = %d is smaller than the number of GPUs (%d): every rank needs a chain" % (ntot, size))
        lo, hi = dist.shard_bounds(ntot, rank, size)
        W = hi - lo
        steps = int(self.num_samples / float(max(1, ntot)))
        S = max(int(self.mpi_comm_freq), 1)
        # an unseeded run draws its key on rank 0; every rank must use the same one
        seed = int(self.seed) if self.seed is not None else \
            dist.broadcast_int(int(np.random.SeedSequence().entropy & (2 ** 63 - 1)))
        kw = dict(dtype=torch.float64, device=dev)
        rl = 2 + nd
        yr = 1 if self.yield_rejected else 0
        per_step = 2 if yr else 1                    # a rejected step may emit the rejected point AND a flush row
        cap = max(steps * per_step, 1)
        # the whole row history stays on the device: the proposal update needs the last half of every chain's
        # rows (:249-253) and so does gelman_rubin(); the host receives it block by block
        need = W * cap * rl * 8
        free = torch.cuda.mem_get_info(dev)[0]
        if need > 0.8 * free:
            raise MemoryError("row buffer for %d chains x %d steps needs %.1f GB; lower num_samples per call"
                              % (W, steps, need / 1e9))
        cur_x = torch.tensor(self.x0, **kw).repeat(W, 1).contiguous()
        cur_lnl = prog.lnl(cur_x)
        cur_w = torch.zeros(W, **kw)
        test_x = torch.empty_like(cur_x)
        test_lnl = torch.empty(W, **kw)
        rows = torch.empty((W, cap, rl), **kw)
        nrows = torch.zeros(W, dtype=torch.int32, device=dev)
        accepted = torch.zeros(W, dtype=torch.uint8, device=dev)
        margin = torch.zeros(W, **kw)
        nacc = torch.zeros((), dtype=torch.int64, device=dev)
        step_ctr = torch.zeros(1, dtype=torch.int32, device=dev)          # Philox step of the next MH step
        m = 1 + nd + nd * nd
        mom_part = torch.empty((592, m), **kw)
        mom = torch.zeros(m, **kw)
        ref = torch.tensor(self.x0, **kw)                                # shift of the fused moments: the start point
        cov_dev = torch.empty((nd, nd), **kw)
        streams = self.streams or {}
        zs = streams.get("z"); ds = streams.get("delta"); us = streams.get("u")
        inj = ds if ds is not None else zs
        F = torch.from_numpy(proposal_factor(self.cov_est, nd, self.proposal_scale)).to(dev).contiguous()
        gauss = prog.gauss_direct is not None and prog.struct.nscalar == 0 and not yr
        # injected normals: the factor must be numpy's (host SVD) so that proposals equal the reference's
        device_adapt = zs is None
        # injected streams of one block live in fixed staging buffers (stable pointers for the graph)
        inj_dev = torch.empty((S, W, nd), **kw) if inj is not None else None
        u_dev = torch.empty((S, W), **kw) if us is not None else None
        ws = prog.workspace(max(W, 1))
        state = L.MHState(ndim=nd, W=W, cap=cap, chain0=lo, seed=seed, temp=float(self.temp),
                          max_weight=float(self.max_weight), yield_rejected=yr, inj_is_delta=1 if ds is not None else 0,
                          cur_x=cur_x.data_ptr(), cur_lnl=cur_lnl.data_ptr(), cur_weight=cur_w.data_ptr(),
                          test_x=test_x.data_ptr(), test_lnl=test_lnl.data_ptr(), rows=rows.data_ptr(),
                          nrows=nrows.data_ptr(), accepted=accepted.data_ptr(), margin=margin.data_ptr(),
                          naccepted=nacc.data_ptr(), F=F.data_ptr(), step=step_ctr.data_ptr(),
                          inj_z=L.ptr(inj_dev), inj_u=L.ptr(u_dev), ws_coef=ws["coef"].data_ptr(),
                          ws_prior=ws["prior"].data_ptr(), ws_R=L.ptr(ws["R"]), ws_chi2=L.ptr(ws["chi2"]),
                          ws_part=L.ptr(ws["part"]))
        use_graph = (not gauss) and os.environ.get("CSL_NO_GRAPH") is None and not getattr(self, "no_graph", False)
        graph = C.c_void_p()
        writer = None
        file_all = self.output_file is not None and size > 1          # rank 0 writes every rank's chains
        if self.output_file is not None and rank == 0:
            writer = _FileBlocks(self.output_file, ["lnl", "weight"] + list(self.sampled) + list(self.output_extra_params), S,
                                 single=(ntot == 1),
                                 dtypes=["f8"] * (2 + nd) + [np.dtype(d).str for d in self.output_extra_params.values()])
        if dist.is_master() and self.print_level >= 1 and self.output_file:
            print("Starting MCMC chain (%s)..." % self.output_file)
        consume = emit or self.output_file is not None
        capb = S * per_step
        stream = RowBlockStream(torch, dev, W, capb, rl, all_ranks=file_all, total=ntot) if consume else None
        local_stream = RowBlockStream(torch, dev, W, capb, rl) if (consume and file_all and emit and rank != 0) else None
        prev_n = torch.zeros(W, dtype=torch.int64, device=dev)
        ar = torch.arange(capb, dtype=torch.int64, device=dev)
        all_ids = np.arange(ntot)
        st = L.stream_ptr()
        done = 0
        t_host = 0.0
        pending = None

        def hand_over(ticket):
            """host side of one block: file records on rank 0 (all ranks' chains), local rows to the consumer"""
            slot, lslot = ticket
            got = stream.collect(slot)
            mine = None
            if got is not None:
                blocks = blocks_of(got[0], got[1], all_ids if file_all else lo + np.arange(W))
                if extras is not None:
                    blocks = append_extras(blocks, extras, nd)
                if writer is not None:
                    for cid, r in blocks:
                        writer.add(cid, r)
                    writer.flush()
                mine = blocks[lo:hi] if file_all else blocks
            if lslot is not None:
                got = local_stream.collect(lslot)
                mine = blocks_of(got[0], got[1], lo + np.arange(W))
                if extras is not None:
                    mine = append_extras(mine, extras, nd)
            if mine is not None and self.print_level >= 1:
                self._print_block(lo, mine)
            return mine

        hooks = getattr(self, "hooks", None) or {}      # bench.py: callables run right before the first / after the last block
        try:
            if "start" in hooks:
                hooks["start"]()
            while done < steps:
                nb = min(S, steps - done)
                t0 = time.perf_counter()
                if inj is not None:
                    inj_dev[:nb].copy_(torch.from_numpy(np.ascontiguousarray(inj[done:done + nb, lo:hi])))
                if us is not None:
                    u_dev[:nb].copy_(torch.from_numpy(np.ascontiguousarray(us[done:done + nb, lo:hi])))
                if gauss:
                    g = prog.gauss_direct
                    L.check(lib.csl_mh_gauss_run(C.byref(prog.struct), W, nb, g["L"].data_ptr(), g["mean"].data_ptr(),
                                                 g["perm"].data_ptr(), cur_x.data_ptr(), cur_lnl.data_ptr(),
                                                 cur_w.data_ptr(), L.ptr(inj_dev), 1 if ds is not None else 0,
                                                 L.ptr(u_dev), F.data_ptr(), seed, done, lo, float(self.temp),
                                                 float(self.max_weight), rows.data_ptr(), nrows.data_ptr(), cap, st),
                            "csl_mh_gauss_run")
                elif use_graph and nb == S:
                    if not graph:
                        L.check(lib.csl_mh_graph_create(C.byref(prog.struct), C.byref(state), S, C.byref(graph)),
                                "csl_mh_graph_create")
                    L.check(lib.csl_graph_launch(graph, st), "csl_graph_launch")
                else:
                    L.check(lib.csl_mh_run(C.byref(prog.struct), C.byref(state), nb, st), "csl_mh_run")
                done += nb
                # ---- pooled proposal update (metropolis_hastings.py:200-201, 249-253), on the device
                # (ignored for a single chain, as the reference ignores it without MPI: docstring :69-72)
                if (self.proposal_update and ntot > 1 and done % S == 0 and done > self.proposal_update_start
                        and done < steps):
                    if device_adapt:
                        L.check(lib.csl_mh_moments_fused(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, ref.data_ptr(),
                                                         mom_part.data_ptr(), mom.data_ptr(), st), "csl_mh_moments_fused")
                        dist.all_reduce_sum(mom)                      # ONE collective: [sum w, sum w dx, sum w dx dx^T]
                        L.check(lib.csl_mh_adapt_factor(nd, mom.data_ptr(), float(self.proposal_scale), ref.data_ptr(),
                                                        cov_dev.data_ptr(), F.data_ptr(), None, st), "csl_mh_adapt_factor")
                        self._adapt.append((done, cov_dev.clone()))
                    else:
                        L.check(lib.csl_mh_moments(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, None,
                                                   mom_part.data_ptr(), mom.data_ptr(), st), "csl_mh_moments")
                        dist.all_reduce_sum(mom[:1 + nd])
                        mean = (mom[1:1 + nd] / mom[0]).contiguous()
                        L.check(lib.csl_mh_moments(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, mean.data_ptr(),
                                                   mom_part.data_ptr(), mom.data_ptr(), st), "csl_mh_moments")
                        dist.all_reduce_sum(mom[1 + nd:])
                        h = mom.cpu().numpy()
                        self.cov_est = h[1 + nd:].reshape(nd, nd) / (h[0] - 1)
                        self._adapt.append((done, self.cov_est.copy()))
                        F.copy_(torch.from_numpy(proposal_factor(self.cov_est, nd, self.proposal_scale)))
                t_host += time.perf_counter() - t0
                # ---- hand the block's new rows to the consumer / the chain file: the block is gathered out of the
                # history on the device and copied to pinned memory asynchronously; the PREVIOUS block is consumed
                # on the host while this one is still being computed
                if consume:
                    idx = (prev_n[:, None] + ar[None, :]).clamp_(max=cap - 1)
                    blk = torch.gather(rows, 1, idx[:, :, None].expand(-1, -1, rl))
                    now = nrows.to(torch.int64)
                    cnt = now - prev_n
                    ticket = (stream.submit(blk, cnt), local_stream.submit(blk, cnt) if local_stream is not None else None)
                    prev_n = now
                    if pending is not None:
                        out = hand_over(pending)
                        if emit and out is not None:
                            yield out
                    pending = ticket
            if "end" in hooks:
                hooks["end"]()
            if pending is not None:
                out = hand_over(pending)
                pending = None
                if emit and out is not None:
                    yield out
            if int(nrows.max().item()) > cap:
                raise RuntimeError("row buffer overflow")
        finally:
            graph_nodes = int(lib.csl_graph_nodes(graph)) if graph else 0
            if graph:
                lib.csl_graph_destroy(graph)
            if writer is not None:
                writer.finish()
        if self._adapt and not isinstance(self._adapt[-1][1], np.ndarray):
            self.cov_est = self._adapt[-1][1].cpu().numpy()
        self.stats = dict(cur_x=cur_x, cur_lnl=cur_lnl, cur_weight=cur_w, rows=rows, nrows=nrows, steps=steps,
                          chains=(lo, hi), seed=seed, last_margin=margin, last_accepted=accepted,
                          accepted=int(nacc.item()) if not gauss else None,
                          host_ms_per_step=1e3 * t_host / max(steps, 1), graph_nodes=graph_nodes,
                          loop="csl_mh_gauss_run (one launch per block)" if gauss else
                               ("CUDA graph per block (csl_mh_graph_create)" if use_graph else "csl_mh_run"))

    def _print_block(self, lo, blocks):
        for cid, r in blocks[:8]:
            if len(r):
                tot = r[:, 1].sum()
                print("\033[1m\033[95mChain %i:\033[0m %i/%i(%.1f%%) best=%.2f" %
                      (cid, (r[:, 1] > 0).sum(), tot, 100.0 * (r[:, 1] > 0).sum() / max(tot, 1), r[:, 0].min()))


class _FileBlocks(object):
    """Chain-file block structure of the reference.

    single=True is the single-process branch (metropolis_hastings.py:219-244): one
    record per ``freq`` yielded rows holding ``samples[:i]`` -- i.e. WITHOUT the block's
    last row (:237) -- and a final short (possibly empty) record.
    single=False is the MPI branch (:209-212, 232): every chain's rows are all written,
    ``freq`` rows per record, under the chain's id."""

    def __init__(self, path, names, freq, single, dtypes=None):
        self.w = ChainWriter(path, names, "ndarray", dtypes=dtypes)
        self.freq, self.single = freq, single
        self.buf = {}
        self.ncol = len(names)

    def add(self, cid, rows):
        if len(rows) == 0 and cid in self.buf:
            return
        b = self.buf.get(cid)
        b = rows.copy() if b is None else np.concatenate([b, rows])
        while len(b) >= self.freq:
            blk, b = b[:self.freq], b[self.freq:]
            self.w.write(0 if self.single else cid, blk[:-1] if self.single else blk)
        self.buf[cid] = b

    def flush(self):
        self.w.flush()

    def finish(self):
        if self.single and not self.buf:
            self.buf[0] = np.zeros((0, self.ncol))
        for cid, b in self.buf.items():
            if self.single:
                self.w.write(0, b[:-1] if len(b) else b.reshape(0, self.ncol))
            elif len(b):
                self.w.write(cid, b)
        self.w.close()
