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

Extra keywords (not in the reference): ``nchains``, ``seed`` (Philox key for the
on-device streams) and ``streams`` (dict of injected host arrays for parity
runs: 'z' [steps, nchains, ndim] standard normals or 'delta' offsets, and 'u'
[steps, nchains] uniforms).
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from .. import _lib as L
from .. import dist
from ..chains import ChainWriter
from ..core import SlikSampler, arguments, sample
from .utils import append_extras, extras_of, initialize_covariance, program_of, proposal_factor


class metropolis_hastings(SlikSampler):

    def __init__(self, params, output_file=None, output_extra_params=None, num_samples=100, print_level=0,
                 cov_est=None, proposal_scale=2.4, proposal_update=True, proposal_update_start=1000,
                 mpi_comm_freq=100, max_weight=10, debug_output=False, temp=1, reseed=True,
                 yield_rejected=False, nchains=1, seed=None, streams=None):
        if output_extra_params is None:
            output_extra_params = []
        super().__init__(**arguments(exclude=["params"]))
        if yield_rejected:
            raise NotImplementedError("yield_rejected is not supported by the batched sampler")
        # name -> dtype name, as the reference (:103); the values are scalar expressions of the sampled
        # parameters evaluated on the device (Slik.extras_program), so only float columns exist
        self.output_extra_params = OrderedDict(k if isinstance(k, tuple) else (k, np.dtype("float").name)
                                               for k in self.output_extra_params)
        for k, dt in self.output_extra_params.items():
            if np.dtype(dt) != np.dtype("float"):
                raise NotImplementedError("output_extra_params: column '%s' has dtype %s; the batched sampler "
                                          "writes float64 columns only" % (k, dt))
        self.sampled = params.find_sampled()
        self.x0 = [params[k].start for k in self.sampled]
        self.cov_est = initialize_covariance(self.sampled, self.cov_est)
        self.adaptations = []       # (step, cov_est) history, for inspection
        self.stats = {}

    # ------------------------------------------------------------------
    def sample(self, lnl):
        """Generator of ``sample(x, lnl, weight)`` objects (attribute ``chain`` = chain id),
        chain by chain within each block of ``mpi_comm_freq`` steps."""
        nd, extra_names = len(self.x0), list(self.output_extra_params)
        for blocks in self._run(program_of(lnl), emit=True, extras=extras_of(lnl, extra_names)):
            for cid, rows in blocks:
                for r in rows:
                    s = sample(r[2:2 + nd].copy(), r[0], r[1], dict(zip(extra_names, r[2 + nd:])) if extra_names else None)
                    s.chain = cid
                    yield s

    def run(self, lnl_or_program, keep_rows=True):
        """Batched entry point: run to completion and return a dict with the final state
        (device tensors) and, if ``keep_rows``, the device row buffer
        ``rows [nchains_local, cap, 2+ndim]`` with per-chain counts ``nrows``."""
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
        lo, hi = dist.shard_bounds(ntot, rank, size)
        W = hi - lo
        steps = int(self.num_samples / float(max(1, ntot)))
        S = int(self.mpi_comm_freq)
        seed = int(self.seed) if self.seed is not None else int(np.random.SeedSequence().entropy & (2 ** 63 - 1))
        kw = dict(dtype=torch.float64, device=dev)
        rl = 2 + nd
        cap = max(steps, 1)
        need = W * cap * rl * 8
        free = torch.cuda.mem_get_info(dev)[0]
        if need > 0.8 * free:
            raise MemoryError("row buffer for %d chains x %d steps needs %.1f GB; lower num_samples per call"
                              % (W, steps, need / 1e9))
        cur_x = torch.tensor(self.x0, **kw).repeat(W, 1).contiguous()
        cur_lnl = prog.lnl(cur_x)
        cur_w = torch.zeros(W, **kw)
        test_x = torch.empty_like(cur_x)
        rows = torch.empty((W, cap, rl), **kw)
        nrows = torch.zeros(W, dtype=torch.int32, device=dev)
        accepted = torch.zeros(W, dtype=torch.uint8, device=dev)
        margin = torch.zeros(W, **kw)
        m = 1 + nd + nd * nd
        mom_part = torch.empty((592, m), **kw)
        mom = torch.zeros(m, **kw)
        streams = self.streams or {}
        zs = streams.get("z"); ds = streams.get("delta"); us = streams.get("u")
        inj = ds if ds is not None else zs
        mode = 0 if ds is not None else (1 if zs is not None else 2)
        F = torch.from_numpy(proposal_factor(self.cov_est, nd, self.proposal_scale)).to(dev)
        fused = prog.gauss_direct is not None and prog.struct.nscalar == 0
        writer = None
        if self.output_file is not None:
            path = self.output_file if rank == 0 else "%s.rank%d" % (self.output_file, rank)
            writer = _FileBlocks(path, ["lnl", "weight"] + list(self.sampled) + list(self.output_extra_params), S,
                                 single=(ntot == 1))
        if dist.is_master() and self.print_level >= 1 and self.output_file:
            print("Starting MCMC chain (%s)..." % self.output_file)
        prev_n = np.zeros(W, dtype=np.int64)
        nacc = 0
        st = L.stream_ptr()
        done = 0
        try:
            while done < steps:
                nb = min(S, steps - done)
                if fused:
                    zb = ub = None
                    if inj is not None:
                        zb = torch.from_numpy(np.ascontiguousarray(inj[done:done + nb, lo:hi])).to(dev)
                    if us is not None:
                        ub = torch.from_numpy(np.ascontiguousarray(us[done:done + nb, lo:hi])).to(dev)
                    g = prog.gauss_direct
                    L.check(lib.csl_mh_gauss_run(C.byref(prog.struct), W, nb, g["L"].data_ptr(), g["mean"].data_ptr(),
                                                 g["perm"].data_ptr(), cur_x.data_ptr(), cur_lnl.data_ptr(),
                                                 cur_w.data_ptr(), L.ptr(zb), 1 if ds is not None else 0, L.ptr(ub),
                                                 F.data_ptr(), seed, done, lo, float(self.temp),
                                                 float(self.max_weight), rows.data_ptr(), nrows.data_ptr(), cap, st),
                            "csl_mh_gauss_run")
                else:
                    for t in range(done, done + nb):
                        it = None
                        if inj is not None:
                            it = torch.from_numpy(np.ascontiguousarray(inj[t, lo:hi])).to(dev)
                        ut = torch.from_numpy(np.ascontiguousarray(us[t, lo:hi])).to(dev) if us is not None else None
                        L.check(lib.csl_mh_propose(W, nd, cur_x.data_ptr(), test_x.data_ptr(), mode, L.ptr(it),
                                                   F.data_ptr(), seed, t, lo, st), "csl_mh_propose")
                        test_lnl = prog.lnl(test_x)
                        L.check(lib.csl_mh_accept(W, nd, cur_x.data_ptr(), cur_lnl.data_ptr(), cur_w.data_ptr(),
                                                  test_x.data_ptr(), test_lnl.data_ptr(), L.ptr(ut), seed, t, lo,
                                                  float(self.temp), float(self.max_weight), rows.data_ptr(),
                                                  nrows.data_ptr(), cap, accepted.data_ptr(), margin.data_ptr(), st),
                                "csl_mh_accept")
                done += nb
                # ---- pooled proposal update (metropolis_hastings.py:200-201, 249-253)
                # (ignored for a single chain, as the reference ignores it without MPI: docstring :69-72)
                if (self.proposal_update and ntot > 1 and done % S == 0 and done > self.proposal_update_start
                        and done < steps):
                    L.check(lib.csl_mh_moments(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, None,
                                               mom_part.data_ptr(), mom.data_ptr(), st), "csl_mh_moments")
                    dist.all_reduce_sum(mom[:1 + nd])
                    mean = (mom[1:1 + nd] / mom[0]).contiguous()
                    L.check(lib.csl_mh_moments(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, mean.data_ptr(),
                                               mom_part.data_ptr(), mom.data_ptr(), st), "csl_mh_moments")
                    dist.all_reduce_sum(mom[1 + nd:])
                    h = mom.cpu().numpy()
                    self.cov_est = h[1 + nd:].reshape(nd, nd) / (h[0] - 1)
                    self.adaptations.append((done, self.cov_est.copy()))
                    F = torch.from_numpy(proposal_factor(self.cov_est, nd, self.proposal_scale)).to(dev)
                # ---- hand the block's new rows to the consumer / the chain file
                if emit or writer is not None:
                    n_now = nrows.cpu().numpy().astype(np.int64)
                    if n_now.max(initial=0) > cap:
                        raise RuntimeError("row buffer overflow")
                    a, b = (int(prev_n.min()), int(n_now.max())) if W else (0, 0)
                    slab = rows[:, a:b].cpu().numpy() if b > a else np.zeros((W, 0, rl))
                    blocks = [(lo + c, slab[c, prev_n[c] - a:n_now[c] - a]) for c in range(W)]
                    prev_n = n_now
                    if extras is not None:
                        blocks = append_extras(blocks, extras, nd)
                    if writer is not None:
                        for cid, r in blocks:
                            writer.add(cid, r)
                        writer.flush()
                    if self.print_level >= 1:
                        self._print_block(lo, blocks)
                    if emit:
                        yield blocks
        finally:
            if writer is not None:
                writer.finish()
        self.stats = dict(cur_x=cur_x, cur_lnl=cur_lnl, cur_weight=cur_w, rows=rows, nrows=nrows, steps=steps,
                          chains=(lo, hi), seed=seed, last_margin=margin, last_accepted=accepted)

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

    def __init__(self, path, names, freq, single):
        self.w = ChainWriter(path, names, "ndarray")
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
