"""Block hand-off of chain rows from the device to the host (and, for the chain file, to rank 0).

The reference's samplers hand their rows over in blocks: ``mpi_comm_freq`` rows per chain to the master
process (metropolis_hastings.py:219-241), ``output_freq`` steps per walker to the file (emcee.py:77-84).
Here a block is a device tensor ``rows [W, capb, 2+ndim]`` with per-chain counts; ``RowBlockStream`` moves
it to PINNED host memory with one asynchronous copy on a side stream, ordered by CUDA events, so the next
block's kernels run while the host consumes this one:

    submit(rows_blk, counts)   compute stream: pack rows + counts into one staging tensor (two of them
                               alternate), [all-gather the ranks' staging tensors], record an event;
                               side stream: wait for it, ONE cudaMemcpyAsync into pinned memory.
    collect(ticket)            host: wait for that copy only, return numpy views.

With ``all_ranks=True`` every rank's block reaches rank 0 (one chain file for all ranks, as the reference's
master writes one file, metropolis_hastings.py:181-186, 209-212); the gather is a device collective, the
host copy happens on rank 0 only.
"""
import numpy as np

from .. import dist


class RowBlockStream(object):

    def __init__(self, torch, device, W, capb, rl, all_ranks=False, total=None, counts=None):
        self.torch, self.dev = torch, device
        self.W, self.capb, self.rl = int(W), int(capb), int(rl)
        self.width = self.capb * self.rl + 1               # rows of one chain, then its count
        self.all_ranks = bool(all_ranks) and dist.get_size() > 1
        self.total = int(total) if self.all_ranks else self.W
        self.counts = counts                               # chains of every rank (None: shard_bounds shares)
        self.to_host = (not self.all_ranks) or dist.get_rank() == 0
        kw = dict(dtype=torch.float64, device=device)
        self.stage = [torch.empty((self.W, self.width), **kw) for _ in range(2)]
        hrows = self.total if self.to_host else 0
        self.host = [torch.empty((hrows, self.width), dtype=torch.float64).pin_memory() if hrows else None
                     for _ in range(2)]
        self.side = torch.cuda.Stream(device=device)
        self.done = [None, None]
        self.status_host = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.has_status = [False, False]
        self.k = 0
        self.bytes_per_block = hrows * self.width * 8

    def submit(self, rows_blk, counts, status=None):
        """rows_blk [W, capb, rl] float64 and counts [W] (any integer dtype), both on the device, as they
        stand in stream order NOW (they may be overwritten by later launches).  `status`: optional device
        int32 word (the sticky exchange-status word of a sharded ensemble) copied along; collect() raises if
        it is non-zero.  Returns a ticket."""
        torch = self.torch
        slot = self.k % 2
        self.k += 1
        cur = torch.cuda.current_stream(self.dev)
        if self.done[slot] is not None:
            cur.wait_event(self.done[slot])                 # the staging tensor's previous copy has left
        st = self.stage[slot]
        if self.W:
            st[:, :-1].copy_(rows_blk.reshape(self.W, -1))
            st[:, -1].copy_(counts)
        src = dist.all_gather_rows(st, self.total, self.counts) if self.all_ranks else st
        self.has_status[slot] = status is not None
        if self.to_host or status is not None:
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                if self.to_host:
                    self.host[slot].copy_(src, non_blocking=True)     # one copy per block
                if status is not None:
                    self.status_host[slot].copy_(status[:1], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.side)
            if src is not st:
                src.record_stream(self.side)
            self.done[slot] = done
        return slot

    def collect(self, slot):
        """(rows [n, capb, rl], counts [n]) of a submitted block as numpy views of the pinned buffer
        (n = all ranks' chains on the gathering rank 0); None on ranks that receive nothing."""
        if self.done[slot] is not None and (self.to_host or self.has_status[slot]):
            self.done[slot].synchronize()
        if self.has_status[slot] and int(self.status_host[slot][0]) != 0:
            from .._lib import CslError
            raise CslError("cosmoslik_b200: a peer did not reach a stretch-move half-step within "
                           "CSL_BARRIER_TIMEOUT_S; the exchange timed out and this run's results are void")
        if not self.to_host:
            return None
        h = self.host[slot].numpy()
        counts = h[:, -1].astype(np.int64)
        if counts.size and counts.max() > self.capb:
            raise RuntimeError("row block overflow: a chain produced %d rows in a block of capacity %d"
                               % (int(counts.max()), self.capb))
        return h[:, :-1].reshape(-1, self.capb, self.rl), counts


def blocks_of(rows, counts, ids):
    """[(chain id, its new rows [m, rl])] from a collected block"""
    return [(int(ids[c]), rows[c, :counts[c]]) for c in range(len(ids))]
