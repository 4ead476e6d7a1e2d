"""Walker sharding across the GPUs of one box.

Replaces the reference's MPI layer (cosmoslik/mpi/mpi.py:7-72 and the
master/worker protocol of metropolis_hastings.py:188-241): one process per GPU
(``torchrun``), ``torch.distributed`` with the NCCL backend over NVLink for
device tensors (gloo for the CPU tests), walkers / chains row-sharded.

  proposal adaptation   all_reduce(sum) of [sum w, sum w x, sum w xx^T]   (was M3-M5)
  stretch move          all_gather of the complementary half's positions  (was mpi_map, M2)
  Gelman-Rubin          all_gather of per-chain (n, mean, var)
"""
import os


def _dist():
    import torch.distributed as d
    return d if (d.is_available() and d.is_initialized()) else None


def get_rank():
    d = _dist()
    return d.get_rank() if d else 0


def get_size():
    d = _dist()
    return d.get_world_size() if d else 1


def is_master():
    return get_rank() == 0


def init_from_env(backend=None):
    """Initialise the process group from torchrun's environment (RANK, WORLD_SIZE,
    LOCAL_RANK, MASTER_ADDR/PORT) and bind this process to its GPU.  No-op when
    WORLD_SIZE is 1 or unset."""
    import torch
    import torch.distributed as d
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or d.is_initialized():
        return get_rank(), get_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        d.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        d.init_process_group(backend)
    return d.get_rank(), d.get_world_size()


def shutdown():
    """Tear the process group down (torchrun scripts call this last)."""
    try:
        import torch.distributed as d
        if d.is_available() and d.is_initialized():
            d.destroy_process_group()
    except Exception:
        pass


def shard_bounds(n, rank=None, size=None):
    """[lo, hi) of the contiguous slice of n items owned by `rank` (equal shares; the first
    n % size ranks take one more)."""
    rank = get_rank() if rank is None else rank
    size = get_size() if size is None else size
    base, extra = divmod(n, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sum(t):
    d = _dist()
    if d:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def all_gather_rows(local, total_rows, counts=None):
    """Concatenate the ranks' row blocks [n_r, ...] in rank order into [total_rows, ...].  `counts` = rows of
    every rank (default: the shares shard_bounds gives).  Equal shares go through all_gather_into_tensor;
    ragged shares are padded to the largest share, gathered and trimmed."""
    d = _dist()
    if not d:
        return local
    import torch
    size = d.get_world_size()
    tail = tuple(local.shape[1:])
    if counts is None:
        counts = [hi - lo for lo, hi in (shard_bounds(total_rows, r, size) for r in range(size))]
    counts = [int(c) for c in counts]
    if sum(counts) != total_rows or counts[d.get_rank()] != local.shape[0]:
        raise ValueError("all_gather_rows: shares %s do not add up to %d rows (local block has %d)"
                         % (counts, total_rows, local.shape[0]))
    if len(set(counts)) == 1:
        out = torch.empty((total_rows,) + tail, dtype=local.dtype, device=local.device)
        d.all_gather_into_tensor(out, local.contiguous())
        return out
    biggest = max(counts)
    padded = torch.zeros((biggest,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(size)]
    d.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def barrier():
    d = _dist()
    if d:
        d.barrier()


def _coll_device():
    """device collectives' tensors must live on: the GPU with NCCL, the host with gloo"""
    import torch
    d = _dist()
    if d and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_int(v, src=0):
    """rank `src`'s integer on every rank (seeds drawn from entropy must agree across ranks)"""
    d = _dist()
    if not d:
        return int(v)
    import torch
    t = torch.tensor([int(v)], dtype=torch.int64, device=_coll_device())
    d.broadcast(t, src)
    return int(t.item())


def all_agree(flag):
    """True iff `flag` is true on EVERY rank (ranks must not pick different code paths)"""
    d = _dist()
    if not d:
        return bool(flag)
    import torch
    t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=_coll_device())
    d.all_reduce(t, op=d.ReduceOp.MIN)
    return bool(t.item())


class EnsembleReplica(object):
    """Every rank's copy of ALL walkers' positions in NVLink symmetric memory
    (torch.distributed._symmetric_memory): the stretch-move accept kernel stores each local walker's
    final position straight into every rank's copy (csl_stretch_accept_publish), so no all-gather
    collective is issued; `barrier()` is a device-side barrier on the current stream.
    Only with the NCCL backend and one rank per GPU; `EnsembleReplica.create` returns None otherwise
    and the sampler falls back to NCCL all_gather."""

    def __init__(self, tensor, handle, peers_dev, multicast, flags=None, flag_handle=None, flag_ptrs=None, sync=None):
        self.tensor, self.handle, self.peers_dev, self.multicast = tensor, handle, peers_dev, multicast
        self.npeers = handle.world_size
        self.flags, self.flag_handle, self.flag_ptrs = flags, flag_handle, flag_ptrs
        self.sync = sync              # int32[2] on the device: sticky status word, ticket counter (csl_ensemble_t.sync)
        self._epoch = 0

    @classmethod
    def create(cls, nrows, ndim, device):
        d = _dist()
        if d is None or d.get_world_size() == 1 or d.get_backend() != "nccl":
            return None
        if os.environ.get("COSMOSLIK_B200_NO_SYMM", "0") == "1":
            return None
        made, err = None, None
        try:
            import torch
            import torch.distributed._symmetric_memory as symm
            t = symm.empty((nrows, ndim), dtype=torch.float64, device=device)
            h = symm.rendezvous(t, d.group.WORLD)
            ptrs = torch.tensor([int(p) for p in h.buffer_ptrs], dtype=torch.int64, device=device)
            mc = 0
            if os.environ.get("COSMOSLIK_B200_MULTICAST", "0") == "1":
                try:                       # NVSwitch multicast address (NVLS), when the fabric offers one
                    mc = int(h.multicast_ptr or 0)
                except Exception:
                    mc = 0
            # flags for the device-side ordering (csl_peer_barrier and the fused kernels): uint64[world] per rank
            f = symm.empty((h.world_size,), dtype=torch.int64, device=device)
            f.zero_()
            fh = symm.rendezvous(f, d.group.WORLD)
            fptrs = torch.tensor([int(p) for p in fh.buffer_ptrs], dtype=torch.int64, device=device)
            sync = torch.zeros(2, dtype=torch.int32, device=device)
            torch.cuda.synchronize()
            made = cls(t, h, ptrs, mc, f, fh, fptrs, sync)
        except Exception as e:      # symmetric memory unavailable on this system
            err = e
        # every rank must take the same exchange path: one failure sends all of them to NCCL all_gather
        if not all_agree(made is not None):
            if get_rank() == 0:
                print("cosmoslik_b200: symmetric-memory replica unavailable (%s); using NCCL all_gather" % (err,))
            return None
        d.barrier()
        return made

    def barrier(self):
        """device-side barrier on the current stream; the host does not wait"""
        from . import _lib as L
        self._epoch += 1
        L.check(L.lib().csl_peer_barrier(self.flag_ptrs.data_ptr(), self.handle.rank, self.npeers, self._epoch,
                                         self.sync.data_ptr(), L.stream_ptr()), "csl_peer_barrier")

    def check(self):
        """Raise if any device-side wait of this replica has timed out (sticky status word): the ensemble may
        have been advanced against stale complementary positions, so its results are void.  Synchronises."""
        if int(self.sync[0].item()) != 0:
            from . import _lib as L
            raise L.CslError("cosmoslik_b200: a peer did not reach a stretch-move half-step within "
                             "CSL_BARRIER_TIMEOUT_S; the exchange timed out and this run's results are void")
