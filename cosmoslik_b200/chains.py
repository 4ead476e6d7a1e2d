"""Chain containers and the chain-file format.

Mirrors the result side of the reference (cosmoslik/chains.py): ``Chain`` is a
dict name -> 1-D array with the special keys 'lnl' and 'weight'
(chains.py:33-63); ``Chains`` is a list of them (:345-372); a chain FILE is a
pickle stream -- a header of column names followed by ``(chain_id, rows)``
records, ``rows`` a structured array with fields f0..fN (lnl, weight, params)
(metropolis_hastings.py:181-186,211,221,237; emcee.py:40-46,82), which
``load_chain`` regroups by chain id (chains.py:930-962).  Files written by this
package load with the reference's ``load_chain`` and vice versa
(tests/test_chains_io.py).  ``load_cosmomc_chain`` reads the text formats
(chains.py:805-880) that ``Chain.savechain`` writes.  Plotting is out of scope.
"""
import os
import pickle
import re
from itertools import chain as _iterchain

import numpy as np

__all__ = ["Chain", "Chains", "get_covariance", "get_correlation", "combine_covs", "load_chain",
           "load_cosmomc_chain", "ChainWriter", "gelman_rubin", "device_summary", "device_add_gauss_prior", "device_thin",
           "device_reweighted"]


def _is_listlike(v):
    return not isinstance(v, str) and hasattr(v, "__iter__")


class _RowCall(object):
    """picklable `row -> func(**row)` for pool.map"""
    def __init__(self, func): self.func = func
    def __call__(self, row): return self.func(**row)


class _Reweight(object):
    def __init__(self, func): self.func = func
    def __call__(self, weight, **row): return {"weight": weight * self.func(weight=weight, **row)}


class Chain(dict):
    """One MCMC chain: parameter name -> array of values, plus 'lnl' and 'weight'."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k in list(self):
            self[k] = np.atleast_1d(self[k])
        if self and "weight" not in self:
            self["weight"] = np.ones(len(next(iter(self.values()))))

    def copy(self):
        return Chain({k: v.copy() for k, v in self.items()})

    def params(self, non_numeric=False, fixed=False):
        out = set()
        for k, v in self.items():
            if k.startswith("_") or k in ("lnl", "weight"):
                continue
            numeric = v.ndim == 1 and np.issubdtype(v.dtype, np.number)
            if not (non_numeric or numeric):
                continue
            if not fixed and np.all(v == v[0]):
                continue
            out.add(k)
        return out

    def sample(self, s, keys=None):
        return Chain((k, self[k][s]) for k in (keys if keys else list(self.keys())))

    def iterrows(self):
        for i in range(self.length()):
            yield {k: v[i] for k, v in self.items()}

    def matrix(self, params=None):
        if params is None:
            params = self.params()
        if _is_listlike(params):
            return np.vstack([self[p] for p in params]).T
        return self[params]

    # weighted statistics (chains.py:78-96)
    def cov(self, params=None):
        return get_covariance(self.matrix(params), self["weight"])

    def corr(self, params=None):
        return get_correlation(self.matrix(params), self["weight"])

    def mean(self, params=None):
        return np.average(self.matrix(params), axis=0, weights=self["weight"])

    def std(self, params=None):
        d = self.matrix(params) - self.mean(params)
        return np.sqrt(np.average(d ** 2, axis=0, weights=self["weight"]))

    def skew(self, params=None):
        d = self.matrix(params) - self.mean(params)
        return np.average(d ** 3, axis=0, weights=self["weight"]) / self.std(params) ** 3

    def confbound(self, param, level=0.95, bound=None):
        """Upper or lower confidence bound from the weighted 1000-bin histogram (chains.py:98-119);
        bound=None picks 'upper' for positively skewed parameters."""
        if bound is None:
            bound = "upper" if self.skew(param) > 0 else "lower"
        if bound == "lower":
            level = 1 - level
        H, x = np.histogram(self[param], weights=self["weight"], bins=1000)
        xc = (x[1:] + x[:-1]) / 2
        b = np.interp(level, np.cumsum(H) / float(H.sum()), xc)
        return (self[param].min(), b) if bound == "upper" else (b, self[param].max())

    def postprocd(self, func, nthreads=1, pool=None):
        """A new chain with the keys `func(**row)` returns added, row by row (chains.py:133-176); rows for which
        a key is missing get nan.  `pool` is anything with a `.map`; nthreads > 1 makes a process pool."""
        own = None
        if pool is None and nthreads != 1:
            import multiprocessing
            own = pool = multiprocessing.Pool(nthreads)
        try:
            mapper = map if pool is None else pool.map
            out = list(mapper(_RowCall(func), self.iterrows()))
        finally:
            if own is not None:
                own.terminate()
        c = self.copy()
        keys = []
        for d in out:
            keys += [k for k in d if k not in keys]
        c.update({k: np.array([d.get(k, np.nan) for d in out]) for k in keys})
        return c

    def reweighted(self, func, nthreads=1, pool=None):
        """A new chain whose weights are multiplied by `func(**row)` (chains.py:178-203)."""
        return self.postprocd(_Reweight(func), nthreads=nthreads, pool=pool)

    def acceptance(self):
        return 1.0 / np.mean(self["weight"])

    def length(self, unique=True):
        return len(self["weight"]) if unique else np.sum(self["weight"])

    def burnin(self, n):
        """Drop the leading rows whose cumulative weight stays below n."""
        first = int(np.searchsorted(np.cumsum(self["weight"]), n, side="left"))
        return self.sample(slice(first, None))

    def thin(self, delta):
        """Keep one row per `delta` units of weight (chains.py:235-243).  With S_k = ceil(cumulative weight
        through row k / delta), row k survives iff it opens a new slot (S_k > S_{k-1}) and its new weight is the
        number of slots it opens."""
        slots = np.ceil(np.cumsum(self["weight"]) / float(delta))
        opened = np.diff(slots, prepend=0.0)
        keep = np.flatnonzero(opened > 0)
        out = self.sample(keep)
        out["weight"] = opened[keep]
        return out

    def thinto(self, num):
        return self.thin(self.length(unique=False) // num)

    def add_gauss_prior(self, params, mean, covstd):
        """Importance-reweight the chain by a Gaussian prior on one parameter (mean, standard deviation) or on
        several (mean vector, covariance matrix): weight *= exp(-penalty), lnl += penalty (chains.py:206-233)."""
        x = self.matrix(params)
        if _is_listlike(params):
            resid = x - np.asarray(mean, dtype=float)
            penalty = np.einsum("ri,ri->r", resid @ np.linalg.inv(covstd), resid) / 2
        else:
            penalty = (x - mean) ** 2 / 2 / covstd ** 2
        out = self.copy()
        out["weight"] = out["weight"] * np.exp(-penalty)
        out["lnl"] = out["lnl"] + penalty
        return out

    def best_fit(self):
        i = int(np.argmin(self["lnl"]))
        return {k: v[i] for k, v in self.items()}

    def savecov(self, file, params=None):
        params = list(params) if params else sorted(self.params())
        with open(file, "wb") as f:
            f.write(("# " + " ".join(params) + "\n").encode())
            np.savetxt(f, self.cov(params))

    def savechain(self, file, params=None):
        keys = ["lnl", "weight"] + list(params if params else sorted(self.params()))
        with open(file, "w") as f:
            f.write("# " + " ".join(keys) + "\n")
            np.savetxt(f, self.matrix(keys))

    def join(self):
        return self

    def __repr__(self):
        return chain_stats(self)

    __str__ = __repr__


class Chains(list):
    """Several chains, e.g. the rows of one batched sampler run."""

    def burnin(self, n):
        return Chains(c.burnin(n) for c in self)

    def join(self):
        return Chain((k, np.concatenate([c[k] for c in self])) for k in self[0].keys())

    def params(self, **kw):
        return self[0].params(**kw)

    def gelman_rubin(self, params=None):
        params = sorted(self.params()) if params is None else list(params)
        return dict(zip(params, gelman_rubin(self, params)))

    def __repr__(self):
        return chain_stats(self)

    __str__ = __repr__


def chain_stats(c):
    lines = [object.__repr__(c)]
    params = c.params(fixed=True)
    width = max([12] + [len(p) for p in params]) + 4
    fmt = "{:>%d}:  {}" % width
    if isinstance(c, Chains):
        lines.append(fmt.format("# of chains", len(c)))
        c = c.join()
    lines.append(fmt.format("# of steps", c.length()))
    lines.append(fmt.format("total weight", "%.2f" % c["weight"].sum()))
    lines.append(fmt.format("acceptance", "%.3g" % c.acceptance()))
    for p in sorted(params):
        lines.append(fmt.format(p, "%10.4g ± %.4g" % (c.mean(p), c.std(p))))
    return "\n".join(lines)


def get_covariance(data, weights=None):
    """Weighted covariance with denominator sum(w) - 1 (chains.py:507-512)."""
    data = np.asarray(data)
    if weights is None:
        return np.cov(data.T)
    weights = np.asarray(weights)
    mu = np.sum(data.T * weights, axis=1) / np.sum(weights)
    z = data - mu
    return np.dot(z.T * weights, z) / (np.sum(weights) - 1)


def get_correlation(data, weights=None):
    cv = get_covariance(data, weights)
    sd = np.sqrt(np.diag(cv))
    return cv / np.outer(sd, sd)


def combine_covs(*covs):
    """Merge covariances given as (names, array), a file ('# names' header then
    the matrix), a Chain, or {name: std}; later entries override the rows and
    columns of the names they repeat (chains.py:891-927)."""
    parsed = []
    for c in covs:
        if isinstance(c, tuple):
            parsed.append((list(c[0]), np.asarray(c[1], dtype=float)))
        elif isinstance(c, str):
            with open(c) as f:
                names = re.sub("#", "", f.readline()).split()
                parsed.append((names, np.atleast_2d(np.loadtxt(f))))
        elif isinstance(c, Chain):
            names = sorted(c.params())
            parsed.append((names, c.cov(names)))
        elif isinstance(c, dict):
            parsed.append((list(c), np.diag([v ** 2 for v in c.values()])))
        else:
            raise ValueError("Unrecognized covariance data type.")
    allnames = list(_iterchain(*[n for n, _ in parsed]))
    out = np.zeros((len(allnames), len(allnames)))
    for names, cov in parsed:
        idx = [allnames.index(n) for n in names]
        for i in idx:
            out[i, :] = 0
            out[:, i] = 0
        out[np.ix_(idx, idx)] = cov
    return allnames, out


def gelman_rubin(chains, params):
    """Potential scale reduction per parameter from several chains (Gelman &
    Rubin 1992) with weighted per-chain means / variances (denominator
    sum(w) - 1).  Host statement of what ``csl_chain_moments`` + one all-reduce
    compute on the device."""
    n, mu, var = [], [], []
    for c in chains:
        w = c["weight"]
        x = c.matrix(params)
        sw = w.sum()
        m = (x * w[:, None]).sum(0) / sw
        n.append(sw); mu.append(m); var.append((w[:, None] * (x - m) ** 2).sum(0) / (sw - 1))
    return rhat_from_moments(np.array(n), np.array(mu), np.array(var))


def rhat_from_moments(n, mu, var):
    m = mu.shape[0]
    nbar = n.mean()
    W = var.mean(0)
    B_over_n = ((mu - mu.mean(0)) ** 2).sum(0) / (m - 1)
    return np.sqrt(((nbar - 1) / nbar * W + B_over_n) / W)


def device_add_gauss_prior(stats, params, mean, covstd, names):
    """``Chain.add_gauss_prior`` (chains.py:206-233) applied to EVERY chain of a finished sampler run, in
    place on the device row buffer: weight *= exp(-dlnl), lnl += dlnl with
    dlnl = (x - mean)^T cov^-1 (x - mean) / 2 over `params` (a name with its standard deviation, or a list
    of names with their covariance).  `names` = the sampled parameter names in column order."""
    import torch
    from . import _lib as L
    rows, nrows = stats["rows"], stats["nrows"]
    W, cap, rl = rows.shape
    names = list(names)
    if _is_listlike(params):
        idx = [names.index(p) for p in params]
        cinv = np.linalg.inv(np.atleast_2d(np.asarray(covstd, dtype=float)))
        mu = np.asarray(mean, dtype=float).ravel()
    else:
        idx = [names.index(params)]
        cinv = np.array([[1.0 / float(covstd) ** 2]])
        mu = np.array([float(mean)])
    if cinv.shape != (len(idx), len(idx)) or mu.shape != (len(idx),):
        raise ValueError("add_gauss_prior: mean / covariance do not match the parameter list")
    dev = rows.device
    d_idx = torch.tensor(idx, dtype=torch.int32, device=dev)
    d_mu = torch.from_numpy(mu).to(dev)
    d_ci = torch.from_numpy(np.ascontiguousarray(cinv)).to(dev)
    L.check(L.lib().csl_rows_gauss_prior(W, rl - 2, rows.data_ptr(), nrows.data_ptr(), cap, len(idx), d_idx.data_ptr(),
                                         d_mu.data_ptr(), d_ci.data_ptr(), L.stream_ptr()), "csl_rows_gauss_prior")
    return stats


def device_thin(stats, delta):
    """``Chain.thin(delta)`` (chains.py:235-243) applied to EVERY chain of a finished sampler run, in place on the
    device row buffer (``csl_rows_thin``: one warp per chain walks its rows, keeps the rows that open a new slot of
    `delta` units of cumulative weight and gives them the number of slots they open); ``nrows`` is updated."""
    from . import _lib as L
    rows, nrows = stats["rows"], stats["nrows"]
    W, cap, rl = rows.shape
    L.check(L.lib().csl_rows_thin(W, rl - 2, rows.data_ptr(), nrows.data_ptr(), cap, float(delta), L.stream_ptr()),
            "csl_rows_thin")
    return stats


def device_reweighted(stats, names, func):
    """``Chain.reweighted(func)`` (chains.py:178-203) on the device: `func(**params)` is called ONCE with symbolic
    parameters (like a script's ``__call__``) and must return a scalar expression of them; it is compiled to the
    coefficient bytecode and evaluated for every stored row of every chain, whose weight it multiplies.
    `names` = the sampled parameter names in column order."""
    import torch
    from . import symbolic as sym
    from .program import ExtrasProgram
    rows, nrows = stats["rows"], stats["nrows"]
    W, cap, rl = rows.shape
    names = list(names)
    expr = func(**{n: sym.Par(i, n) for i, n in enumerate(names)})
    prog = ExtrasProgram(names, [expr], device=rows.device)
    live = torch.arange(cap, device=rows.device)[None, :] < nrows.clamp(max=cap)[:, None]
    x = rows[:, :, 2:][live].contiguous()
    if x.shape[0]:
        factor = prog.values(x)[:, 0]
        w = rows[:, :, 1]
        w[live] = w[live] * factor
    return stats


def device_summary(stats, burn_rows=0):
    """Weighted mean, standard deviation and covariance of ALL chains / walkers of a finished sampler run,
    reduced on the device from its row buffer (`stats` = what ``sampler.run`` returned) -- the numbers
    ``Chains.join().mean()`` / ``.std()`` / ``.cov()`` give (chains.py:78-96, 345-355; covariance
    denominator sum(w) - 1, std with denominator sum(w) like numpy.average), without copying the rows to the
    host.  With several GPUs the sums are all-reduced."""
    import torch
    from . import _lib as L
    from . import dist
    rows, nrows = stats["rows"], stats["nrows"]
    W, cap, rl = rows.shape
    nd = rl - 2
    m = 1 + nd + nd * nd
    part = torch.empty((592, m), dtype=torch.float64, device=rows.device)
    out = torch.zeros(m, dtype=torch.float64, device=rows.device)
    lib, st = L.lib(), L.stream_ptr()
    L.check(lib.csl_rows_moments(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, int(burn_rows), None,
                                 part.data_ptr(), out.data_ptr(), st), "csl_rows_moments")
    dist.all_reduce_sum(out[:1 + nd])
    mean = (out[1:1 + nd] / out[0]).contiguous()
    L.check(lib.csl_rows_moments(W, nd, rows.data_ptr(), nrows.data_ptr(), cap, int(burn_rows), mean.data_ptr(),
                                 part.data_ptr(), out.data_ptr(), st), "csl_rows_moments")
    dist.all_reduce_sum(out[1 + nd:])
    h = out.cpu().numpy()
    c2 = h[1 + nd:].reshape(nd, nd)
    return dict(weight=h[0], mean=h[1:1 + nd] / h[0], cov=c2 / (h[0] - 1), std=np.sqrt(np.diag(c2) / h[0]))


# --------------------------------------------------------------------------
# chain files
# --------------------------------------------------------------------------

class ChainWriter(object):
    """Append-only pickle-stream writer, flushed per block so a running chain can
    be read (doc/quickstart.md:53).

    header_kind 'ndarray' writes the numpy array of names the MH sampler writes
    (metropolis_hastings.py:183-186), 'list' the plain list emcee writes
    (emcee.py:40-44)."""

    def __init__(self, path, names, header_kind="ndarray", dtypes=None):
        """dtypes: numpy dtype per column (default: float64 everywhere, the reference's ``dtype(float).name``);
        ``output_extra_params`` entries ``(name, dtype)`` put other scalar dtypes in their columns
        (metropolis_hastings.py:46-56, :221)."""
        self.f = open(path, "wb")
        self.ncol = len(names)
        kinds = ["f8"] * self.ncol if dtypes is None else [np.dtype(d).str for d in dtypes]
        self.dtype = np.dtype(",".join(kinds))
        self.plain = all(np.dtype(k) == np.dtype("f8") for k in kinds)
        header = np.hstack([names]) if header_kind == "ndarray" else list(names)
        pickle.dump(header, self.f, protocol=pickle.HIGHEST_PROTOCOL)

    def write(self, chain_id, rows):
        """rows: float array [m, ncol] (lnl, weight, params...)"""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self.ncol)
        if self.plain:
            rec = rows.view(self.dtype).reshape(-1)
        else:                                    # columns of other scalar dtypes: cast field by field
            rec = np.empty(len(rows), dtype=self.dtype)
            for k, name in enumerate(self.dtype.names):
                rec[name] = rows[:, k].astype(self.dtype[name])
        pickle.dump((chain_id, rec), self.f, protocol=pickle.HIGHEST_PROTOCOL)

    def flush(self):
        self.f.flush()

    def close(self):
        if not self.f.closed:
            self.f.close()


def load_chain(filename, repack=False):
    """Load a chain file written by this package or by the reference's samplers (chains.py:930-962);
    also accepts a file holding one pickled Chain / Chains.  repack=True rewrites a block-structured file
    as one pickled Chains object (faster to load next time) unless another process still has it open for
    writing (chains.py:964-976)."""
    with open(filename, "rb") as f:
        head = pickle.load(f)
        if isinstance(head, (Chain, Chains)):
            return head
        names = [n.decode() if isinstance(n, bytes) else str(n) for n in head]
        records = []
        while True:
            try:
                records.append(pickle.load(f, encoding="latin1"))
            except Exception:
                break
    ids = []
    for i, _ in records:
        if i not in ids:
            ids.append(i)
    out = Chains()
    for i in sorted(ids):
        blocks = [d for j, d in records if j == i]
        if blocks[0].dtype.kind == "V":
            out.append(Chain({n: np.concatenate([b["f%d" % k] for b in blocks]) for k, n in enumerate(names)}))
        else:
            out.append(Chain(dict(zip(names, np.vstack(blocks).T))))
    if repack and not _open_for_writing(filename):
        tmp = filename + ".repack"
        with open(tmp, "wb") as f:
            pickle.dump(out, f)
        os.replace(tmp, filename)
    return out


def _open_for_writing(filename):
    """True if some process holds `filename` open for writing (the chain is still running)."""
    target = os.path.realpath(filename)
    try:
        pids = [p for p in os.listdir("/proc") if p.isdigit()]
    except OSError:
        return True                       # cannot tell: do not touch the file
    for pid in pids:
        fddir = "/proc/%s/fd" % pid
        try:
            fds = os.listdir(fddir)
        except OSError:
            continue
        for fd in fds:
            try:
                if os.path.realpath(os.path.join(fddir, fd)) != target:
                    continue
                with open("/proc/%s/fdinfo/%s" % (pid, fd)) as g:
                    flags = [ln for ln in g if ln.startswith("flags:")]
                if flags and int(flags[0].split()[1], 8) & 3:      # O_WRONLY or O_RDWR
                    return True
            except (OSError, ValueError):
                continue
    return False


def load_cosmomc_chain(path, paramnames=None):
    """Text chains (chains.py:805-880): a file whose first line is `# name name ...` (what `Chain.savechain`
    writes), a CosmoMC file with a `<root>.paramnames` beside it (columns weight, lnl, parameters...), a
    directory with one file per parameter (WMAP style; last column is the value), or a prefix with
    `<path>_1, <path>_2, ...` -> Chains."""

    def names_from(pn):
        with open(pn) as f:
            return ["weight", "lnl"] + [line.split()[0] for line in f if line.strip()]

    def one(fn, names):
        if os.path.isdir(fn):
            if names is not None:
                raise ValueError("Can't specify custom parameter names if loading chain from a directory.")
            cols = {}
            for k in sorted(os.listdir(fn)):
                try:
                    cols[k] = np.loadtxt(os.path.join(fn, k), usecols=[-1])
                except Exception:
                    pass
            return Chain(cols)
        if names is None:
            d, base = os.path.dirname(fn), os.path.basename(fn)
            pn = [os.path.join(d, f) for f in os.listdir(d)
                  if f.endswith(".paramnames") and base.startswith(f[:-len(".paramnames")])]
            if len(pn) > 1:
                raise ValueError("Found multiple paramnames files for this chain; %s" % pn)
            if pn:
                names = names_from(pn[0])
        elif isinstance(names, str):
            names = names_from(names)
        with open(fn) as f:
            if names is None:
                line = f.readline()
                if not line.startswith("#"):
                    raise ValueError("Couldn't find any paramnames. Specify paramnames=... by hand.")
                names = line.replace("#", "").split()
            try:
                data = np.loadtxt(f, ndmin=2).T
            except Exception:
                return None
        return Chain(zip(names, data))

    path = os.path.abspath(path)
    d, base = os.path.dirname(path), os.path.basename(path)
    files = sorted(os.path.join(d, f) for f in os.listdir(d)
                   if f == base or re.fullmatch(re.escape(base) + r"_[0-9]+(\.txt)?", f))
    if len(files) == 1:
        return one(files[0], paramnames)
    if len(files) > 1:
        return Chains([c for c in (one(f, paramnames) for f in files) if c is not None])
    raise IOError("File not found: " + path)
