"""CPU oracle, part 1: the scalar plugin runtime of the reference, restated.

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, never by the
product package ``cosmoslik_b200``.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` runs this file
side by side with the unmodified reference (via ``oracle/refshim.py``) when
``/root/reference`` is mounted, and ``tests/test_oracle_golden.py`` checks it
against the committed outputs of the reference in ``tests/golden/`` everywhere
else.

Every function names the reference lines it follows (paths relative to the
reference root).  Plain numpy / scipy, scalar per-walker calls, like the
reference itself.
"""
import re
from collections import OrderedDict

import numpy as np


class Param(object):
    """A sampled-parameter marker (cosmoslik/cosmoslik.py:232-238): a bag of
    keyword attributes such as start, scale, min, max, gaussian_prior."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def param_shortcut(*names):
    """cosmoslik/cosmoslik.py:428-445: positional values are zipped onto names."""

    def make(*vals, **kw):
        kw.update(dict(zip(names, vals)))
        return Param(**kw)

    return make


class Tree(object):
    """Attribute tree with dotted-key access, the role SlikDict plays in the
    reference (cosmoslik/cosmoslik.py:177-229).  Stored in a private dict
    rather than by aliasing ``__dict__`` to the mapping itself."""

    def __init__(self, **kw):
        object.__setattr__(self, "_d", OrderedDict())
        for k, v in kw.items():
            self._d[k] = v

    # attribute protocol
    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_d")[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self._d[k] = v

    def __delattr__(self, k):
        del self._d[k]

    # dotted-key protocol (cosmoslik.py:183-216)
    def __getitem__(self, key):
        node = self
        for part in key.split("."):
            node = getattr(node, part)
        return node

    def __setitem__(self, key, v):
        parts = key.split(".")
        node = self
        for part in parts[:-1]:
            node = getattr(node, part)
        setattr(node, parts[-1], v)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except AttributeError:
            return False

    def items(self):
        return self._d.items()

    def clone(self):
        """cosmoslik.py:190-197: nested trees are copied, leaves are shared."""
        new = object.__new__(type(self))
        object.__setattr__(new, "_d", OrderedDict())
        for k, v in self._d.items():
            new._d[k] = v.clone() if isinstance(v, Tree) else v
        return new

    def find_sampled(self):
        """cosmoslik.py:218-229: every Param leaf, keyed by dotted path, in
        lexicographic order of the dotted names."""
        found = {}

        def walk(node, prefix):
            for k, v in node._d.items():
                if isinstance(v, Tree):
                    walk(v, prefix + [k])
                elif isinstance(v, Param):
                    found[".".join(prefix + [k])] = v

        walk(self, [])
        return OrderedDict(sorted(found.items()))


class Plugin(Tree):
    """SlikPlugin (cosmoslik.py:300-313): a Tree that is callable."""

    def __call__(self, *a, **kw):
        raise NotImplementedError


class Priors(Plugin):
    """cosmoslik_plugins/likelihoods/priors.py:6-30.

    Collects, per sampled parameter and in this order: gaussian_prior,
    uniform_prior, range, (min,max) together, then min alone as (min,+inf) and
    max alone as (-inf,max) -- so a parameter with both ``min`` and ``max``
    contributes three boxes (priors.py:13-16)."""

    def __init__(self, params):
        super().__init__()
        g, u = [], []
        for name, p in params.find_sampled().items():
            if hasattr(p, "gaussian_prior"):
                g.append((name,) + tuple(p.gaussian_prior))
            if hasattr(p, "uniform_prior"):
                u.append((name,) + tuple(p.uniform_prior))
            if hasattr(p, "range"):
                u.append((name,) + tuple(p.range))
            if hasattr(p, "min") and hasattr(p, "max"):
                u.append((name, p.min, p.max))
            if hasattr(p, "min"):
                u.append((name, p.min, np.inf))
            if hasattr(p, "max"):
                u.append((name, -np.inf, p.max))
        self.gaussian_priors = g
        self.uniform_priors = u

    def __call__(self, params):
        # priors.py:24-30: unset parameters (still Param objects) are skipped
        def bound(n):
            return n in params and not isinstance(params[n], Param)

        for n, lo, hi in self.uniform_priors:
            if bound(n) and not (lo <= params[n] <= hi):
                return np.inf
        return sum((params[n] - c) ** 2.0 / 2 / w ** 2 for n, c, w in self.gaussian_priors if bound(n))


class Slik(object):
    """cosmoslik/cosmoslik.py:80-145."""

    def __init__(self, script):
        self.params = script
        self._sampled = script.find_sampled()
        self.sampler = getattr(script, "sampler", None)
        if "sampler" in script:
            del script.sampler
        if "priors" not in script:
            script.priors = Priors(script)

    def get_sampled(self):
        return self._sampled

    def get_start(self):
        return {k: v.start for k, v in self._sampled.items()}

    def evaluate(self, *x, **kw):
        """cosmoslik.py:109-141: bind x into a copy of the tree; hard prior hit
        returns (inf, params) without calling the script; otherwise
        script() + priors(params) (priors are evaluated a second time)."""
        params = self.params.clone()
        if x and kw:
            raise ValueError("Expected either *args or **kwargs but not both.")
        if not x:
            missing = [k for k in self._sampled if k not in kw]
            if missing:
                raise ValueError("Missing the following parameters: %s" % missing)
            for k, v in kw.items():
                params[k] = v
        elif len(x) != len(self._sampled):
            raise ValueError("Expected %i parameters, only got %i." % (len(self._sampled), len(x)))
        else:
            for k, v in zip(self._sampled, x):
                params[k] = v
        if np.isinf(params.priors(params)):
            return np.inf, params
        return params() + params.priors(params), params


def lsum(*fs):
    """cosmoslik.py:385-401: running sum that stops at the first +inf."""
    s = 0
    for f in fs:
        s += f()
        if s == np.inf:
            break
    return s


# --------------------------------------------------------------------------
# covariance helpers
# --------------------------------------------------------------------------

def get_covariance(data, weights=None):
    """cosmoslik/chains.py:507-512: weighted covariance with denominator
    sum(w) - 1 (numpy.cov when unweighted)."""
    data = np.asarray(data)
    if weights is None:
        return np.cov(data.T)
    mean = np.sum(data.T * weights, axis=1) / np.sum(weights)
    z = data - mean
    return np.dot(z.T * weights, z) / (np.sum(weights) - 1)


def combine_covs(*covs):
    """cosmoslik/chains.py:891-927.  Inputs: (names, array) tuples, a file
    whose first line is '# n1 n2 ...', or {name: std}.  Later entries wipe
    the rows/columns of names they redefine."""
    norm = []
    for c in covs:
        if isinstance(c, tuple):
            norm.append((list(c[0]), np.asarray(c[1], dtype=float)))
        elif isinstance(c, str):
            with open(c) as f:
                names = re.sub("#", "", f.readline()).split()
                norm.append((names, np.loadtxt(f)))
        elif isinstance(c, dict):
            norm.append((list(c), np.diag([v ** 2 for v in c.values()])))
        else:
            raise ValueError("Unrecognized covariance data type.")
    allnames = [n for names, _ in norm for n in names]
    out = np.zeros((len(allnames), len(allnames)))
    for names, cov in norm:
        idx = [allnames.index(n) for n in names]
        for i in idx:
            out[i, :] = 0
            out[:, i] = 0
        out[np.ix_(idx, idx)] = cov
    return allnames, out


def initialize_covariance(sampled, covs=None):
    """cosmoslik_plugins/samplers/utils.py:4-23."""
    ins = [{k: v.scale for k, v in sampled.items() if hasattr(v, "scale")}]
    if covs is not None:
        ins += covs if isinstance(covs, list) else [covs]
    names, cov = combine_covs(*ins)
    missing = [s for s in sampled if s not in names]
    if missing:
        raise ValueError("Parameters %s not in covariance and no scale given." % missing)
    idx = [names.index(s) for s in sampled]
    return cov[np.ix_(idx, idx)]


def proposal_factor(cov_est, ndim, proposal_scale):
    """The matrix F with ``test_x = cur_x + z @ F`` that
    ``numpy.random.multivariate_normal(cur_x, cov_est/ndim*scale**2)`` uses
    (metropolis_hastings.py:141-142).  numpy's legacy generator factors the
    covariance by SVD, ``(u, s, v) = svd(cov)``, and returns
    ``mean + standard_normal(ndim) @ (sqrt(s)[:, None] * v)``; checked bitwise
    against numpy in tests/test_oracle_golden.py."""
    cov = np.asarray(cov_est, dtype=float) / ndim * proposal_scale ** 2
    _, s, v = np.linalg.svd(cov)
    return np.sqrt(s)[:, None] * v
