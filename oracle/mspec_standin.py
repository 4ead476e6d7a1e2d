"""A stand-in for the third-party package ``mspec`` (github marius311/mspec; not in the reference tree, not in its
requirements.txt, not installable offline) -- just enough of its ``PowerSpectra`` container for the reference's
``likelihoods/mspec_lnl.py`` to RUN its own code here.

TEST INFRASTRUCTURE ONLY (used by ``oracle/make_golden_mspec.py`` and ``tests/test_oracle_vs_reference.py``).
Nothing in the product imports it.

What this buys: the arithmetic that IS in the reference -- ``process_signal`` (cleaning, then resgal, then
calibration as a non-normalised ``lincombo`` of the maps, mspec_lnl.py:35-47), ``get_cl_model`` (:50-60: CMB +
foregrounds per cross-spectrum, sliced at lmax) and ``__call__`` (:64-73: bin, stack in ``use`` order, ``model -
signal``, upper Cholesky factor, ``cho_solve``) -- is executed from the reference's own file.  What it cannot buy: the
container's semantics are stated HERE, from the way mspec_lnl.py uses them, not checked against the real package:

  PowerSpectra(spectra, cov=None, binning=None)   spectra {(a, b): array}; cov = covariance of the spectra concatenated
                                                  in sorted key order (one block per pair of spectra)
  .get_maps()                                     sorted map names
  .lincombo(new_maps, normalize=True, docov=True) spectra (and covariance) of linear combinations of the maps
  a - b                                           spectrum-wise difference (resgal), covariance kept
  .binned(binning)                                {k: W_k @ cl_k[lmin_k:lmax_k]} with the binning's windows / ranges
  .get_as_matrix(lrange=use)                      namedtuple(spec, cov): the spectra named in ``use`` stacked in
                                                  sorted key order, and the matching covariance blocks
  load_signal(x)                                  returns x

The file is stale relative to its own package, so three NAMES have to be supplied before it imports and runs (none of
them touches arithmetic; the reference's files are not edited):
  * ``numpy.product`` (:1, imported and never used; removed in numpy 2) -> ``numpy.prod``;
  * ``cosmoslik.all_kw`` (:3, :20; the reference's core no longer defines it) -> the caller's ``locals()`` minus self;
  * ``M`` -- ``get_cl_model`` (:56) refers to a module-level ``M`` that the file only binds as a LOCAL of ``__init__``
    (``import mspec as M``, :22): a NameError as shipped, with or without the real package -> bound in the module.
"""
import collections
import sys
import types

import numpy as np

Matrix = collections.namedtuple("Matrix", "spec cov")


def _key(spectra, a, b):
    return (a, b) if (a, b) in spectra else (b, a)


class Binning(object):
    """windows {(a, b): array[nbins, lmax - lmin]} and ranges {(a, b): (lmin, lmax)}"""

    def __init__(self, windows, use):
        self.windows, self.use = windows, use


class PowerSpectra(object):

    def __init__(self, spectra=None, cov=None, binning=None):
        self.spectra = collections.OrderedDict(sorted((k, np.asarray(v, dtype=float)) for k, v in (spectra or {}).items()))
        self.cov = None if cov is None else np.asarray(cov, dtype=float)
        self.binning = binning

    def __getitem__(self, k):
        return self.spectra[_key(self.spectra, *k)]

    def get_spectra(self):
        return list(self.spectra)

    def get_maps(self):
        return sorted({m for k in self.spectra for m in k})

    def _offsets(self):
        off, at = {}, 0
        for k, v in self.spectra.items():
            off[k] = (at, at + len(v)); at += len(v)
        return off

    def lincombo(self, new_maps, normalize=True, docov=True):
        weights = {}
        for A, w in new_maps.items():
            tot = float(sum(w.values())) if normalize else 1.0
            weights[A] = {a: float(v) / tot for a, v in w.items()}
        names = sorted(weights)
        new = [(A, B) for i, A in enumerate(names) for B in names[i:]]
        terms = {}                             # new spectrum -> [(old spectrum, weight)]
        for (A, B) in new:
            acc = collections.OrderedDict()
            for a, wa in weights[A].items():
                for b, wb in weights[B].items():
                    k = _key(self.spectra, a, b)
                    if k in self.spectra:      # a cross-spectrum that was never measured contributes nothing
                        acc[k] = acc.get(k, 0.0) + wa * wb
            if acc:
                terms[(A, B)] = list(acc.items())
        spectra = {k: sum(w * self.spectra[o] for o, w in t) for k, t in terms.items()}
        cov = None
        if docov and self.cov is not None:
            old_off = self._offsets()
            keys = sorted(spectra)
            n_new = sum(len(spectra[k]) for k in keys)
            J = np.zeros((n_new, self.cov.shape[0]))
            at = 0
            for k in keys:
                nb = len(spectra[k])
                for o, w in terms[k]:
                    lo, hi = old_off[o]
                    J[at:at + nb, lo:hi] += w * np.eye(nb)
                at += nb
            cov = J @ self.cov @ J.T
        return PowerSpectra(spectra, cov, self.binning)

    def __sub__(self, other):
        get = (lambda k: other[k]) if isinstance(other, PowerSpectra) else (lambda k: other[_key(other, *k)])
        return PowerSpectra({k: v - np.asarray(get(k), dtype=float) for k, v in self.spectra.items()}, self.cov, self.binning)

    def binned(self, binning):
        out = {}
        for k, cl in self.spectra.items():
            kk = _key(binning.windows, *k)
            lmin, lmax = binning.use[_key(binning.use, *k)]
            out[k] = np.dot(binning.windows[kk], cl[lmin:lmax])
        return PowerSpectra(out, None, binning)

    def get_as_matrix(self, lrange=None):
        keys = sorted(_key(self.spectra, *k) for k in (lrange or self.spectra))
        spec = np.concatenate([self.spectra[k] for k in keys])
        cov = None
        if self.cov is not None:
            off = self._offsets()
            pick = np.concatenate([np.arange(*off[k]) for k in keys])
            cov = self.cov[np.ix_(pick, pick)]
        return Matrix(spec, cov)


def load_signal(x):
    return x


def install(reference_module=None):
    """register the stand-in as ``mspec`` / ``mspec.utils``; bind ``M`` in the reference module (see above).
    Call it BEFORE importing the reference module (with no argument) and again with the module."""
    me = sys.modules[__name__]
    if not hasattr(np, "product"):         # mspec_lnl.py:1 imports numpy.product (unused there), gone from numpy 2
        np.product = np.prod
    import cosmoslik                       # the reference package (oracle/refshim.py has put it on the path)
    if not hasattr(cosmoslik, "all_kw"):   # mspec_lnl.py:3, :20: a helper the reference's own core no longer has;
        def all_kw(ns):                    # all it ever did: the caller's locals() minus self, **kwargs merged in
            out = {k: v for k, v in ns.items() if k not in ("self", "__class__", "kwargs")}
            out.update(ns.get("kwargs") or {})
            return out
        cosmoslik.all_kw = all_kw
    pkg = types.ModuleType("mspec")
    pkg.PowerSpectra, pkg.load_signal = PowerSpectra, load_signal
    utils = types.ModuleType("mspec.utils")
    utils.pairs = lambda xs: [(a, b) for i, a in enumerate(xs) for b in xs[i:]]
    pkg.utils = utils
    sys.modules.setdefault("mspec", pkg)
    sys.modules.setdefault("mspec.utils", utils)
    if reference_module is not None:
        reference_module.M = sys.modules["mspec"]
    return me
