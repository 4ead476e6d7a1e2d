"""Bandpower likelihoods: binned power spectra against one dense covariance.

``mspec_lnl`` restates the reference's multi-spectrum likelihood
(likelihoods/mspec_lnl.py:11-73) with the external ``mspec`` container replaced
by plain dicts; ``spt_lowl`` (spt_lowl.py) is its one-spectrum special case and
lives in spt_lowl.py of this package.  ``__call__`` returns a symbolic
BandpowerTerm; the arithmetic (foreground build, window binning on the FP64
tensor pipe, blocked triangular solve) runs in csrc/binning.cu and csrc/chi2.cu.
"""
from collections import OrderedDict, defaultdict

import numpy as np

from ..core import SlikPlugin, arguments
from ..symbolic import BandpowerTerm, SpectrumBlock, as_spectrum


def _spec_key(signal, a, b):
    return (a, b) if (a, b) in signal else (b, a)


def lincombo(signal, cov, new_maps, normalize=True):
    """Spectra and covariance of linear combinations of maps -- what ``mspec``'s ``PowerSpectra.lincombo`` does for
    the reference's ``cleaning`` option (mspec_lnl.py:38-39; the same call, with ``normalize=False``, applies the
    calibration at :44-45).  The external ``mspec`` package is absent; this restates its documented behaviour.

      new_maps  {new_map: {old_map: weight}};  with ``normalize`` the weights of a new map are divided by their sum
      signal    {(a, b): array[nbins]}  auto and cross spectra of the old maps, all with the same binning
      cov       covariance of the spectra concatenated in sorted key order, or None

    The spectrum of new maps A, B is  sum_ab w_Aa w_Bb C_ab  (C_ab = C_ba); returns (new signal dict over the sorted
    pairs of new maps, new covariance in that order)."""
    old_keys = sorted(signal)
    nb = len(np.atleast_1d(signal[old_keys[0]]))
    if any(len(np.atleast_1d(signal[k])) != nb for k in old_keys):
        raise ValueError("lincombo: all spectra must share one binning")
    weights = {}
    for A, w in new_maps.items():
        tot = float(sum(w.values())) if normalize else 1.0
        weights[A] = {a: float(v) / tot for a, v in w.items()}
    names = sorted(weights)
    new_keys = [(A, B) for i, A in enumerate(names) for B in names[i:]]
    T = np.zeros((len(new_keys), len(old_keys)))
    for r, (A, B) in enumerate(new_keys):
        for a, wa in weights[A].items():
            for b, wb in weights[B].items():
                k = _spec_key(signal, a, b)
                if k not in signal:
                    raise ValueError("lincombo: spectrum %s x %s is not in the signal" % (a, b))
                T[r, old_keys.index(k)] += wa * wb
    d = np.array([np.atleast_1d(signal[k]) for k in old_keys], dtype=float)          # [nold, nb]
    new_signal = OrderedDict((k, row) for k, row in zip(new_keys, T @ d))
    new_cov = None
    if cov is not None:
        big = np.kron(T, np.eye(nb))                                                   # acts on the concatenated vector
        new_cov = big @ np.asarray(cov, dtype=float) @ big.T
    return new_signal, new_cov


class mspec_lnl(SlikPlugin):
    """
    Args:
        use:      {(a, b): (lmin, lmax)}  cross-spectra and their ell ranges; the data
                  vector concatenates them in sorted key order (mspec_lnl.py:30, 68)
        windows:  {(a, b): array[nbins, lmax-lmin]}  the binning of spectrum (a, b)
        signal:   {(a, b): array[nbins]}  measured bandpowers
        cov:      array[n, n]  covariance of the concatenated data vector
        cal:      {map_name: value or param}  per-map calibration; spectrum (a, b) of the
                  DATA is scaled by cal[a]*cal[b] (mspec_lnl.py:44-45)
        egfs_kwargs: {(a, b): dict}  extra keywords for the foreground model (mspec_lnl.py:58)
        cleaning: {new_map: {old_map: weight}}  the data (and their covariance) are replaced by those of linear
                  combinations of the maps before anything else (mspec_lnl.py:38-39, :27-30); ``signal`` / ``cov``
                  then describe ALL spectra of the old maps (sorted key order), ``use`` / ``windows`` / ``cal`` the
                  new maps
        resgal:   {(a, b): array[nbins]}  residual Galactic power subtracted from the (cleaned) data (mspec_lnl.py:41-42)
    Map names start with the field letter: 'T150' x 'T220' reads cmb['cl_TT'] (mspec_lnl.py:58).
    """

    def __init__(self, use, windows, signal, cov, cal=None, egfs_kwargs=None, resgal=None, cleaning=None):
        super().__init__(**arguments())
        self.use = OrderedDict(sorted(use.items()))
        self.egfs_kwargs = egfs_kwargs if egfs_kwargs is not None else defaultdict(dict)
        # process_signal(docal=False) of the reference (:27-30, 35-42): constants, done once on the host
        sig = OrderedDict((k, np.asarray(v, dtype=float)) for k, v in sorted(signal.items()))
        cv = np.asarray(cov, dtype=float)
        if cleaning:
            sig, cv = lincombo(sig, cv, cleaning)
        if resgal:
            sig = OrderedDict((k, v - np.asarray(resgal[_spec_key(resgal, *k)], dtype=float)) for k, v in sig.items())
        keys = sorted(sig)
        if list(self.use) != keys:                       # `use` selects spectra: cut the vector and the covariance
            nb = [len(sig[k]) for k in keys]
            off = np.concatenate([[0], np.cumsum(nb)])
            pick = np.concatenate([np.arange(off[keys.index(_spec_key(sig, *k))], off[keys.index(_spec_key(sig, *k)) + 1])
                                   for k in self.use]).astype(int)
            cv = cv[np.ix_(pick, pick)]
        self.signal_processed = OrderedDict((k, sig[_spec_key(sig, *k)]) for k in self.use)
        self.cov_processed = cv

    def get_cl_model(self, cmb, egfs):
        pick = egfs if isinstance(egfs, dict) else None
        out = OrderedDict()
        for (a, b), (lmin, lmax) in self.use.items():
            e = pick[(a, b)] if pick is not None else egfs
            fg = e(spectra="cl_TT", lmax=lmax, spec=(a, b), **self.egfs_kwargs.get((a, b), {}))
            cl = as_spectrum(cmb["cl_%s%s" % (a[0], b[0])], lmax)[:lmax] + fg
            out[(a, b)] = cl[lmin:lmax]
        return out

    def __call__(self, cmb, egfs):
        model = self.get_cl_model(cmb, egfs)
        blocks = []
        for (a, b) in self.use:
            cal = 1.0
            if self.cal:
                cal = self.cal.get(a, 1) * self.cal.get(b, 1)
            blocks.append(SpectrumBlock(as_spectrum(model[(a, b)]), self.windows[(a, b)], self.signal_processed[(a, b)], cal))
        return BandpowerTerm(blocks, self.cov_processed)
