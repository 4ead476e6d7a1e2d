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
    Map names start with the field letter: 'T150' x 'T220' reads cmb['cl_TT'] (mspec_lnl.py:58).
    """

    def __init__(self, use, windows, signal, cov, cal=None, egfs_kwargs=None):
        super().__init__(**arguments())
        self.use = OrderedDict(sorted(use.items()))
        self.egfs_kwargs = egfs_kwargs if egfs_kwargs is not None else defaultdict(dict)

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
            blocks.append(SpectrumBlock(as_spectrum(model[(a, b)]), self.windows[(a, b)], self.signal[(a, b)], cal))
        return BandpowerTerm(blocks, np.asarray(self.cov, dtype=float))
