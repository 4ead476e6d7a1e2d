"""SPT-SZ 2017-style low-L likelihood: bandpowers whose covariance depends on the model through the
beam uncertainty (reference: likelihoods/SPTSZ_lowl_2017/SPTSZ_lowl.py:11-81).

    dl       = cmb['TT'][windowrange] + egfs[windowrange]
    db       = W . dl
    beam_cov = W (beam_corr o dl dl^T) W^T                       (:81)
    lnl      = 1/2 (spec - cal*db)^T (beam_cov + sigma)^-1 (spec - cal*db)      (:66-69, no log-det)

The reference builds ``beam_corr`` as a sum of outer products of beam-error modes (:95-145), so
beam_cov = V V^T with V_k = W (b_k o dl): the modes become extra window rows of the binning kernel and
the per-walker factorisation runs in ``csl_cov_chi2``.  Pass either ``beam_modes`` [K, n_ell] or the
matrix ``beam_corr`` [n_ell, n_ell] (the reference's form; it is factored once into its non-zero eigen
modes).  The reference plugin cannot run as shipped (bytes/str at :31, cwd-relative window path :35-36,
``SPTSZ_beam_errors.tar`` missing from the repository) -- the arrays are therefore arguments here.
``ab_on`` applies the aberration correction of :78-79, 84-89 to the model before it is binned:
dl *= 1 + 0.26 * 1.23e-3 * dln(dl)/dln(l), evaluated per walker in the binning kernel's build stage.
The arithmetic is pinned by tests/golden/sptsz.npz, produced by the reference's own methods.
"""
import numpy as np

from ..core import SlikPlugin
from ..symbolic import LowRankCovBandpowerTerm, as_spectrum


def modes_from_beam_corr(beam_corr, rtol=1e-13):
    """b_k with beam_corr = sum_k b_k b_k^T (eigen-decomposition, eigenvalues below rtol*max dropped)"""
    lam, U = np.linalg.eigh(np.asarray(beam_corr, dtype=float))
    keep = lam > rtol * lam.max()
    return (U[:, keep] * np.sqrt(lam[keep])).T


class SPTSZ_lowl(SlikPlugin):
    """
    Args:
        spec [n], sigma [n, n], windows [n, n_ell], lmin_window: the bandpowers, their (beam-free)
            covariance and window functions over ell = lmin_window .. lmin_window + n_ell - 1
        beam_modes [K, n_ell] or beam_corr [n_ell, n_ell]: fractional beam-error modes dCl/Cl
        egfs: foreground model, called as ``egfs(lmax=...)`` (SPTSZ_lowl.py:74)
        cal: calibration multiplying the MODEL bandpowers (:68)
    """

    def __init__(self, spec, sigma, windows, lmin_window, egfs, beam_modes=None, beam_corr=None, cal=1,
                 ab_on=False, **kwargs):
        super().__init__()
        self.ab_on = bool(ab_on)
        if (beam_modes is None) == (beam_corr is None):
            raise ValueError("pass exactly one of beam_modes / beam_corr")
        self.spec = np.asarray(spec, dtype=float)
        self.sigma = np.asarray(sigma, dtype=float)
        self.windows = np.asarray(windows, dtype=float)
        self.windowrange = slice(int(lmin_window), int(lmin_window) + self.windows.shape[1])
        self.lmax = self.windowrange.stop
        self.beam_modes = np.asarray(beam_modes, dtype=float) if beam_modes is not None \
            else modes_from_beam_corr(beam_corr)
        self.egfs = egfs
        self.cal = cal

    def __call__(self, cmb):
        tt = as_spectrum(cmb["TT"])
        if tt.size < self.lmax:
            raise ValueError("Need the TT spectrum to at least lmax=%i." % self.lmax)
        dl = tt[:self.lmax] + as_spectrum(self.egfs(lmax=self.lmax), self.lmax)[:self.lmax]
        return LowRankCovBandpowerTerm(dl[self.windowrange], self.windows, self.spec, self.cal, self.sigma,
                                       self.beam_modes, aberration=(0.26 * 1.23e-3 if self.ab_on else None),
                                       ell0=self.windowrange.start)
