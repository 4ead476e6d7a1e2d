"""CPU oracle, part 2: foreground models and bandpower likelihoods, restated.

TEST INFRASTRUCTURE ONLY (see oracle/runtime.py for the rules).

Parity status:
  * ``SptLowl`` / ``TemplateEgfs`` / ``ClustPoissonEgfs``: PINNED against the
    unmodified reference (bundled SPT data, tests/golden/spt_lowl_*.npz).
  * ``MspecLike``: the reference plugin (likelihoods/mspec_lnl.py) depends on
    the third-party package ``mspec`` (github marius311/mspec, unpinned, not in
    requirements.txt), which is absent.  The plugin's OWN code is pinned: it runs
    here over a stand-in of that container (oracle/mspec_standin.py) and its
    outputs are tests/golden/mspec.npz (oracle/make_golden_mspec.py; live in
    tests/test_oracle_vs_reference.py).  PARITY UNPINNED only for the container's
    semantics (stacking order, ``lincombo``, ``binned``), which the stand-in
    states from the way mspec_lnl.py uses them; further anchored by the
    1-spectrum case reproducing ``SptLowl``
    (tests/test_oracle_golden.py::test_mspec_one_spectrum_is_spt).
  * ``GaussLike``: the reference has no Gaussian plugin; this is the quickstart
    script body (doc/quickstart.md:41-43) generalised to x^T C^-1 x / 2.
"""
from collections import OrderedDict

import numpy as np
from scipy.linalg import cho_factor, cho_solve, cholesky

from .runtime import Param, Plugin


class TemplateEgfs(Plugin):
    """spt_lowl.py:186-207: three fixed D_ell templates times amplitudes.
    Templates are sliced ``[:lmax]`` by ROW INDEX, whatever ell the file's
    first row holds (spt_lowl.py:201-207)."""

    def __init__(self, template_sz, template_cl, template_ps, Asz=None, Acl=None, Aps=None):
        super().__init__()
        self.Asz = Param(start=5, scale=3, min=0, gaussian_prior=(5.5, 3.0)) if Asz is None else Asz
        self.Acl = Param(start=20, scale=3, min=0, gaussian_prior=(19.3, 3.5)) if Acl is None else Acl
        self.Aps = Param(start=5, scale=3, min=0, gaussian_prior=(5, 2.5)) if Aps is None else Aps
        self.template_sz = np.asarray(template_sz)
        self.template_cl = np.asarray(template_cl)
        self.template_ps = np.asarray(template_ps)

    def __call__(self, lmax, **_):
        return (self.Asz * self.template_sz[:lmax] + self.Acl * self.template_cl[:lmax]
                + self.Aps * self.template_ps[:lmax])


class Egfs(Plugin):
    """models/egfs.py:61-67: calling the plugin with the sampled parameters
    returns a closure; calling that with the data set's keywords sums the
    component dict from ``get_egfs`` (python ``sum`` starting at 0, so an
    empty dict gives the integer 0)."""

    def get_egfs(self, **kw):
        raise NotImplementedError

    def __call__(self, **params):
        def get(**kw):
            kw.update(params)
            return sum(self.get_egfs(**kw).values())

        return get


class ClustPoissonEgfs(Egfs):
    """models/clust_poisson_egfs.py:9-17.  ell runs over arange(lmax), i.e.
    includes ell = 0 (0**ncib)."""

    def get_egfs(self, Aps, Acib, ncib, spectra, lmax, **_):
        if spectra != "cl_TT":
            return {}
        ell = np.arange(lmax) / 3000.0
        return {"ps": Aps * ell ** 2, "cib": Acib * ell ** ncib}


class SptLowl(Plugin):
    """spt_lowl.py:112-133 and 171-182, built from arrays instead of the
    tarball (the loader is spt_lowl.py:65-108; ``make_golden.py`` runs it and
    stores its outputs)."""

    def __init__(self, spec, sigma, windows, lmin_window, egfs, cal=None):
        super().__init__()
        self.spec = np.asarray(spec)
        self.sigma = np.asarray(sigma)
        self.windows = [np.asarray(w) for w in windows]
        nl = len(self.windows[0])
        self.windowrange = slice(int(lmin_window), int(lmin_window) + nl)
        self.lmax = self.windowrange.stop
        self.cho_sigma = cho_factor(self.sigma)  # spt_lowl.py:98
        self.egfs = egfs
        self.cal = cal

    def get_cl_model(self, cmb, egfs):
        if "TT" not in cmb:
            raise ValueError("Need the TT spectrum for spt_lowl.")
        if cmb["TT"].size < self.lmax:
            raise ValueError("Need the TT spectrum to at least lmax=%i for spt_lowl." % self.lmax)
        cl = cmb["TT"][: self.lmax] + egfs(lmax=self.lmax)[: self.lmax]
        return np.array([np.dot(cl[self.windowrange], w) for w in self.windows])

    def __call__(self, cmb, egfs=None, cal=None):
        if egfs is not None:
            self.egfs = egfs
        if cal is not None:
            self.cal = cal
        cl = self.get_cl_model(cmb, self.egfs)
        dcl = self.cal * self.spec - cl  # spt_lowl.py:132: calibration multiplies the DATA
        return np.dot(dcl, cho_solve(self.cho_sigma, dcl)) / 2


class SptszLowl(Plugin):
    """likelihoods/SPTSZ_lowl_2017/SPTSZ_lowl.py:62-89 from arrays (the loader :25-57 and the beam-grid reader
    :95-145 cannot run: py2 bytes/str at :31, cwd-relative path :35-36, SPTSZ_beam_errors.tar missing).
    PINNED by tests/golden/sptsz.npz: the reference's own ``__call__`` / ``get_cl_model_and_beam_cov`` /
    ``DlnClDlnl`` executed on an instance that was handed these arrays (oracle/make_golden_sptsz.py).
      dl (:74-77); aberration dl *= 1 + .26 * 1.23e-3 * dlnCl/dlnl (:78-79, 84-89: forward differences of ln dl
      over ln l, 0 for the first ell); db = W dl and beam_cov = W (beam_corr o dl dl^T) W^T (:81);
      cho_factor(beam_cov + sigma) on every call (:67); deldb = spec - db*cal (:68); no log-det (:69)."""

    def __init__(self, spec, sigma, windows, lmin_window, beam_corr, egfs, cal=1, ab_on=False):
        super().__init__()
        self.spec, self.sigma = np.asarray(spec), np.asarray(sigma)
        self.windows = np.asarray(windows)
        self.windowrange = slice(int(lmin_window), int(lmin_window) + self.windows.shape[1])
        self.lmax = self.windowrange.stop
        self.beam_corr = np.asarray(beam_corr)
        self.egfs, self.cal, self.ab_on = egfs, cal, ab_on

    def dlncl_dlnl(self, y):
        lnx = np.log(np.arange(self.lmax)[self.windowrange])
        lny = np.log(y)
        return np.concatenate([[0.0], (lny[1:] - lny[:-1]) / (lnx[1:] - lnx[:-1])])

    def __call__(self, cmb):
        eg = self.egfs(lmax=self.lmax)[:self.lmax]
        tt = cmb["TT"][:self.lmax]
        dl = tt[self.windowrange] + eg[self.windowrange]
        if self.ab_on:
            dl = dl * (1 + .26 * 1.23e-3 * self.dlncl_dlnl(dl))
        db = np.dot(self.windows, dl)
        beam_cov = np.dot(self.windows, np.dot(self.beam_corr * np.outer(dl, dl), self.windows.T))
        cho = cho_factor(beam_cov + self.sigma)
        deldb = self.spec - db * self.cal
        return np.dot(deldb, cho_solve(cho, deldb)) / 2.0


class MspecLike(Plugin):
    """likelihoods/mspec_lnl.py:11-73 with the external ``mspec`` container
    replaced by plain dicts:

      use      {(a, b): (lmin, lmax)}       per cross-spectrum ell range
      windows  {(a, b): array[nbins, lmax-lmin]}  ``.binned()`` := W @ cl[lmin:lmax]
      signal   {(a, b): array[nbins]}       measured bandpowers
      cov      array[n, n]                  spectra concatenated in sorted key order
                                            (``get_as_matrix().spec``)
      cal      {map_name: value}            per-MAP calibration; a cross spectrum
                                            is scaled by cal[a]*cal[b]
                                            (mspec_lnl.py:44-45, lincombo of maps)

    Model per spectrum (mspec_lnl.py:58):
      cmb['cl_%s%s' % (a[0], b[0])][:lmax] + egfs_ab(spectra='cl_TT', lmax=lmax, spec=(a, b), **kw)
    Residual is model - signal (mspec_lnl.py:72); chi^2 through the UPPER
    Cholesky factor, ``cholesky(cov)`` then ``cho_solve((U, False), d)``
    (mspec_lnl.py:31, 73)."""

    def __init__(self, use, windows, signal, cov, cal=None, egfs_kwargs=None, resgal=None, cleaning=None):
        super().__init__()
        self.use = OrderedDict(sorted(use.items()))
        self.windows = windows
        self.cal = cal
        self.egfs_kwargs = egfs_kwargs or {}
        # mspec_lnl.py:27-30, 35-42: cleaning (lincombo of maps, data AND covariance), then resgal, then `use`
        sig = {k: np.asarray(v, dtype=float) for k, v in signal.items()}
        cov = np.asarray(cov, dtype=float)
        if cleaning:
            sig, cov = lincombo_maps(sig, cov, cleaning)
        if resgal:
            sig = {k: v - np.asarray(resgal[k if k in resgal else k[::-1]], dtype=float) for k, v in sig.items()}
        keys = sorted(sig)
        rows, at = {}, 0
        for k in keys:
            rows[k] = np.arange(at, at + len(sig[k])); at += len(sig[k])
        pick = np.concatenate([rows[k if k in rows else k[::-1]] for k in self.use])
        self.signal = {k: sig[k if k in sig else k[::-1]] for k in self.use}
        self.cov = cov[np.ix_(pick, pick)]
        self.cho_cov = (cholesky(self.cov), False)

    def process_signal(self):
        out = []
        for (a, b) in self.use:
            s = np.asarray(self.signal[(a, b)], dtype=float)
            if self.cal:
                s = s * self.cal.get(a, 1) * self.cal.get(b, 1)
            out.append(s)
        return np.concatenate(out)

    def get_cl_model(self, cmb, egfs):
        pick = egfs if isinstance(egfs, dict) else None
        out = []
        for (a, b), (lmin, lmax) in self.use.items():
            e = pick[(a, b)] if pick is not None else egfs
            cl = cmb["cl_%s%s" % (a[0], b[0])][:lmax] + e(spectra="cl_TT", lmax=lmax, spec=(a, b),
                                                         **self.egfs_kwargs.get((a, b), {}))
            out.append(np.dot(self.windows[(a, b)], cl[lmin:lmax]))
        return np.concatenate(out)

    def __call__(self, cmb, egfs):
        dcl = self.get_cl_model(cmb, egfs) - self.process_signal()
        return np.dot(dcl, cho_solve(self.cho_cov, dcl)) / 2


def lincombo_maps(signal, cov, new_maps, normalize=True):
    """``mspec.PowerSpectra.lincombo`` restated for dicts (the package is absent; PARITY UNPINNED): maps are replaced
    by linear combinations ``new = sum_old w old`` (weights normalised to sum 1 unless told otherwise, as the
    reference's calibration call at mspec_lnl.py:44-45 turns off); every spectrum of two new maps is the bilinear
    form of the old spectra, written here as explicit loops over the bins, and the covariance follows by the same
    linear map.  Deliberately a different formulation from the product's (which builds one matrix with kron)."""
    old = sorted(signal)
    nb = len(signal[old[0]])
    w = {A: {a: v / (sum(m.values()) if normalize else 1.0) for a, v in m.items()} for A, m in new_maps.items()}
    names = sorted(w)
    new = [(A, B) for i, A in enumerate(names) for B in names[i:]]
    pos = {k: i for i, k in enumerate(old)}
    J = np.zeros((len(new) * nb, len(old) * nb))
    out = {}
    for r, (A, B) in enumerate(new):
        acc = np.zeros(nb)
        for a, wa in w[A].items():
            for b, wb in w[B].items():
                k = (a, b) if (a, b) in signal else (b, a)
                acc = acc + wa * wb * np.asarray(signal[k], dtype=float)
                for i in range(nb):
                    J[r * nb + i, pos[k] * nb + i] += wa * wb
        out[(A, B)] = acc
    return out, J @ np.asarray(cov, dtype=float) @ J.T


class GaussLike(Plugin):
    """(x-mu)^T C^-1 (x-mu) / 2 -- doc/quickstart.md:41-43 with a dense
    covariance, evaluated the way the reference evaluates every Gaussian
    chi^2 (cho_factor once, cho_solve per call; spt_lowl.py:98,133)."""

    def __init__(self, mean, cov):
        super().__init__()
        self.mean = np.asarray(mean, dtype=float)
        self.cho = cho_factor(np.asarray(cov, dtype=float))

    def __call__(self, x):
        d = np.asarray(x, dtype=float) - self.mean
        return np.dot(d, cho_solve(self.cho, d)) / 2
