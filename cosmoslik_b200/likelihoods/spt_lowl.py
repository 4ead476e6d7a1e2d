"""The SPT "low-L" bandpower likelihood (reference: likelihoods/spt_lowl/spt_lowl.py).

Same constructor and call signature as the reference plugin.  The reference
bundles its data tarballs inside its source tree; this package does not ship
them: pass ``datadir`` (the directory holding ``spt_lps12_20120828.tar.gz`` /
``bandpowers_spt20082009.tar.gz`` and the three ``dl_*.txt`` templates, e.g. the
reference's ``likelihoods/spt_lowl/data``) or ``data`` (a mapping with the
already-parsed arrays, as stored in tests/golden/spt_lowl_data.npz).
"""
import os.path as osp
import tarfile

import numpy as np

from ..core import SlikPlugin, arguments, param
from ..models.egfs import egfs_specs
from ..symbolic import BandpowerTerm, SpectrumBlock, as_spectrum

_TARS = {"s12": "spt_lps12_20120828.tar.gz", "k11": "bandpowers_spt20082009.tar.gz"}


def read_newdat(datadir, which):
    """Parse the bandpower tarball: 47 bandpowers, the 47x47 covariance and the 47
    window functions over ell = 50..3300 (spt_lowl.py:65-79)."""
    nbin = 47
    with tarfile.open(osp.join(datadir, _TARS[which])) as tf:
        members = tf.getnames()
        newdat = [m for m in members if "newdat" in m][0]
        with tf.extractfile(newdat) as f:
            lines = f.read().decode().splitlines()
        start = next(i for i, ln in enumerate(lines) if "TT" in ln) + 1
        spec = np.array([float(lines[start + i].split()[1]) for i in range(nbin)])
        block = np.array([[float(v) for v in lines[start + nbin + i].split()] for i in range(2 * nbin)])
        sigma = block[nbin:]
        wins, lmin = [], None
        for i in range(1, nbin + 1):
            name = [m for m in members if m.endswith("window_%i" % i)][0]
            w = np.loadtxt(tf.extractfile(name))
            if lmin is None:
                lmin = int(w[:, 0].min())
            wins.append(w[:, 1])
    return dict(spec=spec, sigma=sigma, windows=np.array(wins), lmin=lmin)


class spt_lowl(SlikPlugin):
    """
    Args:
        which: 'k11' or 's12'
        lmin/lmax: restrict the data to this ell range
        drop: remove the bins at these indices
        egfs: callable foreground model ``egfs(lmax=..., egfs_specs=...)``; 'default' is
            :class:`spt_lowl_egfs`; None means it is supplied at call time
        cal: calibration multiplying the DATA power spectrum ('s12' only); defaults to a
            sampled parameter with the 2.6% prior (spt_lowl.py:50-51)
        datadir / data: where the bandpowers come from (see module docstring)
    """

    def __init__(self, which="s12", lmin=None, lmax=None, drop=None, egfs="default", cal=None,
                 datadir=None, data=None, **kwargs):
        super().__init__(**arguments(exclude=["data"]))
        if which not in _TARS:
            raise ValueError("spt_lowl: 'which' must be one of ['s12','k11']")
        if data is None:
            if datadir is None:
                raise ValueError("spt_lowl: pass datadir= (directory with the SPT tarballs) or data=")
            data = read_newdat(datadir, which)
            data = {which + "_" + k: v for k, v in data.items()}
        if self.egfs == "default":
            self.egfs = spt_lowl_egfs(datadir=datadir, data=data)
        if which == "s12":
            if cal is None:
                self.cal = param(start=1, scale=0.026, gaussian_prior=(1, 0.026))
        else:
            if not ((cal is None) or cal != 1.0):
                raise ValueError("Can't use a calibration parameter with the 'k11' likelihood, it already has "
                                 "the calibration marginalized into the covariance.")
            self.cal = 1

        self.spec = np.asarray(data[which + "_spec"], dtype=float)
        self.sigma = np.asarray(data[which + "_sigma"], dtype=float)
        windows = np.asarray(data[which + "_windows"], dtype=float)
        w0 = int(data[which + "_lmin"])
        self.windowrange = slice(w0, w0 + windows.shape[1])

        # trimming by ell range: count windows whose leading / trailing run of |w| < 1e-3 ends
        # before lmin / after lmax (spt_lowl.py:81-96)
        small = np.abs(windows) < 1e-3
        lead = np.where(small.all(1), small.shape[1], (~small).argmax(1))
        trail = np.where(small.all(1), small.shape[1], (~small[:, ::-1]).argmax(1))
        bmin = int(_count_while(lead < lmin)) if lmin is not None else 0
        bmax = int(_count_while((windows.shape[1] - trail) < lmax)) if lmax is not None else len(windows) + 1
        self.spec = self.spec[bmin:bmax]
        self.sigma = self.sigma[bmin:bmax, bmin:bmax]
        windows = windows[bmin:bmax]
        if drop is not None:
            self.spec = np.delete(self.spec, drop)
            self.sigma = np.delete(np.delete(self.sigma, drop, 0), drop, 1)
            windows = np.delete(windows, drop, 0)
        self.windows = windows
        self.lmax = self.windowrange.stop
        self.ells = windows @ np.arange(10000)[self.windowrange]
        self.egfs_specs = egfs_specs(kind="TT", freqs=({"dust": 154, "radio": 151, "tsz": 153},) * 2, fluxcut=50)

    def get_cl_model(self, cmb, egfs):
        """CMB + foreground D_ell over [0, lmax) as a symbolic spectrum (spt_lowl.py:171-179)."""
        if "TT" not in cmb:
            raise ValueError("Need the TT spectrum for spt_lowl.")
        tt = as_spectrum(cmb["TT"])
        if tt.size < self.lmax:
            raise ValueError("Need the TT spectrum to at least lmax=%i for spt_lowl." % self.lmax)
        return tt[:self.lmax] + as_spectrum(egfs(lmax=self.lmax, egfs_specs=self.egfs_specs), self.lmax)[:self.lmax]

    def __call__(self, cmb, egfs=None, cal=None):
        if egfs is not None:
            self.egfs = egfs
        if cal is not None:
            self.cal = cal
        self.cmb = cmb
        cl = self.get_cl_model(cmb, self.egfs)
        # r = cal*spec - W.cl[windowrange]; chi2/2 through the factor of sigma (spt_lowl.py:132-133)
        blk = SpectrumBlock(cl[self.windowrange], self.windows, self.spec, self.cal)
        return BandpowerTerm([blk], self.sigma)


def _count_while(mask):
    """number of leading True entries"""
    mask = np.asarray(mask, dtype=bool)
    return len(mask) if mask.all() else int(np.argmin(mask))


class spt_lowl_egfs(SlikPlugin):
    """Default SPT low-L foregrounds: (tSZ+kSZ), clustered CIB and Poisson templates with free
    amplitudes and the Story/Keisler priors (spt_lowl.py:186-207)."""

    def __init__(self,
                 Asz=None, Acl=None, Aps=None, datadir=None, data=None, **kwargs):
        super().__init__()
        self.Asz = param(start=5, scale=3, min=0, gaussian_prior=(5.5, 3.0)) if Asz is None else Asz
        self.Acl = param(start=20, scale=3, min=0, gaussian_prior=(19.3, 3.5)) if Acl is None else Acl
        self.Aps = param(start=5, scale=3, min=0, gaussian_prior=(5, 2.5)) if Aps is None else Aps
        if data is not None and "template_sz" in data:
            sz, cl, ps = data["template_sz"], data["template_cl"], data["template_ps"]
        elif datadir is not None:
            sz = np.loadtxt(osp.join(datadir, "dl_sz.txt"))[:, 1]
            cl = np.loadtxt(osp.join(datadir, "dl_clustered_point_source.txt"))[:, 1]
            ps = np.loadtxt(osp.join(datadir, "dl_poisson_point_source.txt"))[:, 1]
        else:
            raise ValueError("spt_lowl_egfs: pass datadir= or data= with the three templates")
        self.template_sz, self.template_cl, self.template_ps = (np.asarray(t, dtype=float) for t in (sz, cl, ps))

    def __call__(self, lmax, **_):
        # templates are sliced by ROW index, as in the reference (spt_lowl.py:206-207)
        return (self.Asz * self.template_sz[:lmax] + self.Acl * self.template_cl[:lmax]
                + self.Aps * self.template_ps[:lmax])
