"""Generate tests/golden/sptsz.npz by executing the UNMODIFIED reference code of the model-dependent-covariance
likelihood: ``SPTSZ_lowl.__call__``, ``get_cl_model_and_beam_cov`` and ``DlnClDlnl``
(likelihoods/SPTSZ_lowl_2017/SPTSZ_lowl.py:62-89) and its foreground plugin ``SPTSZ_lowl_egfs`` (:147-170).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is mounted:
    python -m oracle.make_golden_sptsz

What cannot run is the reference's ``__init__`` (:13-57): it reads a tar member as bytes and compares it with a
str (:31), opens the window tarball by a path relative to the current directory (:35-36) and needs
``data/SPTSZ_beam_errors.tar``, which is missing from the repository (.MISSING_LARGE_BLOBS:1).  So the instance
is created without ``__init__`` and given exactly the attributes ``__init__`` would set:
  spec, sigma, windows, windowrange, lmax   read HERE from the reference's own data tarballs (the parsing
                                            below restates :27-38 for Python 3)
  beam_corr                                 synthetic: sum of outer products of smooth beam-error modes
                                            (what spt_beam_cor, :95-145, builds from the missing grids)
  egfs                                      the reference's SPTSZ_lowl_egfs() (runs as shipped)
  cal, ab_on                                as passed
Every number stored comes out of the reference's own methods, with aberration off and on.
"""
import importlib
import os
import sys
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402
from oracle.make_golden import OUT, synthetic_tt  # noqa: E402


def beam_modes(lmin, nl, nmode=6, scale=1.0):
    """smooth fractional beam-error modes dCl/Cl (the same shapes tests/sptsz_problem.py uses)"""
    ell = np.arange(lmin, lmin + nl, dtype=float)
    shapes = [0.010 * (ell / 3000.0) ** 2, 0.020 * np.exp(-ell / 2000.0), 0.005 * np.sin(ell / 500.0) + 0.002,
              0.003 * np.ones(nl), 0.004 * (ell / 3000.0), 0.006 * np.cos(ell / 900.0) * (ell / 3000.0),
              0.002 * (ell / 3000.0) ** 3, 0.003 * np.exp(-((ell - 1500) / 600.0) ** 2)]
    return scale * np.array(shapes[:nmode]).reshape(nmode, nl)


def read_reference_data(mod):
    root = os.path.join(os.path.dirname(mod.__file__), "data", "SPTSZ_bandpowers_and_errors", "spt_lps12_20120828")
    with tarfile.open(os.path.join(root, "Spectrum_spt2500deg2_lps12_alternativeCalibrationImplementation_no_beam_cov.newdat.tar")) as tf:
        name = next(f for f in tf.getnames() if "newdat" in f)
        lines = tf.extractfile(name).read().decode().splitlines()
    at = next(i for i, ln in enumerate(lines) if "TT" in ln) + 1
    spec = np.array([np.array(lines[at + i].split(), dtype=float)[1] for i in range(47)])
    sigma = np.array([np.array(lines[at + 47 + i].split(), dtype=float) for i in range(47)])
    with tarfile.open(os.path.join(root, "windows", "window_lps12.tar")) as tf:
        wins = np.array([np.loadtxt(tf.extractfile("window_lps12/window_%i" % i))[:, 1] for i in range(1, 48)])
        ell = np.loadtxt(tf.extractfile("window_lps12/window_1"))[:, 0]
    return spec, sigma, wins, slice(int(ell.min()), int(ell.max() + 1))


def main():
    R = refshim.load()
    mod = importlib.import_module("cosmoslik_plugins.likelihoods.SPTSZ_lowl_2017.SPTSZ_lowl")
    spec, sigma, windows, wr = read_reference_data(mod)
    assert sigma.shape == (47, 47) and windows.shape[0] == 47 and windows.shape[1] == wr.stop - wr.start
    tt = synthetic_tt()
    egfs = mod.SPTSZ_lowl_egfs()
    rng = np.random.RandomState(20260303)
    out = dict(spec=spec, sigma=sigma, windows=windows, lmin_window=np.array(wr.start), tt=tt,
               template_sz=egfs.template_sz[:wr.stop + 64], template_cl=egfs.template_cl[:wr.stop + 64],
               template_ps=egfs.template_ps[:wr.stop + 64])
    for nmode, scale in ((6, 1.0), (8, 5.0)):
        modes = beam_modes(wr.start, wr.stop - wr.start, nmode, scale)
        for ab_on in (False, True):
            like = mod.SPTSZ_lowl.__new__(mod.SPTSZ_lowl)
            R.core.SlikDict.__init__(like)
            like.spec, like.sigma, like.windows = spec.copy(), sigma.copy(), windows.copy()
            like.windowrange, like.lmax = wr, wr.stop
            like.beam_corr = modes.T @ modes
            like.ab_on = ab_on
            X = np.array([1.0, 5.0, 20.0, 5.0]) + rng.standard_normal((48, 4)) * np.array([0.01, 2.0, 3.0, 2.0])
            X[:, 1:] = np.abs(X[:, 1:])
            vals = []
            for cal, asz, acl, aps in X:
                like.cal = cal
                like.egfs = mod.SPTSZ_lowl_egfs(Asz=asz, Acl=acl, Aps=aps)     # reset: __call__ overwrites self.egfs (:74)
                vals.append(like({"TT": tt}))
            tag = "m%d_ab%d" % (nmode, int(ab_on))
            out[tag + "_x"] = X                     # columns: cal, Asz, Acl, Aps
            out[tag + "_lnl"] = np.array(vals)
            out[tag + "_modes"] = modes
            print(tag, "lnl[0] = %r" % vals[0])
    np.savez_compressed(os.path.join(OUT, "sptsz.npz"), **out)


if __name__ == "__main__":
    main()
