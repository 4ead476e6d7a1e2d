"""Generate tests/golden/mspec.npz by running the synthetic multi-spectrum scripts through the UNMODIFIED reference:
its ``Slik.evaluate`` (cosmoslik.py:109-141), ``param`` / priors, the foreground base class ``models/egfs.py`` and
``likelihoods/mspec_lnl.py`` -- whose container package ``mspec`` is absent and is replaced by the stand-in of
``oracle/mspec_standin.py`` (read its header for what that does and does not pin).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is mounted:
    python -m oracle.make_golden_mspec

The scripts are the ones ``cosmoslik_b200.synthetic.build_script`` writes for the product and for the oracle; here
the plugin namespace holds the reference's own classes.  Two adaptations, both outside the reference's files:
  * ``mspec_lnl(use, windows, signal, cov, ...)`` of the scripts is turned into the reference's
    ``mspec_lnl(signal=PowerSpectra(...), use=..., ...)``;
  * the reference keys the calibration dict by ``'_'.join(map)`` (mspec_lnl.py:45 -- its maps are tuples such as
    ('T', '150')); the scripts key it by the map name ('T150'), so the keys are rewritten the same way before each call.
Cases: BASELINE configs[2]-like (6 cross-spectra of 3 frequencies, per-map calibration) and configs[3]-like (TT / TE /
EE, calibration on the data), small sizes; plus the ``cleaning`` (linear combinations of maps), ``resgal`` and
``use``-subset options on the first.  Stored: the points and the reference's lnl.
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import mspec_standin, refshim  # noqa: E402
from oracle.make_golden import OUT  # noqa: E402
from cosmoslik_b200 import synthetic  # noqa: E402  (problem definitions only: numpy, no device code)

NPTS = 24


def reference_namespace(extra=None):
    R = refshim.load()                    # puts the reference on the path
    mspec_standin.install()
    mod = importlib.import_module("cosmoslik_plugins.likelihoods.mspec_lnl")
    mspec_standin.install(mod)
    egfs_mod = importlib.import_module("cosmoslik_plugins.models.egfs")
    extra = extra or {}

    class adapter(mod.mspec_lnl):
        def __init__(self, use, windows, signal, cov, cal=None, **kw):
            sig = mspec_standin.PowerSpectra(signal, cov, mspec_standin.Binning(windows, dict(use)))
            kw = dict(extra, **kw)
            super().__init__(signal=sig, use=dict(use), cal=cal, **kw)

        def __call__(self, cmb, egfs):
            if self.cal:
                self.cal = {"_".join(k): v for k, v in self.cal.items()}
            return super().__call__(cmb, egfs)

    return R, types.SimpleNamespace(SlikPlugin=R.core.SlikPlugin, param=R.core.param, mspec_lnl=adapter, egfs=egfs_mod.egfs)


def points(prob, n, seed):
    names = sorted(prob["truth"])
    x0 = np.array([prob["truth"][k] for k in names]); sc = np.array([prob["scales"][k] for k in names])
    X = x0 + 1.0 * sc * np.random.RandomState(seed).standard_normal((n, len(names)))
    hard = next(i for i, k in enumerate(names) if k.startswith("A") and k != "A_cmb")
    X[n // 2, hard] = -1.0                                   # one point outside a hard prior (A_* >= 0): lnl = inf
    return names, X


def cleaning_case(prob):
    """the 3-frequency problem re-expressed for ``cleaning``: the DATA are the 6 spectra of the three maps; two
    cleaned maps are formed from them (weights summing to 1 after normalisation) and the likelihood uses the 3
    spectra of the cleaned maps; ``resgal`` subtracts a smooth residual from each"""
    maps = sorted({m for k in prob["use"] for m in k})
    assert len(maps) == 3
    new = {maps[1]: {maps[1]: 1.0, maps[0]: -0.2}, maps[2]: {maps[2]: 3.0, maps[1]: 1.0}}    # cleaned maps keep the names
    keys = [(maps[1], maps[1]), (maps[1], maps[2]), (maps[2], maps[2])]
    anyk = sorted(prob["use"])[0]
    use = {k: prob["use"][anyk] for k in keys}
    windows = {k: prob["windows"][anyk] for k in keys}
    nb = prob["windows"][anyk].shape[0]
    resgal = {k: 0.01 * (i + 1) * np.linspace(1.0, 2.0, nb) for i, k in enumerate(keys)}
    return new, use, windows, resgal


def main():
    out = {}
    # ---- configs[2]-like: 3 frequencies, 6 cross-spectra, per-map calibration
    prob = synthetic.spt_multifreq_problem(3, nbins=10, lmax=601)
    R, ns = reference_namespace()
    slik = R.core.Slik(synthetic.build_script(ns, prob)())
    names, X = points(prob, NPTS, 11)
    assert list(slik.get_sampled()) == names
    out["spt_x"] = X
    out["spt_lnl"] = np.array([slik.evaluate(*x)[0] for x in X])
    print("spt   lnl[:3] = %r  (%d finite of %d)" % (out["spt_lnl"][:3], np.isfinite(out["spt_lnl"]).sum(), NPTS))
    # ---- the same data, `use` restricted to three of the six spectra
    sub = dict(prob)
    keep = sorted(prob["use"])[::2]
    sub["use"] = {k: prob["use"][k] for k in keep}
    sub["windows"] = {k: prob["windows"][k] for k in keep}
    R, ns = reference_namespace()
    slik = R.core.Slik(synthetic.build_script(ns, sub)())
    out["spt_subset_keys"] = np.array(["%s %s" % k for k in keep])
    out["spt_subset_lnl"] = np.array([slik.evaluate(*x)[0] for x in X])
    print("subset lnl[1] = %r" % out["spt_subset_lnl"][1])
    # ---- cleaning + resgal: two cleaned maps out of the three
    new, use, windows, resgal = cleaning_case(prob)
    cl = dict(prob)
    cl["use"], cl["windows"] = use, windows
    R, ns = reference_namespace(dict(cleaning=new, resgal=resgal))
    script = synthetic.build_script(ns, cl)
    slik = R.core.Slik(script())
    out["clean_lnl"] = np.array([slik.evaluate(*x)[0] for x in X])
    out["clean_resgal"] = np.array([resgal[k] for k in sorted(resgal)])
    print("clean lnl[1] = %r  (%d finite)" % (out["clean_lnl"][1], np.isfinite(out["clean_lnl"]).sum()))
    # ---- configs[3]-like: TT / TE / EE, calibration on the data
    pl = synthetic.planck_lite_problem(4)
    R, ns = reference_namespace()
    slik = R.core.Slik(synthetic.build_script(ns, pl)())
    namesp, Xp = points(pl, 8, 12)
    assert list(slik.get_sampled()) == namesp
    out["planck_x"] = Xp
    out["planck_lnl"] = np.array([slik.evaluate(*x)[0] for x in Xp])
    print("planck lnl = %r" % (out["planck_lnl"],))
    np.savez_compressed(os.path.join(OUT, "mspec.npz"), **out)


if __name__ == "__main__":
    main()
