"""SPT-SZ 2017-style problem (model-dependent covariance) on the oracle and on the product plugins:
the real SPT 's12' bandpowers / windows (tests/golden/spt_lowl_data.npz) with synthetic beam-error
modes (the reference's beam grids are missing from its repository)."""
import numpy as np

from conftest import golden
from cosmoslik_b200 import SlikPlugin, likelihoods, param
from oracle import likes as olikes, runtime as ort


def beam_modes(lmin, nl, nmode=6, scale=1.0):
    ell = np.arange(lmin, lmin + nl, dtype=float)
    shapes = [0.010 * (ell / 3000.0) ** 2, 0.020 * np.exp(-ell / 2000.0), 0.005 * np.sin(ell / 500.0) + 0.002,
              0.003 * np.ones(nl), 0.004 * (ell / 3000.0), 0.006 * np.cos(ell / 900.0) * (ell / 3000.0),
              0.002 * (ell / 3000.0) ** 3, 0.003 * np.exp(-((ell - 1500) / 600.0) ** 2)]
    return scale * np.array(shapes[:nmode]).reshape(nmode, nl)


def scripts(nmode=6, scale=1.0, sampler=None, ab_on=False):
    g = golden("spt_lowl_data.npz")
    tt = golden("spt_lowl_lnl.npz")["tt"]
    nl, lmin = g["s12_windows"].shape[1], int(g["s12_lmin"])
    modes = beam_modes(lmin, nl, nmode, scale)
    B = modes.T @ modes

    class oscript(ort.Plugin):
        def __init__(self):
            super().__init__()
            self.cal = ort.Param(start=1, scale=.01, gaussian_prior=(1, .026))
            self.egfs = olikes.TemplateEgfs(g["template_sz"], g["template_cl"], g["template_ps"])
            self.like = olikes.SptszLowl(g["s12_spec"], g["s12_sigma"], g["s12_windows"], lmin, B, None, ab_on=ab_on)

        def __call__(self):
            self.like.egfs, self.like.cal = self.egfs, self.cal
            return self.like({"TT": tt})

    class bscript(SlikPlugin):
        def __init__(self, by_modes=False):
            super().__init__()
            self.cal = param(start=1, scale=.01, gaussian_prior=(1, .026))
            self.egfs = likelihoods.spt_lowl_egfs(data=g)
            kw = dict(beam_modes=modes) if (by_modes or len(modes) == 0) else dict(beam_corr=B)
            self.like = likelihoods.SPTSZ_lowl(g["s12_spec"], g["s12_sigma"], g["s12_windows"], lmin, None, ab_on=ab_on, **kw)
            self.sampler = sampler(self) if sampler else None

        def __call__(self):
            self.like.egfs, self.like.cal = self.egfs, self.cal
            return self.like({"TT": tt})

    return oscript, bscript


def reference_golden_script(tag):
    """product script on the arrays of tests/golden/sptsz.npz (the reference's own data files and the values its own
    __call__ returned, oracle/make_golden_sptsz.py); sampled parameters: cal, egfs.Acl, egfs.Aps, egfs.Asz"""
    g = golden("sptsz.npz")
    modes = g[tag + "_modes"]

    class script(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.cal = param(start=1, scale=.01)
            self.egfs = likelihoods.spt_lowl_egfs(data=dict(template_sz=g["template_sz"], template_cl=g["template_cl"],
                                                            template_ps=g["template_ps"]))
            self.like = likelihoods.SPTSZ_lowl(g["spec"], g["sigma"], g["windows"], int(g["lmin_window"]), None,
                                               beam_modes=modes, ab_on=tag.endswith("ab1"))
            self.sampler = None

        def __call__(self):
            self.like.egfs, self.like.cal = self.egfs, self.cal
            return self.like({"TT": g["tt"]})

    X = g[tag + "_x"][:, [0, 2, 3, 1]]   # columns cal, Asz, Acl, Aps -> sorted names cal, egfs.Acl, egfs.Aps, egfs.Asz
    # Slik.evaluate adds the priors the default foreground parameters carry (spt_lowl.py:190-192); the golden values
    # are the likelihood plugin's own return value
    prior = ((X[:, 1] - 19.3) ** 2 / 2 / 3.5 ** 2 + (X[:, 2] - 5.0) ** 2 / 2 / 2.5 ** 2 + (X[:, 3] - 5.5) ** 2 / 2 / 3.0 ** 2)
    return script, X, g[tag + "_lnl"] + prior
