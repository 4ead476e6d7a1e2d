"""Reference-style scalar scripts built on the CPU oracle (oracle/), used as the
checker in parity tests.  Mirrors the scripts oracle/make_golden.py runs on the
unmodified reference."""
import numpy as np

from oracle import likes, runtime


def synthetic_tt(n=6000):
    ell = np.arange(n)
    return 5000.0 * np.exp(-(ell / 1200.0) ** 1.2) + 50.0


def spt_script(data, which="s12", tt=None):
    """The spt_lowl script of make_golden.py on oracle plugins."""
    tt = synthetic_tt() if tt is None else tt

    class script(runtime.Plugin):
        def __init__(self):
            super().__init__()
            egfs = likes.TemplateEgfs(data["template_sz"], data["template_cl"], data["template_ps"])
            cal = runtime.Param(start=1, scale=0.026, gaussian_prior=(1, 0.026)) if which == "s12" else 1
            self.spt = likes.SptLowl(data[which + "_spec"], data[which + "_sigma"], data[which + "_windows"],
                                     int(data[which + "_lmin"]), egfs, cal=cal)
            self.cmb = {"TT": tt}

        def __call__(self):
            return self.spt(self.cmb)

    return script


def quick_script():
    class quick(runtime.Plugin):  # doc/quickstart.md:26-43
        def __init__(self):
            super().__init__()
            self.a = runtime.Param(start=0, scale=1)
            self.b = runtime.Param(start=0, scale=1)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    return quick


def derived_script():
    """the script of oracle/make_golden_extras.py (derived quantities stored in __call__)"""
    class derived(runtime.Plugin):
        def __init__(self):
            super().__init__()
            self.a = runtime.Param(start=0, scale=1)
            self.b = runtime.Param(start=0.5, scale=1, min=-4)

        def __call__(self):
            self.r2 = self.a ** 2 + self.b ** 2
            self.ratio = (self.a - 1.5) / (2 + self.b ** 2)
            self.fixed = 3.25
            return self.r2 / 2

    return derived


def gauss_script(cov):
    ndim = cov.shape[0]
    cinv = np.linalg.inv(cov)

    class gauss(runtime.Plugin):
        def __init__(self):
            super().__init__()
            for i in range(ndim):
                self["p%02d" % i] = runtime.Param(start=0, scale=1)

        def __call__(self):
            v = np.array([self["p%02d" % i] for i in range(ndim)])
            return v @ cinv @ v / 2

    return gauss
