"""The same scripts as tests/oracle_scripts.py, written against the product package
(cosmoslik_b200) exactly as a CosmoSlik user would write them."""
import numpy as np

from cosmoslik_b200 import SlikPlugin, param, likelihoods


def synthetic_tt(n=6000):
    ell = np.arange(n)
    return 5000.0 * np.exp(-(ell / 1200.0) ** 1.2) + 50.0


def spt_script(data, which="s12", tt=None, sampler=None):
    tt = synthetic_tt() if tt is None else tt

    class script(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.spt = likelihoods.spt_lowl(which, data=data)
            self.cmb = {"TT": tt}
            self.sampler = sampler(self) if sampler else None

        def __call__(self):
            return self.spt(self.cmb)

    return script


def quick_script(sampler=None):
    class quick(SlikPlugin):   # doc/quickstart.md:26-43
        def __init__(self):
            super().__init__()
            self.a = param(start=0, scale=1)
            self.b = param(start=0, scale=1)
            self.sampler = sampler(self) if sampler else None

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    return quick


def gauss_script(cov, sampler=None, start=0.0):
    ndim = cov.shape[0]

    class gauss(SlikPlugin):
        def __init__(self):
            super().__init__()
            for i in range(ndim):
                self["p%02d" % i] = param(start=start, scale=1)
            self.like = likelihoods.gaussian(np.zeros(ndim), cov)
            self.sampler = sampler(self) if sampler else None

        def __call__(self):
            return self.like([self["p%02d" % i] for i in range(ndim)])

    return gauss


def derived_script(sampler=None):
    """two parameters plus derived quantities stored on the script in __call__ (the script of
    oracle/make_golden_extras.py, for ``output_extra_params``)"""
    class derived(SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = param(start=0, scale=1)
            self.b = param(start=0.5, scale=1, min=-4)
            self.sampler = sampler(self) if sampler else None

        def __call__(self):
            self.r2 = self.a ** 2 + self.b ** 2
            self.ratio = (self.a - 1.5) / (2 + self.b ** 2)
            self.fixed = 3.25
            return self.r2 / 2

    return derived
