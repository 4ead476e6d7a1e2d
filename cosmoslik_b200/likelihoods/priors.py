"""Priors declared on `param` objects.

Reference behaviour (likelihoods/priors.py:6-30) that the device tables must reproduce: a sampled
parameter may carry ``gaussian_prior=(mean, std)``, ``uniform_prior=(lo, hi)``, ``range=(lo, hi)``,
``min`` and/or ``max``.  Boxes are collected per parameter in exactly that order; ``min`` and ``max``
together give one two-sided box AND the two one-sided boxes (three in total), ``min`` alone gives
``(min, +inf)``, ``max`` alone ``(-inf, max)``.  Leaving any box gives +inf; the Gaussian penalties add
``(x - mean)**2 / 2 / std**2`` in list order.

The numbers are evaluated on the device by ``k_prepare`` (csrc/prepare.cu) from the two lists built
here -- this plugin has no host evaluation path.
"""
import math

from ..core import SlikPlugin

# (attribute(s) that must be present, function giving the (lo, hi) box) in the reference's order
_BOX_RULES = (
    (("uniform_prior",), lambda p: tuple(p.uniform_prior)),
    (("range",), lambda p: tuple(p.range)),
    (("min", "max"), lambda p: (p.min, p.max)),
    (("min",), lambda p: (p.min, math.inf)),
    (("max",), lambda p: (-math.inf, p.max)),
)


class priors(SlikPlugin):

    def __init__(self, params):
        super().__init__()
        self.gaussian_priors = []      # [(name, mean, std)]
        self.uniform_priors = []       # [(name, lo, hi)]
        for name, p in params.find_sampled().items():
            if hasattr(p, "gaussian_prior"):
                mean, std = p.gaussian_prior
                self.add_gaussian_prior(name, mean, std)
            for needed, box in _BOX_RULES:
                if all(hasattr(p, a) for a in needed):
                    self.add_uniform_prior(name, *box(p))

    def add_gaussian_prior(self, param, mean, std):
        self.gaussian_priors += [(param, mean, std)]

    def add_uniform_prior(self, param, min, max):
        self.uniform_priors += [(param, min, max)]

    def __call__(self, params):
        raise RuntimeError("priors are evaluated on the device as part of the compiled program "
                           "(Slik.evaluate / Slik.evaluate_batch); there is no host evaluation path")
