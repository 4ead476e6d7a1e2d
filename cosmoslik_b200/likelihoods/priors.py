"""Priors declared on `param` objects (reference: likelihoods/priors.py:6-30).

On the device the boxes and Gaussian penalties are evaluated by ``k_prepare``
(csrc/prepare.cu) from the tables built here; the list ORDER and the way
attributes expand into boxes follow the reference exactly, including the case
where a parameter with both ``min`` and ``max`` contributes three boxes.
"""
from numpy import inf

from ..core import SlikPlugin


class priors(SlikPlugin):

    def __init__(self, params):
        super().__init__()
        self.gaussian_priors = []
        self.uniform_priors = []
        for name, p in params.find_sampled().items():
            if hasattr(p, "gaussian_prior"):
                self.add_gaussian_prior(name, *p.gaussian_prior)
            if hasattr(p, "uniform_prior"):
                self.add_uniform_prior(name, *p.uniform_prior)
            if hasattr(p, "range"):
                self.add_uniform_prior(name, *p.range)
            if hasattr(p, "min") and hasattr(p, "max"):
                self.add_uniform_prior(name, p.min, p.max)
            if hasattr(p, "min"):
                self.add_uniform_prior(name, p.min, inf)
            if hasattr(p, "max"):
                self.add_uniform_prior(name, -inf, p.max)

    def add_gaussian_prior(self, param, mean, std):
        self.gaussian_priors.append((param, mean, std))

    def add_uniform_prior(self, param, min, max):
        self.uniform_priors.append((param, min, max))

    def __call__(self, params):
        raise RuntimeError("priors are evaluated on the device as part of the compiled program "
                           "(Slik.evaluate / Slik.evaluate_batch); there is no host evaluation path")
