"""Dense Gaussian likelihood  1/2 (v - mean)^T C^-1 (v - mean).

The reference has no Gaussian plugin class; its Gaussian likelihoods are script
bodies (doc/quickstart.md:41-43, 103-104).  This plugin is the batched form of
BASELINE config 2 (30 correlated parameters): the factor of C is taken once
(the reference's pattern, spt_lowl.py:98) and chi^2 runs on the device, for MH
inside the fused whole-run kernel ``csl_mh_gauss_run``.
"""
import numpy as np

from ..core import SlikPlugin
from ..symbolic import GaussTerm


class gaussian(SlikPlugin):

    def __init__(self, mean, cov):
        super().__init__()
        self.mean = np.atleast_1d(np.asarray(mean, dtype=float))
        self.cov = np.atleast_2d(np.asarray(cov, dtype=float))

    def __call__(self, values):
        return GaussTerm(list(values), self.mean, self.cov)
