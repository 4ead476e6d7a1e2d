"""Poisson + clustered point-source foregrounds with a free clustering index.

Same model as the reference plugin (models/clust_poisson_egfs.py:9-17):

    D_ell = Aps (ell/3000)^2 + Acib (ell/3000)^ncib        for TT, nothing for other spectra

with ell = 0 .. lmax-1 (the reference's ``arange(lmax)``, so ell = 0 is included).  With batched
parameters ``Aps``, ``Acib`` and ``ncib`` are symbolic: the first component becomes a template term,
the second a power-law term whose ``exp(ncib * log(ell/3000))`` is evaluated per walker on the device.
"""
import numpy as np

from .egfs import egfs

PIVOT = 3000.0


class clust_poisson_egfs(egfs):

    def get_colors(self):
        return dict(ps="g", cl="b")

    def get_egfs(self, Aps, Acib, ncib, spectra, lmax, **kwargs):
        components = {}
        if spectra == "cl_TT":
            x = np.arange(lmax) / PIVOT
            components["ps"] = Aps * x ** 2
            components["cib"] = Acib * x ** ncib
        return components
