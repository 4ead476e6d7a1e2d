"""Poisson + power-law clustered foregrounds (reference: models/clust_poisson_egfs.py:9-17)."""
from numpy import arange

from .egfs import egfs


class clust_poisson_egfs(egfs):

    def get_colors(self):
        return {"ps": "g", "cl": "b"}

    def get_egfs(self, Aps, Acib, ncib, spectra, lmax, **kwargs):
        if spectra != "cl_TT":
            return {}
        ell = arange(lmax) / 3000.0          # includes ell = 0, as in the reference
        return {"ps": Aps * ell ** 2, "cib": Acib * ell ** ncib}
