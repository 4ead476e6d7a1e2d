"""Extragalactic-foreground model base (reference: models/egfs.py:5-67)."""
from ..core import SlikDict, SlikPlugin


class egfs_specs(SlikDict):
    """kind ('TT', ...), freqs (pair of {'dust','radio','tsz'} band centres in GHz), fluxcut (mJy)."""


class egfs(SlikPlugin):
    """Subclass and override ``get_egfs`` to return a dict of component spectra.  Calling the
    plugin with the sampled parameters returns a closure; calling that with the data set's
    keywords (spectra, lmax, freq, fluxcut, ...) returns the summed spectrum (egfs.py:61-67).
    With batched parameters the components are symbolic spectra (symbolic.Spectrum)."""

    def get_egfs(self, *args, **kwargs):
        raise NotImplementedError()

    def get_colors(self):
        return {}

    def __call__(self, **params):
        def get(**kw):
            kw.update(params)
            return sum(self.get_egfs(**kw).values())
        return get
