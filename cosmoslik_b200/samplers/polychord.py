"""The PolyChord wrapper's two callbacks, batched (reference: samplers/polychord.py:9-62).

The reference hands PyPolyChord a per-point ``likelihood(theta) -> (-lnl, [extras])`` and a
``prior(hypercube) -> theta`` that maps the unit cube onto each parameter's ``min``/``max`` (or
``range``) box (:43-52).  The nested-sampling engine itself is the third-party package PyPolyChord,
absent here as in the reference tree; what belongs to this package's path is the evaluation, so this
class keeps the reference's constructor and offers the callbacks for whole batches of live points:

    theta = sampler.prior_transform(cube)              # [N, ndim] unit cube -> parameters (host arithmetic)
    logl, extras = sampler.loglikelihood(slik.evaluate, theta)   # one device batch: -lnl [N], extras [N, nextra]

``sample`` runs PyPolyChord with per-point wrappers around those when the package is importable and
raises ImportError otherwise (there is no CPU evaluation path to fall back to).
"""
import os
import os.path as osp

import numpy as np

from ..core import SlikSampler, arguments
from .utils import extras_of, program_of


class polychord(SlikSampler):

    def __init__(self, params, output_file, output_extra_params=None, nlive=None, num_repeats=None,
                 do_clustering=None, read_resume=None, **kwargs):
        if output_extra_params is None:
            output_extra_params = []
        super().__init__(**arguments(exclude=["params"]))
        self.sampled = params.find_sampled()
        self.boxes = []
        for k, p in self.sampled.items():                       # polychord.py:46-51
            if hasattr(p, "min") and hasattr(p, "max"):
                self.boxes.append((float(p.min), float(p.max)))
            elif hasattr(p, "range"):
                self.boxes.append((float(p.range[0]), float(p.range[1])))
            else:
                raise Exception("To use PolyChord, all parameters must have min/max specified")

    def prior_transform(self, hypercube):
        """UniformPrior(a, b)(x) = a + (b - a) x per parameter (polychord.py:43-52), for [N, ndim] or [ndim]"""
        cube = np.asarray(hypercube, dtype=float)
        lo = np.array([b[0] for b in self.boxes])
        hi = np.array([b[1] for b in self.boxes])
        if cube.shape[-1] != len(lo):
            raise ValueError("hypercube must have %d columns" % len(lo))
        return lo + (hi - lo) * cube

    def loglikelihood(self, lnl, theta):
        """(-lnl [N], extras [N, nextra]) for a batch of points (polychord.py:38-41); a missing extra is nan
        in the reference (``p.get(k, nan)``), here every named extra must be a traced derived quantity"""
        theta = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=float)))
        prog = program_of(lnl)
        vals = prog.lnl_host(theta)
        extras = extras_of(lnl, list(self.output_extra_params))
        ex = extras.values_host(theta) if extras is not None else np.zeros((len(theta), 0))
        return -vals, ex

    def sample(self, lnl):
        try:
            from PyPolyChord import run_polychord
            from PyPolyChord.settings import PolyChordSettings
        except ImportError as e:
            raise ImportError("samplers.polychord needs the third-party package PyPolyChord (absent); the batched "
                              "callbacks prior_transform / loglikelihood are available without it") from e
        names = list(self.output_extra_params)
        settings = PolyChordSettings(len(self.sampled), len(names))
        output_file = "./" + osp.normpath(self.output_file)
        settings.base_dir, settings.file_root = osp.dirname(output_file), osp.basename(output_file)
        if not osp.exists(settings.base_dir):
            os.makedirs(settings.base_dir)
        for k in ("nlive", "num_repeats", "do_clustering", "read_resume"):
            if self[k] is not None:
                setattr(settings, k, self[k])

        def likelihood(theta):
            logl, ex = self.loglikelihood(lnl, [theta])
            return float(logl[0]), [float(v) for v in ex[0]]

        with open(self.output_file + ".paramnames", "w") as f:
            for k in list(self.sampled.keys()) + names:
                f.write(k + " " + k + "\n")
        run_polychord(likelihood, len(self.sampled), len(names), settings, lambda cube: list(self.prior_transform(cube)))
        return []
