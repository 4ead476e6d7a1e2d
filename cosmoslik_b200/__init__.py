"""cosmoslik_b200 -- the batched posterior-evaluation path of CosmoSlik on B200.

``from cosmoslik_b200 import *`` gives a script the names ``from cosmoslik import *``
gives it in the reference (cosmoslik/__init__.py:1-6): the plugin runtime, the
chain containers and the ``samplers`` / ``likelihoods`` / ``models`` namespaces.
"""
from .core import *          # noqa: F401,F403
from .core import __all__ as _core_all
from .chains import (Chain, Chains, load_chain, load_cosmomc_chain, get_covariance, combine_covs,  # noqa: F401
                     gelman_rubin)
from . import samplers, likelihoods, models, dist  # noqa: F401

__all__ = list(_core_all) + ["Chain", "Chains", "load_chain", "get_covariance", "combine_covs", "gelman_rubin",
                             "samplers", "likelihoods", "models", "dist"]
__version__ = "0.1.0"
