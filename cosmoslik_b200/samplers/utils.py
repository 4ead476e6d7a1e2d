from numpy import ix_

from ..chains import combine_covs


def initialize_covariance(sampled, covs=None):
    """ndim x ndim proposal covariance for the parameters in `sampled`: ``scale`` of each
    parameter on the diagonal, overridden by whatever `covs` provides (one or a list of the
    formats chains.combine_covs understands).  Reference: samplers/utils.py:4-23."""
    sources = [{k: v.scale for k, v in sampled.items() if hasattr(v, "scale")}]
    if covs is not None:
        sources += covs if isinstance(covs, list) else [covs]
    names, cov = combine_covs(*sources)
    missing = [s for s in sampled if s not in names]
    if missing:
        raise ValueError("Parameters %s not in covariance and no scale given." % missing)
    order = [names.index(s) for s in sampled]
    return cov[ix_(order, order)]


def proposal_factor(cov_est, ndim, proposal_scale):
    """F with ``test_x = cur_x + z @ F``: the factor numpy's legacy
    ``multivariate_normal(cur_x, cov_est/ndim*scale**2)`` builds from an SVD
    (metropolis_hastings.py:141-142), so that injected normal streams reproduce the
    reference's proposals."""
    import numpy as np
    cov = np.asarray(cov_est, dtype=float) / ndim * proposal_scale ** 2
    _, s, v = np.linalg.svd(cov)
    return np.sqrt(s)[:, None] * v


def program_of(lnl):
    """The compiled device program behind the ``lnl`` callback a SlikSampler receives
    (``Slik.evaluate``).  The GPU samplers never call ``lnl`` point by point."""
    slik = getattr(lnl, "__self__", None)
    if slik is None or not hasattr(slik, "program"):
        raise TypeError("cosmoslik_b200 samplers need the bound Slik.evaluate of a traceable script; "
                        "got %r (there is no per-point CPU path)" % (lnl,))
    return slik.program()
