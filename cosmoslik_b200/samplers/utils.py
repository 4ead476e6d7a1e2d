import numpy as np
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


def extras_of(lnl, names):
    """The device program of a sampler's ``output_extra_params`` (None without any): the derived
    quantities the reference reads off the evaluated tree, ``s.extra[k]`` (metropolis_hastings.py:228),
    compiled from the script's trace (``Slik.extras_program``)."""
    if not names:
        return None
    slik = getattr(lnl, "__self__", None)
    if slik is None or not hasattr(slik, "extras_program"):
        raise TypeError("output_extra_params needs the bound Slik.evaluate of a traceable script; got %r" % (lnl,))
    return slik.extras_program(names)


def append_extras(blocks, extras, ndim):
    """[(chain id, rows [m, 2+ndim])] -> the same with the derived-quantity columns appended
    (rows [m, 2+ndim+nextra]); every new row of the block is evaluated in one device batch."""
    counts = [len(r) for _, r in blocks]
    if sum(counts) == 0:
        nx = len(extras.rows)
        return [(cid, np.zeros((0, r.shape[1] + nx))) for cid, r in blocks]
    X = np.concatenate([r[:, 2:2 + ndim] for _, r in blocks if len(r)])
    E = extras.values_host(X)
    out, at = [], 0
    for (cid, r), m in zip(blocks, counts):
        out.append((cid, np.hstack([r, E[at:at + m]])))
        at += m
    return out


def append_extras_of_next(blocks, pos_now, extras, ndim):
    """Extra columns of the emcee wrapper's file rows.  The reference writes, next to a walker's OLD
    position, the lnl AND the derived quantities of the position it moved TO (emcee.py:66-78: ``l`` and
    ``params`` are ``prob`` / ``blobs`` of ``posnext``).  A walker emits a row exactly when it moves (or at
    the last step), so the position a row moved to is the x of that walker's next row, and for its
    newest row the walker's current position ``pos_now[c]`` (it has not moved since).
    ``blocks[c] = (walker id, rows [m, 2+ndim])`` for local walker c, ``pos_now [W, ndim]``."""
    nxt = [np.vstack([r[1:, 2:2 + ndim], pos_now[c][None, :]]) for c, (_, r) in enumerate(blocks) if len(r)]
    nx = len(extras.rows)
    if not nxt:
        return [(wid, np.zeros((0, r.shape[1] + nx))) for wid, r in blocks]
    E = extras.values_host(np.concatenate(nxt))
    out, at = [], 0
    for wid, r in blocks:
        out.append((wid, np.hstack([r, E[at:at + len(r)]])))
        at += len(r)
    return out
