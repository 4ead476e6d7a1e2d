"""CPU oracle, part 3: the samplers, restated with INJECTED random streams.

TEST INFRASTRUCTURE ONLY (see oracle/runtime.py for the rules).

Parity status:
  * ``mh_mcmc`` / ``mh_single_process_blocks`` / ``pooled_new_cov``: PINNED
    against the unmodified reference run with the same injected streams
    (tests/golden/mh_*.npz, tests/test_oracle_golden.py).
  * ``stretch_half`` / ``ensemble_step``: the reference delegates the stretch
    move to the third-party package ``emcee`` (samplers/emcee.py:63-68; not in
    requirements.txt or setup.py, version unpinned; the 4-tuple returned by
    ``run_mcmc`` implies the 2.x API, last release 2.2.1).  It is absent from
    this image and from the reference tree.  PARITY UNPINNED: restated from
    the published algorithm (Goodman & Weare 2010, eq. 7-10; emcee 2.x
    ``EnsembleSampler.sample`` / ``_propose_stretch``).  The wrapper logic
    around it (``emcee_wrapper_rows``) follows samplers/emcee.py:34-88 and IS
    pinned: the reference's own wrapper file runs over a stand-in of the
    package (oracle/emcee_standin.py) whose move is ``ensemble_step``, and its
    yields and chain file are tests/golden/emcee_wrapper.npz / ref_emcee.chain.
  * ``mh_lockstep``: the reference's adaptive mode is an asynchronous MPI
    master/worker protocol whose timing is not deterministic
    (metropolis_hastings.py:188-241).  The lock-step analogue used by the
    B200 sampler is DEFINED here; its arithmetic (accept rule, row emission,
    ``pooled_new_cov``, ``proposal_factor``) is the pinned reference
    arithmetic, the exchange schedule is new.  Protocol PARITY UNPINNED.
  * ``gelman_rubin``: no such statistic exists in the reference; defined here.
"""
import math

import numpy as np

from .runtime import get_covariance, proposal_factor


class Row(object):
    """cosmoslik.py:241-250 ``sample``."""
    __slots__ = ("x", "lnl", "weight", "extra")

    def __init__(self, x, lnl, weight=1, extra=None):
        self.x, self.lnl, self.weight, self.extra = x, lnl, weight, extra


def mh_mcmc(lnl, x0, nsteps, draw, uniform, max_weight=10, temp=1, yield_rejected=False):
    """metropolis_hastings.py:144-164, one chain.

    ``lnl(*x) -> (value, extra)``; ``draw(cur_x) -> test_x`` consumes the
    proposal stream, ``uniform()`` consumes exactly one number per step, drawn
    AFTER the likelihood call and even when the prior already rejected
    (:151-154).  x0 enters with weight 0 (:148); the final state is never
    emitted; a max_weight flush resets the weight to 0 (:162-164); rows whose
    current lnl is +inf are not emitted on accept (:155)."""
    cur_w, cur_x = 0, x0
    cur_lnl, cur_extra = lnl(*x0)
    for _ in range(int(nsteps)):
        test_x = draw(cur_x)
        test_lnl, test_extra = lnl(*test_x)
        if math.log(uniform()) < temp * (cur_lnl - test_lnl):
            if cur_lnl != np.inf:
                yield Row(cur_x, cur_lnl, cur_w, cur_extra)
            cur_lnl, cur_w, cur_x, cur_extra = test_lnl, 1, test_x, test_extra
        else:
            if yield_rejected:
                yield Row(test_x, test_lnl, 0, test_extra)
            cur_w += 1
            if cur_w >= max_weight:
                yield Row(cur_x, cur_lnl, cur_w, cur_extra)
                cur_w = 0


def mh_single_process_blocks(rows, ndim, comm_freq=100):
    """metropolis_hastings.py:219-244, size == 1 branch.

    Takes the emitted rows of one chain and returns ``(yielded, file_blocks)``:
    every row is yielded to the caller, but the block written to the chain file
    is ``samples[:i]`` with i the index of the LAST row of the block (:237), so
    the file drops one row per block; the loop ends after the first block that
    is not full (:239-241), writing an empty block when the row count is an
    exact multiple of ``comm_freq``."""
    it = iter(rows)
    yielded, blocks = [], []
    while True:
        buf = np.zeros((comm_freq, 2 + ndim))
        i = 0
        for i, s in zip(range(comm_freq), it):
            yielded.append(s)
            buf[i, 0], buf[i, 1] = s.lnl, s.weight
            buf[i, 2:] = s.x
        blocks.append(buf[:i].copy())
        if i < comm_freq - 1:
            break
    return yielded, blocks


def pooled_new_cov(chain_rows, weights):
    """metropolis_hastings.py:249-253 ``get_new_cov``: weighted covariance of
    the last half of each chain's ROWS (``s[s.shape[0]//2:]``), pooled over
    chains, through chains.py:507-512.  ``chain_rows[c]`` is [nrows_c, ndim],
    ``weights[c]`` is [nrows_c]; chains with no rows at all (``None``) are
    skipped."""
    xs, ws = [], []
    for x, w in zip(chain_rows, weights):
        if x is None:
            continue
        x = np.asarray(x, dtype=float).reshape(len(w), -1)
        h = x.shape[0] // 2
        xs.append(x[h:])
        ws.append(np.asarray(w, dtype=float)[h:])
    return get_covariance(np.concatenate(xs), np.concatenate(ws))


def mh_lockstep(lnl_batch, x0, nsteps, cov_est, z=None, delta=None, u=None,
                proposal_scale=2.4, proposal_update=True, proposal_update_start=1000,
                exchange=100, max_weight=10, temp=1.0):
    """Lock-step ensemble of independent MH chains with pooled proposal
    adaptation -- the synchronous analogue of metropolis_hastings.py:177-244.

    All chains take step t together.  ``z[t, c, :]`` are standard normals and
    the proposal is ``cur_x + z @ F`` with ``F = proposal_factor(cov_est)``
    (the matrix numpy's multivariate_normal uses, :141-142); alternatively
    ``delta[t, c, :]`` injects the offset itself.  ``u[t, c]`` is the accept
    uniform.  After every ``exchange`` steps, once more than
    ``proposal_update_start`` steps have been taken (and steps remain), ``cov_est`` is replaced by
    ``pooled_new_cov`` of all rows emitted so far (the master's rule, :200-201).

    Returns per-chain lists of rows (lnl, weight, x), the final state and the
    list of (step, cov_est) adaptation points."""
    x0 = np.asarray(x0, dtype=float)
    nch, ndim = x0.shape
    cur_x = x0.copy()
    cur_lnl = np.asarray(lnl_batch(cur_x), dtype=float).copy()
    cur_w = np.zeros(nch)
    rows = [[] for _ in range(nch)]
    cov_est = np.asarray(cov_est, dtype=float)
    F = proposal_factor(cov_est, ndim, proposal_scale)
    adapt = []
    for t in range(nsteps):
        if delta is not None:
            test_x = cur_x + delta[t]
        else:
            test_x = cur_x + z[t] @ F
        test_lnl = np.asarray(lnl_batch(test_x), dtype=float)
        with np.errstate(invalid="ignore"):
            acc = np.log(u[t]) < temp * (cur_lnl - test_lnl)  # inf-inf=nan -> False
        for c in range(nch):
            if acc[c]:
                if cur_lnl[c] != np.inf:
                    rows[c].append((cur_lnl[c], cur_w[c], cur_x[c].copy()))
                cur_lnl[c], cur_w[c], cur_x[c] = test_lnl[c], 1, test_x[c]
            else:
                cur_w[c] += 1
                if cur_w[c] >= max_weight:
                    rows[c].append((cur_lnl[c], cur_w[c], cur_x[c].copy()))
                    cur_w[c] = 0
        if proposal_update and (t + 1) % exchange == 0 and proposal_update_start < (t + 1) < nsteps:
            xs = [np.array([r[2] for r in rc]) if rc else None for rc in rows]
            ws = [np.array([r[1] for r in rc]) if rc else None for rc in rows]
            cov_est = pooled_new_cov(xs, ws)
            F = proposal_factor(cov_est, ndim, proposal_scale)
            adapt.append((t + 1, cov_est.copy()))
    return rows, (cur_x, cur_lnl, cur_w), adapt


def prior_sampler(lnl, sampled_names, uniform_priors, nsamples, uniform):
    """samplers/priors.py:11-47: i.i.d. draws from the uniform priors, every row weight 1.
    ``uniform_priors`` is the list the priors plugin builds (likelihoods/priors.py:6-16); per parameter the
    FIRST box is used, more than two boxes or none is an error (:27-33).  ``uniform(lo, hi)`` is drawn
    parameter by parameter in sorted-name order, sample by sample (:38-39).  PINNED by
    tests/golden/prior_sampler.npz."""
    box = {}
    for k in sampled_names:
        pr = [(n, lo, hi) for (n, lo, hi) in uniform_priors if n == k]
        if len(pr) > 2:
            raise Exception("Can't sample prior for parameters with multiple priors at once.")
        if not pr:
            raise Exception("Prior for parameter '%s' is improper." % k)
        box[k] = pr[0][1:]
    for _ in range(int(nsamples)):
        x = [uniform(*box[k]) for k in sampled_names]
        val, extra = lnl(*x)
        yield Row(x, val, 1, extra)


# --------------------------------------------------------------------------
# affine-invariant ensemble sampler (third-party emcee 2.x, restated)
# --------------------------------------------------------------------------

def stretch_z(u, a=2.0):
    """z = ((a-1) u + 1)^2 / a  (emcee 2.x ``_propose_stretch``; Goodman & Weare
    2010 eq. 9 sampled by inversion)."""
    return ((a - 1.0) * u + 1) ** 2.0 / a


def stretch_half(s, c, lnprob_s, lnprob_fn, u_z, j, u_acc, a=2.0):
    """One half-ensemble update (emcee 2.x ``_propose_stretch`` + the update in
    ``sample``): for walker k of the active half S with partner c[j[k]] of the
    complementary half,
        q = c_j - z (c_j - s_k)
        accept iff (ndim-1) ln z + lnprob(q) - lnprob(s_k) > ln u_acc.
    Stream order per half: Ns uniforms (z), Ns partner indices, the likelihood,
    Ns uniforms (accept).  Returns updated (s, lnprob_s, accepted mask, q,
    lnprob_q)."""
    s = np.array(s, dtype=float)
    lnprob_s = np.array(lnprob_s, dtype=float)
    zz = stretch_z(np.asarray(u_z), a)
    cj = np.asarray(c)[np.asarray(j)]
    q = cj - zz[:, None] * (cj - s)
    newlnp = np.asarray(lnprob_fn(q), dtype=float)
    with np.errstate(invalid="ignore", divide="ignore"):
        lnpdiff = (s.shape[1] - 1.0) * np.log(zz) + newlnp - lnprob_s
        acc = lnpdiff > np.log(np.asarray(u_acc))
    s[acc] = q[acc]
    lnprob_s[acc] = newlnp[acc]
    return s, lnprob_s, acc, q, newlnp


def ensemble_step(p, lnprob, lnprob_fn, u_z, j, u_acc, a=2.0):
    """One full emcee 2.x iteration: first half [0, k/2) is updated against the
    second, then the second against the UPDATED first.  ``u_z``, ``j``,
    ``u_acc`` have shape [2, k/2] (row 0 feeds the first half)."""
    p = np.array(p, dtype=float)
    lnprob = np.array(lnprob, dtype=float)
    k = p.shape[0]
    h = k // 2
    halves = [(slice(0, h), slice(h, k)), (slice(h, k), slice(0, h))]
    nacc = 0
    for hi, (S0, S1) in enumerate(halves):
        s, l, acc, _, _ = stretch_half(p[S0], p[S1], lnprob[S0], lnprob_fn, u_z[hi], j[hi], u_acc[hi], a)
        p[S0] = s
        lnprob[S0] = l
        nacc += int(acc.sum())
    return p, lnprob, nacc


def emcee_wrapper_rows(pos0, lnprob_fn, nsteps, streams, a=2.0, blob_fn=None):
    """samplers/emcee.py:66-88: the reference's bookkeeping around
    ``run_mcmc(pos, 1)``.

    Per step and walker: if the walker did not move and this is not the last
    step its weight grows by one (:70-71); otherwise a row (-lnprob_next,
    weight, x_old) is emitted -- the lnl written beside the OLD position is
    that of the NEXT position (:68, :74, :78) -- and the weight resets to 1.
    ``streams(i) -> (u_z, j, u_acc)`` supplies step i's numbers.
    ``blob_fn(x)``, if given, stands for the ``params`` blob emcee returns with ``posnext``; like the
    lnl, the extras the reference writes beside the old position are those of the next one (:68, :78),
    and the row tuples get them as a fourth element.
    Returns (rows_per_walker, final_pos, final_lnprob)."""
    pos = np.array(pos0, dtype=float)
    k = pos.shape[0]
    weight = np.ones(k)
    rows = [[] for _ in range(k)]
    lnprob = np.asarray(lnprob_fn(pos), dtype=float)  # run_mcmc re-evaluates; same numbers
    for i in range(nsteps):
        u_z, j, u_acc = streams(i)
        posnext, lnprob, _ = ensemble_step(pos, lnprob, lnprob_fn, u_z, j, u_acc, a)
        for w in range(k):
            if (pos[w] == posnext[w]).all() and i != nsteps - 1:
                weight[w] += 1
            else:
                row = (-lnprob[w], weight[w], pos[w].copy())
                rows[w].append(row if blob_fn is None else row + (blob_fn(posnext[w]),))
                weight[w] = 1
        pos = posnext
    return rows, pos, lnprob


# --------------------------------------------------------------------------
# convergence statistic (new functionality; nothing in the reference)
# --------------------------------------------------------------------------

def chain_moments(x, w):
    """Per-chain weighted count, mean and variance (denominator sum(w) - 1,
    the convention of chains.py:507-512)."""
    x = np.asarray(x, dtype=float)
    w = np.asarray(w, dtype=float)
    n = w.sum()
    mu = (x * w[:, None]).sum(0) / n
    var = (w[:, None] * (x - mu) ** 2).sum(0) / (n - 1)
    return n, mu, var


def gelman_rubin(ns, mus, variances):
    """Gelman & Rubin (1992) potential scale reduction per parameter from
    per-chain (n_c, mean_c, var_c):  W = mean_c var_c,  B/n = var_c(mean_c)
    (ddof 1),  n = mean_c n_c,  R = sqrt(((n-1)/n W + B/n) / W)."""
    ns = np.asarray(ns, dtype=float)
    mus = np.asarray(mus, dtype=float)
    variances = np.asarray(variances, dtype=float)
    m = mus.shape[0]
    n = ns.mean()
    W = variances.mean(0)
    mbar = mus.mean(0)
    B_over_n = ((mus - mbar) ** 2).sum(0) / (m - 1)
    return np.sqrt(((n - 1) / n * W + B_over_n) / W)
