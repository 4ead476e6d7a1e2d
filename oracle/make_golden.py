"""Generate tests/golden/* by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is
mounted:   python -m oracle.make_golden
The outputs are committed; the GPU box (no reference tree) only reads them.

What is produced (all from the reference's own code paths, via oracle/refshim.py):
  spt_lowl_data.npz   arrays the reference's spt_lowl loader builds from its
                      bundled tarballs (spt_lowl.py:65-108) + the three egfs
                      templates (spt_lowl.py:196-203), for 's12' and 'k11'
  spt_lowl_lnl.npz    Slik.evaluate (cosmoslik.py:109-141) on a batch of
                      parameter vectors, incl. hard-prior hits
  egfs.npz            clust_poisson_egfs outputs (clust_poisson_egfs.py:9-17)
  mh_quick.npz        metropolis_hastings._mcmc rows + chain file for the
                      2-parameter quickstart script with INJECTED streams
  mh_gauss30.npz      the same for a 30-parameter correlated Gaussian
  mh_spt.npz          the same for the spt_lowl 's12' script (prior rejections)
  cov.npz             get_covariance / get_new_cov / initialize_covariance
  ref_mh.chain        a chain file written by the reference (pickle stream)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def synthetic_tt(n=6000):
    ell = np.arange(n)
    return 5000.0 * np.exp(-(ell / 1200.0) ** 1.2) + 50.0


def run_injected_mh(R, script, nsteps, z, u, tmpfile=None, **kw):
    """Run the reference sampler with its two random sources replaced by the
    given streams: draw_x -> cur_x + z[t] @ F (the matrix numpy builds), and
    the module-level ``uniform`` (metropolis_hastings.py:2,154)."""
    from oracle.runtime import proposal_factor
    mh = R.mh
    zs, us = iter(z), iter(u)

    class injected(mh.metropolis_hastings):
        def draw_x(self, cur_x):
            F = proposal_factor(self.cov_est, len(self.x0), self.proposal_scale)
            return np.asarray(cur_x) + np.dot(next(zs), F)

    old_uniform = mh.uniform
    mh.uniform = lambda: next(us)
    try:
        s = script()
        s.sampler = injected(s, num_samples=nsteps, reseed=False, output_file=tmpfile, **kw)
        slik = R.core.Slik(s)
        rows = [(r.lnl, r.weight, np.array(r.x, dtype=float)) for r in slik.sample()]
    finally:
        mh.uniform = old_uniform
    names = list(slik.get_sampled().keys())
    return names, rows, slik


def pack_rows(rows):
    return (np.array([r[0] for r in rows]), np.array([r[1] for r in rows], dtype=float),
            np.array([r[2] for r in rows]))


def main():
    R = refshim.load()
    core = R.core
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(20260101)

    # ---------------- spt_lowl data + lnl -------------------------------
    data = {}
    lnl_out = {}
    tt = synthetic_tt()
    for which in ("s12", "k11"):
        like = R.spt_lowl.spt_lowl(which)
        data[which + "_spec"] = like.spec
        data[which + "_sigma"] = like.sigma
        data[which + "_windows"] = np.array(like.windows)
        data[which + "_lmin"] = np.array(like.windowrange.start)
        if which == "s12":
            data["template_sz"] = like.egfs.template_sz[:3400]
            data["template_cl"] = like.egfs.template_cl[:3400]
            data["template_ps"] = like.egfs.template_ps[:3400]

        class script(core.SlikPlugin):
            def __init__(self, which=which):
                super().__init__()
                self.spt = R.spt_lowl.spt_lowl(which)
                self.cmb = {"TT": tt}

            def __call__(self):
                return self.spt(self.cmb)

        slik = core.Slik(_with_dummy_sampler(core, script()))
        names = list(slik.get_sampled().keys())
        start = np.array([slik.get_start()[k] for k in names], dtype=float)
        X = start + rng.standard_normal((96, len(names))) * np.array([slik.get_sampled()[k].scale for k in names])
        X[0] = start
        X[5, names.index("spt.egfs.Aps")] = -0.25  # hard prior (min=0) -> inf
        X[6, names.index("spt.egfs.Asz")] = -1e-9
        L = np.array([slik.evaluate(*x)[0] for x in X])
        lnl_out[which + "_names"] = np.array(names)
        lnl_out[which + "_x"] = X
        lnl_out[which + "_lnl"] = L
        print(which, names, "lnl(start) = %r" % L[0], "n_inf =", np.isinf(L).sum())
    np.savez_compressed(os.path.join(OUT, "spt_lowl_data.npz"), **data)
    np.savez_compressed(os.path.join(OUT, "spt_lowl_lnl.npz"), tt=tt, **lnl_out)

    # ---------------- egfs ------------------------------------------------
    cpe = R.clust_poisson_egfs.clust_poisson_egfs()
    e_small = cpe(Aps=10, Acib=5, ncib=0.8)(spectra="cl_TT", lmax=5)
    pars = rng.uniform([1, 1, 0.2], [30, 20, 1.6], size=(16, 3))
    e_big = np.array([cpe(Aps=a, Acib=b, ncib=c)(spectra="cl_TT", lmax=3301) for a, b, c in pars])
    not_tt = cpe(Aps=10, Acib=5, ncib=0.8)(spectra="cl_EE", lmax=5)
    assert not_tt == 0
    np.savez_compressed(os.path.join(OUT, "egfs.npz"), small=e_small, pars=pars, big=e_big)
    print("clust_poisson small:", e_small)

    # ---------------- MH: quickstart 2-parameter Gaussian ------------------
    class quick(core.SlikPlugin):  # doc/quickstart.md:26-43
        def __init__(self):
            super().__init__()
            self.a = core.param(start=0, scale=1)
            self.b = core.param(start=0, scale=1)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    n = 4000
    z = rng.standard_normal((n, 2))
    u = rng.uniform(size=n)
    chainfile = os.path.join(OUT, "ref_mh.chain")
    names, rows, _ = run_injected_mh(R, quick, n, z, u, tmpfile=chainfile)
    lnl, w, x = pack_rows(rows)
    loaded = R.chains.load_chain(chainfile)
    np.savez_compressed(os.path.join(OUT, "mh_quick.npz"), z=z, u=u, names=np.array(names), lnl=lnl, weight=w, x=x,
                        file_lnl=loaded[0]["lnl"], file_weight=loaded[0]["weight"],
                        file_a=loaded[0]["a"], file_b=loaded[0]["b"])
    print("mh_quick rows", len(rows), "sum w", w.sum(), "file rows", len(loaded[0]["lnl"]))

    # ---------------- MH: 30-parameter correlated Gaussian -----------------
    A = np.random.RandomState(0).standard_normal((30, 30))
    C = A @ A.T / 30 + np.eye(30)
    Cinv = np.linalg.inv(C)

    class gauss30(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            for i in range(30):
                self["p%02d" % i] = core.param(start=0, scale=1)

        def __call__(self):
            v = np.array([self["p%02d" % i] for i in range(30)])
            return v @ Cinv @ v / 2

    n = 1500
    z = rng.standard_normal((n, 30))
    u = rng.uniform(size=n)
    names, rows, _ = run_injected_mh(R, gauss30, n, z, u)
    lnl, w, x = pack_rows(rows)
    np.savez_compressed(os.path.join(OUT, "mh_gauss30.npz"), z=z, u=u, cov=C, names=np.array(names), lnl=lnl,
                        weight=w, x=x)
    print("mh_gauss30 rows", len(rows), "sum w", w.sum())

    # ---------------- MH: spt_lowl script ---------------------------------
    class sptscript(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.spt = R.spt_lowl.spt_lowl("s12")
            self.cmb = {"TT": tt}

        def __call__(self):
            return self.spt(self.cmb)

    n = 400
    z = rng.standard_normal((n, 4))
    u = rng.uniform(size=n)
    names, rows, _ = run_injected_mh(R, sptscript, n, z, u, max_weight=4)
    lnl, w, x = pack_rows(rows)
    np.savez_compressed(os.path.join(OUT, "mh_spt.npz"), z=z, u=u, names=np.array(names), lnl=lnl, weight=w, x=x)
    print("mh_spt rows", len(rows), "sum w", w.sum(), names)

    # ---------------- covariance helpers -----------------------------------
    dat = rng.standard_normal((200, 5)) @ rng.standard_normal((5, 5))
    wts = rng.randint(0, 11, size=200).astype(float)
    cov_w = R.chains.get_covariance(dat, wts)
    cov_u = R.chains.get_covariance(dat)
    # get_new_cov wants structured arrays 'f8,f8,...' (metropolis_hastings.py:221, 249-253)
    chains = []
    for c in range(3):
        m = 40 + 17 * c
        s = np.empty(m, ",".join(["f8"] * 7))
        blk = rng.standard_normal((m, 5)) + c
        for i in range(m):
            s[i] = tuple([rng.uniform(), float(rng.randint(0, 11))] + list(blk[i]))
        chains.append(s)
    new_cov = R.mh.get_new_cov(chains + [None], 5)
    chain_x = [np.array([list(r)[2:] for r in s]) for s in chains]
    chain_w = [np.array([r[1] for r in s]) for s in chains]
    sampled = core.SlikDict(a=core.param(start=0, scale=2.0), b=core.param(start=1, scale=0.5),
                            c=core.param(start=1)).find_sampled()
    init_cov = R.sutils.initialize_covariance(sampled, (["c", "a"], np.array([[4.0, 0.3], [0.3, 9.0]])))
    out = dict(dat=dat, wts=wts, cov_w=cov_w, cov_u=cov_u, new_cov=new_cov, init_cov=init_cov)
    for c in range(3):
        out["chain_x%d" % c] = chain_x[c]
        out["chain_w%d" % c] = chain_w[c]
    np.savez_compressed(os.path.join(OUT, "cov.npz"), **out)

    # ---------------- samplers.priors: i.i.d. draws from the uniform priors (samplers/priors.py:11-47) ----
    import cosmoslik_plugins.samplers.priors as ref_priors

    class boxed(core.SlikPlugin):
        def __init__(self):
            super().__init__()
            self.a = core.param(start=0, scale=1, range=(-2, 3))
            self.b = core.param(start=0, scale=1, uniform_prior=(0.5, 1.5), gaussian_prior=(1.0, 0.2))
            self.sampler = ref_priors.priors(self, num_samples=200, reseed=False)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    np.random.seed(4242)
    rows = [(r.lnl, r.weight, np.array(r.x, dtype=float)) for r in core.Slik(boxed()).sample()]
    lnl, w, x = pack_rows(rows)
    np.savez_compressed(os.path.join(OUT, "prior_sampler.npz"), seed=4242, lnl=lnl, weight=w, x=x)
    print("prior_sampler rows", len(rows))
    print("wrote", sorted(os.listdir(OUT)))


def _with_dummy_sampler(core, s):
    s.sampler = None
    return s


if __name__ == "__main__":
    main()
