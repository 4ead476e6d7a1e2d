"""Synthetic bandpower problems of the BASELINE.json shapes.

CAMB/CLASS and the Planck data are not available offline, so theory spectra are
fixed smooth templates with amplitude / tilt parameters and the data are the
model at a truth point plus correlated Gaussian noise (BASELINE.json
north_star).  The problem DATA (windows, templates, covariance, bandpowers) are
plain numpy arrays; ``build_script`` writes the CosmoSlik script once, against
whatever plugin namespace it is given -- the product package, or the CPU oracle
in tests / bench -- so both sides evaluate literally the same script.

Window structure (state it next to every number):
  planck_lite : top-hat bins, plik-lite layout (delta ell 5/9/17/33), weights ~ l(l+1);
                each ell belongs to exactly one bin -> block-sparse window matrix
  spt_multifreq : Gaussian bumps with 1e-10 tails like the SPT window files -> numerically DENSE
"""
from collections import OrderedDict

import numpy as np


def _cmb_templates(lmax):
    ell = np.arange(lmax, dtype=float)
    env = np.exp(-(ell / 1200.0) ** 1.2)
    tt = 5000.0 * env * (1 + 0.35 * np.cos(ell / 95.0)) + 50.0
    te = 120.0 * env * np.sin(ell / 95.0) * np.sqrt(np.maximum(ell, 1) / 500.0)
    ee = 40.0 * np.exp(-(ell / 1500.0) ** 1.3) * (1 - 0.6 * np.cos(ell / 95.0)) * (ell / 800.0) + 0.05
    return tt, te, ee


def _fg_templates(lmax):
    ell = np.arange(lmax, dtype=float)
    x = np.maximum(ell, 1.0) / 3000.0
    tsz = x ** 0.5 * np.exp(-0.1 * x) / (0.5 + x ** 0.5) * 1.5
    ksz = 1.0 / (1 + (x - 1) ** 2 * 0.3)
    cl = x ** 0.8
    dust = (np.maximum(ell, 2.0) / 500.0) ** -0.4
    return tsz, ksz, cl, dust


def _spd_cov(d, rs, err=0.03, neighbour=0.1, common=0.02):
    """dense SPD covariance: `err` fractional errors, `neighbour` correlation between adjacent
    bins and a `common` calibration-like mode"""
    n = d.size
    R = np.eye(n) + neighbour * (np.eye(n, k=1) + np.eye(n, k=-1))
    v = np.ones(n) + 0.1 * rs.standard_normal(n)
    R += common * np.outer(v, v)
    sig = err * np.abs(d) + 0.5
    return R * np.outer(sig, sig)


def plik_bins(lmin=30, lmax=2508):
    """plik-lite bin edges: widths 5 (l<100), 9 (l<1504), 17 (l<2014), 33"""
    edges, l = [], lmin
    while l <= lmax:
        w = 5 if l < 100 else 9 if l < 1504 else 17 if l < 2014 else 33
        edges.append((l, min(l + w, lmax + 1)))
        l += w
    return edges


def planck_lite_problem(seed=0):
    """BASELINE config 4: TT/TE/EE, 215/199/199 bins = 613, ell 30..2508, dense covariance."""
    rs = np.random.RandomState(seed)
    lmax = 2509
    tt, te, ee = _cmb_templates(lmax)
    tsz, ksz, cl, dust = _fg_templates(lmax)
    bins_tt = plik_bins(30, 2508)
    bins_pol = plik_bins(30, 1996)
    assert len(bins_tt) == 215 and len(bins_pol) == 199

    def tophat(bins, lo, hi):
        w = np.zeros((len(bins), hi - lo))
        for b, (a, e) in enumerate(bins):
            l = np.arange(a, e, dtype=float)
            ww = l * (l + 1)
            w[b, a - lo:e - lo] = ww / ww.sum()
        return w

    T, E = "T143", "E143"
    use = {(T, T): (30, 2509), (T, E): (30, 1997), (E, E): (30, 1997)}
    windows = {(T, T): tophat(bins_tt, 30, 2509), (T, E): tophat(bins_pol, 30, 1997),
               (E, E): tophat(bins_pol, 30, 1997)}
    truth = OrderedDict(A_cmb=1.0, n_tilt=0.0, A_ps=60.0, A_cib=8.0, n_cib=0.8, A_tsz=4.0, A_ksz=2.0,
                        A_dust=0.4, y_cal=1.0)
    scales = OrderedDict(A_cmb=0.002, n_tilt=0.002, A_ps=3.0, A_cib=1.0, n_cib=0.1, A_tsz=1.0, A_ksz=1.0,
                         A_dust=0.1, y_cal=0.002)
    prob = dict(kind="planck_lite", maps=(T, E), use=use, windows=windows, truth=truth, scales=scales,
                templates=dict(tt=tt, te=te, ee=ee, tsz=tsz, ksz=ksz, cl=cl, dust=dust), lpivot=1000.0)
    _attach_data(prob, rs)
    return prob


def spt_multifreq_problem(seed=0, nbins=50, lmin=50, lmax=3301):
    """BASELINE config 3: 3 maps -> 6 cross-spectra x 50 bins = 300, ell 50..3300 (3251 columns),
    numerically dense windows, Poisson / clustered / tSZ / kSZ foregrounds."""
    rs = np.random.RandomState(seed)
    tt, _, _ = _cmb_templates(lmax)
    tsz, ksz, cl, dust = _fg_templates(lmax)
    maps = ("T090", "T150", "T220")
    keys = [(a, b) for i, a in enumerate(maps) for b in maps[i:]]
    nl = lmax - lmin
    centers = lmin + (np.arange(nbins) + 0.5) * nl / nbins
    l = np.arange(lmin, lmax, dtype=float)
    w = np.exp(-0.5 * ((l[None, :] - centers[:, None]) / (0.45 * nl / nbins)) ** 2) + 1e-10
    w /= w.sum(1, keepdims=True)
    use = {k: (lmin, lmax) for k in keys}
    windows = {k: w.copy() for k in keys}
    truth = OrderedDict(A_cmb=1.0, A_cib=8.0, n_cib=0.8, A_tsz=4.0, A_ksz=2.0, cal150=1.0, cal220=1.0)
    scales = OrderedDict(A_cmb=0.002, A_cib=1.0, n_cib=0.1, A_tsz=1.0, A_ksz=1.0, cal150=0.005, cal220=0.01)
    for a, b in keys:
        truth["Aps_%s_%s" % (a[1:], b[1:])] = 20.0 + 10.0 * (int(a[1:]) + int(b[1:])) / 300.0
        scales["Aps_%s_%s" % (a[1:], b[1:])] = 2.0
    prob = dict(kind="spt_multifreq", maps=maps, use=use, windows=windows, truth=truth, scales=scales,
                templates=dict(tt=tt, tsz=tsz, ksz=ksz, cl=cl, dust=dust),
                fsz={"090": 1.6, "150": 1.0, "220": 0.05}, fcib={"090": 0.3, "150": 1.0, "220": 3.0})
    _attach_data(prob, rs)
    return prob


def model_vector(prob, p):
    """numpy evaluation of the binned model at parameter dict p (problem SET-UP only: it makes
    the synthetic data vector; it is not an evaluation path of the samplers)."""
    out = []
    for key in sorted(prob["use"]):
        lmin, lmax = prob["use"][key]
        cl = _model_cl(prob, key, p, lmax)
        out.append(prob["windows"][key] @ cl[lmin:lmax])
    return np.concatenate(out)


def _model_cl(prob, key, p, lmax):
    t = prob["templates"]
    ell = np.arange(lmax, dtype=float)
    a, b = key
    if prob["kind"] == "planck_lite":
        tilt = (np.maximum(ell, 1.0) / prob["lpivot"]) ** p["n_tilt"]
        if a[0] == "T" and b[0] == "T":
            x = ell / 3000.0
            return (p["A_cmb"] * tilt * t["tt"][:lmax] + p["A_ps"] * x ** 2 + p["A_cib"] * x ** p["n_cib"]
                    + p["A_tsz"] * t["tsz"][:lmax] + p["A_ksz"] * t["ksz"][:lmax])
        if a[0] != b[0]:
            return p["A_cmb"] * tilt * t["te"][:lmax] + (0.3 * p["A_dust"]) * t["dust"][:lmax]
        return p["A_cmb"] * tilt * t["ee"][:lmax] + p["A_dust"] * t["dust"][:lmax]
    if prob["kind"] == "unbinned":
        tilt = (np.maximum(ell, 1.0) / prob["lpivot"]) ** p["n_tilt"]
        if a[0] == "T" and b[0] == "T":
            return p["A_cmb"] * tilt * t["tt"][:lmax] + p["A_ps"] * (ell / 3000.0) ** 2
        return p["A_cmb"] * tilt * (t["te"] if a[0] != b[0] else t["ee"])[:lmax]
    x = ell / 3000.0
    fa, fb = a[1:], b[1:]
    return (p["A_cmb"] * t["tt"][:lmax] + p["Aps_%s_%s" % (fa, fb)] * x ** 2
            + (p["A_cib"] * prob["fcib"][fa] * prob["fcib"][fb]) * x ** p["n_cib"]
            + (p["A_tsz"] * prob["fsz"][fa] * prob["fsz"][fb]) * t["tsz"][:lmax] + p["A_ksz"] * t["ksz"][:lmax])


def _attach_data(prob, rs):
    m = model_vector(prob, prob["truth"])
    cov = _spd_cov(m, rs)
    d = m + np.linalg.cholesky(cov) @ rs.standard_normal(m.size)
    prob["cov"] = cov
    signal, i = {}, 0
    for key in sorted(prob["use"]):
        nb = prob["windows"][key].shape[0]
        signal[key] = d[i:i + nb]
        i += nb
    prob["signal"] = signal
    prob["n"] = m.size


def build_script(ns, prob, sampler=None):
    """The CosmoSlik script for `prob`, written against plugin namespace `ns`, which provides
    SlikPlugin, param, mspec_lnl and egfs (the foreground base class).  Returns the class."""
    t = prob["templates"]
    truth, scales = prob["truth"], prob["scales"]
    planck = prob["kind"] == "planck_lite"
    unbinned = prob["kind"] == "unbinned"

    class foregrounds(ns.egfs):
        """per-spectrum foreground model; called by mspec_lnl as
        egfs(spectra='cl_TT', lmax=..., spec=(a, b))  (mspec_lnl.py:58)"""

        def get_egfs(self, spectra, lmax, spec, **p):
            a, b = spec
            ell = np.arange(lmax) / 3000.0
            if unbinned:
                return {"ps": p["A_ps"] * ell ** 2} if (a[0] == "T" and b[0] == "T") else {}
            if planck:
                if a[0] == "T" and b[0] == "T":
                    return {"ps": p["A_ps"] * ell ** 2, "cib": p["A_cib"] * ell ** p["n_cib"],
                            "tsz": p["A_tsz"] * t["tsz"][:lmax], "ksz": p["A_ksz"] * t["ksz"][:lmax]}
                if a[0] != b[0]:
                    return {"dust": (0.3 * p["A_dust"]) * t["dust"][:lmax]}
                return {"dust": p["A_dust"] * t["dust"][:lmax]}
            fa, fb = a[1:], b[1:]
            return {"ps": p["Aps_%s_%s" % (fa, fb)] * ell ** 2,
                    "cib": (p["A_cib"] * prob["fcib"][fa] * prob["fcib"][fb]) * ell ** p["n_cib"],
                    "tsz": (p["A_tsz"] * prob["fsz"][fa] * prob["fsz"][fb]) * t["tsz"][:lmax],
                    "ksz": p["A_ksz"] * t["ksz"][:lmax]}

    class script(ns.SlikPlugin):
        def __init__(self):
            super().__init__()
            for k in truth:
                kw = dict(start=truth[k], scale=scales[k])
                if k.startswith("A") and k != "A_cmb":
                    kw["min"] = 0
                if k == "n_cib":
                    kw["range"] = (0.0, 2.0)
                if k == "n_tilt":
                    kw["range"] = (-0.5, 0.5)
                if k in ("y_cal", "cal150", "cal220"):
                    kw["gaussian_prior"] = (1.0, 5 * scales[k])
                self[k] = ns.param(**kw)
            self.fg = foregrounds()
            self.like = ns.mspec_lnl(prob["use"], prob["windows"], prob["signal"], prob["cov"])
            self.sampler = sampler(self) if sampler is not None else None

        def __call__(self):
            p = {k: self[k] for k in truth}
            lmax = max(v[1] for v in prob["use"].values())
            if planck or unbinned:
                tilt = (np.maximum(np.arange(lmax), 1.0) / prob["lpivot"]) ** p["n_tilt"]
                cmb = {"cl_TT": p["A_cmb"] * (tilt * t["tt"][:lmax]), "cl_TE": p["A_cmb"] * (tilt * t["te"][:lmax]),
                       "cl_EE": p["A_cmb"] * (tilt * t["ee"][:lmax])}
                # plik-lite style calibration: the DATA are scaled by y_cal^2
                self.like.cal = {prob["maps"][0]: p["y_cal"], prob["maps"][1]: p["y_cal"]}
            else:
                cmb = {"cl_TT": p["A_cmb"] * t["tt"][:lmax]}
                self.like.cal = {"T150": p["cal150"], "T220": p["cal220"]}
            return self.like(cmb, self.fg(**p))

    return script


def unbinned_problem(lmax=2501, seed=0):
    """BASELINE config 5: UNBINNED TT/TE/EE over ell = 2..lmax-1 (3 x 2499 = 7497 data points at
    lmax = 2501), identity windows, dense covariance (banded + low rank, stored dense)."""
    rs = np.random.RandomState(seed)
    tt, te, ee = _cmb_templates(lmax)
    T, E = "T143", "E143"
    nl = lmax - 2
    use = {(T, T): (2, lmax), (T, E): (2, lmax), (E, E): (2, lmax)}
    windows = {k: np.eye(nl) for k in use}
    truth = OrderedDict(A_cmb=1.0, n_tilt=0.0, A_ps=60.0, y_cal=1.0)
    scales = OrderedDict(A_cmb=0.001, n_tilt=0.001, A_ps=2.0, y_cal=0.001)
    prob = dict(kind="unbinned", maps=(T, E), use=use, windows=windows, truth=truth, scales=scales,
                templates=dict(tt=tt, te=te, ee=ee), lpivot=1000.0, lmax=lmax)
    m = model_vector(prob, truth)
    n = m.size
    sig = 0.05 * np.abs(m) + 1.0
    cov = np.zeros((n, n))
    idx = np.arange(n)
    cov[idx, idx] = 1.0
    cov[idx[:-1], idx[:-1] + 1] = cov[idx[:-1] + 1, idx[:-1]] = 0.2
    cov[idx[:-nl], idx[:-nl] + nl] = cov[idx[:-nl] + nl, idx[:-nl]] = 0.05
    u = rs.standard_normal((n, 3)) * 0.05
    cov += u @ u.T
    cov *= np.outer(sig, sig)
    d = m + np.linalg.cholesky(cov) @ rs.standard_normal(n)
    prob["cov"], prob["n"] = cov, n
    signal, i = {}, 0
    for key in sorted(use):
        signal[key] = d[i:i + nl]
        i += nl
    prob["signal"] = signal
    return prob
