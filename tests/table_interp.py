"""Numpy interpreter of the tables a CompiledProgram holds (TEST-ONLY).

It walks exactly the data the CUDA kernels walk (bytecode, spectrum / tile records, padded
windows, templates, factor), so CPU tests can check the host-side tracing and compilation
against the oracle without a GPU.  It is not a product path."""
import numpy as np
import scipy.linalg as sla

from cosmoslik_b200 import _lib as L


def run_bytecode(cp, x):
    t = cp.tables
    coef = np.zeros(max(cp.meta["ncoef"], 1))
    st = []
    inv = {v: k for k, v in L.OP.items()}
    for op, arg, val in zip(t["code_op"][:cp.meta["ncode"]], t["code_arg"], t["code_val"]):
        name = inv[int(op)]
        if name == "PUSHP": st.append(float(x[arg]))
        elif name == "PUSHC": st.append(float(val))
        elif name == "STORE": coef[arg] = st.pop()
        elif name in ("NEG", "EXP", "LOG", "SQRT", "SQR"):
            a = st.pop()
            f = {"NEG": lambda: -a, "EXP": lambda: np.exp(a), "LOG": lambda: np.log(a), "SQRT": lambda: np.sqrt(a),
                 "SQR": lambda: a * a}[name]
            st.append(f())
        else:
            b = st.pop(); a = st.pop()
            f = {"ADD": lambda: a + b, "SUB": lambda: a - b, "MUL": lambda: a * b, "DIV": lambda: a / b,
                 "POW": lambda: a ** b}[name]
            st.append(f())
    assert not st
    return coef


def prior(cp, x, coef=None):
    """box / Gaussian priors; negative indices name derived quantities (coefficient row -idx - 1)"""
    t, m = cp.tables, cp.meta
    if coef is None and (np.any(t["box_idx"][:m["nbox"]] < 0) or np.any(t["gauss_idx"][:m["ngauss"]] < 0)):
        coef = run_bytecode(cp, x)
    val = lambda ix: x[ix] if ix >= 0 else coef[-ix - 1]
    for b in range(m["nbox"]):
        v = val(int(t["box_idx"][b]))
        if not (t["box_lo"][b] <= v <= t["box_hi"][b]):
            return np.inf
    s = 0.0
    for k in range(m["ngauss"]):
        d = val(int(t["gauss_idx"][k])) - t["gauss_c"][k]
        s = s + d * d / 2 / (t["gauss_w"][k] * t["gauss_w"][k])
    return s


def residual(cp, coef):
    t, m = cp.tables, cp.meta
    r = np.full(m["n_pad"], np.nan)          # every row, padding included, must be written by some record
    written = np.zeros(m["n_pad"], bool)
    if m["resid_mode"] == 2:
        r[:] = 0.0
        for i in range(m["n"]):
            r[i] = coef[t["direct_coef"][i]] - t["data"][i]
        return r
    if m["resid_mode"] == 3:
        pass
    for ci in range(m.get("nsegclass", 0)):
        NTc, NPc, b0, nb = cp.segcls[ci]
        for rec in t["bin_tab"][b0:b0 + nb]:
            if rec[L.B_ROW] < 0:
                continue
            spec = t["spec_tab"][rec[L.B_SPEC]]
            NT, NP = spec[L.S_NTERM], spec[L.S_NPLAW]
            assert NTc <= NT and NP == NPc          # segmented-sum classes carry the REAL number of template terms
            stride = NT + 4 * NP
            l0, l1 = rec[L.B_LO], rec[L.B_HI]
            ltab = t["templates"][spec[L.S_LTAB]: spec[L.S_LTAB] + spec[L.S_NLPAD] * stride].reshape(-1, stride)[l0:l1]
            cl = np.zeros(l1 - l0)
            for k in range(NT):
                cl += coef[spec[L.S_TERM_COEF + k]] * ltab[:, k]
            for k in range(NP):
                with np.errstate(over="ignore"):
                    cl += coef[spec[L.S_PLAW_AMP + k]] * ltab[:, NT + 4 * k + 1] * np.exp(coef[spec[L.S_PLAW_IDX + k]] * ltab[:, NT + 4 * k])
            woff = (int(rec[L.B_WINOFF_HI]) << 32) | (int(rec[L.B_WINOFF_LO]) & 0xFFFFFFFF)
            cal = 1.0 if spec[L.S_CAL] < 0 else coef[spec[L.S_CAL]]
            bval = np.dot(t["windows"][woff + l0: woff + l1], cl)
            dval = t["data"][rec[L.B_ROW]]
            r[rec[L.B_ROW]] = dval - cal * bval if spec[L.S_CALMODE] else cal * dval - bval
            written[rec[L.B_ROW]] = True
    for tile in t["tile_tab"][:m["ntile"]]:
        spec = t["spec_tab"][tile[L.T_SPEC]]
        l0, l1 = tile[L.T_LBEGIN], tile[L.T_LEND]
        NT, NP = spec[L.S_NTERM], spec[L.S_NPLAW]
        stride = NT + 4 * NP
        full = t["templates"][spec[L.S_LTAB]: spec[L.S_LTAB] + spec[L.S_NLPAD] * stride].reshape(-1, stride)
        lo_eval = 0 if spec[L.S_ABERR] else l0           # the aberration factor needs the previous column
        ltab = full[lo_eval:l1]
        cl = np.zeros(l1 - lo_eval)
        for k in range(NT):
            cl += coef[spec[L.S_TERM_COEF + k]] * ltab[:, k]
        for k in range(NP):
            with np.errstate(over="ignore"):
                cl += coef[spec[L.S_PLAW_AMP + k]] * ltab[:, NT + 4 * k + 1] * np.exp(coef[spec[L.S_PLAW_IDX + k]] * ltab[:, NT + 4 * k])
        if spec[L.S_ABERR]:
            G = t["templates"][spec[L.S_ABTAB]: spec[L.S_ABTAB] + spec[L.S_NLPAD]][:l1]
            with np.errstate(divide="ignore", invalid="ignore"):
                lncl = np.log(cl)
                d = np.concatenate([[0.0], lncl[1:] - lncl[:-1]])
            cl = np.where(G != 0, cl * (1 + G * d), cl)[l0:]
        woff = (int(tile[L.T_WINOFF_HI]) << 32) | (int(tile[L.T_WINOFF_LO]) & 0xFFFFFFFF)
        ld = tile[L.T_WINLD]
        cal = 1.0 if spec[L.S_CAL] < 0 else coef[spec[L.S_CAL]]
        for rr in range(tile[L.T_NROWS]):
            w = t["windows"][woff + rr * ld + l0: woff + rr * ld + l1]
            # the group range must cover every non-zero of the row
            g = rr // 8
            nz = np.flatnonzero(w)
            if nz.size:
                assert tile[L.T_GLO + g] <= l0 + nz[0] and l0 + nz[-1] < tile[L.T_GHI + g]
            row = tile[L.T_ROW0] + rr
            r[row] = t["data"][row] - cal * np.dot(w, cl) if spec[L.S_CALMODE] else cal * t["data"][row] - np.dot(w, cl)
            written[row] = True
    assert written.all(), "residual rows %s are written by no tile / bin record" % np.flatnonzero(~written)
    return r


def lnl(cp, x):
    x = np.asarray(x, dtype=float)
    m, t = cp.meta, cp.tables
    pr = prior(cp, x)
    if pr == np.inf:
        return np.inf
    coef = run_bytecode(cp, x)
    like = 0.0
    for k in range(m["nscalar"]):
        like = like + coef[t["scalar_coef"][k]]
    if m["resid_mode"] == 3:
        r = residual(cp, coef)
        n, K = m["n_base"], m["nmode"]
        V = r[n:n + n * K].reshape(K, n).T
        C = t["Sigma"].reshape(n, n) + V @ V.T
        y = sla.solve_triangular(np.linalg.cholesky(C), r[:n], lower=True)
        like = like + 0.5 * float(y @ y)
    elif m["resid_mode"]:
        r = residual(cp, coef)
        npad = m["n_pad"]
        y = sla.solve_triangular(t["Lmat"].reshape(npad, npad), r, lower=True)
        # diagonal-block inverses must be the inverses of the factor's diagonal blocks
        like = like + 0.5 * float(y @ y)
    return like + pr
