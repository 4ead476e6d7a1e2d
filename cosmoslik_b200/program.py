"""Compile a traced posterior into the device tables of ``csl_program_t``
(include/cosmoslik_b200.h) and run it through the C ABI.

Host-side, one-time work only: Cholesky of the covariance (the reference's
``cho_factor``, spt_lowl.py:98 / ``cholesky``, mspec_lnl.py:31), padding of the
window matrices to the kernels' tile geometry, and the support analysis that
lets the binning kernel skip all-zero window blocks.
"""
import ctypes as C

import numpy as np
import scipy.linalg as sla

from . import _lib as L
from .symbolic import (BandpowerTerm, Bin, Const, Expr, GaussTerm, LnlSum, LnlTerm, LowRankCovBandpowerTerm, Par,
                       SpectrumBlock, Un)


def _rup(v, m):
    return (v + m - 1) // m * m


class _Code(object):
    """bytecode emitter for scalar expressions (ops: include/cosmoslik_b200.h CSL_OP_*)"""

    def __init__(self):
        self.op, self.arg, self.val = [], [], []
        self.ncoef = 0
        self._cache = {}

    def _emit(self, op, arg=0, val=0.0):
        self.op.append(L.OP[op]); self.arg.append(int(arg)); self.val.append(float(val))

    def _walk(self, e):
        """emit e; return the stack depth it needs"""
        if isinstance(e, Const):
            self._emit("PUSHC", val=e.v); return 1
        if isinstance(e, Par):
            self._emit("PUSHP", arg=e.index); return 1
        if isinstance(e, Un):
            d = self._walk(e.a)
            self._emit({"neg": "NEG", "sqr": "SQR", "exp": "EXP", "log": "LOG", "sqrt": "SQRT"}[e.op]); return d
        if isinstance(e, Bin):
            da = self._walk(e.a); db = self._walk(e.b)
            self._emit({"add": "ADD", "sub": "SUB", "mul": "MUL", "div": "DIV", "pow": "POW"}[e.op])
            return max(da, db + 1)
        raise TypeError("not an expression: %r" % (e,))

    def coef(self, e):
        """row index of the coefficient holding expression e"""
        e = Expr.wrap(e)
        key = repr(e)
        if key in self._cache:
            return self._cache[key]
        depth = self._walk(e)
        if depth > L.STACK:
            raise ValueError("parameter expression too deep for the device stack (%d > %d)" % (depth, L.STACK))
        k = self.ncoef
        self._emit("STORE", arg=k)
        self.ncoef += 1
        self._cache[key] = k
        return k


def tile_major(M):
    """Tile-major copy of a row-major matrix [64 a, 16 b] for the bulk-copy (TMA) pipelines: tile (rb, c) = rows
    64 rb .., columns 16 c .. as 64 rows of 16 + 4 (padding) doubles, tiles ordered rb-major.  One tile is
    10 240 contiguous bytes -- one cp.async.bulk -- and lands in shared memory in the conflict-free row stride
    of 20 doubles the DMMA fragment loads want (include/cosmoslik_b200.h, csl_program_t::MinvT)."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    nr, nc = M.shape
    assert nr % L.BM == 0 and nc % L.KC == 0
    T = np.zeros((nr // L.BM, nc // L.KC, L.BM, L.KC + 4))
    T[:, :, :, :L.KC] = M.reshape(nr // L.BM, L.BM, nc // L.KC, L.KC).transpose(0, 2, 1, 3)
    return T.reshape(-1)


def _support(rows):
    """[first, last+1) of the non-zero columns of a block of window rows"""
    nz = np.flatnonzero(np.any(rows != 0.0, axis=0))
    if nz.size == 0:
        return 0, 0
    return int(nz[0]), int(nz[-1]) + 1


class CompiledProgram(object):
    """Host-side result of compiling a traced posterior: the tables of csl_program_t as
    numpy arrays (``tables``) plus its integer fields (``meta``).  No CUDA needed."""

    _PTR_FIELDS = ("code_op", "code_arg", "code_val", "box_idx", "box_lo", "box_hi", "gauss_idx", "gauss_c",
                   "gauss_w", "scalar_coef", "spec_tab", "tile_tab", "templates", "windows", "data",
                   "direct_coef", "Lmat", "Linv_diag", "Minv", "bin_tab", "Sigma", "grp_tab", "MinvT", "LmatT", "LinvT",
                   "winT")

    def __init__(self, names, priors, lnl_expr, chi2_mode="auto", binning="auto", binning_split=None, derived=None):
        self.derived = derived          # name -> value the script's __call__ left on the tree (priors on derived quantities)
        self.chi2_mode = chi2_mode
        self.binning = binning      # 'auto' | 'dmma' | 'segsum' (see _build_bandpower)
        self.binning_split = binning_split     # None (by tile length) | 1 | 2 | 4: ell split of the DMMA binning tiles
        self.names = list(names)
        self.ndim = len(self.names)
        if not (1 <= self.ndim <= L.MAX_DIM):
            raise ValueError("number of sampled parameters must be in 1..%d" % L.MAX_DIM)
        self.tables = {}
        self.meta = dict(ndim=self.ndim, ncoef=0, ncode=0, nbox=0, ngauss=0, nscalar=0, nspec=0, ntile=0,
                         n=0, n_pad=0, resid_mode=0, nclass=0, nsegclass=0, n_base=0, nmode=0, bin_split=1, has_aberr=0)
        self.cls = np.zeros((L.MAX_CLASS, 4), np.int32)
        self.segcls = np.zeros((L.MAX_CLASS, 4), np.int32)
        self.host = {}
        self.flops = {}
        self._build(priors, lnl_expr)

    def _dev(self, name, arr, dtype):
        a = np.ascontiguousarray(arr, dtype=dtype)
        if a.size == 0:
            a = np.zeros(1, dtype=dtype)
        self.tables[name] = a
        return a

    def _build(self, priors, lnl_expr):
        code = _Code()
        parts = lnl_expr.parts if isinstance(lnl_expr, LnlSum) else [lnl_expr]
        scalars, quad = [], []
        for part in parts:
            if isinstance(part, (BandpowerTerm, GaussTerm, LowRankCovBandpowerTerm)):
                quad.append(part)
            elif isinstance(part, LnlTerm):
                raise TypeError("unsupported likelihood term %r" % (part,))
            else:
                scalars.append(code.coef(Expr.wrap(part)))
        if len(quad) > 1:
            raise NotImplementedError("one quadratic-form likelihood per script in this version "
                                      "(sum them into one covariance, as mspec_lnl does)")
        P = self.meta
        # ---- priors (likelihoods/priors.py:6-30).  A prior may name a sampled parameter or any quantity the script
        # sets on itself in __call__ (the reference evaluates priors again after the call, cosmoslik.py:139-141;
        # `param_ok`, priors.py:25): those are compiled as coefficient rows and referred to by negative indices.
        # A name that is neither is skipped, as the reference skips it -- but not silently.
        idx = {n: i for i, n in enumerate(self.names)}
        self.derived_priors = 0

        def subject(n):
            if n in idx:
                return idx[n]
            v = self.derived(n) if self.derived is not None else None
            if v is None:
                import warnings
                warnings.warn("prior on '%s' ignored: it is neither a sampled parameter nor set by the script "
                              "(likelihoods/priors.py:25 skips it the same way)" % (n,))
                return None
            try:
                e = Expr.wrap(v)
            except TypeError:
                raise TypeError("prior on '%s': the script sets it to %r, not a scalar function of the sampled "
                                "parameters" % (n, type(v))) from None
            self.derived_priors += 1
            return -code.coef(e) - 1

        box = [(subject(n), lo, hi) for n, lo, hi in priors.uniform_priors]
        box = [b for b in box if b[0] is not None]
        gau = [(subject(n), c, w) for n, c, w in priors.gaussian_priors]
        gau = [g for g in gau if g[0] is not None]
        P['nbox'], P['ngauss'] = len(box), len(gau)
        self._dev('box_idx', [b[0] for b in box], np.int32)
        self._dev('box_lo', [b[1] for b in box], np.float64)
        self._dev('box_hi', [b[2] for b in box], np.float64)
        self._dev('gauss_idx', [g[0] for g in gau], np.int32)
        self._dev('gauss_c', [g[1] for g in gau], np.float64)
        self._dev('gauss_w', [g[2] for g in gau], np.float64)
        self.host["box"], self.host["gauss"] = box, gau

        self.n = self.n_pad = 0
        self.gauss_direct = None
        if quad:
            q = quad[0]
            if isinstance(q, BandpowerTerm):
                self._build_bandpower(P, code, q)
            elif isinstance(q, LowRankCovBandpowerTerm):
                self._build_lowrank(P, code, q)
            else:
                self._build_gauss(P, code, q)
        P['n'], P['n_pad'] = self.n, self.n_pad
        P['nscalar'] = len(scalars)
        self._dev('scalar_coef', scalars, np.int32)
        P['ncoef'], P['ncode'] = code.ncoef, len(code.op)
        self._dev('code_op', code.op, np.int32)
        self._dev('code_arg', code.arg, np.int32)
        self._dev('code_val', code.val, np.float64)
        self.host["code"] = (list(code.op), list(code.arg), list(code.val))
        self.ncoef = code.ncoef

    def _factor(self, P, cov):
        """lower Cholesky factor, identity-padded to the 64-row block size, and
        the inverses of its diagonal blocks"""
        n = cov.shape[0]
        npad = _rup(n, L.BM)
        Lf = np.eye(npad)
        Lf[:n, :n] = sla.cholesky(cov, lower=True)
        nblk = npad // L.BM
        inv = np.empty((nblk, L.BM, L.BM))
        eye = np.eye(L.BM)
        for b in range(nblk):
            s = slice(b * L.BM, (b + 1) * L.BM)
            inv[b] = sla.solve_triangular(Lf[s, s], eye, lower=True)
        self.n, self.n_pad = n, npad
        self.host["L"] = Lf[:n, :n]
        self._dev('Lmat', Lf, np.float64)
        self._dev('Linv_diag', inv, np.float64)
        self._dev('LmatT', tile_major(Lf), np.float64)
        self._dev('LinvT', tile_major(inv.reshape(nblk * L.BM, L.BM)), np.float64)
        # explicit inverse of the whole factor (chi2_mode 1): double-precision triangular solve, one step
        # of Newton refinement with the residual in extended precision, and a check against forward
        # substitution; used only if it reproduces the substitution chi^2 to 1e-13
        self.meta['chi2_mode'] = 0
        if self.chi2_mode in ("auto", "inverse") and n <= 8192:
            M = sla.solve_triangular(Lf[:n, :n], np.eye(n), lower=True)
            if n <= 1024:
                Ll, Ml = Lf[:n, :n].astype(np.longdouble), M.astype(np.longdouble)
                M = np.tril((Ml + Ml @ (np.eye(n, dtype=np.longdouble) - Ll @ Ml)).astype(np.float64))
            rs = np.random.RandomState(12345)
            probe = Lf[:n, :n] @ rs.standard_normal((n, 8)) + 0.01 * rs.standard_normal((n, 8))
            y_sub = sla.solve_triangular(Lf[:n, :n], probe, lower=True)
            y_inv = M @ probe
            err = np.max(np.abs((y_inv ** 2).sum(0) - (y_sub ** 2).sum(0)) / (y_sub ** 2).sum(0))
            self.host["inverse_check"] = float(err)
            if err < 1e-13 or self.chi2_mode == "inverse":
                Mp = np.eye(npad); Mp[:n, :n] = M
                self._dev('Minv', Mp, np.float64)
                self._dev('MinvT', tile_major(Mp), np.float64)
                self.meta['chi2_mode'] = 1

    def _chi2_executed_flops(self):
        """DMMA flops per walker the chi^2 kernel actually issues (8-row groups x 16-column chunks)"""
        n, npad = self.n, self.n_pad
        if not self.meta.get('chi2_mode'):
            return npad * npad + npad * L.BM            # blocked solve: full blocks incl. the diagonal inverses
        tot = 0
        kmax = _rup(n, L.KC)
        for r0 in range(0, npad, 8):
            if r0 >= n:
                continue
            kend = min((r0 // L.BM + 1) * L.BM, kmax)
            tot += sum(1 for k0 in range(0, kend, L.KC) if r0 + 7 >= k0) * 2 * 8 * L.KC
        return tot

    def _build_gauss(self, P, code, q):
        n = len(q.values)
        if q.cov.shape != (n, n) or q.mean.shape != (n,):
            raise ValueError("gaussian likelihood: mean/cov shape mismatch")
        rows = [code.coef(v) for v in q.values]
        self._factor(P, q.cov)
        d = np.zeros(self.n_pad); d[:n] = q.mean
        P['resid_mode'] = 2
        self._dev('direct_coef', rows, np.int32)
        self._dev('data', d, np.float64)
        if all(isinstance(v, Par) for v in q.values) and n <= 32 and self.ndim <= 32 and not self.derived_priors:
            # whole-run fused kernel is available (csl_mh_gauss_run)
            self.gauss_direct = dict(perm=np.array([v.index for v in q.values], np.int32),
                                     L=np.ascontiguousarray(self.host["L"]), mean=np.ascontiguousarray(q.mean))

    def _build_lowrank(self, P, code, q):
        """model-dependent covariance C = Sigma + V V^T: the residual vector stacks r and the K mode
        vectors V_k = (W diag(b_k)) . cl, all binned by the ordinary binning kernels (the mode rows are
        just more window rows with zero data); chi^2 by a per-walker Cholesky (csl_cov_chi2)."""
        n, nl = q.windows.shape
        K = q.modes.shape[0]
        if q.sigma.shape != (n, n) or q.data.shape != (n,) or q.modes.shape[1] != nl:
            raise ValueError("low-rank covariance likelihood: shape mismatch")
        if n > 64 or K > 32:
            raise ValueError("low-rank covariance likelihood: at most 64 bandpowers and 32 modes")
        wm = (q.windows[None, :, :] * q.modes[:, None, :]).reshape(K * n, nl)
        ab = dict(aberration=q.aberration, ell0=q.ell0)
        blocks = [SpectrumBlock(q.model, q.windows, q.data, q.cal, cal_on_model=True, **ab)]
        if K:
            blocks.append(SpectrumBlock(q.model, wm, np.zeros(K * n), 1.0, **ab))
        self._build_bandpower(P, code, BandpowerTerm(blocks, None), factor=False)
        P['resid_mode'] = 3
        P['n_base'], P['nmode'] = n, K
        self._dev('Sigma', q.sigma, np.float64)
        self.flops.update(cov_chol=2 * (n * (n + 1) // 2 * K + n ** 3 // 6 + n * n // 2))

    def _index_bound(self, e):
        """A bound on |idx| for a power-law index expression, the same for every walker: it fixes the Taylor
        degree of the segmented-sum kernel's recurrence per group of bins (spectrum record CSL_S_IDXBOUND).
        A parameter with a hard prior box (`min`/`max`/`range`) gives its box -- returned NEGATIVE, meaning
        "hard": a walker outside it has lnl = +inf whatever chi^2 is, so the kernel need not treat it
        specially.  Anything else gives 4 ("soft"): walkers outside a soft bound are still evaluated exactly
        (exp at every column), only slower."""
        if isinstance(e, Const):
            return -abs(float(e.v))
        if isinstance(e, Par):
            los = [lo for i, lo, hi in self.host.get("box", ()) if i == e.index]
            his = [hi for i, lo, hi in self.host.get("box", ()) if i == e.index]
            if los and np.isfinite(max(los)) and np.isfinite(min(his)):
                return -float(max(abs(max(los)), abs(min(his)), 1e-300))
        return 4.0

    def _build_bandpower(self, P, code, q, factor=True):
        nspec = len(q.blocks)
        spec_tab = np.zeros((nspec, L.SPEC_STRIDE), np.int32)
        tiles, segbins, templ, wins, data = [], [], [], [], []
        wint, wint_off, max_tile_chunks = [], 0, 0   # tile-major window tiles of the DMMA-binned spectra (bulk-copy kernel)
        ltabs = []                  # ell-table of each spectrum (for the group records)
        segtabs, segoff = [], 0     # weighted column rows of the segmented-sum bins (appended after the ell-tables)
        toff = 0      # running offset into the ell-table buffer (doubles, kept even)
        woff = 0      # running offset into the window buffer
        row0 = 0
        flops_dense = flops_exec = flops_nnz = builds = seg_ops = 0
        zero_row = None

        def pad_class(k, sizes):
            return next(s_ for s_ in sizes if s_ >= k)

        for s, blk in enumerate(q.blocks):
            nb, nl = blk.windows.shape
            if blk.data.shape != (nb,):
                raise ValueError("bandpower likelihood: data/windows shape mismatch in spectrum %d" % s)
            nlp = _rup(nl, L.KC)
            m = blk.model
            if m.size < nl:
                raise ValueError("bandpower likelihood: model spectrum shorter than the window range")
            terms = []
            if m.const is not None and np.any(m.const[:nl] != 0.0):
                terms.append((Const(1.0), m.const[:nl]))
            terms += [(c, t[:nl]) for c, t in m.terms]
            if len(terms) > L.MAX_TERMS or len(m.plaws) > L.MAX_PLAW:
                raise ValueError("spectrum %d: at most %d template and %d power-law terms" % (s, L.MAX_TERMS, L.MAX_PLAW))
            NT, NP = pad_class(len(terms), (0, 2, 4, 8)), pad_class(len(m.plaws), (0, 1, 2, 4))
            if NT + NP == 0:
                NT = 2
            if (len(terms) < NT or len(m.plaws) < NP) and zero_row is None:
                zero_row = code.coef(Const(0.0))     # padding terms multiply an all-zero coefficient
            stride = NT + 4 * NP
            rec = spec_tab[s]
            rec[L.S_NLPAD], rec[L.S_NTERM], rec[L.S_NPLAW], rec[L.S_LTAB] = nlp, NT, NP, toff
            rec[L.S_CALMODE] = 1 if blk.cal_on_model else 0
            ltab = np.zeros((nlp, stride))
            for t in range(NT):
                if t < len(terms):
                    rec[L.S_TERM_COEF + t] = code.coef(terms[t][0]); ltab[:nl, t] = terms[t][1]
                else:
                    rec[L.S_TERM_COEF + t] = zero_row
            for k in range(NP):
                if k < len(m.plaws):
                    amp, idx_e, base, mul = m.plaws[k]
                    with np.errstate(divide="ignore"):
                        lg = np.log(base[:nl])
                    # base 0: exp(idx * -1e300) gives 0**idx = 0 (idx>0), 1 (idx==0), inf (idx<0) like numpy
                    ltab[:nl, NT + 4 * k] = np.where(np.isneginf(lg), -1e300, lg)
                    ltab[:nl, NT + 4 * k + 1] = mul[:nl]
                    # D[l] = logB[l+1] - logB[l] in extended precision (the kernels' power-law recurrence)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        lgl = np.log(base[:nl].astype(np.longdouble))
                        dl = np.zeros(nlp, dtype=np.longdouble)
                        dl[:nl - 1] = lgl[1:] - lgl[:-1]
                    dl = np.where(np.isfinite(dl), dl, 1e300).astype(np.float64)   # next to a zero base: never "small"
                    ltab[:, NT + 4 * k + 2] = dl
                    rec[L.S_PLAW_AMP + k] = code.coef(amp); rec[L.S_PLAW_IDX + k] = code.coef(idx_e)
                    bnd = self._index_bound(Expr.wrap(idx_e))      # round the magnitude up, keep the sign flag
                    bnd = np.copysign(np.nextafter(np.float32(min(abs(bnd), 3e38)), np.float32(np.inf)), np.float32(bnd))
                    rec[L.S_IDXBOUND + k] = bnd.view(np.int32)
                else:
                    rec[L.S_PLAW_AMP + k] = rec[L.S_PLAW_IDX + k] = zero_row
            templ.append(ltab.ravel()); toff += ltab.size
            ltabs.append(ltab)
            if getattr(blk, "aberration", None):
                # G[l] = kappa / (ln ell_l - ln ell_{l-1}), G[0] = 0 and 0 on the padding (SPTSZ_lowl.py:78-79, 84-89)
                gt = np.zeros(nlp)
                lnx = np.log(np.arange(blk.ell0, blk.ell0 + nl, dtype=float))
                gt[1:nl] = float(blk.aberration) / (lnx[1:] - lnx[:-1])
                rec[L.S_ABERR], rec[L.S_ABTAB] = 1, toff
                templ.append(gt); toff += gt.size
                P['has_aberr'] = 1
            cal = blk.cal
            if isinstance(cal, Expr) and not (isinstance(cal, Const) and cal.v == 1.0):
                rec[L.S_CAL] = code.coef(cal)
            elif not isinstance(cal, Expr) and float(cal) != 1.0:
                rec[L.S_CAL] = code.coef(Const(float(cal)))
            else:
                rec[L.S_CAL] = -1
            nbp = _rup(nb, L.BM)
            wp = np.zeros((nbp, nlp)); wp[:nb, :nl] = blk.windows
            supports = [_support(wp[b:b + 1]) for b in range(nb)]
            covered = sum(hi - lo for lo, hi in supports)
            # rows whose supports barely overlap (top-hat bins): every cl[l] is needed by ~one row, so a
            # per-row segmented sum does the minimum work; dense windows go to the DMMA tile product
            seg = self.binning == "segsum" or (self.binning == "auto" and covered <= 2 * nl)
            if getattr(blk, "aberration", None):
                seg = False          # the aberration factor is built in the DMMA kernel's Cl stage only
            per_elem = len(terms) + 20 * len(m.plaws)
            NTr = len(terms) if len(terms) + NP > 0 else NT      # segmented-sum class: real term count
            if seg:
                for b in range(_rup(nb, L.SEG_GROUP)):
                    rec_b = np.zeros(L.BIN_STRIDE, np.int32)
                    if b < nb:
                        off = woff + b * nlp
                        lo32 = off & 0xFFFFFFFF
                        rec_b[L.B_ROW], rec_b[L.B_SPEC] = row0 + b, s
                        rec_b[L.B_LO], rec_b[L.B_HI] = supports[b]
                        rec_b[L.B_WINOFF_LO] = lo32 - (1 << 32) if lo32 >= (1 << 31) else lo32
                        rec_b[L.B_WINOFF_HI] = off >> 32
                        lo_b, hi_b = supports[b]
                        # the bin's columns as the kernel consumes them: [w T_t.., w M_q.., D_q..] per column
                        # (REAL template terms only: the segmented-sum classes are keyed by len(terms))
                        wrow = wp[b, lo_b:hi_b, None]
                        cols = np.zeros((hi_b - lo_b, (NTr + 2 * NP + 1) // 2 * 2))     # stride padded to 16 bytes
                        cols[:, :NTr] = wrow * ltab[lo_b:hi_b, :NTr]
                        cols[:, NTr:NTr + NP] = wrow * ltab[lo_b:hi_b, NT + 1::4][:, :NP]
                        cols[:, NTr + NP:NTr + 2 * NP] = ltab[lo_b:hi_b, NT + 2::4][:, :NP]
                        rec_b[L.B_TABOFF_LO], rec_b[L.B_TABOFF_HI] = segoff, 0      # relative; rebased below
                        segtabs.append(cols.ravel()); segoff += cols.size
                        for k in range(len(m.plaws)):
                            seg_d = np.abs(ltab[lo_b:max(hi_b - 1, lo_b), NT + 4 * k + 2])
                            dmax = np.float32(min(seg_d.max() if seg_d.size else 0.0, 3e38))
                            dmax = np.nextafter(dmax, np.float32(np.inf))            # round up
                            rec_b[L.B_DMAX + k] = dmax.view(np.int32)
                    else:
                        rec_b[L.B_ROW], rec_b[L.B_SPEC] = -1, s     # group padding: skipped by the kernel
                    segbins.append(((NTr, NP), rec_b))
                flops_exec += 2 * covered
                builds += covered * per_elem
                # FP64 instructions k_bin_segsum issues per column and walker: NT template FMAs, and per power
                # law 1 FMA + the advance (y = idx D, degree-6 Horner, E *= P: 8; degree 5 / 10 when |y| allows /
                # demands); the window weight is folded into the host-built column table
                seg_ops += covered * (NTr + 9 * NP)
            for t0 in (range(0, nb, L.BM) if not seg else ()):
                nr = min(L.BM, nb - t0)
                lo, hi = _support(wp[t0:t0 + nr])
                lo, hi = lo // L.KC * L.KC, _rup(hi, L.KC)
                rec_t = np.zeros(L.TILE_STRIDE, np.int32)
                rec_t[L.T_ROW0], rec_t[L.T_NROWS], rec_t[L.T_SPEC] = row0 + t0, nr, s
                rec_t[L.T_LBEGIN], rec_t[L.T_LEND] = lo, hi
                off = woff + t0 * nlp
                lo32 = off & 0xFFFFFFFF
                rec_t[L.T_WINOFF_LO] = lo32 - (1 << 32) if lo32 >= (1 << 31) else lo32
                rec_t[L.T_WINOFF_HI] = off >> 32
                rec_t[L.T_WINLD] = nlp
                # the same 64 rows tile-major: chunk c (columns 16 c ..) is 1280 contiguous doubles at + c * 1280
                lo32 = wint_off & 0xFFFFFFFF
                rec_t[L.T_WINT_LO] = lo32 - (1 << 32) if lo32 >= (1 << 31) else lo32
                rec_t[L.T_WINT_HI] = wint_off >> 32
                wt = tile_major(wp[t0:t0 + L.BM])
                wint.append(wt); wint_off += wt.size
                max_tile_chunks = max(max_tile_chunks, (hi - lo) // L.KC)
                for g8 in range(L.BM // 8):
                    glo, ghi = _support(wp[t0 + 8 * g8: t0 + 8 * g8 + 8])
                    rec_t[L.T_GLO + g8], rec_t[L.T_GHI + g8] = glo, ghi
                    nch = len([c for c in range(lo, hi, L.KC) if c + L.KC > glo and c < ghi])
                    flops_exec += 2 * 8 * L.KC * nch
                builds += (hi - lo) * per_elem
                tiles.append(((NT, NP), rec_t))
            flops_dense += 2 * nb * nl
            flops_nnz += 2 * int(np.count_nonzero(blk.windows))
            wins.append(wp.ravel()); woff += wp.size
            data.append(blk.data)
            row0 += nb
        n = row0
        if factor:
            if q.cov.shape != (n, n):
                raise ValueError("bandpower likelihood: covariance is %s, data vector has %d rows" % (q.cov.shape, n))
            self._factor(P, q.cov)
        else:
            self.n, self.n_pad = n, _rup(n, L.BM)
        d = np.zeros(self.n_pad); d[:n] = np.concatenate(data)
        # rows n .. n_pad of R must read zero (identity padding of the factor): give them records with an
        # empty window and zero data, so the binning kernels themselves write the zeros
        if self.n_pad > n:
            last_spec = nspec - 1
            key = (int(spec_tab[last_spec][L.S_NTERM]), int(spec_tab[last_spec][L.S_NPLAW]))
            if segbins:
                key = segbins[-1][0]
                last_spec = int(segbins[-1][1][L.B_SPEC])
                for i in range(_rup(self.n_pad - n, L.SEG_GROUP)):
                    rec_b = np.zeros(L.BIN_STRIDE, np.int32)
                    rec_b[L.B_ROW] = n + i if n + i < self.n_pad else -1
                    rec_b[L.B_SPEC] = last_spec
                    segbins.append((key, rec_b))
            elif tiles and tiles[-1][1][L.T_ROW0] + tiles[-1][1][L.T_NROWS] == n and \
                    tiles[-1][1][L.T_NROWS] + (self.n_pad - n) <= L.BM:
                tiles[-1][1][L.T_NROWS] += self.n_pad - n      # the last tile's zero-padded window rows
            else:
                rec_t = np.zeros(L.TILE_STRIDE, np.int32)
                rec_t[L.T_ROW0], rec_t[L.T_NROWS], rec_t[L.T_SPEC] = n, self.n_pad - n, last_spec
                tiles.append((key, rec_t))

        # records sorted by kernel class (stable): one launch per class <NT, NP>
        if segtabs:
            base = toff + (toff & 1)                         # 16-byte aligned start of the weighted rows
            templ.append(np.zeros(base - toff))
            templ += segtabs
            for _, r in segbins:
                if r[L.B_ROW] >= 0 and r[L.B_HI] > r[L.B_LO]:
                    off = base + int(r[L.B_TABOFF_LO])
                    lo32 = off & 0xFFFFFFFF
                    r[L.B_TABOFF_LO] = lo32 - (1 << 32) if lo32 >= (1 << 31) else lo32
                    r[L.B_TABOFF_HI] = off >> 32
        # Within a class the launch order is free (every record names its output row), so the groups that
        # become one CTA are sorted longest first: the hardware hands CTAs out in index order, and short
        # CTAs at the end keep the tail of the launch short (plik-lite bins widen 6x from ell 30 to 2500).
        def by_class(records, width, group=1, work=None):
            classes = sorted(set(k for k, _ in records))
            if len(classes) > L.MAX_CLASS:
                raise ValueError("too many distinct spectrum classes")
            ordered, cls = [], np.zeros((L.MAX_CLASS, 4), np.int32)
            for ci, k in enumerate(classes):
                mine = [r for kk, r in records if kk == k]
                if work is not None:
                    assert len(mine) % group == 0
                    chunks = [mine[i:i + group] for i in range(0, len(mine), group)]
                    chunks.sort(key=lambda c: -sum(work(r) for r in c))      # stable
                    mine = [r for c in chunks for r in c]
                cls[ci] = (k[0], k[1], len(ordered), len(mine))
                ordered += mine
            return (np.array(ordered) if ordered else np.zeros((1, width), np.int32)), cls, len(classes), len(ordered)

        tile_arr, self.cls, nclass, ntile = by_class(tiles, L.TILE_STRIDE)
        bin_arr, self.segcls, nsegclass, _ = by_class(segbins, L.BIN_STRIDE, L.SEG_GROUP,
                                                      lambda r: int(r[L.B_HI]) - int(r[L.B_LO]))
        ordered = tile_arr
        P['resid_mode'] = 1
        P['nspec'], P['ntile'], P['nclass'], P['nsegclass'] = nspec, ntile, nclass, nsegclass
        self._dev('spec_tab', spec_tab, np.int32)
        self._dev('tile_tab', tile_arr, np.int32)
        self._dev('bin_tab', bin_arr, np.int32)
        # group records: what one CTA of k_bin_segsum needs about its SEG_GROUP bins, in one 512-byte load
        ngrp = len(bin_arr) // L.SEG_GROUP if segbins else 0
        grp = np.zeros((max(ngrp, 1), L.GRP_STRIDE), np.int32)
        for g in range(ngrp):
            recs = bin_arr[g * L.SEG_GROUP:(g + 1) * L.SEG_GROUP]
            rec_g = grp[g]
            dbl = rec_g.view(np.float64)                      # same buffer, 8-byte view (word 2 k <-> double k)
            s_idx = int(recs[0][L.B_SPEC])
            NTs, NPs = int(spec_tab[s_idx][L.S_NTERM]), int(spec_tab[s_idx][L.S_NPLAW])
            pre, prev_hi, base = 0, -1, -1
            gmax = np.zeros(L.MAX_PLAW, np.float32)
            for b, r in enumerate(recs):
                live = r[L.B_ROW] >= 0
                ncol = max(int(r[L.B_HI]) - int(r[L.B_LO]), 0) if live else 0
                rec_g[L.G_PRE + b] = pre
                rec_g[L.G_LO + b] = int(r[L.B_LO]) if live else 0
                rec_g[L.G_ROW + b] = int(r[L.B_ROW]) if live else -1
                rec_g[L.G_CONT + b] = 1 if (ncol > 0 and int(r[L.B_LO]) == prev_hi) else 0
                if ncol > 0:
                    if base < 0:
                        base = (int(r[L.B_TABOFF_HI]) << 32) | (int(r[L.B_TABOFF_LO]) & 0xFFFFFFFF)
                    prev_hi = int(r[L.B_HI])
                    gmax = np.maximum(gmax, r[L.B_DMAX:L.B_DMAX + L.MAX_PLAW].view(np.float32))
                dbl[L.G_DATA // 2 + b] = d[int(r[L.B_ROW])] if live else 0.0
                lo_b = max(int(r[L.B_LO]), 0) if live else 0
                for k in range(NPs):
                    dbl[L.G_LOGB // 2 + b * L.MAX_PLAW + k] = ltabs[s_idx][lo_b, NTs + 4 * k]
                pre += ncol
            rec_g[L.G_PRE + L.SEG_GROUP] = pre
            bounds = np.abs(spec_tab[s_idx][L.S_IDXBOUND:L.S_IDXBOUND + L.MAX_PLAW].view(np.float32)).astype(np.float64)
            ym = float(np.max(bounds[:NPs] * gmax[:NPs].astype(np.float64))) if NPs else 0.0
            rec_g[L.G_MODE] = 0 if ym <= 0.0035 else 1 if ym <= 0.012 else 2 if ym <= 0.1 else 3
            base = max(base, 0)
            lo32 = base & 0xFFFFFFFF
            rec_g[L.G_BASE_LO] = lo32 - (1 << 32) if lo32 >= (1 << 31) else lo32
            rec_g[L.G_BASE_HI] = base >> 32
            rec_g[L.G_SPEC] = s_idx
        self._dev('grp_tab', grp, np.int32)
        self.host["grp_tab"] = grp
        self._dev('templates', np.concatenate(templ), np.float64)
        self._dev('windows', np.concatenate(wins), np.float64)
        if wint:
            self._dev('winT', np.concatenate(wint), np.float64)
        # ell split of the DMMA binning tiles (cluster size of k_bin_residual_ws): from the longest tile's chunk
        # count only -- a property of the program, so results never depend on the launch size
        # (long tiles only: >= 96 chunks, i.e. dense windows over >= 1536 ell: measured on 6 x 50 bins x 3251 ell,
        # 8192 walkers per launch: 1.18 / 1.09 / 1.04 ms with 1 / 2 / 4 parts; 47 x 3251 ell, 4096 chains: 0.31 / 0.17 / 0.09 ms)
        self.meta['bin_split'] = 4 if max_tile_chunks >= 96 else 1
        if self.binning_split is not None:
            self.meta['bin_split'] = int(self.binning_split)
        self._dev('data', d, np.float64)
        self.host["spec_tab"], self.host["tile_tab"], self.host["bin_tab"] = spec_tab, tile_arr, bin_arr
        # work per evaluation: dense windows, non-zero window entries, DMMA blocks that
        # are actually executed, foreground-build FP64 ops (~1 per template term, ~20 per power law), TRSM
        self.flops = dict(bin_dense=flops_dense, bin_nnz=flops_nnz, bin_exec=flops_exec, build_ops=2 * builds, seg_flops=2 * seg_ops,
                          trsm=self._chi2_executed_flops() if factor else 0, trsm_alg=n * n if factor else 0)


class DeviceProgram(CompiledProgram):
    """A compiled posterior resident on one CUDA device."""

    def __init__(self, names, priors, lnl_expr, device=None, chi2_mode="auto", binning="auto", binning_split=None,
                 derived=None):
        torch = L.require_cuda()
        self.lib = L.lib()
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        super().__init__(names, priors, lnl_expr, chi2_mode=chi2_mode, binning=binning, binning_split=binning_split,
                         derived=derived)
        self._ws = {}
        self.dev = {k: torch.from_numpy(v).to(self.device) for k, v in self.tables.items()}
        P = L.Program()
        for k, v in self.meta.items():
            setattr(P, k, int(v))
        for k in self._PTR_FIELDS:
            setattr(P, k, self.dev[k].data_ptr() if k in self.dev else None)
        for ci in range(L.MAX_CLASS):
            for j in range(4):
                P.cls[ci][j] = int(self.cls[ci, j])
                P.segcls[ci][j] = int(self.segcls[ci, j])
        self.struct = P
        if self.gauss_direct is not None:
            self.gauss_direct = {k: torch.from_numpy(v).to(self.device) for k, v in self.gauss_direct.items()}

    # ------------------------------------------------------------------ run
    def workspace(self, W):
        Wp = _rup(W, L.BN)
        ws = self._ws.get(Wp)
        if ws is None:
            t = self.torch
            kw = dict(dtype=t.float64, device=self.device)
            ws = dict(coef=t.empty((max(self.ncoef, 1), Wp), **kw), prior=t.empty(Wp, **kw),
                      R=t.empty((max(self.n_pad, 1), Wp), **kw) if self.n_pad else None,
                      chi2=t.empty(Wp, **kw) if self.n_pad else None,
                      part=t.empty((self.n_pad // 32, Wp), **kw) if self.n_pad else None,   # one partial per 32 rows
                      x=t.empty((Wp, self.ndim), **kw), lnl=t.empty(Wp, **kw))
            self._ws[Wp] = ws
        return ws

    def lnl(self, x, out=None, timers=None):
        """-ln posterior of every row of the device tensor x [W, ndim] -> tensor [W]
        (batched Slik.evaluate, cosmoslik.py:109-141).  With ``timers`` (a dict) the four stages
        are launched one by one between CUDA events appended to timers[stage]."""
        t = self.torch
        if not (isinstance(x, t.Tensor) and x.is_cuda and x.dtype == t.float64 and x.dim() == 2
                and x.shape[1] == self.ndim and x.is_contiguous()):
            raise ValueError("x must be a contiguous float64 CUDA tensor [W, %d]" % self.ndim)
        W = x.shape[0]
        if out is None:
            out = t.empty(W, dtype=t.float64, device=self.device)
        if W == 0:                      # an empty batch has an empty answer; nothing is launched
            return out
        ws = self.workspace(W)
        st = L.stream_ptr()
        if timers is None:
            L.check(self.lib.csl_lnl(C.byref(self.struct), W, x.data_ptr(), out.data_ptr(), ws["coef"].data_ptr(),
                                     ws["prior"].data_ptr(), L.ptr(ws["R"]), L.ptr(ws["chi2"]), L.ptr(ws["part"]), st),
                    "csl_lnl")
            return out
        Wp = _rup(W, L.BN)
        P = C.byref(self.struct)

        def stage(name, fn):
            a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            a.record()
            L.check(fn(), name)
            b.record()
            timers.setdefault(name, []).append((a, b))

        stage("prepare", lambda: self.lib.csl_eval_prepare(P, W, x.data_ptr(), ws["coef"].data_ptr(), Wp,
                                                           ws["prior"].data_ptr(), st))
        chi2 = None
        if self.meta["resid_mode"] in (1, 3):
            stage("bin_residual", lambda: self.lib.csl_bin_residual(P, Wp, ws["coef"].data_ptr(), Wp,
                                                                    ws["R"].data_ptr(), Wp, st))
        elif self.meta["resid_mode"] == 2:
            stage("direct_residual", lambda: self.lib.csl_direct_residual(P, W, ws["coef"].data_ptr(), Wp,
                                                                          ws["R"].data_ptr(), Wp, st))
        if self.meta["resid_mode"] == 3:
            stage("cov_chi2", lambda: self.lib.csl_cov_chi2(P, W, ws["R"].data_ptr(), Wp, ws["chi2"].data_ptr(), st))
            chi2 = ws["chi2"].data_ptr()
        elif self.meta["resid_mode"] and self.meta.get("chi2_mode"):
            # explicit inverse: the tile product leaves partial sums, the combine kernel adds them (as csl_lnl does)
            stage("chi2", lambda: self.lib.csl_chi2_parts(P, Wp, ws["R"].data_ptr(), Wp, ws["part"].data_ptr(), st))
            stage("combine", lambda: self.lib.csl_combine_parts(P, W, ws["coef"].data_ptr(), Wp, ws["prior"].data_ptr(),
                                                                ws["part"].data_ptr(), Wp, out.data_ptr(), st))
            return out
        elif self.meta["resid_mode"]:
            stage("chi2", lambda: self.lib.csl_chi2(P, Wp, ws["R"].data_ptr(), Wp, ws["chi2"].data_ptr(),
                                                    ws["part"].data_ptr(), st))
            chi2 = ws["chi2"].data_ptr()
        stage("combine", lambda: self.lib.csl_combine(P, W, ws["coef"].data_ptr(), Wp, ws["prior"].data_ptr(), chi2,
                                                      out.data_ptr(), st))
        return out

    def lnl_host(self, x_host, out_host=None):
        """End-to-end evaluation with HOST numpy buffers through csl_lnl_host (H2D, kernels, D2H)."""
        x_host = np.ascontiguousarray(x_host, dtype=np.float64)
        if x_host.ndim != 2 or x_host.shape[1] != self.ndim:
            raise ValueError("x must be an array [W, %d]" % self.ndim)
        W = x_host.shape[0]
        if out_host is None:
            out_host = np.empty(W)
        if W == 0:
            return out_host
        ws = self.workspace(W)
        L.check(self.lib.csl_lnl_host(C.byref(self.struct), W, x_host.ctypes.data, out_host.ctypes.data,
                                      ws["x"].data_ptr(), ws["lnl"].data_ptr(), ws["coef"].data_ptr(),
                                      ws["prior"].data_ptr(), L.ptr(ws["R"]), L.ptr(ws["chi2"]), L.ptr(ws["part"]),
                                      L.stream_ptr()), "csl_lnl_host")
        return out_host


class _NoPriors(object):
    """prior tables of a program that only evaluates parameter expressions"""
    uniform_priors = ()
    gaussian_priors = ()


def compile_extras(names, exprs):
    """Host-side tables of the derived-quantity program (no CUDA needed): every expression becomes
    one coefficient row of the bytecode; returns (CompiledProgram, coefficient row of each expression)."""
    cp = CompiledProgram(list(names), _NoPriors, LnlSum([Expr.wrap(e) for e in exprs]))
    return cp, [int(i) for i in cp.tables["scalar_coef"][:len(exprs)]]


class ExtrasProgram(DeviceProgram):
    """``output_extra_params`` on the device: the derived quantities a script stores on itself during
    ``__call__`` (the reference writes ``s.extra[k]`` next to every sample, metropolis_hastings.py:228)
    are scalar expressions of the sampled parameters; they are compiled to the same coefficient
    bytecode as a likelihood's amplitudes and evaluated for any batch of points by the priors +
    bytecode kernel (csl_eval_prepare) -- no per-point Python."""

    def __init__(self, names, exprs, device=None):
        super().__init__(list(names), _NoPriors, LnlSum([Expr.wrap(e) for e in exprs]), device=device)
        self.rows = [int(i) for i in self.tables["scalar_coef"][:len(exprs)]]

    def values(self, x):
        """device tensor x [N, ndim] -> device tensor [N, nextra]"""
        t = self.torch
        if not (isinstance(x, t.Tensor) and x.is_cuda and x.dtype == t.float64 and x.dim() == 2
                and x.shape[1] == self.ndim and x.is_contiguous()):
            raise ValueError("x must be a contiguous float64 CUDA tensor [N, %d]" % self.ndim)
        N = x.shape[0]
        if N == 0 or not self.rows:
            return t.empty((N, len(self.rows)), dtype=t.float64, device=self.device)
        ws = self.workspace(N)
        Wp = _rup(N, L.BN)
        L.check(self.lib.csl_eval_prepare(C.byref(self.struct), N, x.data_ptr(), ws["coef"].data_ptr(), Wp,
                                          ws["prior"].data_ptr(), L.stream_ptr()), "csl_eval_prepare")
        return ws["coef"][self.rows, :N].t().contiguous()

    def values_host(self, X):
        """numpy [N, ndim] -> numpy [N, nextra]"""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.ndim)
        return self.values(self.torch.from_numpy(X).to(self.device)).cpu().numpy()
