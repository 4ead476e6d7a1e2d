"""Symbolic scalars and spectra.

A CosmoSlik script's ``__call__`` is scalar Python (cosmoslik/cosmoslik.py:300-313
in the reference).  To batch it over thousands of walkers without changing the
script, the sampled parameters are bound as :class:`Par` leaves instead of
floats; ordinary arithmetic on them builds small expression trees.  Plugins of
this package understand those trees and return likelihood terms; ``Slik``
compiles the result into a device program (``program.py``).  No numerics happen
here.
"""
import math
import numbers

import numpy as np


class Expr(object):
    """Scalar function of the sampled parameters."""
    __array_ufunc__ = None       # make ndarray <op> Expr defer to our reflected methods
    __array_priority__ = 1000
    __slots__ = ()

    # -- construction helpers
    @staticmethod
    def wrap(v):
        if isinstance(v, Expr):
            return v
        if isinstance(v, (numbers.Real, np.floating, np.integer)) or (isinstance(v, np.ndarray) and v.ndim == 0):
            return Const(float(v))
        raise TypeError("cannot use %r in a parameter expression" % (type(v),))

    def _bin(self, op, other, swap=False):
        if isinstance(other, np.ndarray) and other.ndim > 0:
            return NotImplemented
        if isinstance(other, (Spectrum, LnlTerm)):
            return NotImplemented
        try:
            o = Expr.wrap(other)
        except TypeError:
            return NotImplemented
        a, b = (o, self) if swap else (self, o)
        if isinstance(a, Const) and isinstance(b, Const):
            return Const(_fold(op, a.v, b.v))
        if op == "mul":                      # 1*x is x exactly
            if isinstance(a, Const) and a.v == 1.0: return b
            if isinstance(b, Const) and b.v == 1.0: return a
        return Bin(op, a, b)

    def __add__(self, o): return self._vec(o, "add") if _is_vec(o) else self._bin("add", o)
    def __radd__(self, o): return self._vec(o, "radd") if _is_vec(o) else self._bin("add", o, True)
    def __sub__(self, o): return self._vec(o, "sub") if _is_vec(o) else self._bin("sub", o)
    def __rsub__(self, o): return self._vec(o, "rsub") if _is_vec(o) else self._bin("sub", o, True)
    def __mul__(self, o): return self._vec(o, "mul") if _is_vec(o) else self._bin("mul", o)
    def __rmul__(self, o): return self._vec(o, "mul") if _is_vec(o) else self._bin("mul", o, True)
    def __truediv__(self, o): return self._vec(o, "div") if _is_vec(o) else self._bin("div", o)
    def __rtruediv__(self, o): return self._vec(o, "rdiv") if _is_vec(o) else self._bin("div", o, True)
    def __neg__(self): return Un("neg", self)
    def __pos__(self): return self

    def __pow__(self, o):
        if isinstance(o, numbers.Real) and float(o) == 2.0:
            return Un("sqr", self)          # x**2 is x*x exactly, as in Python
        return self._bin("pow", o)

    def __rpow__(self, o):
        if _is_vec(o):                       # array ** parameter -> power-law spectrum
            base = np.asarray(o, dtype=float)
            return Spectrum(plaws=[(Const(1.0), self, base, np.ones_like(base))])
        return self._bin("pow", o, True)

    def _vec(self, arr, how):
        arr = np.asarray(arr, dtype=float)
        s = Spectrum(const=arr)
        if how == "mul": return s * self
        if how == "add" or how == "radd": return s + self
        if how == "sub": return (s * -1.0) + self
        if how == "rsub": return s - self
        if how == "div": return Spectrum(const=1.0 / arr) * self
        raise TypeError("array / parameter is not a linear spectrum")

    # -- guard against silent numeric use
    def __float__(self):
        raise TypeError("a sampled parameter was used where a number is required; inside a batched script "
                        "parameters are symbolic (cosmoslik_b200.symbolic)")
    __bool__ = None
    def __lt__(self, o): raise TypeError("comparisons on sampled parameters are not traceable")
    __le__ = __gt__ = __ge__ = __lt__
    __hash__ = object.__hash__


def _is_vec(o):
    return isinstance(o, np.ndarray) and o.ndim > 0


def _fold(op, a, b):
    if op == "add": return a + b
    if op == "sub": return a - b
    if op == "mul": return a * b
    if op == "div": return a / b
    if op == "pow": return a ** b
    raise ValueError(op)


class Const(Expr):
    __slots__ = ("v",)
    def __init__(self, v): self.v = float(v)
    def __float__(self): return self.v
    def __repr__(self): return "Const(%r)" % self.v


class Par(Expr):
    """Sampled parameter number `index` (order = sorted dotted names)."""
    __slots__ = ("index", "name")
    def __init__(self, index, name): self.index, self.name = index, name
    def __repr__(self): return "Par(%d, %r)" % (self.index, self.name)


class Bin(Expr):
    __slots__ = ("op", "a", "b")
    def __init__(self, op, a, b): self.op, self.a, self.b = op, a, b
    def __repr__(self): return "(%r %s %r)" % (self.a, self.op, self.b)


class Un(Expr):
    __slots__ = ("op", "a")
    def __init__(self, op, a): self.op, self.a = op, a
    def __repr__(self): return "%s(%r)" % (self.op, self.a)


def exp(x):
    return Un("exp", x) if isinstance(x, Expr) and not isinstance(x, Const) else math.exp(float(x))


def log(x):
    return Un("log", x) if isinstance(x, Expr) and not isinstance(x, Const) else math.log(float(x))


def sqrt(x):
    return Un("sqrt", x) if isinstance(x, Expr) and not isinstance(x, Const) else math.sqrt(float(x))


class Spectrum(object):
    """A spectrum that is linear in template arrays with symbolic coefficients:

        cl[l] = const[l] + sum_t coef_t * T_t[l] + sum_p amp_p * M_p[l] * B_p[l] ** idx_p

    This is the form of every foreground model on the path: spt_lowl.py:206-207
    (three templates), models/clust_poisson_egfs.py:14-17 (Poisson l^2 and a
    power law with a sampled index)."""
    __array_ufunc__ = None
    __array_priority__ = 2000

    def __init__(self, const=None, terms=None, plaws=None):
        self.const = None if const is None else np.asarray(const, dtype=float)
        self.terms = list(terms or [])
        self.plaws = list(plaws or [])

    # -- shape protocol used by the likelihood plugins
    def _arrays(self):
        out = [] if self.const is None else [self.const]
        out += [t for _, t in self.terms]
        out += [b for _, _, b, _ in self.plaws]
        return out

    @property
    def size(self):
        return min(a.size for a in self._arrays())

    def __len__(self):
        return self.size

    def __getitem__(self, sl):
        if not isinstance(sl, slice):
            raise TypeError("symbolic spectra support slicing only")
        return Spectrum(None if self.const is None else self.const[sl],
                        [(c, t[sl]) for c, t in self.terms],
                        [(a, i, b[sl], m[sl]) for a, i, b, m in self.plaws])

    # -- algebra
    def __add__(self, o):
        if isinstance(o, Spectrum):
            n = min(self.size, o.size) if (self._arrays() and o._arrays()) else None
            a, b = (self[:n], o[:n]) if n is not None and self.size != o.size else (self, o)
            if a.const is None: c = b.const
            elif b.const is None: c = a.const
            else: c = a.const + b.const
            return Spectrum(c, a.terms + b.terms, a.plaws + b.plaws)
        if isinstance(o, np.ndarray) and o.ndim > 0:
            return self + Spectrum(const=o)
        if isinstance(o, numbers.Real):
            if o == 0:
                return self
            return self + Spectrum(const=np.full(self.size, float(o)))
        if isinstance(o, Expr):
            return self + Spectrum(terms=[(o, np.ones(self.size))])
        return NotImplemented
    __radd__ = __add__

    def __neg__(self): return self * -1.0
    def __sub__(self, o): return self + (o * -1.0 if isinstance(o, (Spectrum, Expr)) else -np.asarray(o, dtype=float))
    def __rsub__(self, o): return (self * -1.0) + o

    def __mul__(self, o):
        if isinstance(o, (numbers.Real, Expr)):
            k = Expr.wrap(o)
            terms = [(c * k, t) for c, t in self.terms]
            if self.const is not None:
                if isinstance(k, Const):
                    return Spectrum(self.const * k.v, terms, [(a * k, i, b, m) for a, i, b, m in self.plaws])
                terms = [(k, self.const)] + terms
            return Spectrum(None, terms, [(a * k, i, b, m) for a, i, b, m in self.plaws])
        if isinstance(o, np.ndarray) and o.ndim > 0:
            o = np.asarray(o, dtype=float)
            return Spectrum(None if self.const is None else self.const * o,
                            [(c, t * o) for c, t in self.terms],
                            [(a, i, b, m * o) for a, i, b, m in self.plaws])
        return NotImplemented
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, numbers.Real): return self * (1.0 / o)
        if isinstance(o, Expr): return self * (1.0 / o)
        if isinstance(o, np.ndarray): return self * (1.0 / np.asarray(o, dtype=float))
        return NotImplemented


def as_spectrum(v, n=None):
    """Coerce a model output (array, number or Spectrum) to a Spectrum."""
    if isinstance(v, Spectrum):
        return v
    if isinstance(v, np.ndarray):
        return Spectrum(const=v)
    if isinstance(v, numbers.Real):
        return Spectrum(const=np.full(n, float(v)))
    raise TypeError("cannot interpret %r as a spectrum" % (type(v),))


class LnlTerm(object):
    """Base of the likelihood terms a plugin's ``__call__`` returns; terms and
    scalar expressions add up to an :class:`LnlSum`."""
    __array_ufunc__ = None

    def __add__(self, o): return LnlSum([self]) + o
    def __radd__(self, o): return LnlSum([self]).__radd__(o)


class LnlSum(LnlTerm):
    def __init__(self, parts): self.parts = list(parts)

    def _coerce(self, o):
        if isinstance(o, LnlSum): return o.parts
        if isinstance(o, LnlTerm): return [o]
        if isinstance(o, (Expr, numbers.Real)): return [Expr.wrap(o)]
        return None

    def __add__(self, o):
        p = self._coerce(o)
        return NotImplemented if p is None else LnlSum(self.parts + p)

    def __radd__(self, o):
        p = self._coerce(o)
        return NotImplemented if p is None else LnlSum(p + self.parts)


class SpectrumBlock(object):
    """One binned spectrum of a bandpower likelihood: windows [nbins, nl] applied
    to model columns [lmin, lmin+nl), data [nbins], calibration multiplying the data."""
    def __init__(self, model, windows, data, cal=1.0, cal_on_model=False, aberration=None, ell0=0):
        self.model, self.windows, self.data, self.cal = model, np.asarray(windows, dtype=float), \
            np.asarray(data, dtype=float), cal
        self.cal_on_model = cal_on_model      # False: r = cal*d - b (spt_lowl, mspec); True: r = d - cal*b (SPTSZ)
        # aberration = kappa: the model is multiplied by 1 + kappa * dln(cl)/dln(l) (forward differences over the
        # columns, 0 at the first) before it is binned (SPTSZ_lowl.py:78-79, 84-89); ell0 = ell of column 0
        self.aberration, self.ell0 = aberration, int(ell0)


class BandpowerTerm(LnlTerm):
    """1/2 r^T C^-1 r with r = cal*d - W.cl over one or more spectra and one dense
    covariance (spt_lowl.py:112-133; mspec_lnl.py:63-73)."""
    def __init__(self, blocks, cov):
        self.blocks, self.cov = list(blocks), np.asarray(cov, dtype=float)


class LowRankCovBandpowerTerm(LnlTerm):
    """1/2 r^T (Sigma + V V^T)^-1 r  with  r = d - cal * W.cl  and  V_k = (W diag(b_k)).cl:
    a bandpower likelihood whose covariance depends on the model through rank-one beam modes
    (SPTSZ_lowl.py:62-81, beam_cov = W (B o dl dl^T) W^T with B = sum_k b_k b_k^T)."""
    def __init__(self, model, windows, data, cal, sigma, modes, aberration=None, ell0=0):
        self.aberration, self.ell0 = aberration, int(ell0)
        self.model, self.cal = model, cal
        self.windows, self.data = np.asarray(windows, dtype=float), np.asarray(data, dtype=float)
        self.sigma, self.modes = np.asarray(sigma, dtype=float), np.asarray(modes, dtype=float)


class GaussTerm(LnlTerm):
    """1/2 (v-mu)^T C^-1 (v-mu) for a vector of parameter expressions."""
    def __init__(self, values, mean, cov):
        self.values = [Expr.wrap(v) for v in values]
        self.mean, self.cov = np.asarray(mean, dtype=float), np.asarray(cov, dtype=float)
