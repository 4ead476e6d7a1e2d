"""Script-facing runtime: the SlikPlugin / SlikDict / param / Slik / run_chain API
of the reference (cosmoslik/cosmoslik.py), re-implemented around a batched
device program instead of a per-call Python evaluation.

What is kept call-compatible (reference file:line):
  SlikDict          cosmoslik.py:177-229   dotted keys, deepcopy, find_sampled (sorted names)
  param / shortcut  cosmoslik.py:232-238, 428-445
  sample            cosmoslik.py:241-250
  Slik              cosmoslik.py:80-145    get_sampled / get_start / evaluate / sample
  run_chain         cosmoslik.py:254-296   -> Chain / Chains
  SlikPlugin/Sampler cosmoslik.py:300-319
  lsum / lsumk      cosmoslik.py:385-424
  arguments         cosmoslik.py:492-525
  load_script       cosmoslik.py:27-77     (importlib instead of the removed `imp`)

What changes: ``Slik.evaluate`` does not run the script's Python per point.  The
script's ``__call__`` is executed ONCE with symbolic parameters
(symbolic.py), compiled (program.py) and evaluated on the GPU, for one point or
for a whole batch (``Slik.evaluate_batch``).
"""
import argparse
import copy
import hashlib
import importlib.util
import inspect
import sys
from collections import OrderedDict

import numpy as np

from . import symbolic as sym

__all__ = ["load_script", "Slik", "SlikFunction", "SlikDict", "SlikPlugin", "SlikSampler", "param",
           "param_shortcut", "lsum", "lsumk", "run_chain", "SlikMain", "arguments", "sample"]


class SlikDict(dict):
    """A dict whose items are also attributes and whose string keys may be dotted
    paths into nested SlikDicts."""

    def __init__(self, *a, **kw):
        dict.__init__(self)
        for k, v in dict(*a, **kw).items():
            dict.__setitem__(self, k, v)

    # attributes <-> items
    def __getattr__(self, name):
        try:
            return dict.__getitem__(self, name)
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        dict.__setitem__(self, name, value)

    def __delattr__(self, name):
        try:
            dict.__delitem__(self, name)
        except KeyError:
            raise AttributeError(name) from None

    # dotted keys
    @staticmethod
    def _path(key):
        if not isinstance(key, str):
            raise ValueError("Parameter key must be string, not %s" % type(key))
        return key.split(".")

    def __getitem__(self, key):
        node = self
        for part in self._path(key):
            node = getattr(node, part)
        return node

    def __setitem__(self, key, value):
        *head, last = self._path(key)
        node = self
        for part in head:
            node = getattr(node, part)
        setattr(node, last, value)

    def __contains__(self, key):
        try:
            self[key]
        except AttributeError:
            return False
        return True

    def get(self, key, default=None):
        node = self
        for part in self._path(key):
            node = getattr(node, part, default)
        return node

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            dict.__setitem__(self, k, v)

    def deepcopy(self):
        """Nested SlikDicts are copied, every other value is shared by reference."""
        new = copy.copy(self)
        for k, v in dict.items(self):
            if isinstance(v, SlikDict):
                dict.__setitem__(new, k, v.deepcopy())
        return new

    def __copy__(self):
        new = self.__class__.__new__(self.__class__)
        dict.__init__(new)
        for k, v in dict.items(self):
            dict.__setitem__(new, k, v)
        return new

    def find_sampled(self):
        """All `param` leaves, keyed by dotted name, sorted by that name -- this
        order IS the parameter-vector order everywhere (cosmoslik.py:229)."""
        found = []

        def walk(node, prefix):
            for k, v in dict.items(node):
                if isinstance(v, SlikDict):
                    walk(v, prefix + (k,))
                elif isinstance(v, param):
                    found.append((".".join(prefix + (k,)), v))

        walk(self, ())
        return OrderedDict(sorted(found, key=lambda kv: kv[0]))


class param(object):
    """Marks a parameter to be sampled; keyword attributes such as start, scale,
    min, max, range, uniform_prior, gaussian_prior."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):
        return "param(%s)" % ", ".join("%s=%r" % kv for kv in self.__dict__.items())


def param_shortcut(*names):
    """``param_shortcut('start','scale')(1, 2)`` == ``param(start=1, scale=2)``."""

    class shortcut(param):
        def __init__(self, *vals, **kw):
            kw.update(zip(names, vals))
            param.__init__(self, **kw)

    return shortcut


class sample(object):
    """What a sampler's ``sample`` generator yields."""

    def __init__(self, x, lnl, weight=1, extra=None):
        self.x, self.lnl, self.weight, self.extra = x, lnl, weight, extra


class SlikPlugin(SlikDict):
    def __call__(self, *a, **kw):
        raise NotImplementedError()


class SlikSampler(SlikDict):
    def sample(self, lnl):
        raise NotImplementedError()


def SlikMain(cls):
    if not (isinstance(cls, type) and issubclass(cls, SlikPlugin)):
        raise ValueError("SlikMain used on a class which is not a SlikPlugin.")
    cls._slik_main = True
    return cls


def SlikFunction(f):
    f._slik_function = True
    return f


def arguments(exclude=None, exclude_self=True, include_kwargs=True):
    """The calling function's arguments as a dict (plugins use it as
    ``super().__init__(**arguments())``)."""
    frame = inspect.currentframe().f_back
    info = inspect.getargvalues(frame)
    out = {k: info.locals[k] for k in info.args}
    if info.varargs:
        out[info.varargs] = info.locals[info.varargs]
    if info.keywords:
        extra = info.locals[info.keywords]
        if include_kwargs:
            out.update(extra)
    if exclude_self:
        out.pop("self", None)
    for k in (exclude or ()):
        out.pop(k, None)
    out.pop("__class__", None)
    return out


def lsum(*fs):
    """Sum of the callables' results, stopping at the first +inf."""
    total = 0
    for f in fs:
        total = total + f()
        if not isinstance(total, (sym.LnlTerm, sym.Expr)) and total == np.inf:
            break
    return total


def lsumk(lnls, items):
    """Like lsum for (key, callable) pairs; stores each result under lnls[key]
    (nan for the ones skipped after an inf)."""
    total = 0
    for k, f in items:
        if not isinstance(total, (sym.LnlTerm, sym.Expr)) and total == np.inf:
            lnls[k] = np.nan
        else:
            lnls[k] = f()
            total = total + lnls[k]
    return total


class Slik(object):
    """Binds a script (a SlikPlugin tree with `param` leaves and a `.sampler`) to
    its compiled device program."""

    def __init__(self, params):
        self.params = params
        self._sampled = params.find_sampled()
        self.sampler = params.sampler
        del params.sampler
        if "priors" not in params:
            from .likelihoods.priors import priors
            params.priors = priors(params)
        self._program = None
        self._program_options = {}

    def get_sampled(self):
        return self._sampled

    def get_start(self):
        return {k: v.start for k, v in self._sampled.items()}

    # -- tracing / compilation ------------------------------------------------
    def trace(self):
        """Run the script's __call__ once with symbolic parameters."""
        return self._trace()[1]

    def _trace(self):
        """(tree after the symbolic call, traced result): the tree is what the reference hands back
        as the second element of ``Slik.evaluate`` (cosmoslik.py:141) -- whatever the script stored
        on itself during ``__call__`` is there as an expression of the sampled parameters."""
        tree = self.params.deepcopy()
        for i, name in enumerate(self._sampled):
            tree[name] = sym.Par(i, name)
        try:
            out = tree()
        except TypeError as e:
            raise TypeError("this script's __call__ cannot be batched: %s" % e) from e
        if isinstance(out, (int, float, np.floating)):
            out = sym.Const(float(out))
        if not isinstance(out, (sym.Expr, sym.LnlTerm)):
            raise TypeError("script __call__ returned %r; expected a likelihood term or parameter expression" % type(out))
        return tree, out

    def extra_exprs(self, keys):
        """The derived quantities ``params[k]`` a sampler's ``output_extra_params`` names
        (metropolis_hastings.py:228: ``s.extra[k]`` with ``extra`` the evaluated tree), as scalar
        expressions of the sampled parameters.  KeyError for a name the script never sets, as the
        reference's lookup; TypeError for a value that is not a scalar function of the parameters."""
        tree = self._trace()[0]
        out = []
        for k in keys:
            try:
                v = tree[k]
            except (KeyError, AttributeError):
                raise KeyError("output_extra_params: the script has no '%s' after __call__" % (k,)) from None
            try:
                out.append(sym.Expr.wrap(v))
            except TypeError:
                raise TypeError("output_extra_params: '%s' is %r, not a scalar function of the sampled "
                                "parameters" % (k, type(v))) from None
        return out

    def extras_program(self, keys):
        """Device program evaluating ``output_extra_params`` for a batch of points (cached per key list)."""
        keys = tuple(keys)
        cached = getattr(self, "_extras", None)
        if cached is None or cached[0] != keys:
            from .program import ExtrasProgram
            cached = (keys, ExtrasProgram(list(self._sampled), self.extra_exprs(keys)))
            self._extras = cached
        return cached[1]

    def program(self, **options):
        """The compiled device program (cached).  Options (see program.CompiledProgram):
        chi2_mode = 'auto' | 'trsm' | 'inverse';  binning = 'auto' | 'dmma' | 'segsum'."""
        if self._program is None or (options and options != self._program_options):
            from .program import DeviceProgram
            self._program_options = dict(options)
            tree, out = self._trace()
            self._program = DeviceProgram(list(self._sampled), self.params.priors, out, derived=_tree_lookup(tree),
                                          **options)
        return self._program

    # -- evaluation -----------------------------------------------------------
    def _bind(self, args, kwargs):
        names = list(self._sampled)
        if args and kwargs:
            raise ValueError("Expected either *args or **kwargs but not both.")
        if not args:
            missing = [k for k in names if k not in kwargs]
            if missing:
                raise ValueError("Missing the following parameters: %s" % missing)
            return kwargs
        if len(args) != len(names):
            raise ValueError("Expected %i parameters, only got %i." % (len(names), len(args)))
        return dict(zip(names, args))

    def evaluate(self, *args, **kwargs):
        """(lnl, params) at one point -- a batch of one on the device."""
        bound = self._bind(args, kwargs)
        params = self.params.deepcopy()
        for k, v in bound.items():
            params[k] = v
        x = np.array([[float(bound[k]) for k in self._sampled]])
        lnl = float(self.program().lnl_host(x)[0])
        # the reference hands back the EVALUATED tree (cosmoslik.py:141): scalar quantities the script stored
        # on itself during __call__ are filled in from the trace (none after a hard-prior hit, :135-137)
        derived = self.derived_names()
        if derived and lnl != np.inf:
            for k, v in zip(derived, self.extras_program(derived).values_host(x)[0]):
                params[k] = float(v)
        return lnl, params

    def derived_names(self):
        """Dotted names of the scalar quantities the script's __call__ stores on the tree as functions of
        the sampled parameters (found once, from the symbolic trace)."""
        names = getattr(self, "_derived_names", None)
        if names is None:
            names = []

            def walk(node, prefix):
                for k, v in dict.items(node):
                    if isinstance(v, SlikDict):
                        walk(v, prefix + (k,))
                    elif isinstance(v, sym.Expr) and not isinstance(v, (sym.Par, sym.Const)):
                        names.append(".".join(prefix + (k,)))

            walk(self._trace()[0], ())
            self._derived_names = names = sorted(names)
        return names

    def evaluate_batch(self, X):
        """lnl for every row of X [W, ndim].  numpy in -> numpy out (host buffers,
        copies included); CUDA tensor in -> CUDA tensor out (no copies)."""
        prog = self.program()
        if isinstance(X, np.ndarray):
            return prog.lnl_host(X)
        return prog.lnl(X)

    def sample(self):
        return self.sampler.sample(self.evaluate)


def _tree_lookup(tree):
    """name -> what the script left under that dotted name after its symbolic __call__, or None"""
    def look(name):
        try:
            v = tree[name]
        except (AttributeError, KeyError, ValueError):
            return None
        return None if isinstance(v, param) else v
    return look


def run_chain(main, nchains=1, pool=None, args=None, kwargs=None):
    """Run a script and collect its samples.

    nchains == 1 (and a single-chain sampler): returns a Chain with keys
    'weight', 'lnl', then the sampled names (cosmoslik.py:286-289).  Otherwise a
    Chains, one Chain per chain.  ``nchains > 1`` does not fork processes as the
    reference does (cosmoslik.py:290-296): the chains become rows of the
    sampler's device state (``sampler.nchains``)."""
    from .chains import Chain, Chains
    args = [] if args is None else args
    kwargs = {} if kwargs is None else kwargs
    if isinstance(main, tuple):
        main, args, kwargs = main
    if isinstance(main, str):
        _, main = load_script(main)
    script = main(*args, **kwargs)
    if nchains != 1:
        if "nchains" not in script.sampler:
            raise ValueError("run_chain(nchains=%d): sampler %s has no nchains option" % (nchains, type(script.sampler).__name__))
        script.sampler.nchains = nchains
    slik = Slik(script)
    names = list(slik.get_sampled())
    per_chain = OrderedDict()
    for s in slik.sample():
        per_chain.setdefault(getattr(s, "chain", 0), []).append(np.hstack([s.weight, s.lnl, s.x]))
    keys = ["weight", "lnl"] + names

    def to_chain(rows):
        arr = np.array(rows).reshape(len(rows), len(keys))
        return Chain(dict(zip(keys, arr.T)))

    multi = getattr(slik.sampler, "nchains", 1) != 1 and not getattr(slik.sampler, "pool_chains", False)
    if not multi:
        return to_chain([r for rows in per_chain.values() for r in rows])
    return Chains([to_chain(rows) for rows in per_chain.values()])


_script_namespace = "cosmoslik_b200.scripts"


def load_script(script):
    """(argparse parser, main plugin class) for a script file or plugin class.
    Positional CLI args for __init__ parameters without defaults, ``--x`` /
    ``--no-x`` for booleans, ``nargs`` for lists."""
    if isinstance(script, str):
        with open(script) as f:
            digest = hashlib.md5(f.read().encode()).hexdigest()
        modname = "%s.%s" % (_script_namespace, digest)
        if modname in sys.modules:
            mod = sys.modules[modname]
        else:
            spec = importlib.util.spec_from_file_location(modname, script)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[modname] = mod
            spec.loader.exec_module(mod)
        plugins = [v for v in vars(mod).values()
                   if isinstance(v, type) and issubclass(v, SlikPlugin) and v is not SlikPlugin
                   and v.__module__ == modname]
        mains = [v for v in plugins if getattr(v, "_slik_main", False)]
        if len(mains) > 1:
            raise ValueError("Multiple SlikPlugins in '%s' are marked with @SlikMain." % script)
        if len(mains) == 1:
            main = mains[0]
        elif len(plugins) > 1:
            raise ValueError("Multiple SlikPlugins were found in '%s' but none are marked with @SlikMain." % script)
        elif len(plugins) == 1:
            main = plugins[0]
        else:
            raise ValueError("No SlikPlugins were found in '%s'" % script)
    elif isinstance(script, type) and issubclass(script, SlikPlugin):
        main = script
    else:
        raise ValueError("`script` argument should be filename or SlikPlugin class")

    spec = inspect.getfullargspec(main.__init__)
    names = spec.args[1:]
    defaults = list(spec.defaults or [])
    missing = object()
    defaults = [missing] * (len(names) - len(defaults)) + defaults
    parser = argparse.ArgumentParser(prog="cosmoslik %s" % (script if isinstance(script, str) else main.__name__))
    for name, d in zip(names, defaults):
        if d is missing:
            parser.add_argument(name)
        elif d is True:
            parser.add_argument("--no-" + name, dest=name, action="store_false")
        elif d is False:
            parser.add_argument("--" + name, action="store_true")
        elif isinstance(d, list):
            parser.add_argument("--" + name, default=d, nargs=("+" if d else "*"),
                                type=(type(d[0]) if d else None), help="default: " + str(d))
        else:
            parser.add_argument("--" + name, default=d, type=(type(d) if d is not None else None),
                                help="default: " + str(d))
    return parser, main
