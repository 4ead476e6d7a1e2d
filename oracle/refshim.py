"""Import the UNMODIFIED reference (``/root/reference``) under Python 3.12.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
It is used by ``oracle/make_golden.py`` (run in the build container, where the
reference tree is mounted read-only) to generate the committed fixtures under
``tests/golden/``, and by CPU tests that cross-check the numpy restatement in
``oracle/`` against the real reference when the tree is present.  On the GPU
box ``/root/reference`` does not exist and :func:`available` returns False.

The reference needs three stdlib names that Python >= 3.12 removed
(cosmoslik/cosmoslik.py:4,13 ``imp`` / ``inspect.getargspec``;
cosmoslik/chains.py:643 ``collections.Iterable``).  They are injected here
before the first import; the reference files themselves are never edited.
"""
import collections
import collections.abc
import importlib.util
import inspect
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("COSMOSLIK_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cosmoslik"))


def _install_shims():
    if "imp" not in sys.modules:
        imp = types.ModuleType("imp")
        imp.new_module = types.ModuleType

        def load_source(name, path):
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            return mod

        imp.load_source = load_source
        sys.modules["imp"] = imp
    if not hasattr(inspect, "getargspec"):
        ArgSpec = collections.namedtuple("ArgSpec", "args varargs keywords defaults")

        def getargspec(f):
            fs = inspect.getfullargspec(f)
            return ArgSpec(fs.args, fs.varargs, fs.varkw, fs.defaults)

        inspect.getargspec = getargspec
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable


_ref = None


def load():
    """Return a namespace object holding the reference's public names."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import cosmoslik as pkg  # noqa: F401  (the reference package)
        import cosmoslik.cosmoslik as core
        import cosmoslik.chains as chains
        import cosmoslik_plugins.samplers.metropolis_hastings as mh
        import cosmoslik_plugins.likelihoods.priors as priors
        import importlib
        spt_lowl = importlib.import_module('cosmoslik_plugins.likelihoods.spt_lowl.spt_lowl')
        import cosmoslik_plugins.models.clust_poisson_egfs as cpe
        import cosmoslik_plugins.samplers.utils as sutils
    ns = types.SimpleNamespace(pkg=pkg, core=core, chains=chains, mh=mh, priors=priors,
                               spt_lowl=spt_lowl, clust_poisson_egfs=cpe, sutils=sutils)
    _ref = ns
    return ns
