"""The multi-spectrum cases of tests/golden/mspec.npz (oracle/make_golden_mspec.py ran them through the unmodified
reference over the ``mspec`` stand-in): problem + plugin namespace + points + the reference's lnl, for any namespace
(oracle or product) that synthetic.build_script accepts."""
import types

import numpy as np

from conftest import golden
from cosmoslik_b200 import synthetic
from oracle.make_golden_mspec import cleaning_case

CASES = ("spt", "subset", "clean", "planck")


def case(name, ns):
    """-> (problem, namespace to build the script with, X [n, ndim], reference lnl [n])"""
    g = golden("mspec.npz")
    prob = synthetic.spt_multifreq_problem(3, nbins=10, lmax=601)
    if name == "spt":
        return prob, ns, np.array(g["spt_x"]), np.array(g["spt_lnl"])
    if name == "subset":
        keep = sorted(prob["use"])[::2]
        assert ["%s %s" % k for k in keep] == list(g["spt_subset_keys"])
        sub = dict(prob)
        sub["use"] = {k: prob["use"][k] for k in keep}
        sub["windows"] = {k: prob["windows"][k] for k in keep}
        return sub, ns, np.array(g["spt_x"]), np.array(g["spt_subset_lnl"])
    if name == "clean":
        new, use, windows, resgal = cleaning_case(prob)
        np.testing.assert_array_equal(np.array([resgal[k] for k in sorted(resgal)]), g["clean_resgal"])
        cl = dict(prob)
        cl["use"], cl["windows"] = use, windows
        base = ns.mspec_lnl
        ns2 = types.SimpleNamespace(**vars(ns))
        ns2.mspec_lnl = lambda u, w, s, c: base(u, w, s, c, cleaning=new, resgal=resgal)
        return cl, ns2, np.array(g["spt_x"]), np.array(g["clean_lnl"])
    if name == "planck":
        return synthetic.planck_lite_problem(4), ns, np.array(g["planck_x"]), np.array(g["planck_lnl"])
    raise KeyError(name)
