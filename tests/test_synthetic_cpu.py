"""BASELINE-shaped synthetic problems: host compilation against the oracle (no GPU)."""
import numpy as np
import pytest

import table_interp as TI
from namespaces import ORACLE, PRODUCT
from cosmoslik_b200 import Slik, synthetic
from cosmoslik_b200.program import CompiledProgram
from oracle import runtime as ort


@pytest.fixture(scope="module")
def planck():
    return synthetic.planck_lite_problem(0)


def test_planck_lite_shape(planck):
    assert planck["n"] == 613 and sorted(w.shape[0] for w in planck["windows"].values()) == [199, 199, 215]
    assert np.all(np.linalg.eigvalsh(planck["cov"]) > 0)


def test_planck_lite_compiles_to_the_oracle(planck):
    oslik = ort.Slik(synthetic.build_script(ORACLE, planck)())
    slik = Slik(synthetic.build_script(PRODUCT, planck)())
    cp = CompiledProgram(list(slik.get_sampled()), slik.params.priors, slik.trace())
    assert cp.names == list(oslik.get_sampled())
    assert cp.meta["n"] == 613 and cp.meta["n_pad"] == 640
    x0 = np.array([oslik.get_start()[k] for k in cp.names])
    sc = np.array([planck["scales"][k] for k in cp.names])
    rs = np.random.RandomState(0)
    for i in range(4):
        x = x0 + sc * rs.standard_normal(len(x0)) * (i > 0)
        ref = oslik.evaluate(*x)[0]
        got = TI.lnl(cp, x)
        assert abs(got - ref) <= 1e-10 * abs(ref), (got, ref)
    # top-hat bins: the executed DMMA blocks are a small fraction of the dense product
    assert cp.flops["bin_exec"] < 0.25 * cp.flops["bin_dense"] * (640 / 613.0) * 2


def test_spt_multifreq_shape():
    prob = synthetic.spt_multifreq_problem(0, nbins=10, lmax=701)
    oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    cp = CompiledProgram(list(slik.get_sampled()), slik.params.priors, slik.trace())
    assert cp.meta["nspec"] == 6 and cp.meta["n"] == 60
    x0 = np.array([oslik.get_start()[k] for k in cp.names])
    ref = oslik.evaluate(*x0)[0]
    assert abs(TI.lnl(cp, x0) - ref) <= 1e-10 * abs(ref)
