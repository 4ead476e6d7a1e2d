"""Numerical analysis behind the power-law recurrence of the binning kernels (no GPU): the kernels advance
E[l] = exp(idx logB[l]) along consecutive columns by E[l+1] = E[l] P_deg(idx D[l]) with a Taylor polynomial
whose degree is fixed from a host bound on |idx D| (csrc/binning.cu: degree 5 for |y| <= 0.0035, 6 for
|y| <= 0.012, 10 for |y| <= 0.1, exact exp beyond).  Checked here, with the coefficient table read from the
CUDA source: the stated truncation bounds, in extended precision; and the error the float64 recurrence
accumulates over the longest runs the workloads have, against exp evaluated directly."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def coefficients():
    src = open(os.path.join(ROOT, "cosmoslik_b200", "csrc", "binning.cu")).read()
    body = re.search(r"c_inv_fact\[11\]\s*=\s*\{([^}]*)\}", src).group(1)
    c = [float(v) for v in body.replace("\n", " ").split(",")]
    assert len(c) == 11
    return c


def horner(c, deg, y, dtype):
    y = np.asarray(y, dtype=dtype)
    p = np.full_like(y, dtype(c[deg]))
    for k in range(deg - 1, -1, -1):
        p = p * y + dtype(c[k])
    return p


def test_coefficient_table_is_inverse_factorials():
    for k, v in enumerate(coefficients()):
        assert abs(v * math.factorial(k) - 1.0) < 2.3e-16      # 1/k! to the last bit or the one before it


@pytest.mark.parametrize("deg,bound,stated", [(5, 0.0035, 3e-18), (6, 0.012, 7e-18), (10, 0.1, 3e-19)])
def test_truncation_bounds_stated_in_the_kernel(deg, bound, stated):
    c = coefficients()
    y = np.linspace(-bound, bound, 4001).astype(np.longdouble)
    err = np.abs(horner(c, deg, y, np.longdouble) - np.exp(y)) / np.exp(y)
    # longdouble (64-bit mantissa) resolves 5e-20; the coefficients themselves are doubles (rounding 1e-16 * y^k/k!
    # enters at second order in y), so compare against the stated bound plus that representation floor
    assert float(err.max()) < stated + 2e-19 + 1.2e-16 * bound ** 3


@pytest.mark.parametrize("lo,ncol,idx,deg", [(30, 33, 2.5, 10), (30, 33, -2.5, 10), (500, 33, 1.3, 6), (2000, 33, 3.0, 5),
                                             (50, 16, 0.8, 10), (3000, 8, -1.7, 5)])
def test_recurrence_error_over_a_bin(lo, ncol, idx, deg):
    """float64 recurrence over the longest top-hat bin of config 4 (33 columns) / a 16-column DMMA chunk:
    relative error against exp(idx log(l / 3000)) stays a few ulp -- four orders of magnitude inside the
    1e-10 budget on lnL"""
    c = coefficients()
    ell = np.arange(lo, lo + ncol + 1)
    logB = np.log(ell.astype(np.longdouble) / 3000)
    D = (logB[1:] - logB[:-1]).astype(np.float64)          # the host builds D in extended precision
    bound = {5: 0.0035, 6: 0.012, 10: 0.1}[deg]
    assert abs(idx) * np.abs(D).max() <= bound, "test case outside the mode's range"
    E = float(np.exp(np.float64(idx) * np.float64(logB[0])))
    worst = 0.0
    for k in range(ncol):
        E = E * float(horner(c, deg, np.float64(idx) * D[k], np.float64))
        exact = np.exp(np.longdouble(idx) * logB[k + 1])
        worst = max(worst, abs(float((np.longdouble(E) - exact) / exact)))
    assert worst < (ncol + 4) * 2.3e-16
