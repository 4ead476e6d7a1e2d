"""Property test of the tracer + bytecode compiler (no GPU): random scalar expressions of the sampled
parameters, written as ordinary Python arithmetic, are traced (cosmoslik_b200.symbolic), compiled
(program._Code) and interpreted by tests/table_interp.py; the result must equal what Python computes
on floats -- bit for bit, since every bytecode op is one correctly rounded IEEE operation in the same
order (x**2 is x*x on both sides; exp / log / sqrt are numpy's on both sides)."""
import math

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import table_interp as TI
from cosmoslik_b200 import symbolic as sym
from cosmoslik_b200.program import compile_extras

NPAR = 4
# the float side uses the transcendental functions the table interpreter uses (numpy's), so that the test is
# about the tracer and the compiler, not about libm-vs-numpy last-bit differences
_np_exp = lambda v: float(np.exp(v))
_np_log = lambda v: float(np.log(v))
_np_sqrt = lambda v: float(np.sqrt(v))


def leaf():
    return st.one_of(st.integers(0, NPAR - 1).map(lambda i: ("par", i)),
                     st.floats(-3, 3, allow_nan=False).filter(lambda v: abs(v) > 1e-3).map(lambda v: ("const", v)))


def tree():
    return st.recursive(leaf(), lambda kids: st.one_of(
        st.tuples(st.sampled_from(["add", "sub", "mul", "div"]), kids, kids),
        st.tuples(st.sampled_from(["neg", "sqr", "exp", "sqrtabs", "logabs"]), kids)), max_leaves=12)


def build(t, x, symbolic):
    """evaluate the tree on Par leaves (symbolic) or on floats, through the SAME Python code"""
    kind = t[0]
    if kind == "par":
        return x[t[1]]
    if kind == "const":
        return t[1]
    if len(t) == 3:
        a, b = build(t[1], x, symbolic), build(t[2], x, symbolic)
        return {"add": lambda: a + b, "sub": lambda: a - b, "mul": lambda: a * b, "div": lambda: a / b}[kind]()
    a = build(t[1], x, symbolic)
    # a subtree without parameters is plain float arithmetic on both sides
    ex, lg, sq = (sym.exp, sym.log, sym.sqrt) if isinstance(a, sym.Expr) else (_np_exp, _np_log, _np_sqrt)
    # a parameter squared is traced as x*x (symbolic.Expr.__pow__); on floats write the product out, since
    # libm's pow(x, 2.0) is not guaranteed to be the correctly rounded product
    sq2 = (lambda v: v ** 2) if isinstance(a, sym.Expr) else (lambda v: v * v)
    if kind == "neg":
        return -a
    if kind == "sqr":
        return sq2(a)
    if kind == "exp":
        return ex(a * 0.125)
    if kind == "sqrtabs":
        return sq(sq2(a) + 1.0)
    return lg(sq2(a) + 0.5)


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(tree(), st.lists(st.floats(-2, 2, allow_nan=False), min_size=NPAR, max_size=NPAR))
def test_traced_bytecode_equals_python_arithmetic(t, xs):
    try:
        ref = build(t, xs, symbolic=False)
    except (ZeroDivisionError, OverflowError, ValueError):
        return
    if not isinstance(ref, float) or not math.isfinite(ref):
        return
    pars = [sym.Par(i, "p%d" % i) for i in range(NPAR)]
    expr = build(t, pars, symbolic=True)
    try:
        cp, rows = compile_extras(["p%d" % i for i in range(NPAR)], [expr])
    except ValueError as e:                   # deeper than the 16-entry device stack: refused, not miscompiled
        assert "too deep" in str(e)
        return
    got = TI.run_bytecode(cp, np.array(xs))[rows[0]]
    assert got == ref or (math.isnan(got) and math.isnan(ref)), (t, xs, got, ref)
