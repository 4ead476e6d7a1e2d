"""Host logic of ``output_extra_params`` (no GPU): how the samplers attach derived-quantity columns to
the rows they emit.  The device evaluation itself is covered by tests/test_gpu_xtra_params.py."""
import numpy as np

from cosmoslik_b200.samplers.utils import append_extras, append_extras_of_next
from oracle import samplers as osamp


class FakeExtras(object):
    """stands in for program.ExtrasProgram: two derived quantities of a 3-parameter point"""
    rows = [0, 1]

    @staticmethod
    def f(X):
        X = np.asarray(X, dtype=float).reshape(-1, 3)
        return np.stack([X.sum(1), X[:, 0] * X[:, 2]], axis=1)

    def values_host(self, X):
        return self.f(X)


def test_mh_rows_get_the_extras_of_their_own_position():
    rs = np.random.RandomState(0)
    blocks = [(7, rs.standard_normal((4, 5))), (8, np.zeros((0, 5))), (9, rs.standard_normal((1, 5)))]
    out = append_extras(blocks, FakeExtras(), 3)
    assert [cid for cid, _ in out] == [7, 8, 9]
    for (_, r), (_, e) in zip(blocks, out):
        assert e.shape == (len(r), 7)
        np.testing.assert_array_equal(e[:, :5], r)
        np.testing.assert_array_equal(e[:, 5:], FakeExtras.f(r[:, 2:]))
    empty = append_extras([(0, np.zeros((0, 5)))], FakeExtras(), 3)
    assert empty[0][1].shape == (0, 7)


def test_emcee_file_rows_get_the_extras_of_the_next_position():
    """emcee.py:66-78: lnl and extras written beside x_old are those of the position the walker moved
    to.  The sampler only has the emitted rows and the current positions at collection time; the
    rule 'next row's x, else the current position' must reproduce a direct record of posnext."""
    rs = np.random.RandomState(5)
    k, nd, nsteps, F = 10, 3, 23, 5
    lnprob_fn = lambda q: -0.5 * (np.atleast_2d(q) ** 2).sum(1)
    pos = rs.standard_normal((k, nd))
    lnp = lnprob_fn(pos)
    weight = np.ones(k)
    direct = [[] for _ in range(k)]          # (row, extras at posnext), recorded inside the loop
    pending = [[] for _ in range(k)]         # rows emitted since the last collection
    got = [[] for _ in range(k)]
    for i in range(nsteps):
        u_z, j, u_acc = rs.uniform(size=(2, k // 2)), rs.randint(k // 2, size=(2, k // 2)), rs.uniform(size=(2, k // 2))
        posnext, lnp, _ = osamp.ensemble_step(pos, lnp, lnprob_fn, u_z, j, u_acc, 2.0)
        for w in range(k):
            if (pos[w] == posnext[w]).all() and i != nsteps - 1:
                weight[w] += 1
            else:
                row = np.hstack([-lnp[w], weight[w], pos[w]])
                direct[w].append(np.hstack([row, FakeExtras.f(posnext[w])[0]]))
                pending[w].append(row)
                weight[w] = 1
        pos = posnext
        if (i + 1) % F == 0 or i == nsteps - 1:
            blocks = [(w, np.array(pending[w]).reshape(-1, 2 + nd)) for w in range(k)]
            for w, r in append_extras_of_next(blocks, pos, FakeExtras(), nd):
                got[w].extend(list(r))
            pending = [[] for _ in range(k)]
    assert sum(len(d) for d in direct) > k
    for w in range(k):
        np.testing.assert_array_equal(np.array(got[w]), np.array(direct[w]))


def test_polychord_prior_transform_and_constructor():
    """samplers/polychord.py:43-52: the unit cube maps onto min/max (or range) boxes; a parameter without
    a box is refused; without PyPolyChord ``sample`` raises ImportError instead of falling back"""
    import pytest
    from cosmoslik_b200 import SlikPlugin, param, samplers

    class script(SlikPlugin):
        def __init__(self, bad=False):
            super().__init__()
            self.a = param(start=0, scale=1, min=-2, max=3)
            self.b = param(start=1, scale=1, range=(0.5, 1.5))
            if bad:
                self.c = param(start=0, scale=1, min=0)

        def __call__(self):
            return (self.a ** 2 + self.b ** 2) / 2

    s = script()
    pc = samplers.polychord(s, output_file="chains/pc", nlive=50)
    assert pc.nlive == 50 and pc.num_repeats is None and list(pc.sampled) == ["a", "b"]
    cube = np.array([[0.0, 0.0], [1.0, 1.0], [0.25, 0.5]])
    np.testing.assert_array_equal(pc.prior_transform(cube), [[-2, 0.5], [3, 1.5], [-0.75, 1.0]])
    np.testing.assert_array_equal(pc.prior_transform(cube[2]), [-0.75, 1.0])
    with pytest.raises(ValueError):
        pc.prior_transform(np.zeros((2, 3)))
    with pytest.raises(Exception, match="min/max"):
        samplers.polychord(script(bad=True), output_file="chains/pc")
    with pytest.raises(ImportError):
        pc.sample(lambda *x: None)
