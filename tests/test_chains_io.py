"""Chain containers and the chain-file format (CPU): files written by this package load with the
reference's own ``load_chain`` (when the reference tree is mounted), the reference's file loads here,
and the statistics agree with the oracle."""
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN, golden
from cosmoslik_b200 import Chain, Chains, load_chain, get_covariance, combine_covs
from cosmoslik_b200.chains import ChainWriter, gelman_rubin
from cosmoslik_b200.samplers.metropolis_hastings import _FileBlocks
from oracle import refshim, runtime as ort, samplers as osamp


def test_reference_chain_file_loads_here():
    g = golden("mh_quick.npz")
    c = load_chain(os.path.join(GOLDEN, "ref_mh.chain"))
    assert isinstance(c, Chains) and len(c) == 1
    assert list(c[0].keys()) == ["lnl", "weight", "a", "b"]
    np.testing.assert_array_equal(c[0]["lnl"], g["file_lnl"])
    np.testing.assert_array_equal(c[0]["weight"], g["file_weight"])
    np.testing.assert_array_equal(c[0]["a"], g["file_a"])


def test_single_process_block_structure_matches_the_reference(tmp_path):
    """nchains == 1: records of mpi_comm_freq rows minus the last one (metropolis_hastings.py:237)"""
    g = golden("mh_quick.npz")
    rows = np.column_stack([g["lnl"], g["weight"], g["x"]])
    path = str(tmp_path / "single.chain")
    fb = _FileBlocks(path, ["lnl", "weight", "a", "b"], 100, single=True)
    for i in range(0, len(rows), 37):          # rows arrive in arbitrary pieces
        fb.add(0, rows[i:i + 37])
    fb.finish()
    with open(path, "rb") as f, open(os.path.join(GOLDEN, "ref_mh.chain"), "rb") as fr:
        mine, ref = [], []
        for fh, out in ((f, mine), (fr, ref)):
            while True:
                try:
                    out.append(pickle.load(fh))
                except EOFError:
                    break
    assert list(mine[0]) == list(ref[0])                       # header: ndarray of names
    assert len(mine) == len(ref)
    for (i1, r1), (i2, r2) in zip(mine[1:], ref[1:]):
        assert i1 == i2 and r1.dtype == r2.dtype
        np.testing.assert_array_equal(r1.view(float), r2.view(float))


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("header", ["ndarray", "list"])
def test_files_written_here_load_with_the_reference_loader(tmp_path, header):
    R = refshim.load()
    rs = np.random.RandomState(0)
    path = str(tmp_path / "multi.chain")
    w = ChainWriter(path, ["lnl", "weight", "p.a", "p.b", "q"], header)
    data = {}
    for block in range(3):
        for cid in (2, 0, 1):
            r = rs.standard_normal((4 + cid, 5))
            data.setdefault(cid, []).append(r)
            w.write(cid, r)
        w.flush()
    w.close()
    ref = R.chains.load_chain(path)
    mine = load_chain(path)
    assert len(ref) == len(mine) == 3
    for cid in range(3):
        full = np.concatenate(data[cid])
        for k, name in enumerate(["lnl", "weight", "p.a", "p.b", "q"]):
            np.testing.assert_array_equal(mine[cid][name], full[:, k])
        # the reference loader iterates a set of ids; match by content
        assert any(np.array_equal(rc["q"], full[:, 4]) for rc in ref)


def test_chain_statistics():
    rs = np.random.RandomState(1)
    x = rs.standard_normal((500, 3)) @ rs.standard_normal((3, 3))
    w = rs.randint(1, 8, 500).astype(float)
    c = Chain({"lnl": rs.uniform(size=500), "weight": w, "a": x[:, 0], "b": x[:, 1], "c": x[:, 2]})
    names = ["a", "b", "c"]
    np.testing.assert_allclose(c.cov(names), ort.get_covariance(x, w), rtol=1e-13)
    np.testing.assert_allclose(c.mean(names), np.average(x, axis=0, weights=w))
    assert c.length() == 500 and c.length(unique=False) == w.sum()
    assert abs(c.acceptance() - 1 / w.mean()) < 1e-15
    b = c.burnin(100)
    first = np.searchsorted(np.cumsum(w), 100)
    assert b.length() == 500 - first
    t = c.thin(10)
    assert abs(t["weight"].sum() - np.ceil(w.sum() / 10)) <= 1
    assert c.best_fit()["lnl"] == c["lnl"].min()
    j = Chains([c, c]).join()
    assert j.length() == 1000
    g = c.add_gauss_prior("a", 0.0, 2.0)
    np.testing.assert_allclose(g["lnl"] - c["lnl"], c["a"] ** 2 / 2 / 4)
    assert "acceptance" in str(c)


def test_combine_covs_and_get_covariance_match_the_reference_outputs():
    g = golden("cov.npz")
    np.testing.assert_allclose(get_covariance(g["dat"], g["wts"]), g["cov_w"], rtol=1e-14)
    names, cov = combine_covs({"a": 2.0, "b": 0.5}, (["c", "a"], np.array([[4.0, 0.3], [0.3, 9.0]])))
    idx = [names.index(k) for k in ("a", "b", "c")]
    np.testing.assert_array_equal(cov[np.ix_(idx, idx)], g["init_cov"])


def test_gelman_rubin_host():
    rs = np.random.RandomState(2)
    chains = Chains([Chain({"lnl": np.zeros(2000), "weight": rs.randint(1, 5, 2000).astype(float),
                            "x": rs.standard_normal(2000), "y": rs.standard_normal(2000) * 3}) for _ in range(6)])
    got = gelman_rubin(chains, ["x", "y"])
    mom = [osamp.chain_moments(c.matrix(["x", "y"]), c["weight"]) for c in chains]
    ref = osamp.gelman_rubin([m[0] for m in mom], [m[1] for m in mom], [m[2] for m in mom])
    np.testing.assert_allclose(got, ref, rtol=1e-13)
    assert set(chains.gelman_rubin()) == {"x", "y"}


# ---- text formats and row-wise post-processing -----------------------------
def _toy_chain(seed=0, n=400):
    rs = np.random.RandomState(seed)
    return Chain(lnl=rs.chisquare(3, n), weight=rs.randint(1, 6, n).astype(float), a=rs.standard_normal(n),
                 b=rs.standard_normal(n) ** 2)


def test_savechain_text_round_trip_and_cosmomc_layouts(tmp_path):
    from cosmoslik_b200 import load_cosmomc_chain
    c = _toy_chain()
    p = str(tmp_path / "one.txt")
    c.savechain(p)
    back = load_cosmomc_chain(p)
    assert list(back.keys()) == ["lnl", "weight", "a", "b"]
    for k in back:
        np.testing.assert_allclose(back[k], c[k], rtol=1e-15)
    # CosmoMC layout: root_1.txt, root_2.txt + root.paramnames (columns: weight, lnl, parameters)
    for i in (1, 2):
        cc = _toy_chain(i)
        np.savetxt(str(tmp_path / ("run_%d.txt" % i)), np.column_stack([cc["weight"], cc["lnl"], cc["a"], cc["b"]]))
    (tmp_path / "run.paramnames").write_text("a   a_{label}\nb   b\n")
    cs = load_cosmomc_chain(str(tmp_path / "run"))
    assert isinstance(cs, Chains) and len(cs) == 2
    np.testing.assert_allclose(cs[0]["a"], _toy_chain(1)["a"], rtol=1e-15)
    np.testing.assert_allclose(cs[1]["weight"], _toy_chain(2)["weight"])
    # one file per parameter (WMAP style): the last column is the value
    d = tmp_path / "wmap"; d.mkdir()
    np.savetxt(str(d / "omegabh2"), np.column_stack([np.arange(5), np.linspace(0.02, 0.03, 5)]))
    w = load_cosmomc_chain(str(d))
    np.testing.assert_allclose(w["omegabh2"], np.linspace(0.02, 0.03, 5))
    with pytest.raises(IOError):
        load_cosmomc_chain(str(tmp_path / "nothing_here"))


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_text_chain_written_here_loads_with_the_reference(tmp_path):
    R = refshim.load()
    c = _toy_chain(3)
    p = str(tmp_path / "x.txt")
    c.savechain(p); c.savecov(str(tmp_path / "x.cov"))
    import cosmoslik.chains as rch
    rc = rch.load_cosmomc_chain(p)
    for k in ("lnl", "weight", "a", "b"):
        np.testing.assert_allclose(rc[k], c[k], rtol=1e-15)
    rc2 = rch.Chain(dict(c))
    np.testing.assert_allclose(np.loadtxt(str(tmp_path / "x.cov")), rc2.cov(["a", "b"]), rtol=1e-12)
    # the same statistics through both containers
    np.testing.assert_allclose(c.confbound("b"), rc2.confbound("b"), rtol=1e-12)
    np.testing.assert_allclose(c.confbound("a", level=0.68, bound="lower"), rc2.confbound("a", level=0.68, bound="lower"),
                               rtol=1e-12)
    f = lambda a, **_: np.exp(-a ** 2 / 2)
    np.testing.assert_allclose(c.reweighted(f)["weight"], rc2.reweighted(f)["weight"], rtol=1e-15)
    g = lambda a, b, **_: {"s": a + b} if a > 0 else {}
    np.testing.assert_array_equal(c.postprocd(g)["s"], rc2.postprocd(g)["s"])


def test_reweighted_and_postprocd_without_the_reference():
    c = _toy_chain(4)
    r = c.reweighted(lambda a, **_: np.exp(-a ** 2 / 2))
    np.testing.assert_allclose(r["weight"], c["weight"] * np.exp(-c["a"] ** 2 / 2), rtol=1e-15)
    assert r is not c and np.array_equal(r["a"], c["a"])
    pp = c.postprocd(lambda a, b, **_: {"s": a + b} if a > 0 else {})
    assert np.all(np.isnan(pp["s"][c["a"] <= 0])) and np.allclose(pp["s"][c["a"] > 0], (c["a"] + c["b"])[c["a"] > 0])
    lo, hi = c.confbound("b", level=0.95)
    assert lo == c["b"].min() and c["b"].min() < hi < c["b"].max()          # positively skewed -> upper bound


def test_load_chain_repack(tmp_path):
    path = str(tmp_path / "blocks.chain")
    w = ChainWriter(path, ["lnl", "weight", "a"], "ndarray")
    rs = np.random.RandomState(1)
    for cid in (0, 1, 0, 1):
        w.write(cid, rs.standard_normal((5, 3)))
    # still open for writing: repack must leave the file alone
    w.flush()
    before = open(path, "rb").read()
    c1 = load_chain(path, repack=True)
    assert open(path, "rb").read() == before
    w.close()
    c2 = load_chain(path, repack=True)
    assert open(path, "rb").read() != before
    c3 = load_chain(path)                                   # now a single pickled Chains
    assert isinstance(c3, Chains) and len(c3) == 2
    for x, y, z in zip(c1, c2, c3):
        for k in x:
            np.testing.assert_array_equal(x[k], y[k]); np.testing.assert_array_equal(x[k], z[k])


@pytest.mark.parametrize("nrows", [0, 1, 49, 50, 51, 99, 100, 150, 237])
def test_single_process_blocks_for_every_row_count(tmp_path, nrows):
    """metropolis_hastings.py:219-244 at the boundaries: an exact multiple of mpi_comm_freq ends with an
    empty record, every record drops its last row; rows reach the writer in arbitrary pieces"""
    freq, nd = 50, 3
    rs = np.random.RandomState(nrows)
    rows = rs.standard_normal((nrows, 2 + nd))
    orows = [osamp.Row(r[2:], r[0], r[1]) for r in rows]
    _, blocks = osamp.mh_single_process_blocks(orows, nd, freq)
    path = str(tmp_path / "b.chain")
    fb = _FileBlocks(path, ["lnl", "weight", "a", "b", "c"], freq, single=True)
    at = 0
    while at < nrows:
        m = int(rs.randint(1, 70))
        fb.add(0, rows[at:at + m])
        at += m
    fb.finish()
    recs = []
    with open(path, "rb") as f:
        pickle.load(f)
        while True:
            try:
                recs.append(pickle.load(f))
            except EOFError:
                break
    assert len(recs) == len(blocks)
    for (cid, rec), blk in zip(recs, blocks):
        assert cid == 0
        np.testing.assert_array_equal(rec.view(float).reshape(-1, 2 + nd), blk)


def test_chain_writer_with_typed_extra_columns(tmp_path):
    """output_extra_params entries (name, dtype) (metropolis_hastings.py:46-56): the column is stored in that dtype
    and comes back through load_chain (and through the reference's loader where it is mounted)"""
    path = str(tmp_path / "typed.chain")
    w = ChainWriter(path, ["lnl", "weight", "a", "n", "flag"], "ndarray", dtypes=["f8", "f8", "f8", "i8", "f4"])
    rows = np.array([[1.5, 2, 0.25, 7, 0.5], [2.5, 1, 0.75, -3, 1.5], [0.5, 4, 1.25, 12, 2.5]])
    w.write(0, rows[:2]); w.write(0, rows[2:]); w.close()
    c = load_chain(path)[0]
    assert c["n"].dtype == np.int64 and c["flag"].dtype == np.float32 and c["a"].dtype == np.float64
    np.testing.assert_array_equal(c["n"], [7, -3, 12])
    np.testing.assert_array_equal(c["weight"], [2, 1, 4])
    if refshim.available():
        R = refshim.load()
        rc = R.chains.load_chain(path)
        rc = rc[0] if isinstance(rc, list) else rc
        np.testing.assert_array_equal(rc["n"], [7, -3, 12])
        np.testing.assert_array_equal(rc["flag"], np.float32([0.5, 1.5, 2.5]))
