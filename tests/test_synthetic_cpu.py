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


def test_segmented_sum_tables(planck):
    """Host tables of k_bin_segsum: the weighted column table is window weight x ell-table row, the bins of a
    group are contiguous in it, groups are sorted longest first within a class, and index bounds from prior
    boxes are stored negative (hard)."""
    from cosmoslik_b200 import _lib as L
    slik = Slik(synthetic.build_script(PRODUCT, planck)())
    cp = CompiledProgram(list(slik.get_sampled()), slik.params.priors, slik.trace(), binning="segsum")
    assert cp.meta["nsegclass"] == 2 and cp.meta["nclass"] == 0
    bt, st = cp.host["bin_tab"], cp.host["spec_tab"]
    tpl, win = cp.tables["templates"], cp.tables["windows"]
    for c in range(cp.meta["nsegclass"]):
        NT, NP, b0, nb = (int(v) for v in cp.segcls[c])          # NT: REAL template terms of the class
        RS = (NT + 2 * NP + 1) // 2 * 2                            # staged-row stride, padded to 16 bytes
        assert nb % L.SEG_GROUP == 0
        widths = []
        for g0 in range(b0, b0 + nb, L.SEG_GROUP):
            recs = bt[g0:g0 + L.SEG_GROUP]
            live = [r for r in recs if r[L.B_ROW] >= 0 and r[L.B_HI] > r[L.B_LO]]
            widths.append(sum(int(r[L.B_HI] - r[L.B_LO]) for r in live))
            off = None
            for r in live:
                t = (int(r[L.B_TABOFF_HI]) << 32) | (int(r[L.B_TABOFF_LO]) & 0xFFFFFFFF)
                assert off is None or t == off                       # contiguous, in record order
                lo, hi = int(r[L.B_LO]), int(r[L.B_HI])
                woff = (int(r[L.B_WINOFF_HI]) << 32) | (int(r[L.B_WINOFF_LO]) & 0xFFFFFFFF)
                w = win[woff + lo:woff + hi]
                base = int(st[r[L.B_SPEC]][L.S_LTAB])
                NTP = int(st[r[L.B_SPEC]][L.S_NTERM]); LS = NTP + 4 * NP       # ell-table layout: padded class
                assert NTP >= NT
                rows = tpl[base + lo * LS:base + hi * LS].reshape(hi - lo, LS)
                cols = tpl[t:t + (hi - lo) * RS].reshape(hi - lo, RS)
                np.testing.assert_array_equal(cols[:, :NT], w[:, None] * rows[:, :NT])
                np.testing.assert_array_equal(cols[:, NT:NT + NP], w[:, None] * rows[:, NTP + 1::4][:, :NP])
                np.testing.assert_array_equal(cols[:, NT + NP:NT + 2 * NP], rows[:, NTP + 2::4][:, :NP])
                off = t + (hi - lo) * RS
        assert widths == sorted(widths, reverse=True)
    hb = st[:, L.S_IDXBOUND:L.S_IDXBOUND + 2].view(np.float32)
    assert np.all(hb <= 0) and np.any(np.isclose(hb, -2.0)) and np.any(np.isclose(hb, -0.5))
    assert cp.flops["seg_flops"] > 0
    # group records (one 512-byte load per CTA) restate the bin records of their group
    gt = cp.host["grp_tab"]
    assert gt.shape == (len(bt) // L.SEG_GROUP, L.GRP_STRIDE)
    d = cp.tables["data"]
    for g, rec in enumerate(gt):
        recs = bt[g * L.SEG_GROUP:(g + 1) * L.SEG_GROUP]
        dbl = rec.view(np.float64)
        pre = 0
        for b, r in enumerate(recs):
            live = r[L.B_ROW] >= 0
            ncol = max(int(r[L.B_HI] - r[L.B_LO]), 0) if live else 0
            assert rec[L.G_PRE + b] == pre and rec[L.G_ROW + b] == (r[L.B_ROW] if live else -1)
            if live:
                assert rec[L.G_LO + b] == r[L.B_LO] and dbl[L.G_DATA // 2 + b] == d[r[L.B_ROW]]
            pre += ncol
        assert rec[L.G_PRE + L.SEG_GROUP] == pre and rec[L.G_SPEC] == recs[0][L.B_SPEC]
        assert 0 <= rec[L.G_MODE] <= 3
    # plik-lite bins are contiguous: every bin but the first of a group continues the previous one
    full = [rec for rec in gt if rec[L.G_PRE + L.SEG_GROUP] > 0 and np.all(rec[L.G_ROW:L.G_ROW + 8] >= 0)]
    assert full and all(list(rec[L.G_CONT:L.G_CONT + 8]) == [0] + [1] * 7 for rec in full)
