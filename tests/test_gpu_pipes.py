"""Tile movement is not arithmetic: the bulk-copy (cp.async.bulk + mbarrier, producer warp) variants of the FP64
DMMA kernels must give bit for bit the results of the cp.async variants of round 1, for every tile shape."""
import ctypes as C

import numpy as np
import pytest

from namespaces import PRODUCT
from conftest import golden  # noqa: F401
from cosmoslik_b200 import Slik, synthetic
from cosmoslik_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def set_opts(**kw):
    lib = L.lib()
    for k in ("pipe", "trigemm_rows", "trigemm_stages", "trsm_shape", "bin_variant", "bin_split", "slab_warps", "seg_pair", "seg_small"):
        L.check(lib.csl_set_option(k.encode(), int(kw.get(k, -1))))


@pytest.fixture(autouse=True)
def reset_options():
    yield
    set_opts()


def chi2_of(torch, prog, R0, **opts):
    set_opts(**opts)
    lib = L.lib()
    W = R0.shape[1]
    R = R0.clone()
    chi2 = torch.empty(W, dtype=torch.float64, device="cuda")
    part = torch.zeros((prog.n_pad // 32, W), dtype=torch.float64, device="cuda")
    L.check(lib.csl_chi2(C.byref(prog.struct), W, R.data_ptr(), W, chi2.data_ptr(), part.data_ptr(), L.stream_ptr()))
    torch.cuda.synchronize()
    return chi2.cpu().numpy(), R.cpu().numpy()


@pytest.mark.parametrize("kind", ["planck", "spt"])
@pytest.mark.parametrize("mode", ["inverse", "trsm"])
def test_chi2_bulk_pipeline_is_bitwise_the_cp_async_one(torch, kind, mode):
    prob = synthetic.planck_lite_problem(0) if kind == "planck" else synthetic.spt_multifreq_problem(0)
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    prog = slik.program(chi2_mode=mode)
    assert prog.meta["chi2_mode"] == (1 if mode == "inverse" else 0)
    rs = np.random.RandomState(3)
    for W in (128, 1664):
        R0 = np.zeros((prog.n_pad, W))
        R0[:prog.n] = rs.standard_normal((prog.n, W)) * np.sqrt(np.diag(prob["cov"]))[:, None]
        d = torch.from_numpy(R0).cuda()
        ref, Rref = chi2_of(torch, prog, d, pipe=0, trsm_shape=0)
        # against LAPACK: chi2 = r^T C^-1 r
        want = np.einsum("iw,iw->w", R0[:prog.n], np.linalg.solve(prob["cov"], R0[:prog.n]))
        np.testing.assert_allclose(ref, want, rtol=1e-10)
        variants = [dict(pipe=1), dict(trsm_shape=1), dict(trsm_shape=2), dict(trsm_shape=3), dict(trsm_shape=4),
                    dict(trsm_shape=4, slab_warps=1), dict(trsm_shape=4, slab_warps=5), dict(trsm_shape=4, slab_warps=14)]
        if mode == "inverse":
            variants += [dict(pipe=p, trigemm_rows=r, trigemm_stages=s) for p in (0, 1) for r in (32, 64) for s in (3, 4)]
        for v in variants:
            got, Rgot = chi2_of(torch, prog, d, **v)
            np.testing.assert_array_equal(got, ref, err_msg=str(v))
            if mode == "trsm":
                np.testing.assert_array_equal(Rgot, Rref, err_msg=str(v))      # Y = L^-1 R left in place


@pytest.mark.parametrize("kind", ["spt", "planck_dmma", "spt_lowl"])
def test_binning_warp_specialised_cluster_kernel(torch, spt_data, kind):
    """k_bin_residual_ws (producer / builder / MMA warps, bulk copies, split-ell clusters with the partial tiles
    summed through distributed shared memory) against k_bin_residual: without a split bit for bit, with a split to
    rounding of the re-associated ell sum; every variant within 1e-10 of the oracle's lnL"""
    import b200_scripts as B
    from namespaces import ORACLE
    from oracle import runtime as ort
    if kind == "spt_lowl":
        g = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "spt_lowl_lnl.npz"))
        make = lambda **kw: Slik(B.spt_script(spt_data, "s12", g["tt"])()).program(**kw)
        X = np.array(g["s12_x"]); want = np.array(g["s12_lnl"])
    else:
        prob = synthetic.spt_multifreq_problem(0) if kind == "spt" else synthetic.planck_lite_problem(0)
        kw0 = dict(binning="dmma") if kind == "planck_dmma" else {}
        make = lambda **kw: Slik(synthetic.build_script(PRODUCT, prob)()).program(**dict(kw0, **kw))
        names = sorted(prob["truth"])
        rs = np.random.RandomState(8)
        X = np.array([prob["truth"][k] for k in names]) + np.array([prob["scales"][k] for k in names]) * rs.standard_normal((300, len(names)))
        oslik = ort.Slik(synthetic.build_script(ORACLE, prob)())
        want = np.array([oslik.evaluate(*x)[0] for x in X[:40]])
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    prog = make()
    assert prog.meta["bin_split"] == {"planck_dmma": 1, "spt": 4, "spt_lowl": 4}[kind]
    set_opts(bin_variant=0, bin_split=1)
    old = prog.lnl(Xd).cpu().numpy()
    fin = np.isfinite(want)
    np.testing.assert_allclose(old[:len(want)][fin], want[fin], rtol=1e-10)
    set_opts(bin_variant=2, bin_split=1)
    np.testing.assert_array_equal(prog.lnl(Xd).cpu().numpy(), old)            # same arithmetic, new data movement
    for split in (2, 4):
        per_variant = []
        for variant in (0, 2):
            set_opts(bin_variant=variant, bin_split=split)
            got = prog.lnl(Xd).cpu().numpy()
            f2 = np.isfinite(old)
            np.testing.assert_array_equal(np.isfinite(got), f2)
            np.testing.assert_allclose(got[f2], old[f2], rtol=1e-12)
            np.testing.assert_allclose(got[:len(want)][fin], want[fin], rtol=1e-10)
            per_variant.append(got)
        np.testing.assert_array_equal(per_variant[0], per_variant[1])        # the split fixes the bits, not the kernel
    set_opts()
    # the program's own split (fixed per program): results do not depend on the batch a walker is evaluated in
    a = prog.lnl(Xd).cpu().numpy()
    b = np.concatenate([prog.lnl(Xd[:130].contiguous()).cpu().numpy(), prog.lnl(Xd[130:].contiguous()).cpu().numpy()])
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("W", [128, 384, 2048, 4224, 33152 + 128])
def test_slab_substitution_kernel_over_launch_widths(torch, W):
    """k_trsm_chi2_slab (one warp per 16- / 8-walker slab, the default substitution kernel): CTAs of wide slabs only,
    of wide + narrow slabs (16 consumer warps), a ragged last CTA and a second round of CTAs -- chi^2 and the
    Y = L^-1 R it leaves in place are bit for bit those of the round-1 kernel; run twice (the kernel reads rows it
    wrote itself: a stale read would differ between runs)"""
    prob = synthetic.planck_lite_problem(0)
    prog = Slik(synthetic.build_script(PRODUCT, prob)()).program(chi2_mode="trsm")
    g = torch.Generator(device="cuda"); g.manual_seed(W)
    d = torch.zeros((prog.n_pad, W), dtype=torch.float64, device="cuda")
    d[:prog.n] = torch.randn((prog.n, W), dtype=torch.float64, device="cuda", generator=g)
    ref, Rref = chi2_of(torch, prog, d, trsm_shape=0)
    for rep in range(2):
        got, Rgot = chi2_of(torch, prog, d)                  # the default shape
        np.testing.assert_array_equal(got, ref)
        np.testing.assert_array_equal(Rgot, Rref)


def test_two_segmented_sum_classes_in_one_launch(torch):
    """k_bin_segsum_pair (TT and TE / EE classes of the plik-lite-like problem in one grid) against one launch per class:
    the same bits, with one and with two walkers per thread, ragged walker counts"""
    prob = synthetic.planck_lite_problem(0)
    slik = Slik(synthetic.build_script(PRODUCT, prob)())
    prog = slik.program()
    assert prog.meta["nsegclass"] == 2 and prog.meta["nclass"] == 0
    lib = L.lib()
    names = sorted(prob["truth"])
    rs = np.random.RandomState(5)
    for W, small in ((300, -1), (1000, 0), (9000, -1)):
        X = np.array([prob["truth"][k] for k in names]) + np.array([prob["scales"][k] for k in names]) * rs.standard_normal((W, len(names)))
        Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
        set_opts(seg_pair=0, seg_small=small)
        assert lib.csl_bin_residual_launches(prog.struct) == 2
        ref = prog.lnl(Xd).cpu().numpy()
        set_opts(seg_small=small)
        assert lib.csl_bin_residual_launches(prog.struct) == 1
        np.testing.assert_array_equal(prog.lnl(Xd).cpu().numpy(), ref)
