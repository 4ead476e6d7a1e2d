"""The DEVICE source of the random streams (cosmoslik_b200/csrc/philox.cuh) compiled for the host by g++
(tests/host_twin/philox_twin.cpp defines the CUDA qualifiers away) and checked on the CPU against the
Random123 known-answer vectors and the numpy statement oracle/philox.py: counters, keys, stream ids,
word-to-double conversion and partner indices bit for bit; Box-Muller normals to libm rounding."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import philox

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("twin") / "philox_twin.so")
    subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "host_twin", "philox_twin.cpp")])
    lib = C.CDLL(so)
    lib.twin_u01.restype = C.c_double
    lib.twin_u01.argtypes = [C.c_uint32, C.c_uint32]
    return lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_device_philox_source_known_answers(twin):
    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        ctr, key, out = np.array(c, np.uint32), np.array(k, np.uint32), np.zeros(4, np.uint32)
        twin.twin_philox4x32_10(ptr(ctr), ptr(key), ptr(out))
        assert " ".join("%08x" % int(v) for v in out) == want
    rs = np.random.RandomState(0)
    for hi, lo in rs.randint(0, 2 ** 32, size=(200, 2), dtype=np.uint64):
        assert twin.twin_u01(int(hi), int(lo)) == float(philox.u01(np.uint32(hi), np.uint32(lo)))
    # never 0 (the samplers take log u); the largest mantissa rounds to exactly 1.0 on both sides
    assert twin.twin_u01(0, 0) == 2.0 ** -54 and twin.twin_u01(0xffffffff, 0xffffffff) == 1.0
    assert float(philox.u01(np.uint32(0xffffffff), np.uint32(0xffffffff))) == 1.0


@pytest.mark.parametrize("nd", [1, 2, 9, 30, 64])
def test_device_mh_streams_equal_the_oracle(twin, nd):
    seed, step, c0, W = 0x1234567890ABCDEF, 41, 17, 300
    z, u = np.zeros((W, nd)), np.zeros(W)
    twin.twin_mh(C.c_uint64(seed), C.c_uint32(step), C.c_uint32(c0), W, nd, ptr(z), ptr(u))
    oz, ou = philox.mh_streams(seed, step, c0 + np.arange(W), nd)
    np.testing.assert_array_equal(u, ou)
    np.testing.assert_allclose(z, oz, rtol=0, atol=1e-14)         # sqrt / log / sincos: libm vs numpy last bits


def test_device_stretch_streams_equal_the_oracle(twin):
    seed, step, Ns, Nc, w0 = 99, 7, 257, 32768, 1000
    for half in (0, 1):
        uz, j, ua = np.zeros(Ns), np.zeros(Ns, np.int64), np.zeros(Ns)
        twin.twin_stretch(C.c_uint64(seed), C.c_uint32(step), C.c_uint32(half), C.c_uint32(w0), Ns, C.c_uint32(Nc),
                          ptr(uz), ptr(j), ptr(ua))
        ouz, oj, oua = philox.stretch_streams(seed, step, half, w0 + np.arange(Ns), Nc)
        np.testing.assert_array_equal(uz, ouz)
        np.testing.assert_array_equal(j, oj)
        np.testing.assert_array_equal(ua, oua)
        assert j.min() >= 0 and j.max() < Nc
