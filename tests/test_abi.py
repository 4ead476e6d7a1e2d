"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/cosmoslik_b200.h declares; the Python binding agrees with the header's constants."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from cosmoslik_b200 import _lib


def header():
    with open(os.path.join(ROOT, "include", "cosmoslik_b200.h")) as f:
        return f.read()


def test_every_declared_symbol_is_exported_and_bound():
    names = re.findall(r"^(?:int|const char\*)\s+(csl_\w+)\s*\(", header(), flags=re.M)
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "symbol %s has no ctypes signature" % n
    assert set(_lib.SIGNATURES) == set(names)


def test_constants_match_header():
    h = header()
    defs = dict(re.findall(r"#define\s+CSL_(\w+)\s+(-?\d+)", h))
    assert int(defs["BM"]) == _lib.BM and int(defs["BN"]) == _lib.BN and int(defs["KC"]) == _lib.KC
    assert int(defs["MAX_DIM"]) == _lib.MAX_DIM and int(defs["MAX_TERMS"]) == _lib.MAX_TERMS
    assert int(defs["MAX_PLAW"]) == _lib.MAX_PLAW and int(defs["SPEC_STRIDE"]) == _lib.SPEC_STRIDE
    assert int(defs["TILE_STRIDE"]) == _lib.TILE_STRIDE and int(defs["STACK"]) == _lib.STACK
    enums = dict((k, int(v)) for k, v in re.findall(r"CSL_(\w+)\s*=\s*(\d+)", h))
    for k, v in _lib.OP.items():
        assert enums["OP_" + k] == v
    for k in ("S_NLPAD", "S_NTERM", "S_NPLAW", "S_CAL", "S_LTAB", "S_CALMODE", "S_TERM_COEF", "S_PLAW_AMP", "S_PLAW_IDX",
              "T_ROW0", "T_NROWS", "T_SPEC", "T_LBEGIN", "T_LEND", "T_WINOFF_LO",
              "T_WINOFF_HI", "T_WINLD", "T_GLO", "T_GHI", "B_ROW", "B_SPEC", "B_LO", "B_HI", "B_WINOFF_LO", "B_WINOFF_HI",
              "B_DMAX"):
        assert enums[k] == getattr(_lib, k), k


def test_library_loads_and_struct_layout_agrees():
    lib = _lib.lib()
    assert lib.csl_abi_version() == 2
    assert lib.csl_sizeof_program() == ctypes.sizeof(_lib.Program)
    assert lib.csl_sizeof_ensemble() == ctypes.sizeof(_lib.Ensemble)
    assert lib.csl_sizeof_mh() == ctypes.sizeof(_lib.MHState)


def test_argument_validation_without_a_gpu():
    """bad arguments are rejected before any CUDA call"""
    lib = _lib.lib()
    rc = lib.csl_dgemm_dmma(63, 128, 16, None, None, None, None)
    assert rc == -1 and b"null" in lib.csl_last_error()
    rc = lib.csl_mh_propose(0, 3, None, None, 0, None, None, 0, 0, 0, None)
    assert rc == -1


def test_integration_md_struct_matches_the_binding():
    """the ctypes stub shown to reference maintainers is the real layout"""
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        src = f.read()
    code = src[src.index("class csl_program_t(C.Structure):"):src.index("lib.csl_lnl.restype")]
    ns = {"C": ctypes}
    exec(code, ns)
    assert ctypes.sizeof(ns["csl_program_t"]) == ctypes.sizeof(_lib.Program)
    assert [f[0] for f in ns["csl_program_t"]._fields_] == [f[0] for f in _lib.Program._fields_]


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/cosmoslik_b200.h must compile as C99 and as C++ on its own (no CUDA,
    no torch types), and a C program can take the address of every declared entry point"""
    import shutil
    import subprocess
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    if gcc is None or gxx is None:
        pytest.skip("no gcc / g++")
    header = os.path.join(ROOT, "include", "cosmoslik_b200.h")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", header])
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", header])
    names = sorted(set(re.findall(r"\b(csl_[a-z0-9_]+)\s*\(", open(header).read())))
    assert len(names) >= 20
    src = tmp_path / "use.c"
    src.write_text('#include "cosmoslik_b200.h"\nvoid* table[] = {%s};\nint main(void) { return table[0] == 0; }\n'
                   % ", ".join("(void*)%s" % n for n in names))
    exe = str(tmp_path / "use")
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                           "-L", os.path.join(ROOT, "cosmoslik_b200"), "-lcosmoslik_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "cosmoslik_b200")])


def test_every_dynamic_shared_memory_launch_is_bounded_or_opts_in():
    """a launch whose dynamic shared memory grows with the problem must either opt in to the size
    (cudaFuncSetAttribute) or prove at compile time that it stays under the 48 KB default (static_assert):
    k_prepare missed both and refused to launch from 48 parameters on until the ndim = 64 tests found it"""
    csrc = os.path.join(ROOT, "cosmoslik_b200", "csrc")
    checked = 0
    for fn in sorted(os.listdir(csrc)):
        if not fn.endswith(".cu"):
            continue
        src = open(os.path.join(csrc, fn)).read()
        for m in re.finditer(r"<<<(.*?)>>>", src, flags=re.S):
            args = [a.strip() for a in re.split(r",(?![^()]*\))", m.group(1))]
            if len(args) < 3 or args[2] == "0":
                continue
            # the enclosing function: from the previous line starting in column 0 with a letter up to the launch
            start = max(src.rfind("\nstatic ", 0, m.start()), src.rfind("\nextern ", 0, m.start()),
                        src.rfind("\nint ", 0, m.start()), src.rfind("\ntemplate ", 0, m.start()))
            body = src[start:m.start()]
            assert "cudaFuncSetAttribute" in body or "static_assert" in body, \
                "%s: launch with dynamic shared memory %r neither opts in nor bounds it" % (fn, args[2])
            checked += 1
    assert checked >= 8
