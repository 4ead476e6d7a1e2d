"""ctypes binding of the C ABI declared in include/cosmoslik_b200.h.

There is no CPU fallback: if the shared library is missing or a CUDA device is
needed and absent, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcosmoslik_b200.so")

# geometry constants (keep in sync with the header; checked in tests/test_abi.py)
BM, BN, KC = 64, 128, 16
MAX_DIM, MAX_TERMS, MAX_PLAW = 64, 8, 4
SPEC_STRIDE, TILE_STRIDE, STACK = 64, 32, 16
OP = dict(PUSHP=0, PUSHC=1, ADD=2, SUB=3, MUL=4, DIV=5, NEG=6, EXP=7, LOG=8, POW=9, SQRT=10, STORE=11, SQR=12)
S_NLPAD, S_NTERM, S_NPLAW, S_CAL = 0, 1, 2, 3
S_LTAB, S_CALMODE, S_TERM_COEF, S_PLAW_AMP, S_PLAW_IDX, S_IDXBOUND = 4, 5, 8, 24, 28, 32
S_ABERR, S_ABTAB = 6, 7
MAX_CLASS = 8
BIN_STRIDE, SEG_GROUP, GRP_STRIDE = 16, 8, 128
G_PRE, G_LO, G_ROW, G_CONT, G_MODE, G_BASE_LO, G_BASE_HI, G_SPEC, G_DATA, G_LOGB = 0, 9, 17, 25, 33, 34, 35, 36, 40, 56
B_ROW, B_SPEC, B_LO, B_HI, B_WINOFF_LO, B_WINOFF_HI, B_TABOFF_LO, B_TABOFF_HI, B_DMAX = 0, 1, 2, 3, 4, 5, 6, 7, 8
T_ROW0, T_NROWS, T_SPEC, T_LBEGIN, T_LEND, T_WINOFF_LO, T_WINOFF_HI, T_WINLD, T_GLO, T_GHI = 0, 1, 2, 3, 4, 5, 6, 7, 8, 16
T_WINT_LO, T_WINT_HI = 24, 25
ATILE = 64 * 20      # doubles of one tile-major A tile

_i32p, _f64p, _vp = C.c_void_p, C.c_void_p, C.c_void_p


class Program(C.Structure):
    """csl_program_t"""
    _fields_ = ([(n, C.c_int32) for n in
                 ("ndim", "ncoef", "ncode", "nbox", "ngauss", "nscalar", "nspec", "ntile",
                  "n", "n_pad", "resid_mode", "nclass")] +
                [("cls", (C.c_int32 * 4) * 8)] +
                [(n, C.c_void_p) for n in
                 ("code_op", "code_arg", "code_val", "box_idx", "box_lo", "box_hi",
                  "gauss_idx", "gauss_c", "gauss_w", "scalar_coef", "spec_tab", "tile_tab",
                  "templates", "windows", "data", "direct_coef", "Lmat", "Linv_diag", "Minv")] +
                [("chi2_mode", C.c_int32), ("nsegclass", C.c_int32), ("segcls", (C.c_int32 * 4) * 8),
                 ("bin_tab", C.c_void_p), ("Sigma", C.c_void_p), ("n_base", C.c_int32), ("nmode", C.c_int32),
                 ("grp_tab", C.c_void_p), ("MinvT", C.c_void_p), ("LmatT", C.c_void_p), ("LinvT", C.c_void_p),
                 ("winT", C.c_void_p), ("bin_split", C.c_int32), ("has_aberr", C.c_int32)])


class Ensemble(C.Structure):
    """csl_ensemble_t"""
    _fields_ = ([(n, C.c_int32) for n in ("ndim", "n0", "n1", "h0", "h1")] +
                [("gid0", C.c_uint32), ("gid1", C.c_uint32), ("cap", C.c_int32), ("a", C.c_double),
                 ("seed", C.c_uint64)] +
                [(n, C.c_void_p) for n in ("pos", "lnl", "q", "lnl_q", "zz", "weight", "pos_old", "rows", "nrows",
                                           "accepted", "naccepted", "ws_coef", "ws_prior", "ws_R", "ws_chi2",
                                           "ws_part", "replica", "peer_ptrs", "flag_ptrs")] +
                [("multicast_ptr", C.c_uint64), ("epoch", C.c_uint64), ("npeers", C.c_int32), ("rank", C.c_int32),
                 ("sync", C.c_void_p)])


class MHState(C.Structure):
    """csl_mh_t"""
    _fields_ = ([(n, C.c_int32) for n in ("ndim", "W", "cap")] +
                [("chain0", C.c_uint32), ("seed", C.c_uint64), ("temp", C.c_double), ("max_weight", C.c_double),
                 ("yield_rejected", C.c_int32), ("inj_is_delta", C.c_int32)] +
                [(n, C.c_void_p) for n in ("cur_x", "cur_lnl", "cur_weight", "test_x", "test_lnl", "rows", "nrows",
                                           "accepted", "margin", "naccepted", "F", "step", "inj_z", "inj_u",
                                           "ws_coef", "ws_prior", "ws_R", "ws_chi2", "ws_part")])


_P = C.POINTER(Program)
_int, _u64, _u32, _dbl = C.c_int, C.c_uint64, C.c_uint32, C.c_double

# name -> (restype, argtypes); every symbol include/cosmoslik_b200.h declares
SIGNATURES = {
    "csl_abi_version": (_int, []),
    "csl_last_error": (C.c_char_p, []),
    "csl_sizeof_program": (_int, []),
    "csl_sizeof_ensemble": (_int, []),
    "csl_device_info": (_int, [_int, _vp]),
    "csl_eval_prepare": (_int, [_P, _int, _vp, _vp, _int, _vp, _vp]),
    "csl_bin_residual": (_int, [_P, _int, _vp, _int, _vp, _int, _vp]),
    "csl_direct_residual": (_int, [_P, _int, _vp, _int, _vp, _int, _vp]),
    "csl_chi2": (_int, [_P, _int, _vp, _int, _vp, _vp, _vp]),
    "csl_cov_chi2": (_int, [_P, _int, _vp, _int, _vp, _vp]),
    "csl_combine": (_int, [_P, _int, _vp, _int, _vp, _vp, _vp, _vp]),
    "csl_combine_parts": (_int, [_P, _int, _vp, _int, _vp, _vp, _int, _vp, _vp]),
    "csl_chi2_parts": (_int, [_P, _int, _vp, _int, _vp, _vp]),
    "csl_lnl": (_int, [_P, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "csl_lnl_host": (_int, [_P, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "csl_mh_propose": (_int, [_int, _int, _vp, _vp, _int, _vp, _vp, _u64, _u32, _u32, _vp]),
    "csl_mh_accept": (_int, [_int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _dbl, _dbl,
                             _vp, _vp, _int, _vp, _vp, _vp]),
    "csl_mh_gauss_run": (_int, [_P, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp,
                                _u64, _u32, _u32, _dbl, _dbl, _vp, _vp, _int, _vp]),
    "csl_mh_moments": (_int, [_int, _int, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "csl_rows_moments": (_int, [_int, _int, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp]),
    "csl_rows_gauss_prior": (_int, [_int, _int, _vp, _vp, _int, _int, _vp, _vp, _vp, _vp]),
    "csl_rows_thin": (_int, [_int, _int, _vp, _vp, _int, _dbl, _vp]),
    "csl_stretch_propose": (_int, [_int, _int, _int, _vp, _vp, _dbl, _vp, _vp, _u64, _u32, _u32, _u32,
                                   _vp, _vp, _vp]),
    "csl_stretch_accept": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _u32,
                                  _vp, _vp, _vp]),
    "csl_stretch_accept_publish": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _u32,
                                          _vp, _vp, _vp, _int, _u64, C.c_int64, _vp]),
    "csl_peer_barrier": (_int, [_vp, _int, _int, _u64, _vp, _vp]),
    "csl_peer_wait": (_int, [_vp, _int, _int, _u64, _vp, _vp]),
    "csl_host_gate": (_int, [_vp, _dbl, _vp]),
    "csl_bin_residual_launches": (_int, [_P]),
    "csl_sizeof_mh": (_int, []),
    "csl_mh_run": (_int, [_P, C.POINTER(MHState), _int, _vp]),
    "csl_mh_graph_create": (_int, [_P, C.POINTER(MHState), _int, C.POINTER(C.c_void_p)]),
    "csl_graph_launch": (_int, [_vp, _vp]),
    "csl_graph_nodes": (_int, [_vp]),
    "csl_graph_destroy": (_int, [_vp]),
    "csl_mh_moments_fused": (_int, [_int, _int, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "csl_mh_adapt_factor": (_int, [_int, _vp, _dbl, _vp, _vp, _vp, _vp, _vp]),
    "csl_stretch_run": (_int, [_P, C.POINTER(Ensemble), _int, _u32, C.c_int64, _vp]),
    "csl_emcee_rows": (_int, [_int, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _int, _vp]),
    "csl_chain_moments": (_int, [_int, _int, _vp, _vp, _int, _int, _vp, _vp]),
    "csl_philox_mh": (_int, [_int, _int, _u64, _u32, _u32, _vp, _vp, _vp]),
    "csl_philox_stretch": (_int, [_int, _int, _u64, _u32, _u32, _u32, _vp, _vp, _vp, _vp]),
    "csl_dgemm_dmma": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp]),
    "csl_set_option": (_int, [C.c_char_p, _int]),
    "csl_get_option": (_int, [C.c_char_p]),
}

_lib = None


class CslError(RuntimeError):
    pass


def lib():
    """The loaded shared library (built by ``__graft_entry__.build()`` /
    ``cosmoslik_b200/csrc/build.sh``).  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "cosmoslik_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or cosmoslik_b200/csrc/build.sh -- there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        if (L.csl_sizeof_program() != C.sizeof(Program) or L.csl_sizeof_ensemble() != C.sizeof(Ensemble)
                or L.csl_sizeof_mh() != C.sizeof(MHState)):
            raise ImportError("cosmoslik_b200: struct layout mismatch between header and binding")
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().csl_last_error()
        raise CslError("%s failed (code %d): %s" % (what or "csl call", rc, msg.decode() if msg else ""))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CslError("cosmoslik_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
    return torch


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
