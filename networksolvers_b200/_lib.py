"""ctypes binding of libnsb200.so (include/nsb200.h).  There is no CPU fallback: importing works
without a GPU (so that the ABI can be inspected), but creating a context raises if no CUDA device
is present or the library has not been built."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnsb200.so")

NSB_OK = 0
NSB_F64, NSB_C128 = 0, 1
NSB_SITE, NSB_SITE_OUT = -1, -2
NSB_SOLVER_RK, NSB_SOLVER_KRYLOV = 0, 1
NSB_EXPAND_NONE, NSB_EXPAND_DENSITYMATRIX, NSB_EXPAND_ORTHO = 0, 1, 2
NSB_NUM_TIMERS = 8
TIMER_NAMES = ["gauge", "theta", "expand", "env", "matvec", "krylov", "factorize", "other"]
INT64_MAX = 2**63 - 1

ERROR_NAMES = {-1: "NSB_EINVAL", -2: "NSB_ENOMEM", -3: "NSB_ECUDA", -4: "NSB_ENCCL", -5: "NSB_ENOTCONV",
               -6: "NSB_EUNSUPPORTED", -7: "NSB_EINTERNAL"}


class NsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Trunc(C.Structure):
    _fields_ = [("cutoff", C.c_double), ("mindim", C.c_int64), ("maxdim", C.c_int64)]


class Expand(C.Structure):
    _fields_ = [("algorithm", C.c_int32), ("north_pass", C.c_int32), ("expansion_factor", C.c_double),
                ("max_expand", C.c_int64)]


class Krylov(C.Structure):
    _fields_ = [("krylovdim", C.c_int32), ("maxiter", C.c_int32), ("tol", C.c_double), ("which", C.c_int32),
                ("eager", C.c_int32), ("rk_order", C.c_int32), ("reserved", C.c_int32)]


class ExtractInfo(C.Structure):
    _fields_ = [("expanded", C.c_int32), ("env_builds", C.c_int32), ("qr_steps", C.c_int32),
                ("local_rank", C.c_int32), ("local_numel", C.c_int64)]


class SolveInfo(C.Structure):
    _fields_ = [("nmatvec", C.c_int32), ("krylovdim", C.c_int32), ("converged", C.c_int32),
                ("reserved", C.c_int32), ("residual", C.c_double)]


class InsertInfo(C.Structure):
    _fields_ = [("newdim", C.c_int64), ("truncerr", C.c_double), ("decomp", C.c_int32),
                ("jacobi_sweeps", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("gemm_calls", C.c_uint64), ("gemm_flops", C.c_double),
                ("permute_bytes", C.c_uint64), ("matvecs", C.c_uint64), ("env_builds", C.c_uint64),
                ("qr_calls", C.c_uint64), ("svd_calls", C.c_uint64), ("jacobi_sweeps", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


P = C.POINTER
_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol declared in include/nsb200.h
SIGNATURES = {
    "nsb_ctx_create": (C.c_int, [C.c_int, P(_vp)]),
    "nsb_ctx_destroy": (C.c_int, [_vp]),
    "nsb_last_error": (C.c_char_p, [_vp]),
    "nsb_version": (C.c_char_p, []),
    "nsb_ctx_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "nsb_ctx_counters": (C.c_int, [_vp, P(Counters)]),
    "nsb_ctx_counters_reset": (C.c_int, [_vp]),
    "nsb_ctx_synchronize": (C.c_int, [_vp]),
    "nsb_event_tic": (C.c_int, [_vp]),
    "nsb_event_toc": (C.c_int, [_vp, P(_dbl)]),
    "nsb_timers_enable": (C.c_int, [_vp, C.c_int]),
    "nsb_timers_get": (C.c_int, [_vp, P(_dbl)]),
    "nsb_timers_reset": (C.c_int, [_vp]),
    "nsb_mem_info": (C.c_int, [_vp, P(_i64), P(_i64), P(_i64)]),
    "nsb_profiler": (C.c_int, [_vp, _i32]),
    "nsb_gemm_profile_enable": (C.c_int, [_vp, _i32]),
    "nsb_gemm_profile_read": (C.c_int, [_vp, _i64, P(_dbl), P(_dbl), P(_i64), P(_i64)]),
    "nsb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "nsb_comm_init": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int]),
    "nsb_comm_destroy": (C.c_int, [_vp]),
    "nsb_net_set_shard": (C.c_int, [_vp, _i32, P(_i32)]),
    "nsb_shard_emulate": (C.c_int, [_vp, _i32, _vp, P(_i32)]),
    "nsb_peer_window_create": (C.c_int, [_vp, _i64, C.c_char_p]),
    "nsb_peer_window_open": (C.c_int, [_vp, _i32, C.c_char_p]),
    "nsb_multi_create": (C.c_int, [P(_i32), _i32, P(_vp)]),
    "nsb_multi_destroy": (C.c_int, [_vp]),
    "nsb_multi_last_error": (C.c_char_p, [_vp]),
    "nsb_multi_ndev": (C.c_int, [_vp, P(_i32)]),
    "nsb_multi_ctx": (C.c_int, [_vp, _i32, P(_vp)]),
    "nsb_multi_net": (C.c_int, [_vp, _i32, P(_vp)]),
    "nsb_multi_network_create": (C.c_int, [_vp, _i32, P(_i32), _i32, P(_i64), _i32]),
    "nsb_multi_site_upload": (C.c_int, [_vp, _i32, _i32, P(_i32), P(_i64), _vp]),
    "nsb_multi_mpo_upload": (C.c_int, [_vp, _i32, _i32, P(_i32), P(_i64), _vp]),
    "nsb_multi_site_fill_random": (C.c_int, [_vp, _i32, _i32, P(_i32), P(_i64), C.c_uint64, _dbl]),
    "nsb_multi_set_ortho_region": (C.c_int, [_vp, P(_i32), _i32]),
    "nsb_multi_set_shard": (C.c_int, [_vp, _i32, P(_i32)]),
    "nsb_multi_extract": (C.c_int, [_vp, P(_i32), _i32, P(Trunc), P(Expand), P(ExtractInfo)]),
    "nsb_multi_update_eigsolve": (C.c_int, [_vp, P(Krylov), P(_dbl), P(SolveInfo)]),
    "nsb_multi_update_exp": (C.c_int, [_vp, _dbl, _dbl, _i32, P(Krylov), _i32, _i32, P(SolveInfo)]),
    "nsb_multi_insert": (C.c_int, [_vp, P(Trunc), _i32, _i32, P(InsertInfo)]),
    "nsb_multi_matvec_device": (C.c_int, [_vp, _i32]),
    "nsb_multi_local_download": (C.c_int, [_vp, _vp]),
    "nsb_multi_synchronize": (C.c_int, [_vp]),
    "nsb_local_sync": (C.c_int, [_vp]),
    "nsb_network_create": (C.c_int, [_vp, _i32, P(_i32), _i32, P(_i64), _i32, P(_vp)]),
    "nsb_network_destroy": (C.c_int, [_vp]),
    "nsb_site_upload": (C.c_int, [_vp, _i32, _i32, P(_i32), P(_i64), _vp]),
    "nsb_fit_target_upload": (C.c_int, [_vp, _i32, _i32, P(_i32), P(_i64), _vp]),
    "nsb_update_fit": (C.c_int, [_vp, P(_dbl)]),
    "nsb_site_info": (C.c_int, [_vp, _i32, P(_i32), P(_i32), P(_i64)]),
    "nsb_site_download": (C.c_int, [_vp, _i32, _vp]),
    "nsb_site_fill_random": (C.c_int, [_vp, _i32, _i32, P(_i32), P(_i64), C.c_uint64, _dbl]),
    "nsb_mpo_upload": (C.c_int, [_vp, _i32, _i32, P(_i32), P(_i64), _vp]),
    "nsb_set_ortho_region": (C.c_int, [_vp, P(_i32), _i32]),
    "nsb_get_ortho_region": (C.c_int, [_vp, P(_i32), P(_i32)]),
    "nsb_linkdim": (C.c_int, [_vp, _i32, _i32, P(_i64)]),
    "nsb_maxlinkdim": (C.c_int, [_vp, P(_i64)]),
    "nsb_env_drop_all": (C.c_int, [_vp]),
    "nsb_env_count": (C.c_int, [_vp, P(_i32)]),
    "nsb_qn_enable": (C.c_int, [_vp, _i32, P(_i32)]),
    "nsb_qn_set_site": (C.c_int, [_vp, _i32, P(_i32)]),
    "nsb_qn_set_link": (C.c_int, [_vp, _i32, _i32, P(_i32)]),
    "nsb_qn_get_link": (C.c_int, [_vp, _i32, _i32, P(_i32)]),
    "nsb_qn_project": (C.c_int, [_vp, _i32]),
    "nsb_extract": (C.c_int, [_vp, P(_i32), _i32, P(Trunc), P(Expand), P(ExtractInfo)]),
    "nsb_update_eigsolve": (C.c_int, [_vp, P(Krylov), P(_dbl), P(SolveInfo)]),
    "nsb_update_exp": (C.c_int, [_vp, _dbl, _dbl, _i32, P(Krylov), _i32, _i32, P(SolveInfo)]),
    "nsb_insert": (C.c_int, [_vp, P(Trunc), _i32, _i32, P(InsertInfo)]),
    "nsb_local_info": (C.c_int, [_vp, P(_i32), P(_i32), P(_i64)]),
    "nsb_local_download": (C.c_int, [_vp, _vp]),
    "nsb_local_upload": (C.c_int, [_vp, _vp]),
    "nsb_matvec_host": (C.c_int, [_vp, _vp, _vp]),
    "nsb_matvec_host_slab": (C.c_int, [_vp, _vp, _vp]),
    "nsb_env_bytes": (C.c_int, [_vp, P(C.c_int64), P(C.c_int64)]),
    "nsb_shard_range": (C.c_int, [_vp, P(C.c_int64), P(C.c_int64), P(C.c_int64)]),
    "nsb_matvec_device": (C.c_int, [_vp, _i32, _vp]),
    "nsb_matvec_flops": (C.c_int, [_vp, P(_dbl)]),
    "nsb_matvec_flops_executed": (C.c_int, [_vp, P(_dbl)]),
    "nsb_norm": (C.c_int, [_vp, P(_dbl)]),
    "nsb_gemm_host": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32]),
    "nsb_gemm_bench": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _i64, _i64, _i32, _i32, P(_dbl)]),
    "nsb_dmma_peak": (C.c_int, [_vp, P(_dbl)]),
    "nsb_eigh_host": (C.c_int, [_vp, _i32, _i64, _vp, P(_dbl), _vp]),
    "nsb_factorize_host": (C.c_int, [_vp, _i32, _i64, _i64, _vp, P(Trunc), _vp, _vp, P(_dbl), P(InsertInfo)]),
    "nsb_qr_host": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp, _vp]),
    "nsb_qr_bench": (C.c_int, [_vp, _i32, _i64, _i64, _i32, P(_dbl)]),
    "nsb_range_finder_host": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp, _i64, _i32, _i32, _dbl, _dbl, C.c_uint64, _vp, P(_i64)]),
    "nsb_range_finder_heff": (C.c_int, [_vp, _vp, C.c_uint64, _i64, _i32, _i32, _dbl, _dbl, _vp, P(_i64)]),
    "nsb_expand_set_probe": (C.c_int, [_vp, _i64, _i64, _vp]),
    "nsb_sbr_chase_host": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _i64]),
}

_lib = None


def load():
    """Load libnsb200.so and bind every symbol of the header.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C networksolvers_b200/csrc).  networksolvers_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, ctx=None):
    if code != NSB_OK:
        msg = load().nsb_last_error(ctx).decode(errors="replace")
        raise NsbError(code, msg)
