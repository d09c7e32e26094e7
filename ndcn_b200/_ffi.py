"""ctypes binding of libndcn_b200.so (the C ABI declared in include/ndcn_b200.h).

There is NO fallback: if the library is missing or does not export a declared symbol,
loading raises.  Every compute entry point of the package goes through ``lib()``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _build

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

RHS_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
EXCHANGE_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)

# status codes (include/ndcn_b200.h)
OK = 0
E_ARG, E_WORKSPACE, E_METHOD = -1, -2, -3
E_NONFINITE, E_DT_UNDERFLOW, E_MAX_STEPS, E_PEER_TIMEOUT = -10, -11, -12, -13

RHS_NDCN, RHS_HEAT, RHS_GENE, RHS_MUTUAL, RHS_CALLBACK_KIND = 0, 1, 2, 3, 4
F_NO_GRAPH, F_NO_CONTROL, F_NO_RELU = 1, 2, 4
EULER, MIDPOINT, RK4, DOPRI5 = 0, 1, 2, 3
METHODS = {"euler": EULER, "midpoint": MIDPOINT, "rk4": RK4, "dopri5": DOPRI5}
O_TERMINAL_ONLY, O_FORCED_DT, O_TIME_KERNELS = 1, 2, 4
GATHER_LOCAL, GATHER_EXTERNAL = 0, 1
K_STAGE, K_ALGEBRA, K_CONTROL, K_EMIT, K_INIT, K_GATHER, K_EXCHANGE = 0, 1, 2, 3, 4, 5, 6
IMPL_AUTO, IMPL_SIMT, IMPL_UMMA = 0, 1, 2
CFG_STAGE_IMPL, CFG_GATHER_CW, CFG_UMMA_MIN_ROWS, CFG_GATHER_VERSION, CFG_SMALL_SOLVER = 0, 1, 2, 3, 4


class RhsDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("flags", C.c_uint32), ("H", C.c_int32), ("reserved", C.c_int32),
        ("W", C.c_void_p), ("b", C.c_void_p), ("p", C.c_float * 8),
        ("callback", RHS_CALLBACK), ("callback_user", C.c_void_p),
        ("prepared", C.c_void_p),
    ]


class SolveOpts(C.Structure):
    _fields_ = [
        ("method", C.c_int32), ("flags", C.c_uint32), ("rtol", C.c_double), ("atol", C.c_double),
        ("forced_dt", C.c_double), ("max_num_steps", C.c_int64),
        ("exchange", EXCHANGE_CALLBACK), ("exchange_user", C.c_void_p),
        ("safety", C.c_double), ("ifactor", C.c_double), ("dfactor", C.c_double),
        ("first_step", C.c_double),
        ("gather_mode", C.c_int32), ("z_block_cols", C.c_int32),
        ("dec_W", C.c_void_p), ("dec_b", C.c_void_p), ("dec_classes", C.c_int32), ("reserved", C.c_int32),
    ]


class PeerConfig(C.Structure):
    """``ndcn_peer_config_t``: the peer-push multi-GPU scheme (partition.PushPartition)."""

    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("pad", C.c_void_p * 8), ("delta_bytes", C.c_int64 * 8)]


class FeaturePeerConfig(C.Structure):
    """``ndcn_feature_peer_config_t``: feature-sharded peer push (partition.FeaturePushPartition)."""

    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("pad", C.c_void_p * 8), ("xcs", C.c_void_p * 8),
                ("z", C.c_void_p * 8), ("row_bounds", C.c_int64 * 9)]


class GatherRequest(C.Structure):
    _fields_ = [("src_dev", C.c_void_p), ("z_dev", C.c_void_p)]


class SolveStats(C.Structure):
    _fields_ = [
        ("nfe", C.c_int64), ("n_accepted", C.c_int64), ("n_rejected", C.c_int64), ("n_launches", C.c_int64),
        ("first_step", C.c_double), ("last_dt", C.c_double), ("t_final", C.c_double),
        ("status", C.c_int32), ("reserved", C.c_int32),
        ("class_ms", C.c_double * 8), ("class_launches", C.c_int64 * 8),
    ]


# name -> (restype, argtypes); kept in sync with include/ndcn_b200.h (tests/test_abi.py checks)
PROTOTYPES = {
    "ndcn_graph_create": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_void_p)]),
    "ndcn_graph_destroy": (C.c_int, [C.c_void_p]),
    "ndcn_spmm_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "ndcn_rhs_eval_f32": (C.c_int, [C.c_void_p, C.POINTER(RhsDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "ndcn_rhs_vjp_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RhsDesc), C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ndcn_prepared_weights_bytes": (C.c_size_t, [C.c_int32]),
    "ndcn_prepare_weights_f32": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ndcn_weight_grads_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p]),
    "ndcn_fixed_grid_adjoint_small_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RhsDesc), C.c_int32, c_double_p,
                                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p]),
    "ndcn_solver_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int32, C.c_int32]),
    "ndcn_solver_create": (C.c_int, [C.c_void_p, C.POINTER(RhsDesc), C.c_int32, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_void_p)]),
    "ndcn_solver_destroy": (C.c_int, [C.c_void_p]),
    "ndcn_odeint_f32": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int32, C.c_void_p,
                                  C.POINTER(SolveOpts), C.POINTER(SolveStats), C.c_void_p]),
    "ndcn_odeint_small_f32": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int32, C.c_void_p,
                                        C.POINTER(SolveOpts), C.POINTER(SolveStats), C.c_void_p]),
    "ndcn_odeint_staged_f32": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int32, C.c_void_p,
                                         C.POINTER(SolveOpts), C.POINTER(SolveStats), C.c_void_p]),
    "ndcn_rk_combine_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), c_double_p, C.c_int32,
                                      C.c_float, C.c_int64, C.c_void_p]),
    "ndcn_error_ratio_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int64,
                                       C.c_void_p, C.c_void_p]),
    "ndcn_pack_rows_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "ndcn_pack_cols_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ndcn_solver_set_peers": (C.c_int, [C.c_void_p, C.POINTER(PeerConfig)]),
    "ndcn_solver_set_feature_peers": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(FeaturePeerConfig)]),
    "ndcn_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "ndcn_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "ndcn_peer_close": (C.c_int, [C.c_void_p]),
    "ndcn_peer_free": (C.c_int, [C.c_void_p]),
    "ndcn_peer_enable_access": (C.c_int, [C.c_int]),
    "ndcn_config_set": (C.c_int, [C.c_int32, C.c_int64]),
    "ndcn_config_get": (C.c_int64, [C.c_int32]),
    "ndcn_debug_umma_trace": (C.c_int, [C.c_void_p]),
    "ndcn_sizeof": (C.c_int, [C.c_int32]),
    "ndcn_version": (C.c_char_p, []),
    "ndcn_sm_arch": (C.c_int, []),
}

_lib: Optional[C.CDLL] = None


class NdcnLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The loaded library.  Raises NdcnLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        raise NdcnLibraryError(
            "libndcn_b200.so is not built (%s). Run `python -m ndcn_b200._build` or "
            "__graft_entry__.build(); there is no CPU/PyTorch fallback for the hot path." % path)
    try:
        handle = C.CDLL(path)
    except OSError as exc:  # pragma: no cover
        raise NdcnLibraryError("cannot load %s: %s" % (path, exc)) from exc
    for name, (restype, argtypes) in PROTOTYPES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as exc:
            if os.environ.get("NDCN_B200_LIB"):
                continue  # A/B run against an older build (scripts/): symbols added since are simply absent
            raise NdcnLibraryError("libndcn_b200.so lacks symbol %s" % name) from exc
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return handle


_MESSAGES = {
    E_ARG: "invalid argument",
    E_WORKSPACE: "workspace too small",
    E_METHOD: "unsupported method",
    E_NONFINITE: "non-finite values in state `y`",
    E_DT_UNDERFLOW: "underflow in dt",
    E_MAX_STEPS: "max_num_steps exceeded",
    E_PEER_TIMEOUT: "peer push: another rank did not reach the barrier within 20 s",
}


def check(rc: int, what: str = "ndcn call") -> None:
    """Map a status code to the exception type the reference raises for the same condition."""
    if rc == OK:
        return
    if rc in (E_NONFINITE, E_DT_UNDERFLOW, E_MAX_STEPS):
        # the reference signals these with bare ``assert`` (dopri5.py:89,100,102)
        raise AssertionError(_MESSAGES[rc])
    if rc < 0:
        raise ValueError("%s: %s (status %d)" % (what, _MESSAGES.get(rc, "error"), rc))
    raise RuntimeError("%s: CUDA error %d" % (what, rc))


def configure(stage_impl: Optional[int] = None, gather_cw: Optional[int] = None,
              umma_min_rows: Optional[int] = None, gather_version: Optional[int] = None,
              small_solver: Optional[int] = None) -> dict:
    """Process-wide kernel-family knobs (``ndcn_config_set``); returns the previous values."""
    h = lib()
    prev = {"stage_impl": int(h.ndcn_config_get(CFG_STAGE_IMPL)), "gather_cw": int(h.ndcn_config_get(CFG_GATHER_CW)),
            "umma_min_rows": int(h.ndcn_config_get(CFG_UMMA_MIN_ROWS)),
            "gather_version": int(h.ndcn_config_get(CFG_GATHER_VERSION)),
            "small_solver": int(h.ndcn_config_get(CFG_SMALL_SOLVER))}
    if stage_impl is not None:
        check(h.ndcn_config_set(CFG_STAGE_IMPL, int(stage_impl)), "ndcn_config_set")
    if gather_cw is not None:
        check(h.ndcn_config_set(CFG_GATHER_CW, int(gather_cw)), "ndcn_config_set")
    if umma_min_rows is not None:
        check(h.ndcn_config_set(CFG_UMMA_MIN_ROWS, int(umma_min_rows)), "ndcn_config_set")
    if gather_version is not None:
        check(h.ndcn_config_set(CFG_GATHER_VERSION, int(gather_version)), "ndcn_config_set")
    if small_solver is not None:
        check(h.ndcn_config_set(CFG_SMALL_SOLVER, int(small_solver)), "ndcn_config_set")
    return prev
