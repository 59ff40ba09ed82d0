"""ctypes binding of ``libtdgl_b200.so`` (the C ABI declared in ``include/tdgl_b200.h``).

There is no CPU fallback: if the shared library is missing or a CUDA call fails, the
error is raised to the caller.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (TDGL_B200_LIB: development only — a side-by-side build of the same sources, see
# tools/build_variant.sh; there is still no fallback if it is missing)
LIB_PATH = os.environ.get("TDGL_B200_LIB") or os.path.join(_HERE, "libtdgl_b200.so")

TDGL_OK, TDGL_E_STEP_FAILED, TDGL_E_MU_SOLVER, TDGL_E_CUDA, TDGL_E_INVALID = range(5)


class TDGLLibraryError(RuntimeError):
    pass


class tdgl_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("device", C.c_int32),
        ("mu_rtol", C.c_double),
        ("mu_max_iter", C.c_int32),
        ("amg_theta", C.c_double),
        ("amg_max_coarse", C.c_int32),
        ("use_graph", C.c_int32),
        ("reorder", C.c_int32),
        ("running_capacity", C.c_int32),
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("replicate_below", C.c_int32),
        ("fuse_coarse", C.c_int32),
    ]


class tdgl_advance_info(C.Structure):
    _fields_ = [
        ("steps_done", C.c_int64),
        ("step", C.c_int64),
        ("time", C.c_double),
        ("dt", C.c_double),
        ("tentative_dt", C.c_double),
        ("finished", C.c_int32),
        ("status", C.c_int32),
        ("failed_step", C.c_int64),
        ("failed_dt", C.c_double),
        ("retries", C.c_int64),
        ("mu_iterations", C.c_int64),
        ("mu_rel_residual", C.c_double),
        ("device_ms", C.c_double),
        ("screening_iterations", C.c_int64),
        ("screening_error", C.c_double),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_D = C.c_double

# name -> (restype, argtypes); every symbol include/tdgl_b200.h declares
SIGNATURES = {
    "tdgl_version": (C.c_char_p, []),
    "tdgl_create": (C.c_int, [C.POINTER(_P), _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _I64,
                              _I32, _P, _D, _D, _P, _I64, C.POINTER(tdgl_config)]),
    "tdgl_destroy": (None, [_P]),
    "tdgl_last_error": (C.c_char_p, [_P]),
    "tdgl_set_link_exponents": (C.c_int, [_P, _P]),
    "tdgl_set_epsilon": (C.c_int, [_P, _P]),
    "tdgl_set_mu_boundary": (C.c_int, [_P, _P]),
    "tdgl_set_dA_dt": (C.c_int, [_P, _P]),
    "tdgl_set_vector_potential_ramp": (C.c_int, [_P, _P, _I32, _P, _P]),
    "tdgl_set_state": (C.c_int, [_P, _P, _P]),
    "tdgl_set_terminal_current_table": (C.c_int, [_P, _I32, _P, _P, _I32, _P, _P]),
    "tdgl_set_epsilon_table": (C.c_int, [_P, _P, _P, _I32, _P, _P]),
    "tdgl_set_screening": (C.c_int, [_P, _I32, _D, _P, _P, _D, _I32, _D, _D]),
    "tdgl_set_induced_vector_potential": (C.c_int, [_P, _P]),
    "tdgl_get_induced_vector_potential": (C.c_int, [_P, _P]),
    "tdgl_get_running_screening": (C.c_int, [_P, _I64, _P]),
    "tdgl_set_stepper": (C.c_int, [_P, _D, _D, _I32, _I32, _I32, _D]),
    "tdgl_advance": (C.c_int, [_P, _I64, _D, _I64, _D, C.POINTER(tdgl_advance_info)]),
    "tdgl_update": (C.c_int, [_P, _P, _P, _I64, _D, _P, _P, _P, _P,
                              C.POINTER(tdgl_advance_info)]),
    "tdgl_local_maps": (C.c_int, [_P, _P, _P, _P]),
    "tdgl_update_local": (C.c_int, [_P, _P, _P, _I64, _D, _P, _P, _P, _P,
                                    C.POINTER(tdgl_advance_info)]),
    "tdgl_host_alloc": (C.c_void_p, [_I64]),
    "tdgl_host_free": (None, [_P]),
    "tdgl_get_state": (C.c_int, [_P, _P, _P]),
    "tdgl_get_currents": (C.c_int, [_P, _P, _P]),
    "tdgl_snapshot_begin": (C.c_int, [_P, _I32]),
    "tdgl_snapshot_wait": (C.c_int, [_P, _I32, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                     C.POINTER(_P)]),
    "tdgl_get_running": (C.c_int, [_P, _I64, _P, _P, _P]),
    "tdgl_op_psi_laplacian": (C.c_int, [_P, _P, _P]),
    "tdgl_op_psi_step": (C.c_int, [_P, _P, _P, _D, _P, _P, C.POINTER(_I32)]),
    "tdgl_op_mu_rhs": (C.c_int, [_P, _P, _P]),
    "tdgl_op_mu_laplacian": (C.c_int, [_P, _P, _P]),
    "tdgl_op_mu_solve": (C.c_int, [_P, _P, _P, C.POINTER(_I32), C.POINTER(_D)]),
    "tdgl_time_kernel": (C.c_int, [_P, _I32, _I32, _I32, C.POINTER(_D)]),
    "tdgl_time_cusparse": (C.c_int, [_P, _I32, _I32, _I32, C.POINTER(_D)]),
    "tdgl_get_info": (C.c_int, [_P, C.POINTER(_I64), _I32]),
    "tdgl_get_trace": (C.c_int, [_P, _I32, C.POINTER(_I32), _P, _P, _P, _P, _P]),
    "tdgl_comm_export": (C.c_int, [_P, _P]),
    "tdgl_comm_connect_ipc": (C.c_int, [_P, _P, _I32]),
    "tdgl_comm_connect_local": (C.c_int, [_P, C.POINTER(_P), _I32]),
    "tdgl_shard_info": (C.c_int, [_P, C.POINTER(_I64), _I32]),
    "tdgl_stage_outputs": (C.c_int, [_P, _I32, C.POINTER(_P), C.POINTER(_I64)]),
    "tdgl_fetch_outputs": (C.c_int, [_P, _P, _P, _P, _P]),
    "tdgl_host_shard_probe": (C.c_int, [_I64, _I64, _P, _P, _P, _P, _I32, _D, _I32, _I64,
                                        C.POINTER(_I32), _P, _P, _P, _P, _P, _I32, _D,
                                        C.POINTER(_I32)]),
    "tdgl_host_shard_lists": (C.c_int, [_I64, _I64, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P,
                                        _P]),
    "tdgl_host_mesh_dual": (C.c_int, [_I64, _I64, _P, _P, C.POINTER(_I64), _P, _P, _P, _P, _P,
                                      _P, _P, _P]),
    "tdgl_host_amg_probe": (C.c_int, [_I64, _I64, _P, _P, _P, _D, _I32, C.POINTER(_I32),
                                      C.POINTER(_I64), C.POINTER(_I64), _P, _P, _I32, _D,
                                      C.POINTER(_I32)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (built by ``__graft_entry__.build()``); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TDGLLibraryError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a)."
            " tdgl_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_f64(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and out.shape != shape:
        raise ValueError(f"expected shape {shape}, got {out.shape}")
    return out


def as_i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def as_c128(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.complex128)
    if shape is not None and out.shape != shape:
        raise ValueError(f"expected shape {shape}, got {out.shape}")
    return out
