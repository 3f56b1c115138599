"""ctypes binding of liblanczos_b200.so (the C ABI declared in include/lanczos_b200.h).

There is deliberately no fallback: if the shared library is missing or the process has no
CUDA device, importing works (so that CPU-only tooling can inspect the package) but the
first call raises.  The product never routes through NumPy/SciPy or the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LANCZOS_B200_LIB: load another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("LANCZOS_B200_LIB") or os.path.join(_HERE, "liblanczos_b200.so")

LZ_OK, LZ_ERR_INVALID, LZ_ERR_CUDA, LZ_ERR_NOMEM, LZ_ERR_BREAKDOWN, LZ_ERR_UNSUPPORTED, LZ_ERR_PEER = range(7)
LZ_BC_PERIODIC, LZ_BC_DIRICHLET = 0, 1
LZ_FMT_CSR, LZ_FMT_SELL, LZ_FMT_SELL_VALUES = 0, 1, 2
LZ_REORTH_NONE, LZ_REORTH_FULL, LZ_REORTH_SELECTIVE = 0, 1, 2
LZ_SWEEP_CPU, LZ_SWEEP_GPU = 0, 1


class LanczosBreakdown(ArithmeticError):
    """beta vanished: the Krylov space is exhausted (the reference would divide by zero,
    Lanczos.py:113).  ``steps_done`` says how many steps are valid."""

    def __init__(self, msg, steps_done=0):
        super().__init__(msg)
        self.steps_done = steps_done


class RunOpts(C.Structure):
    _fields_ = [("reorth", C.c_int32), ("cgs_passes", C.c_int32), ("ref_compat", C.c_int32),
                ("profile", C.c_int32), ("step_kernel", C.c_int32), ("flags", C.c_int32),
                ("breakdown_tol", C.c_double), ("select_tol", C.c_double)]


class RunInfo(C.Structure):
    _fields_ = [("steps_done", C.c_int32), ("reorth_count", C.c_int32), ("launches", C.c_int32),
                ("apply_launches", C.c_int32), ("update_launches", C.c_int32),
                ("dots_launches", C.c_int32), ("gsupd_launches", C.c_int32),
                ("fused_launches", C.c_int32),
                ("gpu_ms", C.c_float), ("apply_ms", C.c_float), ("update_ms", C.c_float),
                ("dots_ms", C.c_float), ("gsupd_ms", C.c_float), ("fused_ms", C.c_float),
                ("step_kernel", C.c_int32), ("gsfused_launches", C.c_int32), ("gsfused_ms", C.c_float),
                ("border_launches", C.c_int32), ("border_ms", C.c_float), ("alpha_in_update", C.c_int32),
                ("overlap", C.c_int32), ("graph", C.c_int32)]


_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_P = C.POINTER

# name -> (restype, argtypes); every symbol include/lanczos_b200.h declares
SIGNATURES = {
    "lz_abi_version": (C.c_int, []),
    "lz_last_error": (C.c_char_p, []),
    "lz_device_count": (C.c_int, [_P(C.c_int)]),
    "lz_ctx_create": (C.c_int, [C.c_int, _vp, _P(_vp)]),
    "lz_ctx_destroy": (C.c_int, [_vp]),
    "lz_ctx_sync": (C.c_int, [_vp]),
    "lz_op_stencil_create": (C.c_int, [_vp, C.c_int, _P(_i64), C.c_int, _dbl, _P(_dbl), _vp, _P(_vp)]),
    "lz_op_stencil27_create": (C.c_int, [_vp, _P(_i64), C.c_int, _P(_dbl), _vp, _P(_vp)]),
    "lz_op_csr_create": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, C.c_int, C.c_int, _P(_vp)]),
    "lz_op_csr_shard_create": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, C.c_int, C.c_int, _P(_vp)]),
    "lz_op_csr_create_dev": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, C.c_int, C.c_int, _P(_vp)]),
    "lz_potential_eval": (C.c_int, [_vp, _P(_i64), _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp]),
    "lz_op_rows": (C.c_int, [_vp, _P(_i64)]),
    "lz_op_nnz": (C.c_int, [_vp, _P(_i64), _P(_i64)]),
    "lz_op_value_free": (C.c_int, [_vp, _P(_i32)]),
    "lz_op_windowed": (C.c_int, [_vp, _P(_i32)]),
    "lz_op_apply": (C.c_int, [_vp, _vp, _vp]),
    "lz_op_export_csr": (C.c_int, [_vp, _P(_i64), _vp, _vp, _vp]),
    "lz_op_destroy": (C.c_int, [_vp]),
    "lz_lanczos_run": (C.c_int, [_vp, _vp, _vp, _i32, _P(RunOpts), _vp, _vp, _vp, _i64, _vp, _P(RunInfo)]),
    "lz_basis_normalize": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _vp]),
    "lz_reorthogonalize": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _i32, _i32]),
    "lz_ritz_vectors": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _i32, _vp, _i64]),
    "lz_dot": (C.c_int, [_vp, _vp, _vp, _i64, _P(_dbl)]),
    "lz_comm_bytes": (C.c_int, [C.c_int, _i32, _i64, _i64, _P(_i64)]),
    "lz_comm_alloc": (C.c_int, [_vp, _i64, _P(_vp), _vp]),
    "lz_comm_open": (C.c_int, [_vp, _vp, _P(_vp)]),
    "lz_comm_close": (C.c_int, [_vp, _vp]),
    "lz_comm_free": (C.c_int, [_vp, _vp]),
    "lz_team_create": (C.c_int, [C.c_int, C.c_int, _P(C.c_int), _P(_vp), _i64, _i32, _i64, _i64, _P(_vp)]),
    "lz_team_attach": (C.c_int, [_vp, C.c_int, _P(_vp), C.c_int, C.c_int]),
    "lz_team_set_ghosts": (C.c_int, [_vp, C.c_int, _i32, _vp, _vp, _vp]),
    "lz_team_lanczos_run": (C.c_int, [_vp, _P(_vp), _P(_vp), _i32, _P(RunOpts), _vp, _vp, _P(_vp), _P(_i64), _vp, _P(RunInfo)]),
    "lz_team_apply_dots": (C.c_int, [_vp, _P(_vp), _P(_vp), _P(_vp), _vp]),
    "lz_team_destroy": (C.c_int, [_vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m lanczos_b200.build` "
            "(lanczos_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status == LZ_OK:
        return
    msg = load().lz_last_error().decode("utf-8", "replace")
    if status == LZ_ERR_INVALID:
        raise ValueError(msg)
    if status == LZ_ERR_BREAKDOWN:
        raise LanczosBreakdown(msg)
    if status == LZ_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(f"lanczos_b200 error {status}: {msg}")


def device_count() -> int:
    n = C.c_int(0)
    check(load().lz_device_count(C.byref(n)))
    return n.value
