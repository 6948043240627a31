"""ctypes binding of libmicrotorch_b200.so (the C ABI declared in include/microtorch_b200.h).

This is the ONLY place the shared library is touched.  There is no fallback: if the library
is missing, cannot be loaded, or no B200 is visible, `lib()` / `require_device()` raise - the
CUDA path never silently runs somewhere else.
"""

from __future__ import annotations

import ctypes as C
import functools
import os
from typing import Optional, Sequence

from . import _build

_LIB: Optional[C.CDLL] = None
_DEVICE_READY = False

c_f32p = C.POINTER(C.c_float)
c_i64p = C.POINTER(C.c_int64)
c_sizep = C.POINTER(C.c_size_t)
c_u64p = C.POINTER(C.c_uint64)
c_voidpp = C.POINTER(C.c_void_p)
VP = C.c_void_p

MLP_MAX_LAYERS = 8


class MlpDesc(C.Structure):
    """mt_mlp_desc of the header."""
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MLP_MAX_LAYERS + 1)), ("act", C.c_int32 * MLP_MAX_LAYERS),
                ("weight", C.c_void_p * MLP_MAX_LAYERS), ("bias", C.c_void_p * MLP_MAX_LAYERS),
                ("grad_weight", C.c_void_p * MLP_MAX_LAYERS)]


# name -> (argtypes, doc)   every symbol include/microtorch_b200.h declares
PROTOTYPES = {
    "mt_device_count": [C.POINTER(C.c_int)],
    "mt_init": [C.c_int],
    "mt_shutdown": [],
    "mt_is_initialized": [],
    "mt_device_info": [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), c_sizep, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "mt_sync": [],
    "mt_stream": [c_voidpp],
    "mt_event_create": [c_voidpp],
    "mt_event_record": [VP],
    "mt_event_sync": [VP],
    "mt_event_elapsed_ms": [VP, VP, C.POINTER(C.c_float)],
    "mt_event_destroy": [VP],
    "mt_launch_count": [c_u64p],
    "mt_launch_count_reset": [],
    "mt_graph_begin": [],
    "mt_graph_end": [c_voidpp],
    "mt_graph_launch": [VP],
    "mt_graph_destroy": [VP],
    "mt_l2_flush": [],
    "mt_prof_enable": [C.c_int],
    "mt_prof_reset": [],
    "mt_prof_gemm": [c_u64p, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "mt_prof_gemm_units": [c_u64p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "mt_malloc": [c_voidpp, C.c_size_t],
    "mt_free": [VP],
    "mt_pool_stats": [c_sizep, c_sizep, c_sizep, c_u64p, c_u64p],
    "mt_pool_trim": [],
    "mt_host_alloc": [c_voidpp, C.c_size_t],
    "mt_host_free": [VP],
    "mt_h2d": [VP, VP, C.c_size_t],
    "mt_d2h": [VP, VP, C.c_size_t],
    "mt_h2d_async": [VP, VP, C.c_size_t],
    "mt_d2h_async": [VP, VP, C.c_size_t],
    "mt_h2d_prefetch": [VP, VP, C.c_size_t],
    "mt_prefetch_wait": [],
    "mt_d2d": [VP, VP, C.c_size_t],
    "mt_memset_zero": [VP, C.c_size_t],
    "mt_memset_multi": [c_voidpp, c_sizep, C.c_int],
    "mt_fill_f32": [VP, C.c_size_t, C.c_float],
    "mt_copy_strided_f32": [VP, VP, C.c_int, c_i64p, c_i64p],
    "mt_scatter_strided_f32": [VP, c_i64p, VP, C.c_int, c_i64p],
    "mt_gather_rows_f32": [VP, VP, VP, C.c_int64, C.c_int64, C.c_int64],
    "mt_scatter_rows_assign_f32": [VP, VP, VP, C.c_int64, C.c_int64, C.c_int64],
    "mt_ew_binary_f32": [C.c_int, VP, VP, VP, C.c_int, c_i64p, c_i64p, c_i64p, C.c_float],
    "mt_ew_inplace_f32": [C.c_int, VP, c_i64p, VP, c_i64p, C.c_int, c_i64p],
    "mt_ew_unary_f32": [C.c_int, VP, VP, C.c_size_t, C.c_float],
    "mt_ew_where_f32": [VP, VP, VP, VP, C.c_int, c_i64p, c_i64p, c_i64p, c_i64p],
    "mt_reduce_sum_f32": [VP, VP, C.c_int64, C.c_int64, C.c_int64, VP, C.c_int64, C.c_int64, C.c_int64],
    "mt_reduce_extreme_f32": [C.c_int, VP, VP, C.c_int64, C.c_int64, C.c_int64],
    "mt_linear_fwd_f32": [VP, VP, VP, VP, C.c_int64, C.c_int64, C.c_int64, C.c_int],
    "mt_linear_bwd_f32": [VP, VP, VP, VP, VP, VP, C.c_int64, C.c_int64, C.c_int64, C.c_int],
    "mt_matmul_f32": [VP, VP, VP] + [C.c_int64] * 10,
    "mt_gemm_set_precision": [C.c_int],
    "mt_gemm_last_engine": [C.POINTER(C.c_int)],
    "mt_gemm_set_engine": [C.c_int],
    "mt_loss_fwd_bwd_f32": [C.c_int, VP, VP, VP, VP, C.c_size_t, C.c_float],
    "mt_sgd_multi_f32": [c_voidpp, c_voidpp, c_sizep, C.c_int, C.c_float, C.c_int],
    "mt_mlp_fit_f32": [C.POINTER(MlpDesc), VP, VP, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                       C.c_int, VP],
    "mt_comm_load": [C.c_char_p],
    "mt_comm_unique_id": [C.c_char_p],
    "mt_comm_init_rank": [C.c_char_p, C.c_int, C.c_int],
    "mt_comm_world": [C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "mt_allreduce_sum_f32": [VP, C.c_size_t],
    "mt_allreduce_sgd_f32": [VP, VP, C.c_size_t, C.c_float],
    "mt_rng_uniform_f32": [VP, C.c_size_t, C.c_ulonglong, C.c_ulonglong, C.c_float, C.c_float],
    "mt_rng_normal_f32": [VP, C.c_size_t, C.c_ulonglong, C.c_ulonglong],
    "mt_p2p_create": [C.c_size_t, C.c_char_p],
    "mt_p2p_open": [C.c_int, C.c_int, C.c_char_p],
    "mt_p2p_active": [C.POINTER(C.c_int)],
    "mt_p2p_error": [C.POINTER(C.c_uint)],
    "mt_p2p_allreduce_sgd_f32": [VP, VP, C.c_size_t, C.c_float, C.c_ulonglong],
    "mt_p2p_destroy": [],
    "mt_comm_wait": [],
    "mt_comm_check": [],
    "mt_comm_destroy": [],
    "mt_trace_enable": [C.c_int],
    "mt_trace_reset": [],
    "mt_trace_count": [C.POINTER(C.c_int)],
    "mt_trace_get": [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_float)],
}
# return `const char*`
STRING_FUNCS = ("mt_last_error", "mt_version")

# enums of the header
BIN_ADD, BIN_MUL, BIN_SUB, BIN_DIV, BIN_TANH_BWD, BIN_POW_BWD, BIN_EXP_BWD = range(7)
BIN_LT, BIN_GT, BIN_LE, BIN_GE, BIN_EQ, BIN_NE, BIN_MAX, BIN_MIN = range(7, 15)
UN_NEG, UN_LOG, UN_TANH, UN_EXP, UN_POW, UN_SCALE, UN_ADDS, UN_ABS = range(8)
ACT_NONE, ACT_TANH = 0, 1
LOSS_MSE, LOSS_CE_REF, LOSS_BCE = 0, 1, 2
ERR_UNSUPPORTED = 6


class MicrotorchB200Error(RuntimeError):
    """A C-ABI call returned a non-zero mt_status."""


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen the shared library and declare every prototype.  Raises if it is absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = path or os.environ.get("MICROTORCH_B200_LIB") or _build.lib_path()
    if not os.path.exists(path):
        raise MicrotorchB200Error(
            f"{path} not found: build it with `python -m microtorch_b200._build` "
            "(there is no CPU fallback for device='cuda')")
    lib = C.CDLL(path)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name in STRING_FUNCS:
        fn = getattr(lib, name)
        fn.argtypes = []
        fn.restype = C.c_char_p
    _LIB = lib
    return lib


def lib() -> C.CDLL:
    return _LIB if _LIB is not None else load_library()


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().mt_last_error().decode(errors="replace")
        raise MicrotorchB200Error(f"{what or 'libmicrotorch_b200'} failed (status {status}): {msg}")


def device_count() -> int:
    """Number of visible CUDA devices; 0 when there is no driver/GPU (never raises for that)."""
    try:
        n = C.c_int(0)
        if lib().mt_device_count(C.byref(n)) != 0:
            return 0
        return n.value
    except (MicrotorchB200Error, OSError):
        return 0


def require_device(device: Optional[int] = None) -> None:
    """Initialise the library on this process's GPU (LOCAL_RANK by default)."""
    global _DEVICE_READY
    if _DEVICE_READY:
        return
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    check(lib().mt_init(device), "mt_init")
    _DEVICE_READY = True
    level = os.environ.get("MICROTORCH_B200_MATMUL_PRECISION")
    if level:
        set_float32_matmul_precision(level)


MATMUL_PRECISIONS = {"high": 0, "higher": 1, "highest": 2}


def set_float32_matmul_precision(level: str) -> None:
    """fp32 emulation level of the tensor-core GEMMs: "high" (fastest, opt-in), "higher" (default: 3xTF32 forward
    GEMMs), "highest" (SGEMM-class everywhere).  See mt_gemm_set_precision in the header for the measured errors and
    costs."""
    if level not in MATMUL_PRECISIONS:
        raise ValueError(f"matmul precision must be one of {sorted(MATMUL_PRECISIONS)}, got {level!r}")
    check(lib().mt_gemm_set_precision(MATMUL_PRECISIONS[level]), "mt_gemm_set_precision")


def device_ready() -> bool:
    return _DEVICE_READY


@functools.lru_cache(maxsize=8192)
def _i64_tuple(values: tuple):
    return (C.c_int64 * max(len(values), 1))(*values)


def i64(values: Sequence[int]):
    """int64[] argument for shapes / strides.  The arrays are read-only on both sides of the ABI, so one ctypes array
    per distinct tuple is built once and shared (this call sits on the per-launch host path)."""
    return _i64_tuple(values if type(values) is tuple else tuple(values))
