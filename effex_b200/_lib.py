"""ctypes binding of libeffex_fx.so (C ABI declared in include/effex_fx.h).

There is no fallback: if the shared library cannot be loaded this module
raises, and every product entry point that needs the GPU fails with it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EFFEX_FX_LIB") or os.path.join(HERE, "libeffex_fx.so")   # env: experiment variants only

FX_OK = 0
FX_ERR_INVALID = -1
FX_ERR_CUDA = -2
FX_ERR_UNSUPPORTED = -3
FX_ERR_STATE = -4
FX_ERR_COMM = -5
FX_ABI_VERSION = 2            # must equal FX_ABI_VERSION of include/effex_fx.h (checked in load())
FX_COMM_TOKEN_BYTES = 128
FX_FLAG_FORCE_GENERIC = 1
FX_FLAG_LOCKSTEP_KERNEL = 2
FX_FLAG_CROSS_ONLY = 4


class FxConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("ntaps", C.c_int32), ("nbins", C.c_int32),
                ("dc_remove", C.c_int32), ("num_samp", C.c_int64), ("max_blocks", C.c_int32),
                ("flags", C.c_int32)]


# name -> (restype, argtypes); mirrors include/effex_fx.h one to one
_VP = C.c_void_p
SIGNATURES = {
    "fx_abi_version": (C.c_int, []),
    "fx_device_count": (C.c_int, []),
    "fx_create": (C.c_int, [C.POINTER(FxConfig), C.POINTER(_VP)]),
    "fx_destroy": (C.c_int, [_VP]),
    "fx_last_error": (C.c_char_p, [_VP]),
    "fx_sync": (C.c_int, [_VP]),
    "fx_uses_fused": (C.c_int, [_VP]),
    "fx_set_taps": (C.c_int, [_VP, C.POINTER(C.c_double), C.c_size_t]),
    "fx_set_rot": (C.c_int, [_VP, C.POINTER(C.c_double), C.c_size_t]),
    "fx_process": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP]),
    "fx_integrate": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP]),
    "fx_process_acc": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "fx_span_sums": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.POINTER(C.c_uint64)]),
    "fx_integrate_stream": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, C.POINTER(C.c_uint64), C.c_int64, _VP, _VP, _VP, _VP]),
    "fx_comm_export": (C.c_int, [_VP, C.c_int, C.c_size_t, _VP]),
    "fx_comm_attach": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "fx_comm_fence": (C.c_int, [_VP]),
    "fx_process_reduce": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP, C.c_int, _VP]),
    "fx_integrate_stream_reduce": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, C.POINTER(C.c_uint64), C.c_int64, C.c_int, _VP]),
    "fx_reduce_f64": (C.c_int, [_VP, _VP, C.c_size_t, C.c_int]),
    "fx_reduce_f32": (C.c_int, [_VP, _VP, C.c_size_t, C.c_int]),
    "fx_process_host": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP, _VP]),
    "fx_copy_probe": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP]),
    "fx_pfb_c64": (C.c_int, [_VP, _VP, _VP]),
    "fx_pfb_u8": (C.c_int, [_VP, _VP, _VP]),
    "fx_lag_c64": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "fx_lag_u8": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "fx_lag_fft_len": (C.c_int64, [_VP]),
    "fx_lag_accumulate_u8": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, C.c_int]),
    "fx_lag_accumulate_c64": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, C.c_int]),
    "fx_lag_finish": (C.c_int, [_VP, _VP, C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "fx_lag_finish_async": (C.c_int, [_VP, _VP, _VP, _VP]),
    "fx_csv_rows_bound": (C.c_size_t, [C.c_int64, C.c_int64]),
    "fx_csv_format_rows": (C.c_int, [_VP, C.c_int64, C.c_int64, C.c_int, _VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "fx_csv_format_double": (C.c_int, [C.c_double, C.c_int, C.c_char_p]),
    "fx_dev_alloc": (C.c_int, [_VP, C.c_size_t, C.POINTER(_VP)]),
    "fx_dev_free": (C.c_int, [_VP, _VP]),
    "fx_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(_VP)]),
    "fx_host_free_pinned": (C.c_int, [_VP]),
    "fx_memcpy_h2d": (C.c_int, [_VP, _VP, _VP, C.c_size_t]),
    "fx_memcpy_d2h": (C.c_int, [_VP, _VP, _VP, C.c_size_t]),
    "fx_memset": (C.c_int, [_VP, _VP, C.c_int, C.c_size_t]),
    "fx_reset_counters": (C.c_int, [_VP]),
    "fx_kernel_launches": (C.c_int64, [_VP]),
    "fx_enable_timing": (C.c_int, [_VP, C.c_int]),
    "fx_dominant_kernel_time": (C.c_int, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fx_stream": (_VP, [_VP]),
    "fx_stream_aux": (_VP, [_VP]),
}

_lib = None


def load() -> C.CDLL:
    """Load libeffex_fx.so (building it first if nvcc is here and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    default_lib = "EFFEX_FX_LIB" not in os.environ
    if not os.path.exists(LIB_PATH):
        from . import build as _build          # raises if nvcc is missing
        _build.build()
    elif default_lib:
        # a stale binary would be called through mismatched argtypes: rebuild when the sources are newer
        # and a compiler is here (a GPU box gets the prebuilt .so and usually the sources with it)
        from . import build as _build
        try:
            if _build._stale():
                _build.build()
        except Exception:                      # no nvcc, read-only tree, ...: the ABI version check below is the guard
            pass
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:                       # pragma: no cover - environment specific
        raise RuntimeError(
            f"libeffex_fx.so could not be loaded ({e}); the FX hot path has no CPU fallback") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)                # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    got = lib.fx_abi_version()
    if got != FX_ABI_VERSION:
        raise RuntimeError(f"libeffex_fx.so has ABI version {got}, this package binds version {FX_ABI_VERSION}: "
                           "rebuild with `python -m effex_b200.build --force`")
    _lib = lib
    return lib
