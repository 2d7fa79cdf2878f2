"""ctypes binding of libivl_b200.so (the C ABI in include/ivl_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libivl_b200.so")

IVL_ERR_LAUNCH = -5
IVL_DTYPE_F32 = 0
IVL_DTYPE_BF16 = 1

_lib = None

# name -> (restype, argtypes); must list every symbol include/ivl_b200.h declares
SIGNATURES = {
    "ivl_abi_version": (c_int, []),
    "ivl_strerror": (c_char_p, [c_int]),
    "ivl_last_cuda_error": (c_char_p, []),
    "ivl_stream_init": (c_int, [c_void_p]),
    "ivl_stream_release": (c_int, [c_void_p]),
    "ivl_gdn_chunk_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ivl_gdn_chunk_fwd": (c_int, [c_void_p] * 6 + [c_int, c_void_p, c_void_p, c_int] + [c_int] * 5
                          + [c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "ivl_gdn_chunk_fwd_fused": (c_int, [c_void_p] * 14 + [c_int, c_void_p, c_void_p, c_int] + [c_int] * 5 + [c_float, c_void_p, c_size_t, c_void_p]),
    "ivl_gdn_chunk_fwd_varlen": (c_int, [c_void_p] * 6 + [c_int, c_void_p, c_void_p, c_int] + [c_int] * 4
                                 + [c_float, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t,
                                    c_void_p]),
    "ivl_gdn_chunk_prep": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "ivl_gdn_chunk_scan": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int] + [c_int] * 3
                           + [c_void_p, c_size_t, c_void_p]),
    "ivl_gdn_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ivl_gdn_bwd": (c_int, [c_void_p] * 14 + [c_int] * 5 + [c_float, c_void_p, c_size_t, c_void_p]),
    "ivl_gdn_decode_step": (c_int, [c_void_p] * 16 + [c_int, c_void_p] + [c_int] * 4 + [c_float, c_float, c_void_p]),
    "ivl_gdn_recurrent_fwd": (c_int, [c_void_p] * 6 + [c_int, c_void_p, c_void_p, c_int] + [c_int] * 5
                              + [c_float, c_int, c_void_p]),
    "ivl_swa_fwd": (c_int, [c_void_p] * 8 + [c_int] * 7 + [c_float, c_void_p]),
    "ivl_ipc_export": (c_int, [c_void_p, c_void_p, ctypes.POINTER(ctypes.c_uint64)]),
    "ivl_ipc_open": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "ivl_ipc_close": (c_int, [c_void_p]),
    "ivl_peer_put": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, ctypes.c_uint32, c_void_p, c_void_p]),
    "ivl_stream_wait_value32": (c_int, [c_void_p, c_void_p, ctypes.c_uint32]),
    "ivl_swa_fwd_varlen": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ivl_swa_fwd_pos": (c_int, [c_void_p] * 8 + [c_int] * 7 + [c_float, ctypes.c_int64, c_void_p]),
    "ivl_swa_decode_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ivl_swa_decode_fwd": (c_int, [c_void_p] * 6 + [c_int] * 6 + [c_float, c_void_p, c_size_t, c_void_p]),
    "ivl_swa_ring_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ivl_swa_ring_state_bytes": (c_size_t, [c_int, c_int]),
    "ivl_swa_ring_append": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "ivl_swa_ring_decode": (c_int, [c_void_p] * 9 + [c_int] * 6 + [c_float, c_void_p, c_size_t, c_void_p]),
    "ivl_swa_ring_fwd": (c_int, [c_void_p] * 7 + [c_int] * 7 + [c_float, c_void_p]),
    "ivl_short_conv_fwd": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "ivl_short_conv_fwd_varlen": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ivl_gdn_gate_fwd": (c_int, [c_void_p] * 6 + [ctypes.c_int64, c_int, c_void_p]),
    "ivl_rmsnorm_gated_fwd": (c_int, [c_void_p] * 4 + [ctypes.c_int64, c_int, c_float, c_void_p]),
    "ivl_mrope_apply": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p]),
}


class IvlError(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load the shared library (building it in-tree first if it is absent and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise IvlError(f"{LIB_PATH} not found: run `python -m infinitevl_b200.build` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        lib = load()
        msg = lib.ivl_strerror(code).decode()
        detail = lib.ivl_last_cuda_error().decode() if code == IVL_ERR_LAUNCH else ""
        raise IvlError(f"{what} failed: {msg} (code {code})" + (f": {detail}" if detail else ""))
