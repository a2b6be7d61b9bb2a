"""ctypes binding of libdctn_b200.so (the C ABI declared in include/dctn_b200.h).

The library is loaded lazily on first use and the product path fails loudly when it is missing —
there is deliberately no fallback implementation.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_int, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCTN_B200_LIB") or os.path.join(_HERE, "libdctn_b200.so")   # override: A/B comparisons of two builds

F32, F64 = 0, 1
VARIANT_AUTO, VARIANT_FFMA, VARIANT_TC3, VARIANT_TC1, VARIANT_DIRECT, VARIANT_TCH3 = 0, 1, 2, 3, 4, 5
VARIANTS = {"auto": 0, "ffma": 1, "tc3": 2, "tc1": 3, "direct": 4, "tch3": 5}
WS_FORWARD, WS_BACKWARD_CORE, WS_BACKWARD_INPUT, WS_BACKWARD_INPUT_SAVED, WS_FORWARD_STATS = 0, 1, 2, 3, 4
FAMILY_CUDA_CORE, FAMILY_TCGEN05, FAMILY_STREAMING = 1, 2, 4

# every symbol include/dctn_b200.h declares: (name, restype, argtypes)
SYMBOLS = {
    "dctn_version": (c_int, []),
    "dctn_last_error": (c_char_p, []),
    "dctn_launch_count": (c_ulonglong, []),
    "dctn_eps_plan_get": (c_void_p, [c_int] * 6),
    "dctn_eps_plan_describe": (c_char_p, [c_void_p]),
    "dctn_eps_kernel_family": (c_int, [c_void_p, c_int, c_int, c_int, c_int]),
    "dctn_eps_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "dctn_eps_forward": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "dctn_eps_backward_core": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "dctn_eps_backward_input": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "dctn_eps_saved_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "dctn_eps_forward_train": (c_int, [c_void_p] * 5 + [c_size_t] + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "dctn_eps_backward_input_saved": (c_int, [c_void_p] * 5 + [c_size_t, c_void_p] + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "dctn_eps_forward_stats": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "dctn_window_stats_workspace_bytes": (c_size_t, [c_int] * 4),
    "dctn_window_stats": (c_int, [c_void_p] + [c_int] * 7 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "dctn_eps_forward_from_pixels": (c_int, [c_void_p, c_void_p, ctypes.c_double, c_void_p, c_void_p] + [c_int] * 3 + [c_void_p]),
    "dctn_logmatmulexp_forward": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "dctn_logmatmulexp_backward": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_void_p]),
    "dctn_logmatmulexp_workspace_bytes": (c_size_t, [c_int] * 4),
    "dctn_logmatmulexp_forward_ws": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p, c_size_t, c_void_p]),
    "dctn_logmatmulexp_backward_ws": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_void_p, c_size_t, c_void_p]),
    "dctn_logmatmulexp_batched_forward": (c_int, [c_void_p] * 3 + [ctypes.c_longlong] + [c_int] * 4 + [c_void_p]),
    "dctn_logmatmulexp_batched_backward": (c_int, [c_void_p] * 6 + [ctypes.c_longlong] + [c_int] * 4 + [c_void_p]),
    "dctn_eps_forward_host_device_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "dctn_eps_forward_host": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
}

_lib = None
_lock = threading.Lock()


class DctnLibraryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Returns the loaded library, loading it on first call."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise DctnLibraryError(
                        f"{LIB_PATH} is not built. Run `make -C dctn_b200/csrc` (or __graft_entry__.build()). "
                        "dctn_b200 has no CPU or PyTorch fallback."
                    )
                handle = ctypes.CDLL(LIB_PATH)
                for name, (restype, argtypes) in SYMBOLS.items():
                    if os.environ.get("DCTN_B200_LIB") and not hasattr(handle, name):
                        continue  # A/B runs against an OLDER build (tools/kbench.py): it may predate newer entry points
                    fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
                    fn.restype = restype
                    fn.argtypes = argtypes
                _lib = handle
    return _lib


def last_error() -> str:
    return lib().dctn_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")


def launch_count() -> int:
    return int(lib().dctn_launch_count())
