"""ctypes loader of libacsolver_b200.so -- the only bridge between the Python host code
and the CUDA kernels.  There is NO CPU fallback: if the library is missing, or no CUDA
device is visible, every compute entry point raises."""

from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libacsolver_b200.so")

ACS_OK = 0
ROW_OK, ROW_ASSERT, ROW_INDEX = 0, 1, 2
OP_ACMOVE, OP_CONCAT_RAW, OP_CONJ_RAW, OP_SIMPLIFY_RELATOR, OP_SIMPLIFY_PRESENTATION = range(5)
FLAG_CYCLICAL, FLAG_NORMALIZED = 1, 2


class AcsError(RuntimeError):
    pass


class SearchResult(C.Structure):
    _fields_ = [
        ("solved", C.c_int32),
        ("status", C.c_int32),
        ("budget_hit", C.c_int32),
        ("path_len", C.c_int32),
        ("n_visited", C.c_int64),
        ("n_expanded", C.c_int64),
        ("n_moves", C.c_int64),
        ("frontier_left", C.c_int64),
        ("n_levels", C.c_int32),
        ("n_minlen", C.c_int32),
        ("minlen_log", C.c_int32 * 128),
        ("seconds_device", C.c_double),
    ]


_lib = None
_lock = threading.Lock()
_ctx = {}

_P = C.c_void_p
_SIGS = {
    "acs_version": (C.c_int, []),
    "acs_last_error": (C.c_char_p, []),
    "acs_device_count": (C.c_int, []),
    "acs_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "acs_ctx_destroy": (None, [_P]),
    "acs_moves_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "acs_moves_batch_host": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int]),
    "acs_env_step_batch": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _P]),
    "acs_env_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_int64)]),
    "acs_validate_batch": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    "acs_validate_batch_host": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int]),
    "acs_generic_batch": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, _P]),
    "acs_generic_host": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int]),
    "acs_bfs_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(_P)]),
    "acs_bfs_run": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(SearchResult)]),
    "acs_bfs_visited": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "acs_bfs_destroy": (None, [_P]),
    "acs_greedy_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "acs_greedy_run": (C.c_int, [_P, _P, _P, _P]),
    "acs_greedy_visited": (C.c_int, [_P, C.c_int, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "acs_greedy_destroy": (None, [_P]),
}


def exported_symbols():
    """Names include/acsolver_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGS)


def lib():
    """Load the shared library (no device needed for this step)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise AcsError(
                    f"{LIB_PATH} is missing: build it with `python -m ac_solver_b200.build` "
                    "(there is no CPU fallback for the AC-move kernels)"
                )
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc: int):
    if rc != ACS_OK:
        msg = lib().acs_last_error()
        raise AcsError(f"libacsolver_b200 error {rc}: {msg.decode() if msg else ''}")


def ctx(device: int = 0):
    """Per-device context (streams + scratch).  Raises without a GPU."""
    L = lib()
    with _lock:
        if device not in _ctx:
            h = _P()
            rc = L.acs_ctx_create(int(device), C.byref(h))
            if rc != ACS_OK:
                msg = L.acs_last_error()
                raise AcsError(f"cannot create a CUDA context on device {device}: {msg.decode() if msg else rc}")
            _ctx[device] = h
        return _ctx[device]


def default_device() -> int:
    """LOCAL_RANK under torchrun, else 0."""
    return int(os.environ.get("LOCAL_RANK", "0"))
