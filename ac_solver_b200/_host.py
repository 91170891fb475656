"""Thin numpy <-> C-ABI glue for the *_host entry points (host buffers in, host buffers out)."""

from __future__ import annotations

import numpy as np

from . import _lib


def _i8(a):
    a = np.asarray(a)
    if a.dtype != np.int8:
        if a.size and (a.min() < -128 or a.max() > 127):
            raise ValueError("letters must fit in int8")
        a = a.astype(np.int8)
    return np.ascontiguousarray(a)


def generic_call(op, rows, actions=None, i=0, j=1, sign=1, cyclical=True, device=None):
    """Run one generic byte-domain op over ``rows`` [n, width] -> (out, aux, status)."""
    L = _lib.lib()
    ctx = _lib.ctx(_lib.default_device() if device is None else device)
    rows = _i8(rows)
    n, width = rows.shape
    out = np.empty_like(rows)
    per = 2 if op in (_lib.OP_ACMOVE, _lib.OP_SIMPLIFY_PRESENTATION) else 1
    aux = np.zeros((n, per) if per == 2 else (n,), dtype=np.int32)
    status = np.zeros(n, dtype=np.uint8)
    act = None
    if actions is not None:
        act = np.ascontiguousarray(actions, dtype=np.uint8)
    _lib.check(
        L.acs_generic_host(
            ctx, int(op), rows.ctypes.data, None if act is None else act.ctypes.data, out.ctypes.data,
            aux.ctypes.data, status.ctypes.data, n, width, int(i), int(j), int(sign), int(bool(cyclical)),
        )
    )
    return out, aux, status


def validate_rows(rows, device=None):
    """flags per row: bit0 valid presentation, bit1 alphabet {0,+-1,+-2}, bit2 right-padded."""
    L = _lib.lib()
    ctx = _lib.ctx(_lib.default_device() if device is None else device)
    rows = _i8(rows)
    n, width = rows.shape
    flags = np.zeros(n, dtype=np.uint8)
    _lib.check(L.acs_validate_batch_host(ctx, rows.ctypes.data, flags.ctypes.data, n, width // 2))
    return flags
