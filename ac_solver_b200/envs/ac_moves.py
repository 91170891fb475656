"""The 12 AC' moves -- drop-in for the reference's ``ac_solver/envs/ac_moves.py``.

Single-presentation calls go through the generic byte kernel (any integer alphabet, exactly
the reference's array semantics); the batched entry point ``ac_moves_batch`` runs the packed
TMA-staged kernel that is the hot path (csrc/moves_kernel.cu).
"""

from __future__ import annotations

import numpy as np

from .. import _lib
from .._host import _i8, generic_call, validate_rows


def _raw(op, presentation, max_relator_length, i, j, sign):
    p = np.asarray(presentation)
    assert p.size == 2 * max_relator_length
    out, aux, _ = generic_call(op, _i8(p)[None, :], i=i, j=j, sign=sign)
    if int(aux[0]) == -2:  # the reference indexes relator_nonzero[0] on an empty array
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    return out[0].astype(p.dtype), int(aux[0])


def concatenate_relators(presentation, max_relator_length, i, j, sign, lengths):
    """envs/ac_moves.py:4-76 -- r_i <- r_i r_j^{sign}; rejected (unchanged) if longer than mrl.
    Like the reference, an accepted move updates the caller's ``lengths`` in place."""
    assert all([i in [0, 1], j in [0, 1], i == 1 - j]), f"expect i and j to be 0 or 1 and i != j; got i = {i}, j = {j}"
    assert sign in [1, -1], f"expect sign to be +1 or -1, received {sign}"
    out, new_size = _raw(_lib.OP_CONCAT_RAW, presentation, max_relator_length, i, j, sign)
    if new_size >= 0:
        lengths[i] = new_size
    return out, lengths


def conjugate(presentation, max_relator_length, i, j, sign, lengths):
    """envs/ac_moves.py:79-156 -- r_i <- x_j^{sign} r_i x_j^{-sign}."""
    assert all([i in [0, 1], j in [1, 2]]), f"expect i to be 0 and 1 and j to be 1 or 2; got i = {i}, j = {j}"
    assert sign in [1, -1], f"expect sign to be +1 or -1, received {sign}"
    out, new_size = _raw(_lib.OP_CONJ_RAW, presentation, max_relator_length, i, j, sign)
    if new_size >= 0:
        lengths = lengths.copy()
        lengths[i] = new_size
    return out, lengths


def ACMove(move_id, presentation, max_relator_length, lengths, cyclical=True):
    """envs/ac_moves.py:159-231.  ``lengths`` is accepted and ignored exactly like the
    reference (both output lengths are recomputed from the data)."""
    assert move_id in range(0, 12), f"Expect n to be in range 0-11 (both inclusive); got {move_id}"
    p = np.asarray(presentation)
    assert p.size == 2 * max_relator_length
    out, aux, status = generic_call(_lib.OP_ACMOVE, _i8(p)[None, :], actions=[move_id], cyclical=cyclical)
    if status[0] == _lib.ROW_ASSERT:
        raise AssertionError(f"{p} is not a valid presentation. Expect all zeros to be padded to the right.")
    if status[0] == _lib.ROW_INDEX:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    return out[0].astype(p.dtype), [int(aux[0, 0]), int(aux[0, 1])]


def ac_moves_batch(states, actions, cyclical=True, validate=True, normalized=False, device=None):
    """Batched ACMove over HOST arrays: states [N, 2*mrl] int8, actions [N] in 0..11 ->
    (next_states int8 [N, 2*mrl], lengths uint8 [N,2], status uint8 [N]).

    status[k] != 0 marks rows for which the reference raises (1 AssertionError, 2 IndexError);
    those rows are returned unchanged.  Runs the packed kernel through ``acs_moves_batch_host``
    (chunked H2D -> kernel -> D2H pipeline).  ``validate`` checks once, on the GPU, that rows
    are right-padded words over {+-1,+-2}.  ``normalized=True`` promises that every word is
    already a normal form for ``cyclical`` (as all outputs of this function are) and selects
    the steady-state kernel variant that skips re-simplifying the untouched relator."""
    L = _lib.lib()
    ctx = _lib.ctx(_lib.default_device() if device is None else device)
    s = _i8(states)
    if s.ndim != 2 or s.shape[1] % 2:
        raise ValueError("states must be [N, 2*max_relator_length]")
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    n, w = s.shape
    if a.shape != (n,):
        raise ValueError("actions must be [N]")
    if validate and n:
        flags = validate_rows(s, device)
        if not ((flags & 6) == 6).all():
            bad = int(np.flatnonzero((flags & 6) != 6)[0])
            raise ValueError(f"row {bad} is not a right-padded word pair over {{+-1,+-2}}")
    if w // 2 > 64:  # beyond the packed kernels' width: the generic byte kernel (any width <= 127)
        out, aux, status = generic_call(_lib.OP_ACMOVE, s, actions=a, cyclical=cyclical, device=device)
        return out, aux.astype(np.uint8), status
    out = np.empty_like(s)
    lens = np.zeros((n, 2), np.uint8)
    status = np.zeros(n, np.uint8)
    _lib.check(
        L.acs_moves_batch_host(ctx, s.ctypes.data, a.ctypes.data, out.ctypes.data, lens.ctypes.data,
                               status.ctypes.data, n, w // 2,
                               (_lib.FLAG_CYCLICAL if cyclical else 0) | (_lib.FLAG_NORMALIZED if normalized else 0))
    )
    return out, lens, status
