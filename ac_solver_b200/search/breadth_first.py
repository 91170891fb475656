"""Breadth-first search of the AC graph -- drop-in for the reference's
``ac_solver/search/breadth_first.py`` (``bfs``), executed entirely on the GPU
(csrc/pbfs.cu, the partitioned engine with a world of one): same return value, same visited set in the same order at any node budget,
same console output.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ..envs.utils import is_array_valid_presentation


def _raise_for(status):
    if status == _lib.ROW_ASSERT:
        raise AssertionError("a move produced an invalid presentation (empty relator): "
                             "the reference raises AssertionError here (envs/utils.py:261-263)")
    if status == _lib.ROW_INDEX:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")


def bfs_device(presentation, max_nodes_to_explore=10000, cyclically_reduce_after_moves=False,
               want_visited=False, device=None):
    """Run the device search -> (solved, path|None, info).  ``info`` carries the counters the
    reference keeps implicitly (visited, expanded, moves, successive minimal lengths) and,
    on request, the visited states in insertion order as int8 rows."""
    L = _lib.lib()
    dev = _lib.default_device() if device is None else device
    p = np.asarray(presentation)
    if p.ndim != 1 or p.size % 2 or p.size == 0:
        raise AssertionError(f"{presentation} is not a valid presentation")
    if p.size and (np.abs(p).max() > 2):
        raise ValueError("the GPU search supports the two-generator alphabet {+-1, +-2} only")
    p8 = np.ascontiguousarray(p, dtype=np.int8)
    mrl = p8.size // 2
    h = C.c_void_p()
    _lib.check(L.acs_bfs_create(None, dev, mrl, int(max_nodes_to_explore), int(bool(cyclically_reduce_after_moves)),
                                C.byref(h)))
    try:
        cap = 1 << 16
        path = np.zeros((cap, 2), np.int32)
        res = _lib.SearchResult()
        _lib.check(L.acs_bfs_run(h, p8.ctypes.data, path.ctypes.data, cap, C.byref(res)))
        info = {
            "n_visited": int(res.n_visited), "n_expanded": int(res.n_expanded), "n_moves": int(res.n_moves),
            "frontier_left": int(res.frontier_left), "budget_hit": bool(res.budget_hit),
            "n_levels": int(res.n_levels), "status": int(res.status),
            "minlen_log": [int(res.minlen_log[i]) for i in range(res.n_minlen)],
            "seconds_device": float(res.seconds_device),
        }
        if want_visited and res.status == 0:
            vis = np.zeros((max(int(res.n_visited), 1), 2 * mrl), np.int8)
            n_out = C.c_int64(0)
            _lib.check(L.acs_bfs_visited(h, vis.ctypes.data, vis.shape[0], C.byref(n_out)))
            info["visited"] = vis[: n_out.value]
    finally:
        L.acs_bfs_destroy(h)
    _raise_for(res.status)
    plist = [(int(a), int(l)) for a, l in path[: res.path_len]] if res.solved else None
    return bool(res.solved), plist, info


def bfs(presentation, max_nodes_to_explore=10000, verbose=False, cyclically_reduce_after_moves=False):
    """search/breadth_first.py:15-97.

    Returns ``(True, path)`` with ``path = [(-1, L0), (action, total_length), ...]`` ending in a
    state of total length 2, or ``(False, None)`` when the node budget is exhausted."""
    assert is_array_valid_presentation(presentation), f"{presentation} is not a valid presentation"
    solved, path, info = bfs_device(presentation, max_nodes_to_explore, cyclically_reduce_after_moves)
    if verbose:
        for m in info["minlen_log"]:
            print(f"New minimal length found: {m}")
    if solved:
        return True, path
    if info["budget_hit"]:
        print(f"Exiting search as number of explored nodes = {info['n_visited']} has exceeded the limit "
              f"{max_nodes_to_explore}")
    return False, None
