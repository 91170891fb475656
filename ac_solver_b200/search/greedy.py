"""Greedy (best-first) search of the AC graph -- drop-in for the reference's
``ac_solver/search/greedy.py`` (``greedy_search``), executed on the GPU (csrc/greedy.cu), plus
the batched form used for the Miller-Schupp sweep (one CTA per presentation, one launch: a whole
(length, depth) bucket of the frontier is expanded per round, csrc/greedy_bucket.cuh).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .breadth_first import _raise_for


def greedy_search_batch(presentations, max_nodes_to_explore=10000, cyclically_reduce_after_moves=False,
                        want_visited=False, device=None, path_cap=None):
    """Run greedy search on every row of ``presentations`` [S, 2*mrl] (same mrl) concurrently.

    Returns a list of ``(solved, path, info)`` per row, each identical to what the reference's
    ``greedy_search`` returns for that row (``path`` is the reference's failure path when not
    solved).  Rows whose first move raises in the reference carry ``info["status"] != 0``."""
    L = _lib.lib()
    dev = _lib.default_device() if device is None else device
    P = np.asarray(presentations)
    if P.ndim != 2 or P.shape[1] % 2 or P.shape[1] == 0:
        raise ValueError("presentations must be [S, 2*max_relator_length]")
    if P.size and np.abs(P).max() > 2:
        raise ValueError("the GPU search supports the two-generator alphabet {+-1, +-2} only")
    P8 = np.ascontiguousarray(P, dtype=np.int8)
    S, w = P8.shape
    mrl = w // 2
    budget = int(max_nodes_to_explore)
    if path_cap is None:
        path_cap = int(min(budget + 4, 1 << 14))
    path_cap = max(path_cap, 4)
    h = C.c_void_p()
    _lib.check(L.acs_greedy_create(dev, S, mrl, budget, int(bool(cyclically_reduce_after_moves)), path_cap,
                                   C.byref(h)))
    out = []
    try:
        paths = np.zeros((S, path_cap, 2), np.int32)
        res = (_lib.SearchResult * S)()
        _lib.check(L.acs_greedy_run(h, P8.ctypes.data, paths.ctypes.data, res))
        for s in range(S):
            r = res[s]
            info = {
                "n_visited": int(r.n_visited), "n_expanded": int(r.n_expanded), "n_moves": int(r.n_moves),
                "frontier_left": int(r.frontier_left), "budget_hit": bool(r.budget_hit), "status": int(r.status),
                "minlen_log": [int(r.minlen_log[i]) for i in range(r.n_minlen)],
                "seconds_device": float(r.seconds_device),
                # bucket rounds of the CTA-per-search kernel; -1: served by the one-warp heap kernel
                "rounds": int(r.n_levels),
            }
            if r.path_len > path_cap:  # deeper than the buffer: the reference has no limit -- run again with room
                L.acs_greedy_destroy(h)
                h = None
                return greedy_search_batch(presentations, max_nodes_to_explore, cyclically_reduce_after_moves, want_visited,
                                           device, path_cap=int(max(res[i].path_len for i in range(S))) + 4)
            if want_visited and r.status == 0:
                vis = np.zeros((max(int(r.n_visited), 1), w), np.int8)
                n_out = C.c_int64(0)
                _lib.check(L.acs_greedy_visited(h, s, vis.ctypes.data, vis.shape[0], C.byref(n_out)))
                info["visited"] = vis[: n_out.value]
            plist = [(int(a), int(l)) for a, l in paths[s, : r.path_len]]
            out.append((bool(r.solved), plist, info))
    finally:
        if h is not None:
            L.acs_greedy_destroy(h)
    return out


def greedy_search_groups(groups, max_nodes_to_explore=10000, cyclically_reduce_after_moves=False, device=None,
                         path_cap=4096, threads=8):
    """Sweep helper: ``groups`` is a list of presentation arrays [S_k, 2*mrl_k] (one array per max_relator_length).
    All engines are CREATED first, then the groups search concurrently (one stream and one host thread each), then
    the engines are destroyed -- ``cudaMalloc`` stalls behind kernels that are already running (measured: 0.02-1.5 s
    per engine when interleaved with other groups' searches, 5 ms when not), so interleaving allocation and search
    costs a sweep more than the searches themselves.  Returns a list (per group) of ``greedy_search_batch`` results."""
    from concurrent.futures import ThreadPoolExecutor

    L = _lib.lib()
    dev = _lib.default_device() if device is None else device
    budget = int(max_nodes_to_explore)
    arrays, handles = [], []
    try:
        for P in groups:
            P8 = np.ascontiguousarray(np.asarray(P), dtype=np.int8)
            if P8.ndim != 2 or P8.shape[1] % 2 or P8.shape[1] == 0:
                raise ValueError("every group must be [S, 2*max_relator_length]")
            if P8.size and np.abs(P8).max() > 2:
                raise ValueError("the GPU search supports the two-generator alphabet {+-1, +-2} only")
            h = C.c_void_p()
            _lib.check(L.acs_greedy_create(dev, P8.shape[0], P8.shape[1] // 2, budget, int(bool(cyclically_reduce_after_moves)),
                                           int(path_cap), C.byref(h)))
            arrays.append(P8)
            handles.append(h)

        def run(k):
            P8, h = arrays[k], handles[k]
            S = P8.shape[0]
            paths = np.zeros((S, path_cap, 2), np.int32)
            res = (_lib.SearchResult * S)()
            _lib.check(L.acs_greedy_run(h, P8.ctypes.data, paths.ctypes.data, res))
            return paths, res

        with ThreadPoolExecutor(max_workers=max(1, min(threads, len(arrays) or 1))) as pool:
            raw = list(pool.map(run, range(len(arrays))))
    finally:
        for h in handles:
            L.acs_greedy_destroy(h)
    out = []
    for k, (paths, res) in enumerate(raw):
        S = arrays[k].shape[0]
        if any(res[s].path_len > path_cap for s in range(S)):  # deeper than the buffer: redo this group with room
            out.append(greedy_search_batch(arrays[k], budget, cyclically_reduce_after_moves, device=dev,
                                           path_cap=int(max(res[s].path_len for s in range(S))) + 4))
            continue
        rows = []
        for s in range(S):
            r = res[s]
            info = {"n_visited": int(r.n_visited), "n_expanded": int(r.n_expanded), "n_moves": int(r.n_moves),
                    "frontier_left": int(r.frontier_left), "budget_hit": bool(r.budget_hit), "status": int(r.status),
                    "minlen_log": [int(r.minlen_log[i]) for i in range(r.n_minlen)], "seconds_device": float(r.seconds_device),
                    "rounds": int(r.n_levels)}
            rows.append((bool(r.solved), [(int(a), int(l)) for a, l in paths[s, : r.path_len]], info))
        out.append(rows)
    return out


def greedy_search(presentation, max_nodes_to_explore=10000, verbose=False, cyclically_reduce_after_moves=False):
    """search/greedy.py:15-121.

    Returns ``(True, path)`` or ``(False, path_of_last_expanded_node + [(11, length)])`` exactly as
    the reference does (its failure value is a path, not ``None``)."""
    p = np.array(presentation, dtype=np.int8)
    mrl = p.size // 2
    solved, path, info = greedy_search_batch(p[None, :], max_nodes_to_explore, cyclically_reduce_after_moves)[0]
    if verbose:
        for m in info["minlen_log"]:
            print(f"New minimal length found: {m}")
    _raise_for(info["status"])
    if solved:
        if verbose:  # greedy.py:92-99: the three lines of the reference, from a replay of the path
            from ..envs.ac_moves import ACMove

            state, lens = p.copy(), [0, 0]
            for action, _ in path[1:]:
                state, lens = ACMove(action, state, mrl, lens, cyclical=cyclically_reduce_after_moves)
            print(f"Found {state[0:1], state[mrl:mrl + 1]} after exploring "
                  f"{info['n_visited'] - info['frontier_left']} nodes")
            print(f"Path to a trivial state: (tuples are of form (action, length of a state)) {path}")
            print(f"Total path length: {len(path)}")
        return True, path
    if info["budget_hit"]:
        print(f"Exiting search as number of explored nodes = {info['n_visited']} has exceeded the limit "
              f"{max_nodes_to_explore}")
    return False, path
