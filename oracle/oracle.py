"""ctypes binding of the CPU ORACLE (oracle/ac_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, from bench.py's cpu_baseline /
``--impl reference`` leg and from ``__graft_entry__.smoke()`` -- never from the product
package ``ac_solver_b200``.  Parity status: pinned, see ac_oracle.h.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libac_oracle.so")

ACO_OK, ACO_ASSERT, ACO_INDEX = 0, 1, 2


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
        ("n_minlen", C.c_int32),
        ("minlen_log", C.c_int32 * 128),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile (system gcc)."""
    src = os.path.join(_HERE, "ac_oracle.c")
    hdr = os.path.join(_HERE, "ac_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libac_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i8p, u8p, i32p = C.c_void_p, C.c_void_p, C.c_void_p
        L.aco_simplify_relator.argtypes = [i8p, C.c_int, C.c_int]
        L.aco_simplify_relator.restype = C.c_int
        L.aco_is_valid_presentation.argtypes = [i8p, C.c_int]
        L.aco_is_valid_presentation.restype = C.c_int
        for f in (L.aco_concatenate_relators, L.aco_conjugate):
            f.argtypes = [i8p, C.c_int, C.c_int, C.c_int, C.c_int]
            f.restype = C.c_int
        L.aco_acmove.argtypes = [C.c_int, i8p, C.c_int, C.c_int, i8p, C.c_void_p]
        L.aco_acmove.restype = C.c_int
        L.aco_moves_batch.argtypes = [i8p, u8p, i8p, u8p, u8p, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.aco_moves_batch.restype = None
        L.aco_env_step_batch.argtypes = [
            i8p, u8p, i32p, u8p, u8p, i32p, u8p, u8p, C.c_int64, C.c_int, C.c_int, C.c_int,
        ]
        L.aco_env_step_batch.restype = None
        for f in (L.aco_bfs, L.aco_greedy):
            f.argtypes = [
                i8p, C.c_int, C.c_int64, C.c_int, i32p, C.c_int, i8p, C.c_int64,
                C.POINTER(SearchResult),
            ]
            f.restype = C.c_int
        L.aco_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _raise(status):
    if status == ACO_ASSERT:
        raise AssertionError("oracle: reference raises AssertionError (invalid presentation)")
    if status == ACO_INDEX:
        raise IndexError("oracle: reference raises IndexError (conjugate on an empty relator)")


def num_threads() -> int:
    return lib().aco_num_threads()


def simplify_relator(relator, max_relator_length, cyclical=False, padded=True):
    """utils.py:175-240"""
    rel = np.ascontiguousarray(relator, dtype=np.int8).copy()
    n = int(np.count_nonzero(rel))
    assert (rel[n:] == 0).all(), "expect all zeros to be at the right end"
    m = lib().aco_simplify_relator(_p(rel), n, int(bool(cyclical)))
    out = rel[:m]
    if padded:
        out = np.pad(out, (0, max_relator_length - m))
    assert max_relator_length >= m
    return out, m


def _raw_move(fn, presentation, max_relator_length, i, j, sign, lengths, copy_lengths):
    p = np.ascontiguousarray(presentation, dtype=np.int8).copy()
    ns = fn(_p(p), int(max_relator_length), int(i), int(j), int(sign))
    if ns == -3:
        raise AssertionError("bad move arguments")
    if ns == -2:
        raise IndexError("conjugate on an empty relator")
    if ns >= 0:
        if copy_lengths:
            lengths = lengths.copy()
        lengths[i] = ns
    return p, lengths


def concatenate_relators(presentation, max_relator_length, i, j, sign, lengths):
    """ac_moves.py:4-76 (mutates the caller's lengths on acceptance, like the reference)."""
    return _raw_move(lib().aco_concatenate_relators, presentation, max_relator_length, i, j, sign, lengths, False)


def conjugate(presentation, max_relator_length, i, j, sign, lengths):
    """ac_moves.py:79-156"""
    return _raw_move(lib().aco_conjugate, presentation, max_relator_length, i, j, sign, lengths, True)


def acmove(move_id, presentation, max_relator_length, cyclical=True):
    """ac_moves.py:159-231 -> (next_state int8, [len0, len1]); raises like the reference."""
    p = np.ascontiguousarray(presentation, dtype=np.int8)
    assert p.size == 2 * max_relator_length
    out = np.empty_like(p)
    lens = (C.c_int * 2)()
    st = lib().aco_acmove(int(move_id), _p(p), int(max_relator_length), int(bool(cyclical)), _p(out), lens)
    _raise(st)
    return out, [lens[0], lens[1]]


def moves_batch(states, actions, cyclical=True, nthreads=0):
    """Batched ACMove -> (next_states, lens[N,2], status[N])."""
    s = np.ascontiguousarray(states, dtype=np.int8)
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    n, w = s.shape
    out = np.empty_like(s)
    lens = np.zeros((n, 2), np.uint8)
    status = np.zeros(n, np.uint8)
    lib().aco_moves_batch(_p(s), _p(a), _p(out), _p(lens), _p(status), n, w // 2, int(bool(cyclical)), nthreads)
    return out, lens, status


def env_step_batch(state, actions, step_count, horizon, nthreads=0):
    """In-place ACEnv.step over rows -> (reward, done, truncated, lens, status)."""
    assert state.dtype == np.int8 and state.flags.c_contiguous
    assert step_count.dtype == np.int32 and step_count.flags.c_contiguous
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    n, w = state.shape
    reward = np.zeros(n, np.int32)
    done = np.zeros(n, np.uint8)
    trunc = np.zeros(n, np.uint8)
    lens = np.zeros((n, 2), np.uint8)
    status = np.zeros(n, np.uint8)
    lib().aco_env_step_batch(
        _p(state), _p(a), _p(reward), _p(done), _p(trunc), _p(step_count), _p(lens), _p(status),
        n, w // 2, int(horizon), nthreads,
    )
    return reward, done, trunc, lens, status


def _search(fn, presentation, max_nodes_to_explore, cyclical, want_visited, path_cap=1 << 16):
    p = np.ascontiguousarray(presentation, dtype=np.int8)
    mrl = p.size // 2
    path = np.zeros((path_cap, 2), np.int32)
    cap = int(max_nodes_to_explore) + 16 if want_visited else 0
    visited = np.zeros((cap, 2 * mrl), np.int8) if want_visited else None
    res = SearchResult()
    rc = fn(_p(p), mrl, int(max_nodes_to_explore), int(bool(cyclical)), _p(path), path_cap,
            _p(visited), cap, C.byref(res))
    if rc < 0:
        raise MemoryError("oracle search failed")
    info = {
        "n_visited": res.n_visited,
        "n_expanded": res.n_expanded,
        "n_moves": res.n_moves,
        "frontier_left": res.frontier_left,
        "budget_hit": bool(res.budget_hit),
        "minlen_log": [res.minlen_log[i] for i in range(res.n_minlen)],
        "status": res.status,
    }
    if want_visited:
        info["visited"] = visited[: min(res.n_visited, cap)]
    _raise(res.status)
    pl = [(int(a), int(l)) for a, l in path[: res.path_len]]
    return bool(res.solved), pl, info


def bfs(presentation, max_nodes_to_explore=10000, cyclically_reduce_after_moves=False, want_visited=False):
    """breadth_first.py:15-97 -> (solved, path|None, info)."""
    solved, path, info = _search(lib().aco_bfs, presentation, max_nodes_to_explore,
                                 cyclically_reduce_after_moves, want_visited)
    return solved, (path if solved else None), info


def greedy_search(presentation, max_nodes_to_explore=10000, cyclically_reduce_after_moves=False,
                  want_visited=False):
    """greedy.py:15-121 -> (solved, path, info)."""
    return _search(lib().aco_greedy, presentation, max_nodes_to_explore,
                   cyclically_reduce_after_moves, want_visited)
