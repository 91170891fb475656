"""GPU parity tests for the device BFS against the reference-generated fixtures and the oracle."""

import hashlib

import numpy as np
import pytest

from conftest import ms_row
from oracle import oracle as O

pytestmark = pytest.mark.gpu

AK2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0])
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int8).tobytes()).hexdigest()


# reference tests/search/test_bfs.py:7-27 (known answer) and :30-37
def test_bfs_on_AK2():
    from ac_solver_b200 import bfs

    expected = (True, [(-1, 11), (4, 11), (11, 11), (2, 12), (4, 12), (11, 12), (9, 12), (0, 11), (5, 11), (7, 11),
                       (3, 13), (11, 13), (9, 13), (2, 12), (8, 12), (9, 12), (3, 7), (0, 5), (0, 3), (3, 2)])
    assert bfs(presentation=AK2, max_nodes_to_explore=int(1e6), verbose=False) == expected


def test_bfs_max_nodes_reached(capsys):
    from ac_solver_b200 import bfs

    assert bfs(presentation=AK2, max_nodes_to_explore=10, verbose=False) == (False, None)
    assert "Exiting search as number of explored nodes = 12 has exceeded the limit 10" in capsys.readouterr().out


def test_bfs_golden_cases(search_cases, search_visited):
    """Every bfs fixture recorded from the REAL reference: result, visited count, visited
    states in insertion order (sha256 and, for small budgets, the full array), stdout."""
    from ac_solver_b200.search.breadth_first import bfs_device

    for c in search_cases["bfs"]:
        solved, path, info = bfs_device(np.array(c["presentation"]), c["budget"], c["cyclical"], want_visited=True)
        assert solved == c["solved"], c
        assert path == (None if c["path"] is None else [tuple(x) for x in c["path"]]), c
        assert info["n_visited"] == c["n_visited"], c
        assert info["n_moves"] == c["n_moves"], c
        assert _sha(info["visited"]) == c["visited_sha256"], c
        if "visited_key" in c:
            assert np.array_equal(info["visited"], search_visited[c["visited_key"]])
        mins = [int(l.rsplit(" ", 1)[1]) for l in c["stdout"].splitlines() if l.startswith("New minimal")]
        assert info["minlen_log"] == mins, c
        assert info["budget_hit"] == ("Exiting search" in c["stdout"]), c


def test_bfs_verbose_stdout(search_cases, capsys):
    from ac_solver_b200 import bfs

    c = [c for c in search_cases["bfs"] if c["budget"] == 5000][0]
    bfs(np.array(c["presentation"]), c["budget"], verbose=True)
    assert capsys.readouterr().out == c["stdout"]


@pytest.mark.parametrize("budget", [1, 2, 13, 500, 4097, 50_000, 300_000, 2_000_000])
def test_bfs_ak3_vs_oracle(budget):
    """AK(3), mrl 24 (BASELINE config 5 at one GPU) at budgets the oracle finishes in seconds:
    identical visited ARRAY (order included) and counters; multi-chunk levels are exercised."""
    from ac_solver_b200.search.breadth_first import bfs_device

    solved, path, info = bfs_device(AK3, budget, want_visited=True)
    es, ep, ei = O.bfs(AK3, budget, want_visited=True)
    assert (solved, path) == (es, ep)
    for k in ("n_visited", "n_expanded", "n_moves", "frontier_left", "budget_hit", "minlen_log"):
        assert info[k] == ei[k], k
    assert np.array_equal(info["visited"], ei["visited"])


@pytest.mark.parametrize("mrl_pad", [30, 36, 48])
def test_bfs_wide_keys(mrl_pad):
    """mrl > 29 uses 32-byte keys (two words per relator)."""
    from ac_solver_b200.search.breadth_first import bfs_device

    m = 7
    p = np.zeros(2 * mrl_pad, np.int8)
    p[:m], p[mrl_pad : mrl_pad + m] = AK2[:m], AK2[m:]
    solved, path, info = bfs_device(p, 30000, want_visited=True)
    es, ep, ei = O.bfs(p, 30000, want_visited=True)
    assert (solved, path) == (es, ep)
    assert info["n_visited"] == ei["n_visited"] and np.array_equal(info["visited"], ei["visited"])


def test_bfs_miller_schupp_rows(miller_schupp):
    """reference tests/search/miller_schupp/test_miller_schupp.py:159-190 flavour: rows of the
    shipped dataset at budget 1e4 vs the oracle (one of n=1 rows solves with a 5-entry path)."""
    from ac_solver_b200.search.breadth_first import bfs_device

    n_solved = 0
    for k in list(range(0, 12)) + [170, 400, 533, 1189]:
        p = ms_row(miller_schupp, k)
        solved, path, info = bfs_device(p, 10000, want_visited=True)
        es, ep, ei = O.bfs(p, 10000, want_visited=True)
        assert (solved, path) == (es, ep), k
        assert info["n_visited"] == ei["n_visited"] and np.array_equal(info["visited"], ei["visited"]), k
        n_solved += solved
    assert n_solved >= 1


def test_bfs_error_paths():
    from ac_solver_b200 import bfs
    from ac_solver_b200.search.breadth_first import bfs_device

    with pytest.raises(AssertionError):  # invalid root (breadth_first.py:36-38)
        bfs(np.array([1, 0, 2, 0, 0, 0, 1, 0]))
    # r1 == r0: move 1 (r0 -> r0 r1^-1) empties r0 -> the reference raises AssertionError
    p = np.array([1, 2, 0, 0, 1, 2, 0, 0])
    with pytest.raises(AssertionError):
        O.bfs(p, 1000)
    with pytest.raises(AssertionError):
        bfs_device(p, 1000)
    with pytest.raises(ValueError):
        bfs_device(np.array([1, 3, 0, 0, 2, 0, 0, 0]), 100)
