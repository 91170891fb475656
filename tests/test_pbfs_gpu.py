"""GPU parity tests of the native partitioned BFS (csrc/pbfs.cu) against the CPU oracle: result,
path, counters and the visited ARRAY in insertion order, for world sizes 1..8.  With one visible
GPU the ranks of a world > 1 are simulated inside one process on one device (same kernels, same
peer-pointer exchange, same flag protocol); scripts/run_partitioned.py runs the same checks with
one process per GPU under torchrun."""

import numpy as np
import pytest

from conftest import ms_row
from oracle import oracle as O

pytestmark = pytest.mark.gpu

AK2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0])
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)

CASES = [(AK2, 1, False), (AK2, 10, False), (AK2, 5000, False), (AK2, 3000, True), (AK2, 1000000, False),
         (AK3, 50, False), (AK3, 7777, False), (AK3, 300000, False), (AK3, 2000000, False)]


def _compare(pres, budget, cyc, world, chunk):
    from ac_solver_b200.search.partitioned import bfs_partitioned

    solved, path, info = bfs_partitioned(pres, budget, cyc, want_visited=True, sim_world=world, chunk_parents=chunk,
                                         verbose=None)
    es, ep, ei = O.bfs(pres, budget, cyc, want_visited=True)
    assert (solved, path) == (es, ep)
    for k in ("n_visited", "n_expanded", "n_moves", "frontier_left", "budget_hit", "minlen_log"):
        assert info[k] == ei[k], k
    assert np.array_equal(info["visited"], ei["visited"])
    assert sum(info["n_local"]) == info["n_visited"]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("pres,budget,cyc", CASES)
def test_pbfs_vs_oracle(pres, budget, cyc, world):
    _compare(pres, budget, cyc, world, 50000)


@pytest.mark.parametrize("world,chunk", [(1, 0), (1, 1024), (4, 0), (5, 3000), (16, 40000)])
def test_pbfs_chunk_sizes(world, chunk):
    """default chunk size (one chunk per level) and very small chunks (many chunks per level)"""
    _compare(AK3, 400000, False, world, chunk)


@pytest.mark.parametrize("world", [1, 4])
def test_pbfs_wide_keys_and_errors(miller_schupp, world):
    from ac_solver_b200.search.partitioned import bfs_partitioned

    p = ms_row(miller_schupp, 1189)  # mrl 36 -> 32-byte keys
    solved, path, info = bfs_partitioned(p, 20000, want_visited=True, sim_world=world, verbose=None)
    es, ep, ei = O.bfs(p, 20000, want_visited=True)
    assert (solved, path) == (es, ep) and np.array_equal(info["visited"], ei["visited"])
    with pytest.raises(AssertionError):  # r1 == r0: move 1 empties r0 -> the reference raises
        bfs_partitioned(np.array([1, 2, 0, 0, 1, 2, 0, 0]), 1000, sim_world=world)
    with pytest.raises(AssertionError):  # invalid root (breadth_first.py:36-38)
        bfs_partitioned(np.array([1, 0, 2, 0, 0, 0, 1, 0]), 1000, sim_world=world)
    with pytest.raises(ValueError):
        bfs_partitioned(np.array([1, 3, 0, 0, 2, 0, 0, 0]), 100, sim_world=world)


def test_pbfs_engine_reuse():
    """one engine, several runs (epochs and flags carry over)"""
    from ac_solver_b200.search.partitioned import PartitionedBfs

    with PartitionedBfs(24, 100000, sim_world=3, chunk_parents=20000) as eng:
        for _ in range(3):
            solved, path, info = eng.run(AK3, want_visited=True)
            es, ep, ei = O.bfs(AK3, 100000, want_visited=True)
            assert (solved, path) == (es, ep) and np.array_equal(info["visited"], ei["visited"])


def test_pbfs_miller_schupp_rows(miller_schupp):
    from ac_solver_b200.search.partitioned import bfs_partitioned

    for k in list(range(0, 6)) + [170, 533]:
        p = ms_row(miller_schupp, k)
        solved, path, info = bfs_partitioned(p, 10000, want_visited=True, sim_world=2, verbose=None)
        es, ep, ei = O.bfs(p, 10000, want_visited=True)
        assert (solved, path) == (es, ep), k
        assert np.array_equal(info["visited"], ei["visited"]), k
