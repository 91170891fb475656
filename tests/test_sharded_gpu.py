"""GPU tests of the sharded BFS kernels (csrc/sbfs.cu) through the real chunk loop.  With one
visible GPU the world size is 1 (all records are self-owned); scripts/run_sharded.py runs the
same checks under torchrun with 2+ GPUs."""

import numpy as np
import pytest

from conftest import ms_row
from oracle import oracle as O

pytestmark = pytest.mark.gpu

AK2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0])
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)


@pytest.mark.parametrize("pres,budget,cyc", [(AK2, 1, False), (AK2, 10, False), (AK2, 5000, False), (AK2, 3000, True),
                                             (AK2, 1000000, False), (AK3, 50, False), (AK3, 7777, False),
                                             (AK3, 300000, False), (AK3, 2000000, False)])
def test_sharded_world1_vs_oracle(pres, budget, cyc):
    from ac_solver_b200.search.sharded import bfs_sharded

    solved, path, info = bfs_sharded(pres, budget, cyc, want_visited=True, chunk_parents=50000)
    es, ep, ei = O.bfs(pres, budget, cyc, want_visited=True)
    assert (solved, path) == (es, ep)
    for k in ("n_visited", "n_expanded", "n_moves", "frontier_left", "budget_hit", "minlen_log"):
        assert info[k] == ei[k], k
    assert np.array_equal(info["visited"], ei["visited"])


def test_sharded_world1_wide_keys_and_errors(miller_schupp):
    from ac_solver_b200.search.sharded import bfs_sharded

    p = ms_row(miller_schupp, 1189)  # mrl 36 -> 32-byte keys
    solved, path, info = bfs_sharded(p, 20000, want_visited=True)
    es, ep, ei = O.bfs(p, 20000, want_visited=True)
    assert (solved, path) == (es, ep) and np.array_equal(info["visited"], ei["visited"])
    with pytest.raises(AssertionError):
        bfs_sharded(np.array([1, 2, 0, 0, 1, 2, 0, 0]), 1000)
    with pytest.raises(AssertionError):
        bfs_sharded(np.array([1, 0, 2, 0, 0, 0, 1, 0]), 1000)
