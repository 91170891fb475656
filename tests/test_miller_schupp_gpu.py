"""Miller-Schupp generator + sweep against the reference's data and tests
(tests/search/miller_schupp/test_miller_schupp.py)."""

import numpy as np
import pytest

from conftest import ms_path, ms_row
from oracle import oracle as O

pytestmark = pytest.mark.gpu


# reference test_miller_schupp.py:13-56
@pytest.mark.parametrize("n,max_w_len", [(n, m) for n in range(1, 4) for m in range(1, 4)])
def test_generate_structure(n, max_w_len):
    from ac_solver_b200.search.miller_schupp import generate_miller_schupp_presentations

    ms = generate_miller_schupp_presentations(n=n, max_w_len=max_w_len)
    assert isinstance(ms, dict) and list(ms.keys()) == list(range(1, max_w_len + 1))
    for w_len, group in ms.items():
        for p in group:
            assert len(p) % 2 == 0
            m = len(p) // 2
            assert np.count_nonzero(p[:m]) == 2 * n + 3 and p[m] == -1


def test_generated_set_equals_shipped_dataset(miller_schupp):
    """n = 1..7, |w| <= 7 reproduces the 1190 presentations of all_presentations.txt: 170 per n
    (reference test :42-56), the same lists as the data file (which orders them solved-first)."""
    from ac_solver_b200.search.miller_schupp import generate_miller_schupp_presentations

    shipped = {tuple(int(v) for v in ms_row(miller_schupp, k)) for k in range(1190)}
    mine = set()
    for n in range(1, 8):
        ms = generate_miller_schupp_presentations(n=n, max_w_len=7)
        if n == 1:
            assert (len(ms[1]), len(ms[2]), len(ms[3])) == (2, 2, 2)
        assert sum(len(v) for v in ms.values()) == 170
        mine |= {tuple(p) for g in ms.values() for p in g}
    assert mine == shipped


def test_trivialize_through_greedy_and_bfs(tmp_path, capsys):
    """reference test :59-190 flavour: n in {1,2}, |w| <= 2: greedy@1e6 solves all 8, bfs@1e4
    solves exactly one with the path [(-1,7),(1,7),(7,5),(0,4),(5,2)]; files are written in the
    reference's naming scheme; everything equals the oracle."""
    from ac_solver_b200 import bfs, greedy_search
    from ac_solver_b200.search.miller_schupp import trivialize_miller_schupp_through_search
    from ac_solver_b200.search.miller_schupp.miller_schupp import load_presentations_from_text_file

    solved, unsolved, paths = trivialize_miller_schupp_through_search(
        min_n=1, max_n=2, min_w_len=1, max_w_len=2, max_nodes_to_explore=int(1e6), search_fn=greedy_search,
        write_output_to_file=True, output_dir=str(tmp_path))
    assert len(solved) == 8 and len(unsolved) == 0
    for pres, path in zip(solved, paths):
        ok, epath, _ = O.greedy_search(np.array(pres, np.int8), int(1e6))
        assert ok and path == epath
    base = tmp_path / "n-1-to-2_lenw-1-to-2-max-nodes-1000000-greedy_search"
    assert load_presentations_from_text_file(str(base) + "_solved.txt") == solved
    assert [[tuple(x) for x in p] for p in load_presentations_from_text_file(str(base) + "_paths.txt")] == paths
    solved_b, unsolved_b, paths_b = trivialize_miller_schupp_through_search(
        min_n=1, max_n=2, min_w_len=1, max_w_len=2, max_nodes_to_explore=int(1e4), search_fn=bfs)
    assert len(solved_b) == 1 and len(unsolved_b) == 7
    assert paths_b[0] == [(-1, 7), (1, 7), (7, 5), (0, 4), (5, 2)]
    assert "Applying bfs to presentations of n = 1, lenw = 1" in capsys.readouterr().out


def test_bfs_solved_file_reproduced_on_the_gpu(miller_schupp):
    """The reference's bfs_solved_presentations.txt (278 rows, settings not recorded upstream) is what
    bfs(p, 1e6, cyclically_reduce_after_moves=True) solves: every one of the first 533 rows and a sample of
    the rest through the GPU search."""
    import contextlib
    import io

    from ac_solver_b200.search.breadth_first import bfs_device

    listed = {int(k) for k in miller_schupp["bfs_solved_index"]}
    assert len(listed) == 278 and max(listed) < 533
    got = set()
    with contextlib.redirect_stdout(io.StringIO()):
        for k in list(range(533)) + list(range(533, 1190, 13)):
            if bfs_device(np.array(ms_row(miller_schupp, k), dtype=np.int8), 1_000_000, True)[0]:
                got.add(k)
    assert got == listed
