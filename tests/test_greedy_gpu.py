"""GPU parity tests for the batched greedy search (csrc/greedy.cu)."""

import hashlib

import numpy as np
import pytest

from conftest import ms_path, ms_row
from oracle import oracle as O

pytestmark = pytest.mark.gpu

AK2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0])
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int8).tobytes()).hexdigest()


# reference tests/search/test_gs.py:7-27 (known answer: pins the heap tie-breaking)
def test_gs_on_AK2():
    from ac_solver_b200 import greedy_search

    expected = (True, [(-1, 11), (11, 11), (4, 11), (2, 12), (11, 12), (4, 12), (9, 12), (0, 11), (1, 13), (8, 13),
                       (6, 13), (9, 13), (0, 12), (1, 11), (6, 11), (8, 11), (3, 8), (6, 8), (2, 5), (5, 5), (1, 3),
                       (8, 3), (2, 2)])
    assert greedy_search(presentation=AK2, max_nodes_to_explore=int(1e6), verbose=False) == expected


def test_greedy_golden_cases(search_cases, search_visited):
    """Every greedy fixture recorded from the REAL reference, incl. the failure-path quirk
    (False, path + [(11, len)]), visited count and visited states in insertion order."""
    from ac_solver_b200.search.greedy import greedy_search_batch

    for c in search_cases["greedy"]:
        solved, path, info = greedy_search_batch(np.array(c["presentation"])[None, :], c["budget"], c["cyclical"],
                                                 want_visited=True)[0]
        assert solved == c["solved"], c
        assert path == [tuple(x) for x in c["path"]], c
        assert info["n_visited"] == c["n_visited"], c
        assert info["n_moves"] == c["n_moves"], c
        assert _sha(info["visited"]) == c["visited_sha256"], c
        if "visited_key" in c:
            assert np.array_equal(info["visited"], search_visited[c["visited_key"]])
        mins = [int(l.rsplit(" ", 1)[1]) for l in c["stdout"].splitlines() if l.startswith("New minimal")]
        assert info["minlen_log"] == mins, c
        assert info["budget_hit"] == ("Exiting search" in c["stdout"]), c


@pytest.mark.parametrize("budget", [1, 2, 50, 777, 20000, 200000])
def test_greedy_ak3_vs_oracle(budget):
    from ac_solver_b200.search.greedy import greedy_search_batch

    solved, path, info = greedy_search_batch(AK3[None, :], budget, want_visited=True)[0]
    es, ep, ei = O.greedy_search(AK3, budget, want_visited=True)
    assert (solved, path) == (es, ep)
    for k in ("n_visited", "n_expanded", "n_moves", "frontier_left", "budget_hit", "minlen_log"):
        assert info[k] == ei[k], k
    assert np.array_equal(info["visited"], ei["visited"])


def test_greedy_miller_schupp_batch(miller_schupp):
    """Batched sweep flavour of BASELINE config 3: all rows of one mrl group in ONE launch.
    Solved rows must reproduce the paths shipped with the reference
    (greedy_search_paths.txt, stored as action+1); every row must equal the oracle."""
    from ac_solver_b200.search.greedy import greedy_search_batch

    ms = miller_schupp
    budget = 30000
    for mrl in (18, 24):
        rows = [k for k in range(len(ms["mrl"])) if ms["mrl"][k] == mrl][:120]
        P = np.stack([ms_row(ms, k) for k in rows])
        got = greedy_search_batch(P, budget)
        for k, (solved, path, info) in zip(rows, got):
            es, ep, ei = O.greedy_search(ms_row(ms, k), budget)
            assert (solved, path) == (es, ep), k
            assert info["n_visited"] == ei["n_visited"], k
            if solved and k < 533:
                assert path == ms_path(ms, k), k


def test_greedy_wide_keys_and_errors():
    from ac_solver_b200 import greedy_search
    from ac_solver_b200.search.greedy import greedy_search_batch

    p = np.zeros(72, np.int8)
    p[:7], p[36:43] = AK2[:7], AK2[7:]
    solved, path, info = greedy_search_batch(p[None, :], 5000, want_visited=True)[0]
    es, ep, ei = O.greedy_search(p, 5000, want_visited=True)
    assert (solved, path) == (es, ep) and np.array_equal(info["visited"], ei["visited"])
    q = np.array([1, 2, 0, 0, 1, 2, 0, 0])  # r1 == r0: move 1 empties r0 -> AssertionError
    with pytest.raises(AssertionError):
        O.greedy_search(q, 100)
    with pytest.raises(AssertionError):
        greedy_search(q, 100)


def test_greedy_verbose_stdout(search_cases, capsys):
    """verbose=True prints the reference's lines (greedy.py:82-99).  The fixture was recorded under
    numpy 2, which prints lengths as np.int64(11); the reference's pinned numpy 1.24 prints 11."""
    import re
    from ac_solver_b200 import greedy_search

    c = [c for c in search_cases["greedy"] if c["solved"] and len(c["presentation"]) == 14 and not c["cyclical"]][0]
    ok, path = greedy_search(np.array(c["presentation"]), c["budget"], verbose=True)
    assert ok
    expected = re.sub(r"np\.int64\((\d+)\)", r"\1", c["stdout"])
    assert capsys.readouterr().out == expected


def test_moves_batch_beyond_packed_width():
    """max_relator_length > 64 is served by the generic byte kernel."""
    from ac_solver_b200 import ac_moves_batch
    from ac_solver_b200.synthetic import random_actions, random_presentations

    S = random_presentations(3000, 100, seed=5)
    A = random_actions(3000, seed=6)
    out, lens, status = ac_moves_batch(S, A, cyclical=True)
    eo, el, es = O.moves_batch(S, A, cyclical=True)
    assert np.array_equal(status, es) and np.array_equal(out, eo)
    assert np.array_equal(lens[es == 0], el[es == 0])
