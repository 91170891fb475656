"""Pins the CPU oracle (oracle/ac_oracle.c) against fixtures produced by the REAL reference
(oracle/gen_golden.py).  CPU only."""

import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from conftest import ms_path, ms_row


def _alphabet_ok(*arrays):
    return all(np.abs(np.asarray(a)).max(initial=0) <= 127 for a in arrays)


# reference tests/test_ac_env.py:18-83
def test_simplify_relator_vectors(unit_vectors):
    for c in unit_vectors["simplify_relator"]:
        r, l = O.simplify_relator(np.array(c["relator"]), c["mrl"], cyclical=c["cyclical"], padded=c["padded"])
        assert r.tolist() == c["expected_relator"], c
        assert l == c["expected_length"], c


# reference tests/test_ac_env.py:86-104
def test_is_valid_presentation_vectors(unit_vectors):
    for c in unit_vectors["is_array_valid_presentation"]:
        p = np.array(c["presentation"], dtype=np.int8)
        if p.size == 0 or p.size % 2:
            continue  # length guards live in the Python shim, not in the C core
        assert bool(O.lib().aco_is_valid_presentation(p.ctypes.data, p.size // 2)) == c["expected"], c


# reference tests/test_ac_env.py:184-326 and :329-477
@pytest.mark.parametrize("name", ["concatenate_relators", "conjugate"])
def test_raw_move_vectors(unit_vectors, name):
    fn = getattr(O, name)
    for c in unit_vectors[name]:
        r, l = fn(np.array(c["rels"]), c["mrl"], c["i"], c["j"], c["sign"], list(c["lengths"]))
        assert r.tolist() == c["expected"], c
        assert list(l) == c["expected_lengths"], c


# reference tests/test_ac_env.py:495-538 (all 12 ids, both cyclical flags)
def test_acmove_vectors(unit_vectors):
    for c in unit_vectors["ACMove"]:
        r, l = O.acmove(c["move_id"], np.array(c["presentation"]), c["mrl"], cyclical=c["cyclical"])
        assert r.tolist() == c["expected"], c
        assert l == c["expected_lengths"], c


@pytest.mark.parametrize("mrl", [4, 7, 10, 12, 18, 24, 36])
def test_acmove_random_differential(acmove_random, mrl):
    g = acmove_random
    S, A, Cy, Oo, Ln, St = g[f"s{mrl}"], g[f"a{mrl}"], g[f"c{mrl}"], g[f"o{mrl}"], g[f"l{mrl}"], g[f"t{mrl}"]
    for cyc in (0, 1):
        m = Cy == cyc
        out, lens, status = O.moves_batch(S[m], A[m], cyclical=bool(cyc), nthreads=2)
        assert np.array_equal(status, St[m])
        ok = St[m] == 0
        assert np.array_equal(out[ok], Oo[m][ok])
        assert np.array_equal(lens[ok], Ln[m][ok])


def test_env_traces(env_traces):
    t = env_traces
    names = sorted({k.rsplit("_", 1)[0] for k in t.files if k.endswith("_init")})
    assert len(names) == 9
    for n in names:
        init, H = t[n + "_init"], int(t[n + "_horizon"])
        acts = t[n + "_actions"]
        state = init[None].copy()
        sc = np.zeros(1, np.int32)
        for k, a in enumerate(acts):
            r, d, tr, lens, st = O.env_step_batch(state, np.array([a], np.uint8), sc, H)
            assert st[0] == 0
            assert np.array_equal(state[0], t[n + "_states"][k])
            assert int(r[0]) == int(t[n + "_rewards"][k])
            assert bool(d[0]) == bool(t[n + "_dones"][k])
            assert bool(tr[0]) == bool(t[n + "_truncs"][k])


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int8).tobytes()).hexdigest()


@pytest.mark.parametrize("kind", ["bfs", "greedy"])
def test_search_cases(search_cases, search_visited, kind):
    fn = O.bfs if kind == "bfs" else O.greedy_search
    for c in search_cases[kind]:
        solved, path, info = fn(np.array(c["presentation"], np.int8), c["budget"], c["cyclical"], want_visited=True)
        assert solved == c["solved"], c
        exp_path = None if c["path"] is None else [tuple(x) for x in c["path"]]
        assert path == exp_path, c
        assert info["n_visited"] == c["n_visited"], c
        assert info["n_moves"] == c["n_moves"], c
        # same visited states in the same insertion order
        assert _sha(info["visited"]) == c["visited_sha256"], c
        if "visited_key" in c:
            assert np.array_equal(info["visited"], search_visited[c["visited_key"]])
        # the verbose stdout of the reference: successive minima and the budget line
        mins = [int(l.rsplit(" ", 1)[1]) for l in c["stdout"].splitlines() if l.startswith("New minimal")]
        assert info["minlen_log"] == mins, c
        assert info["budget_hit"] == ("Exiting search" in c["stdout"]), c


# reference tests/search/test_bfs.py:15 and tests/search/test_gs.py:15 (known answers)
def test_known_answer_paths():
    ak2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0], np.int8)
    ok, path, _ = O.bfs(ak2, int(1e6))
    assert ok and path == [(-1, 11), (4, 11), (11, 11), (2, 12), (4, 12), (11, 12), (9, 12), (0, 11), (5, 11),
                           (7, 11), (3, 13), (11, 13), (9, 13), (2, 12), (8, 12), (9, 12), (3, 7), (0, 5), (0, 3), (3, 2)]
    ok, path, _ = O.greedy_search(ak2, int(1e6))
    assert ok and path == [(-1, 11), (11, 11), (4, 11), (2, 12), (11, 12), (4, 12), (9, 12), (0, 11), (1, 13),
                           (8, 13), (6, 13), (9, 13), (0, 12), (1, 11), (6, 11), (8, 11), (3, 8), (6, 8), (2, 5),
                           (5, 5), (1, 3), (8, 3), (2, 2)]
    assert O.bfs(ak2, 10)[:2] == (False, None)  # tests/search/test_bfs.py:30-37


def test_miller_schupp_stored_greedy_paths(miller_schupp):
    """greedy_search_paths.txt row k is the budget-1e6 greedy path of presentation k
    (stored as action+1).  All 533 are replayed by the C oracle (fast)."""
    ms = miller_schupp
    for k in range(533):
        ok, path, info = O.greedy_search(ms_row(ms, k), int(1e6))
        assert ok, k
        assert path == ms_path(ms, k), k


def test_miller_schupp_unsolved_sample(miller_schupp):
    ms = miller_schupp
    for k in (533, 700, 900, 1189):
        ok, path, info = O.greedy_search(ms_row(ms, k), 20000)
        assert not ok and info["budget_hit"]
