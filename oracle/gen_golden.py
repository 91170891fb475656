"""Generate tests/golden/* by running the REAL reference (/root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    python oracle/gen_golden.py

The reference imports ``gymnasium`` (absent here, no network); a 10-line stub with
``Env``, ``spaces.Discrete`` and ``spaces.Box`` -- the whole surface ac_env.py:8-9,74,77
touches -- is written to a temp dir and put on sys.path.  Nothing is copied from the
reference: only inputs we choose and the outputs the reference computes are stored.

Fixtures written (all small):
  ref_unit_vectors.json   the parametrised unit vectors of the reference's own
                          tests/test_ac_env.py, harvested from the pytest marks, with the
                          reference's actual outputs beside the tests' expected values
  acmove_random.npz       random (state, move, cyclical) triples incl. non-reduced words,
                          empty relators and every raising case
  env_traces.npz          ACEnv.step trajectories (state, reward, done, truncated)
  search_cases.json       bfs / greedy results, visited counts, ordered visited digests
  miller_schupp.npz       the 1190 shipped presentations, the 533 stored greedy paths and
                          the bfs-solved subset (data, used as inputs / known answers)
"""

from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys
import tempfile
import time

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _install_stub():
    d = tempfile.mkdtemp(prefix="gym_stub_")
    os.makedirs(os.path.join(d, "gymnasium"))
    with open(os.path.join(d, "gymnasium", "__init__.py"), "w") as f:
        f.write("class Env:\n    pass\nfrom . import spaces\n")
    with open(os.path.join(d, "gymnasium", "spaces.py"), "w") as f:
        f.write(
            "import numpy as np\n"
            "class Discrete:\n    def __init__(self, n):\n        self.n = n\n"
            "class Box:\n    def __init__(self, low, high, dtype=None):\n"
            "        self.low, self.high, self.dtype = low, high, dtype\n"
            "        self.shape = np.asarray(low).shape\n"
        )
    sys.path.insert(0, d)
    sys.path.insert(0, REF)
    # ac_solver/__init__.py also imports the PPO trainer (torch, wandb, tqdm ...): load the
    # four hot-path modules by name after planting a bare package object instead.
    import types

    pkg = types.ModuleType("ac_solver")
    pkg.__path__ = [os.path.join(REF, "ac_solver")]
    sys.modules["ac_solver"] = pkg
    for sub in ("envs", "search"):
        m = types.ModuleType(f"ac_solver.{sub}")
        m.__path__ = [os.path.join(REF, "ac_solver", sub)]
        sys.modules[f"ac_solver.{sub}"] = m


def _tolist(x):
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (list, tuple)):
        return [_tolist(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.bool_,)):
        return bool(x)
    return x


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int8).tobytes()).hexdigest()


# --------------------------------------------------------------------------------------
def gen_unit_vectors():
    from ac_solver.envs import utils as U
    from ac_solver.envs import ac_moves as M

    spec = importlib.util.spec_from_file_location("ref_test_ac_env", os.path.join(REF, "tests", "test_ac_env.py"))
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)

    def params(fn):
        mark = [m for m in fn.pytestmark if m.name == "parametrize"][0]
        return mark.args[1]

    out = {}
    rows = []
    for rel, mrl, cyc, padded, exp_rel, exp_len in params(T.test_simplify_relator):
        r, l = U.simplify_relator(rel.copy(), mrl, cyclical=cyc, padded=padded)
        assert np.array_equal(r, exp_rel) and l == exp_len
        rows.append(dict(relator=_tolist(rel), mrl=mrl, cyclical=cyc, padded=padded,
                         expected_relator=_tolist(exp_rel), expected_length=exp_len))
    out["simplify_relator"] = rows  # tests/test_ac_env.py:18-83

    out["is_array_valid_presentation"] = [
        dict(presentation=_tolist(p), expected=bool(e)) for p, e in params(T.test_is_array_valid_presentation)
    ]  # :86-104
    out["is_presentation_trivial"] = [
        dict(presentation=_tolist(p), expected=bool(e)) for p, e in params(T.test_is_presentation_trivial)
    ]  # :107-124

    rows = []
    for pres, mrl, lens, exp, exp_l in params(T.test_simplify_presentation):
        r, l = U.simplify_presentation(pres.copy(), mrl, list(lens))
        assert np.array_equal(r, exp) and list(l) == list(exp_l)
        rows.append(dict(presentation=_tolist(pres), mrl=mrl, lengths=_tolist(lens),
                         expected=_tolist(exp), expected_lengths=_tolist(exp_l)))
    out["simplify_presentation"] = rows  # :140-181

    for name, fn, test in (("concatenate_relators", M.concatenate_relators, T.test_concatenate_relators),
                           ("conjugate", M.conjugate, T.test_conjugate)):
        rows = []
        for rels, nrel, i, j, sign, lens, exp, exp_l in params(test):
            r, l = fn(rels.copy(), nrel, i, j, sign, list(lens))
            assert np.array_equal(r, exp) and list(l) == list(exp_l), (name, rels)
            rows.append(dict(rels=_tolist(rels), mrl=nrel, i=i, j=j, sign=sign, lengths=_tolist(lens),
                             expected=_tolist(exp), expected_lengths=_tolist(exp_l)))
        out[name] = rows  # :184-326, :329-477

    # test_ACMove (:495-538): one state, ids 0,1,2,4,5 with deliberately wrong lengths;
    # we record all 12 ids for both cyclical flags.
    init = np.array([1, 2, 0, 0, -2, 0, 0, 0])
    rows = []
    for cyc in (True, False):
        for mid in range(12):
            r, l = M.ACMove(mid, init.copy(), 4, [4, 4], cyclical=cyc)
            rows.append(dict(move_id=mid, presentation=_tolist(init), mrl=4, cyclical=cyc,
                             expected=_tolist(r), expected_lengths=_tolist(l)))
    out["ACMove"] = rows

    # notebooks/Stable-AK3.ipynb: 53 moves (1-indexed) at mrl 15, cyclical=False, must land on AK(3)
    nb = json.load(open(os.path.join(REF, "notebooks", "Stable-AK3.ipynb")))
    src = "\n".join("".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code")
    out["stable_ak3_notebook_source_sha"] = hashlib.sha256(src.encode()).hexdigest()
    return out, src


def _random_word(rng, length, reduced):
    if length == 0:
        return []
    letters = [-2, -1, 1, 2]
    w = [int(rng.choice(letters))]
    while len(w) < length:
        c = int(rng.choice(letters))
        if reduced and c == -w[-1]:
            continue
        w.append(c)
    return w


def gen_acmove_random():
    from ac_solver.envs.ac_moves import ACMove

    rng = np.random.default_rng(20240817)
    groups = {}
    for mrl, count in ((4, 1500), (7, 2500), (10, 3000), (12, 3000), (18, 2500), (24, 3000), (36, 4000)):
        S = np.zeros((count, 2 * mrl), np.int8)
        A = np.zeros(count, np.uint8)
        Cy = np.zeros(count, np.uint8)
        O = np.zeros((count, 2 * mrl), np.int8)
        Ln = np.zeros((count, 2), np.uint8)
        St = np.zeros(count, np.uint8)
        for k in range(count):
            mode = rng.integers(0, 10)
            reduced = mode >= 3  # 30 % non-reduced inputs
            l0 = int(rng.integers(1, mrl + 1))
            l1 = int(rng.integers(1, mrl + 1))
            if mode == 9:  # short words: many cancellations / emptied relators
                l0, l1 = int(rng.integers(1, 4)), int(rng.integers(1, 4))
            w0 = _random_word(rng, l0, reduced)
            w1 = _random_word(rng, l1, reduced)
            if mode == 8:  # r1 = r0^{+-1} prefix games -> AssertionError cases
                w1 = list(w0) if rng.integers(0, 2) else [-c for c in reversed(w0)]
            if mode == 7 and rng.integers(0, 4) == 0:  # empty relator inputs
                if rng.integers(0, 2):
                    w0 = []
                else:
                    w1 = []
            S[k, : len(w0)] = w0
            S[k, mrl : mrl + len(w1)] = w1
            A[k] = rng.integers(0, 12)
            Cy[k] = rng.integers(0, 2)
            try:
                r, l = ACMove(int(A[k]), S[k].copy(), mrl, [0, 0], cyclical=bool(Cy[k]))
                O[k] = r
                Ln[k] = l
            except AssertionError:
                St[k] = 1
            except IndexError:
                St[k] = 2
        groups[f"s{mrl}"] = S
        groups[f"a{mrl}"] = A
        groups[f"c{mrl}"] = Cy
        groups[f"o{mrl}"] = O
        groups[f"l{mrl}"] = Ln
        groups[f"t{mrl}"] = St
        print(f"  acmove mrl={mrl}: {count} cases, {int((St==1).sum())} assert, {int((St==2).sum())} index")
    return groups


def gen_env_traces():
    from ac_solver.envs.ac_env import ACEnv, ACEnvConfig

    rng = np.random.default_rng(7)
    out = {}
    ak2 = [1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0]
    ak3_36 = [1, 1, 1, -2, -2, -2, -2] + [0] * 29 + [1, 2, 1, -2, -1, -2] + [0] * 30
    triv = [1, 2, 0, 0, -1, 0, 0, 0]  # reaches done quickly
    cases = [("ak2", ak2, 40, 120), ("ak3", ak3_36, 200, 300), ("short", triv, 10, 60)]
    for name, init, horizon, steps in cases:
        for rep in range(3):
            env = ACEnv(ACEnvConfig(initial_state=np.array(init, dtype=np.int8), horizon_length=horizon))
            acts = rng.integers(0, 12, size=steps).astype(np.uint8)
            states, rewards, dones, truncs = [], [], [], []
            n_done = 0
            for a in acts:
                try:
                    s, r, d, t, info = env.step(int(a))
                except AssertionError:
                    break
                states.append(np.array(s, dtype=np.int8))
                rewards.append(int(r))
                dones.append(bool(d))
                truncs.append(bool(t))
                if d:
                    assert info["actions"] == [int(x) for x in acts[: len(states)]]
                    n_done += 1
            k = f"{name}_{rep}"
            n = len(states)
            out[k + "_init"] = np.array(init, np.int8)
            out[k + "_horizon"] = np.array(horizon)
            out[k + "_actions"] = acts[:n]
            out[k + "_states"] = np.array(states, np.int8)
            out[k + "_rewards"] = np.array(rewards, np.int64)
            out[k + "_dones"] = np.array(dones, np.uint8)
            out[k + "_truncs"] = np.array(truncs, np.uint8)
            print(f"  env trace {k}: {n} steps, {n_done} done flags, max_reward {env.max_reward}")
    return out


def _run_search(mod, fn_name, pres, budget, cyc):
    """Run a reference search, recording visited states in insertion order by wrapping
    the module's ACMove (visited = root + every generated child except a solving one)."""
    import io
    from contextlib import redirect_stdout

    fn = getattr(mod, fn_name)
    real = mod.ACMove
    seen = {}
    root = tuple(np.array(pres, dtype=np.int8).tolist())
    seen[root] = None
    n_calls = [0]

    def wrapped(*a, **k):
        s, l = real(*a, **k)
        n_calls[0] += 1
        if sum(l) != 2:
            seen.setdefault(tuple(np.array(s, dtype=np.int8).tolist()), None)
        return s, l

    mod.ACMove = wrapped
    buf = io.StringIO()
    try:
        with redirect_stdout(buf):
            t0 = time.time()
            err = None
            try:
                ok, path = fn(presentation=np.array(pres), max_nodes_to_explore=budget, verbose=True,
                              cyclically_reduce_after_moves=cyc)
            except AssertionError:
                ok, path, err = None, None, "AssertionError"
            dt = time.time() - t0
    finally:
        mod.ACMove = real
    visited = np.array(list(seen.keys()), dtype=np.int8)
    return dict(
        presentation=list(map(int, pres)), budget=int(budget), cyclical=bool(cyc),
        solved=ok, path=_tolist(path), error=err, n_visited=len(seen), n_moves=n_calls[0],
        visited_sha256=_sha(visited), stdout=buf.getvalue(), seconds=round(dt, 3),
    ), visited


def gen_search(ms_rows):
    from ac_solver.search import breadth_first as B
    from ac_solver.search import greedy as G

    ak2 = [1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0]
    ak3 = [1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18
    nonred = [1, -1, 2, 2, 1, 0, 0, 0, 2, 1, -1, -2, -2, 1, 0, 0]  # non-reduced root
    cases = {"bfs": [], "greedy": []}
    small_visited = {}

    def add(kind, mod, fn, pres, budget, cyc, keep=False):
        rec, vis = _run_search(mod, fn, pres, budget, cyc)
        if keep:
            key = f"{kind}_{len(cases[kind])}"
            small_visited[key] = vis
            rec["visited_key"] = key
        cases[kind].append(rec)
        print(f"  {kind} budget={budget} cyc={cyc} L={len(pres)//2}: solved={rec['solved']} "
              f"visited={rec['n_visited']} ({rec['seconds']} s)")

    for b in (1, 10, 137, 1000, 5000, 10000, 20000, 1000000):
        add("bfs", B, "bfs", ak2, b, False, keep=b <= 1000)
    add("bfs", B, "bfs", ak2, 3000, True, keep=True)
    for b in (50, 1000, 7777):
        add("bfs", B, "bfs", ak3, b, False, keep=b <= 1000)
    add("bfs", B, "bfs", nonred, 500, False, keep=True)
    for idx in (0, 5, 170, 400, 533, 700, 1189):
        add("bfs", B, "bfs", ms_rows[idx], 2000, False)
    for b in (500, 1000000):
        add("greedy", G, "greedy_search", ak2, b, False, keep=b <= 500)
    add("greedy", G, "greedy_search", ak2, 2000, True, keep=True)
    add("greedy", G, "greedy_search", ak3, 3000, False, keep=True)
    add("greedy", G, "greedy_search", nonred, 500, False, keep=True)
    for idx in (0, 200, 400, 533, 600, 1189):
        add("greedy", G, "greedy_search", ms_rows[idx], 10000 if idx < 533 else 3000, False)
    return cases, small_visited


def gen_miller_schupp():
    import ast

    d = os.path.join(REF, "ac_solver", "search", "miller_schupp", "data")
    rows = [ast.literal_eval(l) for l in open(os.path.join(d, "all_presentations.txt")) if l.strip()]
    solved = [ast.literal_eval(l) for l in open(os.path.join(d, "greedy_solved_presentations.txt")) if l.strip()]
    paths = [ast.literal_eval(l) for l in open(os.path.join(d, "greedy_search_paths.txt")) if l.strip()]
    bfs_solved = [ast.literal_eval(l) for l in open(os.path.join(d, "bfs_solved_presentations.txt")) if l.strip()]
    assert len(rows) == 1190 and len(solved) == 533 and len(paths) == 533
    assert rows[:533] == solved
    mrl = np.array([len(r) // 2 for r in rows], np.int32)
    P = np.zeros((len(rows), 72), np.int8)  # each row re-padded to mrl 36 (utils.py:148-172)
    for k, r in enumerate(rows):
        m = len(r) // 2
        P[k, :m] = r[:m]
        P[k, 36 : 36 + m] = r[m:]
    index = {tuple(r): k for k, r in enumerate(rows)}
    bfs_idx = np.array(sorted(index[tuple(r)] for r in bfs_solved), np.int32)
    flat = np.array([x for p in paths for x in p], np.int16).reshape(-1, 2)
    offs = np.cumsum([0] + [len(p) for p in paths]).astype(np.int64)
    return rows, dict(presentations36=P, mrl=mrl, greedy_path_flat=flat, greedy_path_offsets=offs,
                      bfs_solved_index=bfs_idx)


def main():
    os.makedirs(OUT, exist_ok=True)
    _install_stub()
    t0 = time.time()
    print("unit vectors ...")
    uv, _ = gen_unit_vectors()
    json.dump(uv, open(os.path.join(OUT, "ref_unit_vectors.json"), "w"), indent=0)
    print("miller-schupp data ...")
    rows, ms = gen_miller_schupp()
    np.savez_compressed(os.path.join(OUT, "miller_schupp.npz"), **ms)
    print("random ACMove triples ...")
    np.savez_compressed(os.path.join(OUT, "acmove_random.npz"), **gen_acmove_random())
    print("env traces ...")
    np.savez_compressed(os.path.join(OUT, "env_traces.npz"), **gen_env_traces())
    print("searches ...")
    cases, vis = gen_search(rows)
    json.dump(cases, open(os.path.join(OUT, "search_cases.json"), "w"), indent=0)
    np.savez_compressed(os.path.join(OUT, "search_visited.npz"), **vis)
    print(f"done in {time.time()-t0:.1f} s")


if __name__ == "__main__":
    main()
