"""GPU parity tests for K1 (batched ACMove / ACEnv.step) and the generic byte kernel,
all through the C ABI, against the CPU oracle and the reference-generated golden vectors."""

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def random_rows(rng, n, mrl, reduced=True, min_len=1):
    """Rows of two random words over {+-1,+-2}; reduced => freely AND cyclically reduced."""
    from ac_solver_b200.synthetic import random_presentations

    seed = int(rng.integers(0, 2**31))
    if reduced:
        return random_presentations(n, mrl, seed=seed, min_len=min_len)
    r = np.random.default_rng(seed)
    S = np.zeros((n, 2 * mrl), np.int8)
    for h in range(2):
        L = r.integers(min_len, mrl + 1, size=n)
        w = r.choice(np.array([-2, -1, 1, 2], np.int8), size=(n, mrl))
        S[:, h * mrl : (h + 1) * mrl] = np.where(np.arange(mrl)[None, :] < L[:, None], w, 0)
    return S


@pytest.mark.parametrize("mrl", [4, 7, 10, 12, 18, 24, 36])
def test_moves_batch_golden(acmove_random, mrl):
    """Packed kernel == the REAL reference on the stored random triples (incl. non-reduced
    words, empty relators and every raising case)."""
    from ac_solver_b200 import ac_moves_batch

    g = acmove_random
    S, A, Cy, Oo, Ln, St = g[f"s{mrl}"], g[f"a{mrl}"], g[f"c{mrl}"], g[f"o{mrl}"], g[f"l{mrl}"], g[f"t{mrl}"]
    for cyc in (0, 1):
        m = Cy == cyc
        out, lens, status = ac_moves_batch(S[m], A[m], cyclical=bool(cyc))
        assert np.array_equal(status, St[m])
        ok = St[m] == 0
        assert np.array_equal(out[ok], Oo[m][ok])
        assert np.array_equal(lens[ok], Ln[m][ok])
        assert np.array_equal(out[~ok], S[m][~ok])  # raising rows are returned unchanged


@pytest.mark.parametrize("mrl,n", [(36, 200_000), (24, 100_000), (7, 50_000), (18, 50_000), (33, 20_000), (64, 20_000), (1, 1000), (2, 5000)])
@pytest.mark.parametrize("cyclical", [True, False])
def test_moves_batch_vs_oracle(mrl, n, cyclical):
    """Seeded differential test at sizes the oracle finishes in seconds (bit-exact)."""
    from ac_solver_b200 import ac_moves_batch

    rng = np.random.default_rng(1000 + mrl)
    S = random_rows(rng, min(n, 20000), mrl, reduced=True)
    S = np.concatenate([S, random_rows(rng, 2000, mrl, reduced=False)])
    reps = -(-n // len(S))
    S = np.tile(S, (reps, 1))[:n]
    A = rng.integers(0, 12, size=n).astype(np.uint8)
    out, lens, status = ac_moves_batch(S, A, cyclical=cyclical)
    eo, el, es = O.moves_batch(S, A, cyclical=cyclical)
    assert np.array_equal(status, es)
    assert np.array_equal(out, eo)
    ok = es == 0
    assert np.array_equal(lens[ok], el[ok])


@pytest.mark.parametrize("mrl,n", [(36, 300_000), (24, 100_000), (28, 50_000), (7, 50_000), (18, 50_000), (61, 20_000), (16, 20_000), (3, 5000)])
@pytest.mark.parametrize("cyclical", [True, False])
def test_moves_batch_normalized_variant(mrl, n, cyclical):
    """The steady-state kernel variant (ACS_FLAG_NORMALIZED) on normal-form inputs, chained
    for several steps so that its own outputs feed it again: bit-exact vs the oracle."""
    from ac_solver_b200 import ac_moves_batch

    rng = np.random.default_rng(2000 + mrl)
    S = random_rows(rng, n, mrl, reduced=True)  # freely and cyclically reduced
    if mrl >= 8:  # exercise rejection / long words: half the rows near the maximum length
        S[: n // 2] = random_rows(rng, n // 2, mrl, reduced=True, min_len=max(1, mrl - 3))
    for step in range(4):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        out, lens, status = ac_moves_batch(S, A, cyclical=cyclical, normalized=True, validate=False)
        eo, el, es = O.moves_batch(S, A, cyclical=cyclical)
        assert np.array_equal(status, es)
        assert np.array_equal(out, eo)
        ok = es == 0
        assert np.array_equal(lens[ok], el[ok])
        S = out


def test_moves_batch_edge_cases():
    from ac_solver_b200 import ac_moves_batch

    # empty batch
    out, lens, status = ac_moves_batch(np.zeros((0, 72), np.int8), np.zeros(0, np.uint8))
    assert out.shape == (0, 72)
    # ragged tail (N not a multiple of the 128-row tile) and maximum-length words
    rng = np.random.default_rng(5)
    for n in (1, 127, 129, 1000):
        S = random_rows(rng, n, 36, min_len=36)
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        out, lens, status = ac_moves_batch(S, A)
        eo, el, es = O.moves_batch(S, A)
        assert np.array_equal(out, eo) and np.array_equal(status, es)
    # invalid move ids raise AssertionError in the reference (ac_moves.py:188-190)
    S = random_rows(rng, 4, 8)
    out, lens, status = ac_moves_batch(S, np.array([12, 200, 3, 255], np.uint8))
    assert status.tolist()[0] == 1 and status[1] == 1 and status[3] == 1
    # boundary validation: mis-padded rows and foreign letters are refused loudly
    bad = S.copy()
    bad[0, :3] = [1, 0, 2]
    with pytest.raises(ValueError):
        ac_moves_batch(bad, np.zeros(4, np.uint8))
    bad = S.copy()
    bad[1, 0] = 3
    with pytest.raises(ValueError):
        ac_moves_batch(bad, np.zeros(4, np.uint8))


def test_full_size_properties():
    """BASELINE config 2 size (1 Mi rows, mrl 36): size-independent properties.
    (a) without cyclic reduction, g(.)g^-1 followed by g^-1(.)g is the identity whenever the first
    move is accepted; (b) with cyclic reduction a conjugation of a cyclically reduced word is a
    rotation by at most one letter (lengths, letter multisets and the other relator preserved);
    (c) lengths returned == non-zero counts of the returned rows; (d) r1 -> r1 r0 followed by
    r1 -> r1 r0^-1 restores every row where the first move was accepted."""
    from ac_solver_b200 import ac_moves_batch

    rng = np.random.default_rng(0)
    base = random_rows(rng, 4096, 36)
    n = 1 << 20
    S = np.tile(base, (n // len(base), 1))
    inv = {4: 8, 8: 4, 5: 9, 9: 5, 6: 10, 10: 6, 7: 11, 11: 7}
    A = rng.choice(np.array(list(inv), np.uint8), size=n)
    # (a) without cyclic reduction g(.)g^-1 then g^-1(.)g is the identity whenever the first
    # move is accepted (target relator of at most 34 letters)
    out1, lens1, st1 = ac_moves_batch(S, A, cyclical=False, validate=False)
    assert not st1.any()
    assert np.array_equal(lens1[:, 0], np.count_nonzero(out1[:, :36], axis=1))
    assert np.array_equal(lens1[:, 1], np.count_nonzero(out1[:, 36:], axis=1))
    Ainv = np.vectorize(inv.get, otypes=[np.uint8])(A)
    out2, _, st2 = ac_moves_batch(out1, Ainv, cyclical=False, validate=False)
    assert not st2.any()
    tgt_len = np.where(A % 2 == 1, np.count_nonzero(S[:, :36], axis=1), np.count_nonzero(S[:, 36:], axis=1))
    acc = tgt_len <= 34
    assert acc.sum() > n // 2
    assert np.array_equal(out2[acc], S[acc])
    # with cyclic reduction a conjugation of a cyclically reduced word is a rotation by at
    # most one letter: lengths and the untouched relator are preserved
    cyc, lens_c, st_c = ac_moves_batch(S, A, cyclical=True, validate=False)
    assert not st_c.any()
    assert np.array_equal(lens_c[:, 0], np.count_nonzero(S[:, :36], axis=1))
    assert np.array_equal(lens_c[:, 1], np.count_nonzero(S[:, 36:], axis=1))
    assert np.array_equal(np.sort(cyc[:, :36], axis=1), np.sort(S[:, :36], axis=1))
    assert np.array_equal(np.sort(cyc[:, 36:], axis=1), np.sort(S[:, 36:], axis=1))
    # concat round trip with cyclical=False
    A0 = np.zeros(n, np.uint8)
    c1, l1, s1 = ac_moves_batch(S, A0, cyclical=False, validate=False)
    acc = (c1 != S).any(axis=1)
    c2, l2, s2 = ac_moves_batch(c1, np.full(n, 2, np.uint8), cyclical=False, validate=False)
    ok = acc & (s1 == 0) & (s2 == 0)
    assert ok.sum() > n // 8
    assert np.array_equal(c2[ok], S[ok])


@pytest.mark.parametrize("n,steps", [(5000, 60), (600_001, 3)])  # the second spans 3 pipeline chunks (262 144 rows each)
def test_env_step_batch_vs_oracle(n, steps):
    """acs_env_step_host (device-resident state, host actions in, host results out, chunked
    copy/compute pipeline) == oracle ACEnv.step."""
    import ctypes as C
    import torch
    from ac_solver_b200 import _lib

    L = _lib.lib()
    ctx = _lib.ctx(0)
    rng = np.random.default_rng(11)
    mrl, H = 36, 50
    S = random_rows(rng, n, mrl)
    ref_state = S.copy()
    ref_sc = np.zeros(n, np.int32)
    d_state = torch.from_numpy(S.copy()).cuda()
    d_sc = torch.zeros(n, dtype=torch.int32, device="cuda")
    obs = np.zeros_like(S)
    rew = np.zeros(n, np.int32)
    done = np.zeros(n, np.uint8)
    trunc = np.zeros(n, np.uint8)
    nbad = C.c_int64(0)
    for step in range(steps):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        _lib.check(L.acs_env_step_host(ctx, d_state.data_ptr(), d_sc.data_ptr(), A.ctypes.data, obs.ctypes.data,
                                       rew.ctypes.data, done.ctypes.data, trunc.ctypes.data, n, mrl, H,
                                       2 if step >= steps // 2 else 0,  # both kernel variants (states are normalized)
                                       C.byref(nbad)))
        er, ed, et, el, es = O.env_step_batch(ref_state, A, ref_sc, H)
        ok = es == 0
        assert nbad.value == int((~ok).sum())
        assert np.array_equal(obs, ref_state)
        assert np.array_equal(rew[ok], er[ok]) and np.array_equal(done[ok], ed[ok]) and np.array_equal(trunc[ok], et[ok])
        assert np.array_equal(d_sc.cpu().numpy(), ref_sc)


@pytest.mark.parametrize("mrl", [36, 24, 20])
def test_env_step_lens_carry(mrl):
    """Device-pointer env step with the lengths carried beside the state
    (ACS_FLAG_NORMALIZED | ACS_FLAG_LENS_VALID), chained for 40 steps, vs the oracle."""
    import torch
    from ac_solver_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(100 + mrl)
    n, H = 70_001, 33
    S = random_rows(rng, n, mrl)
    S[: n // 3] = random_rows(rng, n // 3, mrl, min_len=mrl - 2)
    ref_state, ref_sc = S.copy(), np.zeros(n, np.int32)
    d_state = torch.from_numpy(S.copy()).cuda()
    d_sc = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_rew = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_done = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_tr = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_lens = torch.from_numpy(np.stack([np.count_nonzero(S[:, :mrl], axis=1), np.count_nonzero(S[:, mrl:], axis=1)],
                                       axis=1).astype(np.uint8)).cuda()
    err = torch.tensor([0, -1], dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for step in range(40):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        d_a = torch.from_numpy(A).cuda()
        _lib.check(L.acs_env_step_batch(d_state.data_ptr(), d_a.data_ptr(), d_rew.data_ptr(), d_done.data_ptr(),
                                        d_tr.data_ptr(), d_sc.data_ptr(), d_lens.data_ptr(), None, err.data_ptr(),
                                        n, mrl, H, _lib.FLAG_NORMALIZED | _lib.FLAG_LENS_VALID, stream))
        er, ed, et, el, es = O.env_step_batch(ref_state, A, ref_sc, H)
        ok = es == 0
        assert int(err[0]) == int((~ok).sum())
        err[0] = 0
        assert np.array_equal(d_state.cpu().numpy(), ref_state)
        assert np.array_equal(d_lens.cpu().numpy()[ok], el[ok])
        assert np.array_equal(d_rew.cpu().numpy()[ok], er[ok])
        assert np.array_equal(d_done.cpu().numpy()[ok], ed[ok]) and np.array_equal(d_tr.cpu().numpy()[ok], et[ok])
        assert np.array_equal(d_sc.cpu().numpy(), ref_sc)


def test_env_traces_golden(env_traces):
    """ACEnv (single env, reference API) reproduces the reference's trajectories."""
    from ac_solver_b200 import ACEnv, ACEnvConfig

    t = env_traces
    for n in sorted({k.rsplit("_", 1)[0] for k in t.files if k.endswith("_init")}):
        env = ACEnv(ACEnvConfig(initial_state=t[n + "_init"], horizon_length=int(t[n + "_horizon"])))
        acts = t[n + "_actions"]
        for k, a in enumerate(acts[:80]):
            s, r, d, tr, info = env.step(int(a))
            assert np.array_equal(s, t[n + "_states"][k])
            assert int(r) == int(t[n + "_rewards"][k]) and bool(d) == bool(t[n + "_dones"][k])
            assert bool(tr) == bool(t[n + "_truncs"][k])
            if d:
                assert info["actions"] == [int(x) for x in acts[: k + 1]]


# ---- the reference's own unit vectors through the drop-in API (tests/test_ac_env.py) ----
def test_reference_unit_vectors(unit_vectors):
    from ac_solver_b200.envs.ac_moves import ACMove, concatenate_relators, conjugate
    from ac_solver_b200.envs.utils import simplify_presentation, simplify_relator

    for c in unit_vectors["simplify_relator"]:  # :18-83 (letters up to +-3)
        r, l = simplify_relator(np.array(c["relator"]), c["mrl"], cyclical=c["cyclical"], padded=c["padded"])
        assert r.tolist() == c["expected_relator"] and l == c["expected_length"], c
    for c in unit_vectors["simplify_presentation"]:  # :140-181
        r, l = simplify_presentation(np.array(c["presentation"]), c["mrl"], c["lengths"])
        assert r.tolist() == c["expected"] and l == c["expected_lengths"], c
    for c in unit_vectors["concatenate_relators"]:  # :184-326
        r, l = concatenate_relators(np.array(c["rels"]), c["mrl"], c["i"], c["j"], c["sign"], list(c["lengths"]))
        assert r.tolist() == c["expected"] and list(l) == c["expected_lengths"], c
    for c in unit_vectors["conjugate"]:  # :329-477
        r, l = conjugate(np.array(c["rels"]), c["mrl"], c["i"], c["j"], c["sign"], list(c["lengths"]))
        assert r.tolist() == c["expected"] and list(l) == c["expected_lengths"], c
    for c in unit_vectors["ACMove"]:  # :495-538, all 12 ids, lengths deliberately wrong
        r, l = ACMove(c["move_id"], np.array(c["presentation"]), c["mrl"], [4, 4], cyclical=c["cyclical"])
        assert r.tolist() == c["expected"] and l == c["expected_lengths"], c
    with pytest.raises(AssertionError):  # r1 = r0 -> r0 r1^-1 empties r0 (utils.py:261-263)
        ACMove(1, np.array([1, 2, 0, 1, 2, 0]), 3, [2, 2])
    with pytest.raises(AssertionError):
        ACMove(12, np.array([1, 0, 2, 0]), 2, [1, 1])


def test_generic_kernel_vs_oracle_any_alphabet():
    """Generic byte kernel with letters up to +-5 against the oracle."""
    from ac_solver_b200 import _lib
    from ac_solver_b200._host import generic_call

    rng = np.random.default_rng(3)
    n, mrl = 4000, 9
    S = np.zeros((n, 2 * mrl), np.int8)
    for k in range(n):
        for h in range(2):
            L = int(rng.integers(0 if k % 50 == 0 else 1, mrl + 1))
            w = rng.integers(1, 6, size=L) * rng.choice([-1, 1], size=L)
            S[k, h * mrl : h * mrl + L] = w
    A = rng.integers(0, 12, size=n).astype(np.uint8)
    for cyc in (True, False):
        out, aux, status = generic_call(_lib.OP_ACMOVE, S, actions=A, cyclical=cyc)
        eo, el, es = O.moves_batch(S, A, cyclical=cyc)
        assert np.array_equal(status, es)
        assert np.array_equal(out, eo)
        assert np.array_equal(aux[es == 0], el[es == 0])
