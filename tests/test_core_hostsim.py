"""CPU check of the DEVICE headers (csrc/ac_core.cuh, csrc/ac_pack.cuh) compiled for the host
with shimmed intrinsics, against the oracle and the reference-generated golden triples.
Covers both kernel variants (general / ACS_FLAG_NORMALIZED) and both load paths."""

import numpy as np
import pytest

from oracle import oracle as O
import hostsim
from ac_solver_b200.synthetic import random_presentations


@pytest.mark.parametrize("mrl", [4, 7, 10, 12, 18, 24, 36])
@pytest.mark.parametrize("use_words", [True, False])
def test_hostsim_golden(acmove_random, mrl, use_words):
    g = acmove_random
    S, A, Cy, Oo, Ln, St = g[f"s{mrl}"], g[f"a{mrl}"], g[f"c{mrl}"], g[f"o{mrl}"], g[f"l{mrl}"], g[f"t{mrl}"]
    for cyc in (0, 1):
        m = Cy == cyc
        out, lens, status = hostsim.moves(S[m], A[m], cyclical=bool(cyc), trusted=False, use_words=use_words)
        assert np.array_equal(status, St[m])
        ok = St[m] == 0
        assert np.array_equal(out[ok], Oo[m][ok])
        assert np.array_equal(lens[ok], Ln[m][ok])


@pytest.mark.parametrize("mrl", [1, 2, 3, 5, 8, 15, 16, 17, 20, 28, 31, 32, 33, 36, 40, 47, 48, 49, 60, 61, 64])
@pytest.mark.parametrize("cyclical", [True, False])
def test_hostsim_normalized_chain(mrl, cyclical):
    """ACS_FLAG_NORMALIZED variant chained on its own outputs, vs the oracle."""
    rng = np.random.default_rng(mrl)
    n = 6000
    S = random_presentations(n, mrl, seed=mrl)
    if mrl >= 6:
        S[: n // 2] = random_presentations(n // 2, mrl, seed=mrl + 100, min_len=mrl - 2)
    for step in range(5):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        eo, el, es = O.moves_batch(S, A, cyclical=cyclical)
        for trusted in (True, False):
            out, lens, status = hostsim.moves(S, A, cyclical=cyclical, trusted=trusted, use_words=True)
            assert np.array_equal(status, es), (step, trusted)
            assert np.array_equal(out, eo), (step, trusted)
            assert np.array_equal(lens[es == 0], el[es == 0])
        S = eo


@pytest.mark.parametrize("mrl", [6, 16, 36, 50])
def test_hostsim_unreduced_inputs(mrl):
    rng = np.random.default_rng(7 + mrl)
    n = 8000
    S = np.zeros((n, 2 * mrl), np.int8)
    for h in range(2):
        L = rng.integers(0, mrl + 1, size=n)
        w = rng.choice(np.array([-2, -1, 1, 2], np.int8), size=(n, mrl))
        S[:, h * mrl : (h + 1) * mrl] = np.where(np.arange(mrl)[None, :] < L[:, None], w, 0)
    A = rng.integers(0, 13, size=n).astype(np.uint8)
    for cyc in (True, False):
        eo, el, es = O.moves_batch(S, np.minimum(A, 12), cyclical=cyc)
        out, lens, status = hostsim.moves(S, A, cyclical=cyc, trusted=False, use_words=mrl % 4 == 0)
        assert np.array_equal(status, es)
        assert np.array_equal(out, eo)
