"""CPU-side checks of the boundary: the shared library loads without a GPU, exports every
symbol include/acsolver_b200.h declares, and compute entry points fail loudly (no fallback)."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "acsolver_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ac_solver_b200 import _lib
    from ac_solver_b200.build import build

    build()
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    # the ctypes signature table covers the same set
    assert sorted(_lib.exported_symbols()) == names
    assert _lib.lib().acs_version() >= 100


def test_no_cpu_fallback():
    import torch
    from ac_solver_b200 import ACMove, _lib, ac_moves_batch, bfs, greedy_search

    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only hosts")
    ak2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0])
    with pytest.raises(_lib.AcsError):
        ACMove(0, ak2, 7, [5, 6])
    with pytest.raises(_lib.AcsError):
        ac_moves_batch(ak2[None, :].astype(np.int8), np.zeros(1, np.uint8))
    with pytest.raises(_lib.AcsError):
        bfs(ak2, 100)
    with pytest.raises(_lib.AcsError):
        greedy_search(ak2, 100)


def test_host_helpers_match_reference_vectors(unit_vectors):
    """Pure host logic (predicates / layout converters), reference tests/test_ac_env.py:86-137
    and tests/envs/test_envs_utils.py:6-14."""
    from ac_solver_b200.envs.utils import (change_max_relator_length_of_presentation,
                                           convert_relators_to_presentation, generate_trivial_states,
                                           is_array_valid_presentation, is_presentation_trivial)

    for c in unit_vectors["is_array_valid_presentation"]:
        assert is_array_valid_presentation(np.array(c["presentation"])) == c["expected"], c
    for c in unit_vectors["is_presentation_trivial"]:
        assert is_presentation_trivial(np.array(c["presentation"])) == c["expected"], c
    for m in (1, 2, 3, 4):
        st = generate_trivial_states(m)
        assert st.shape == (8, 2 * m)
        for s in st:
            assert np.count_nonzero(s) == 2 and abs(s[0]) != abs(s[m]) and s[0] != 0 and s[m] != 0
    ak2 = convert_relators_to_presentation([1, 1, -2, -2, -2], [1, 2, 1, -2, -1, -2], 7)
    assert ak2.dtype == np.int8
    assert ak2.tolist() == [1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0]
    wide = change_max_relator_length_of_presentation(ak2, 10)
    assert wide.tolist() == [1, 1, -2, -2, -2, 0, 0, 0, 0, 0, 1, 2, 1, -2, -1, -2, 0, 0, 0, 0]


def test_env_config_errors():
    from ac_solver_b200 import ACEnv, ACEnvConfig

    with pytest.raises(TypeError):
        ACEnvConfig(initial_state=(1, 0, 2, 0))
    with pytest.raises(ValueError):
        ACEnvConfig(initial_state=np.array([[1, 0], [2, 0]]))
    with pytest.raises(ValueError):
        ACEnvConfig(initial_state=np.array([1, 0, 2]))
    with pytest.raises(ValueError):
        ACEnvConfig(initial_state=np.array([1, 0, 0, 0]))
    with pytest.raises(NotImplementedError):
        ACEnv(ACEnvConfig(use_supermoves=True))
    cfg = ACEnvConfig.from_dict({"initial_state": [1, 1, 0, 2, 0, 0], "horizon_length": 7})
    env = ACEnv(cfg)
    assert env.max_relator_length == 3 and env.max_reward == 7 * 3 * 2 and env.lengths == [2, 1]
    assert env.observation_space.shape == (6,) and env.action_space.n == 12


def test_shard_owner_is_independent_of_table_slot():
    """The owner rank of a key must not pin bits of its table slot (a correlation clustered the
    per-rank tables at 8 GPUs): for keys of one owner, every 3-bit window of the low 32 hash bits
    stays uniform."""
    from ac_solver_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(0)
    hs = rng.integers(0, 2**63, size=40000, dtype=np.int64).astype(np.uint64)
    for world in (2, 3, 8):
        owners = np.array([L.acs_sbfs_owner(int(h), world) for h in hs])
        counts = np.bincount(owners, minlength=world)
        assert counts.min() > 0.8 * len(hs) / world
        mine = hs[owners == 0]
        for shift in range(0, 30):
            win = ((mine >> np.uint64(shift)) & np.uint64(7)).astype(np.int64)
            c = np.bincount(win, minlength=8)
            assert c.min() > 0.6 * len(mine) / 8, (world, shift, c)


def test_vector_env_normal_form_mask_matches_oracle():
    """The host predicate that gates the steady-state kernel variant in ACVectorEnv must agree with
    "simplify_presentation(cyclical=True) is the identity" for every row."""
    from ac_solver_b200.envs.vector_env import _lens_of, _normal_form_mask
    from ac_solver_b200.synthetic import random_presentations
    from oracle import oracle as O

    rng = np.random.default_rng(3)
    mrl, n = 12, 6000
    S = np.zeros((n, 2 * mrl), np.int8)
    for h in range(2):
        L = rng.integers(1, mrl + 1, size=n)
        w = rng.choice(np.array([-2, -1, 1, 2], np.int8), size=(n, mrl))
        S[:, h * mrl : (h + 1) * mrl] = np.where(np.arange(mrl)[None, :] < L[:, None], w, 0)
    S[: n // 3] = random_presentations(n // 3, mrl, seed=1)  # plenty of genuine normal forms
    mask = _normal_form_mask(S)
    expect = np.zeros(n, bool)
    for k in range(n):
        ok = True
        for h in range(2):
            w = S[k, h * mrl : (h + 1) * mrl]
            r, _ = O.simplify_relator(w, mrl, cyclical=True)
            ok = ok and np.array_equal(r, w)
        expect[k] = ok
    assert np.array_equal(mask, expect)
    assert 0 < mask.sum() < n
    lens = _lens_of(S)
    assert np.array_equal(lens[:, 0], np.count_nonzero(S[:, :mrl], axis=1))
    assert np.array_equal(lens[:, 1], np.count_nonzero(S[:, mrl:], axis=1))
