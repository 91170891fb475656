"""GPU-resident vector env vs a per-environment loop of the oracle's ACEnv.step with gymnasium
0.28.1 SyncVectorEnv auto-reset rules (unpinned by the reference's tests)."""

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _initial_states(ms, n, mrl=36):
    rows = ms["presentations36"][: n - 6].astype(np.int8)
    easy = np.zeros((6, 2 * mrl), np.int8)
    for k, (a, b) in enumerate([([1, 2], [-1]), ([2], [1, 2]), ([1, 1, 2], [1]), ([-2, 1], [-2]), ([1], [2, 2, 1]), ([1, 2, 2], [2])]):
        easy[k, : len(a)], easy[k, mrl : mrl + len(b)] = a, b
    return np.concatenate([easy, rows])


@pytest.mark.parametrize("numpy_io", [True, False])
def test_vector_env_matches_looped_oracle(miller_schupp, numpy_io):
    import torch
    from ac_solver_b200.envs.vector_env import ACVectorEnv

    n, H, clip = 200, 17, (-10, 1000)
    init = _initial_states(miller_schupp, n)
    env = ACVectorEnv(init, horizon_length=H, clip_rewards=clip)
    assert env.single_observation_space.shape == (72,) and env.single_action_space.n == 12
    assert env.envs[0].max_reward == H * 36 * 2
    obs0, _ = env.reset()
    assert np.array_equal(obs0, init)
    ref_state = init.copy()
    ref_sc = np.zeros(n, np.int32)
    ref_log = [[] for _ in range(n)]
    rng = np.random.default_rng(5)
    n_done = n_trunc = 0
    for step in range(90):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        A[:6] = rng.choice([0, 1, 2, 3], size=6)  # concatenations solve the easy rows quickly
        if step == 40:  # curriculum hook: plant new starting states into a few envs
            for i in (7, 9):
                env.envs[i].reset(options={"starting_state": init[i + 20]})
                ref_state[i], ref_sc[i], ref_log[i] = init[i + 20], 0, []
        out = env.step(A if numpy_io else torch.from_numpy(A).cuda())
        obs, rew, done, trunc, infos = out
        if not numpy_io:
            obs, rew, done, trunc = (x.cpu().numpy() for x in (obs, rew, done, trunc))
        for i in range(n):
            ref_log[i].append(int(A[i]))
        er, ed, et, el, es = O.env_step_batch(ref_state, A, ref_sc, H)
        assert not es.any()
        exp_rew = np.clip(er.astype(np.float64), *clip)
        fin = ed.astype(bool) | et.astype(bool)
        exp_obs = ref_state.copy()
        for i in np.flatnonzero(fin):
            assert infos["_final_observation"][i] and infos["_final_info"][i]
            assert np.array_equal(infos["final_observation"][i], ref_state[i])
            assert infos["final_info"][i] == ({"actions": ref_log[i]} if ed[i] else {})
            ref_state[i], ref_sc[i], ref_log[i] = init[i], 0, []  # auto-reset to the env's own initial state
            exp_obs[i] = init[i]
        if fin.any():
            assert int(infos["_final_info"].sum()) == int(fin.sum())
        else:
            assert infos == {}
        assert np.array_equal(obs, exp_obs)
        assert np.array_equal(rew, exp_rew) and rew.dtype == np.float64
        assert np.array_equal(done, ed.astype(bool)) and np.array_equal(trunc, et.astype(bool))
        n_done += int(ed.sum())
        n_trunc += int(et.sum())
    assert n_done > 0 and n_trunc > 0


def test_step_device_matches_looped_oracle(miller_schupp):
    """The sync-free path (fused step + auto-reset kernels, lengths carried, action log on the
    device) against the same per-environment oracle loop."""
    import torch
    from ac_solver_b200.envs.vector_env import ACVectorEnv

    n, H = 300, 13
    init = _initial_states(miller_schupp, n)
    env = ACVectorEnv(init, horizon_length=H)
    env.reset()
    ref_state, ref_sc = init.copy(), np.zeros(n, np.int32)
    ref_log = [[] for _ in range(n)]
    rng = np.random.default_rng(9)
    n_done = 0
    for step in range(70):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        A[:6] = rng.choice([0, 1, 2, 3], size=6)
        obs, rew, done, trunc = env.step_device(torch.from_numpy(A).cuda())
        for i in range(n):
            ref_log[i].append(int(A[i]))
        er, ed, et, el, es = O.env_step_batch(ref_state, A, ref_sc, H)
        assert not es.any()
        fin = ed.astype(bool) | et.astype(bool)
        final_obs = env.final_obs.cpu().numpy()
        final_steps = env.final_steps.cpu().numpy()
        for i in np.flatnonzero(fin):
            assert np.array_equal(final_obs[i], ref_state[i])
            assert final_steps[i] == len(ref_log[i])
            if ed[i]:
                assert env.final_actions(i) == ref_log[i]
                n_done += 1
            ref_state[i], ref_sc[i], ref_log[i] = init[i], 0, []
        assert np.array_equal(obs.cpu().numpy(), ref_state)
        assert np.array_equal(rew.cpu().numpy(), er)
        assert np.array_equal(done.cpu().numpy(), ed) and np.array_equal(trunc.cpu().numpy(), et)
        assert np.array_equal(env.step_count.cpu().numpy(), ref_sc)
        assert np.array_equal(env.lens.cpu().numpy()[:, 0], np.count_nonzero(ref_state[:, :36], axis=1))
    env.check_errors()
    assert n_done > 0


def test_vector_env_errors(miller_schupp):
    from ac_solver_b200.envs.vector_env import ACVectorEnv

    with pytest.raises(ValueError):
        ACVectorEnv(np.array([[1, 0, 0, 0]]))
    with pytest.raises(NotImplementedError):
        ACVectorEnv(np.array([[1, 0, 2, 0]]), use_supermoves=True)
    env = ACVectorEnv(np.array([[1, 2, 0, 1, 2, 0]]), horizon_length=5)
    with pytest.raises(AssertionError):  # r0 -> r0 r1^-1 empties r0
        env.step(np.array([1], np.uint8))
