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


def _pad(rows, mrl_from, mrl_to):
    out = np.zeros((len(rows), 2 * mrl_to), np.int8)
    out[:, :mrl_from] = rows[:, :mrl_from]
    out[:, mrl_to : mrl_to + mrl_from] = rows[:, mrl_from:]
    return out


@pytest.mark.parametrize("mrl", [7, 13, 36])
def test_step_device_any_width_and_non_normal_initial_states(mrl):
    """The sync-free path for max_relator_length % 4 != 0 and for initial states that are not normal
    forms (general kernel variant), auto-reset included, vs the looped oracle."""
    import torch
    from ac_solver_b200.envs.vector_env import ACVectorEnv

    rng = np.random.default_rng(mrl)
    n, H = 300, 9
    init = np.zeros((n, 2 * mrl), np.int8)
    for i in range(n):
        for h in range(2):
            ln = int(rng.integers(1, min(mrl, 6) + 1))
            init[i, h * mrl : h * mrl + ln] = rng.choice([-2, -1, 1, 2], size=ln)  # may be non-reduced
    init[0, :mrl] = 0
    init[0, 0], init[0, mrl] = 1, 2  # <x, y>: solved by any move that keeps total length 2
    env = ACVectorEnv(init, horizon_length=H)
    env.reset()
    ref_state, ref_sc = init.copy(), np.zeros(n, np.int32)
    for step in range(40):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        obs, rew, done, trunc = env.step_device(torch.from_numpy(A).cuda())
        er, ed, et, el, es = O.env_step_batch(ref_state, A, ref_sc, H)
        ok = es == 0
        env.err[0] = 0  # raising rows keep their state; their outputs are cleared
        fin = (ed.astype(bool) | et.astype(bool)) & ok
        assert np.array_equal(env.final_obs.cpu().numpy()[fin], ref_state[fin])
        ref_state[fin], ref_sc[fin] = init[fin], 0
        assert np.array_equal(obs.cpu().numpy(), ref_state)
        assert np.array_equal(rew.cpu().numpy()[ok], er[ok]) and np.array_equal(done.cpu().numpy()[ok], ed[ok])
        assert np.array_equal(trunc.cpu().numpy()[ok], et[ok])
        assert np.array_equal(env.step_count.cpu().numpy(), ref_sc)


def test_reward_normalize_and_clip(miller_schupp):
    """NormalizeReward (gymnasium 0.28.1 formulas, one wrapper per environment) + TransformReward clip on
    the device vs a numpy restatement."""
    import torch
    from ac_solver_b200.envs.vector_env import ACVectorEnv

    n, H, gamma, clip = 64, 12, 0.99, (-10.0, 1000.0)
    init = _initial_states(miller_schupp, n)
    env = ACVectorEnv(init, horizon_length=H, clip_rewards=clip, norm_rewards=True, gamma=gamma)
    env.reset()
    returns, mean, var, count = np.zeros(n), np.zeros(n), np.ones(n), np.full(n, 1e-4)
    rng = np.random.default_rng(2)
    for step in range(60):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        A[:6] = rng.choice([0, 1, 2, 3], size=6)
        env.step_device(torch.from_numpy(A).cuda())
        got = env.transformed_reward().cpu().numpy()
        r = env.reward.cpu().numpy().astype(np.float64)
        d = env.done.cpu().numpy().astype(np.float64)
        returns = returns * gamma * (1 - d) + r
        delta, tot = returns - mean, count + 1
        mean = mean + delta / tot
        var = (var * count + delta * delta * count / tot) / tot
        count = tot
        exp = np.clip(r / np.sqrt(var + 1e-8), *clip)
        assert np.allclose(got, exp.astype(np.float32), rtol=1e-6, atol=1e-6)


def test_device_curriculum_round_one(miller_schupp):
    """The device-side curriculum reset against a host emulation of the reference's loop
    (training.py:169-224) while round one lasts (sequential hand-out of the unprocessed states):
    same states, same current-state indices, same success record and ACMoves_hist."""
    import torch
    from ac_solver_b200.envs.vector_env import ACVectorEnv

    n, H, mrl = 16, 8, 36
    pool = np.concatenate([_initial_states(miller_schupp, 200), _initial_states(miller_schupp, 200)[::-1]])
    env = ACVectorEnv(pool[:n], horizon_length=H)
    env.reset()
    env.enable_curriculum(pool, repeat_solved_prob=0.25, seed=3)
    ref_state, ref_sc = pool[:n].copy(), np.zeros(n, np.int32)
    cur, nxt = list(range(n)), n
    solved, hist, logs = set(), {}, [[] for _ in range(n)]
    rng = np.random.default_rng(9)
    for step in range(150):
        A = rng.integers(0, 12, size=n).astype(np.uint8)
        A[::3] = rng.choice([0, 1, 2, 3], size=len(A[::3]))
        env.step_device(torch.from_numpy(A).cuda())
        for i in range(n):
            logs[i].append(int(A[i]))
        er, ed, et, el, es = O.env_step_batch(ref_state, A, ref_sc, H)
        assert not es.any()
        for i in range(n):
            if ed[i]:
                solved.add(cur[i])
                if cur[i] not in hist or len(logs[i]) < len(hist[cur[i]]):
                    hist[cur[i]] = list(logs[i])
            if ed[i] or et[i]:
                assert nxt < len(pool), "the test must stay inside round one"
                cur[i], nxt = nxt, nxt + 1
                ref_state[i], ref_sc[i], logs[i] = pool[cur[i]], 0, []
        assert np.array_equal(env.state.cpu().numpy(), ref_state)
        assert np.array_equal(env._curriculum["cur_state"].cpu().numpy(), np.array(cur))
    env.check_errors()
    assert env.success_record()["solved"] == solved
    got = env.acmoves_hist()
    assert set(got) == set(hist) and all(len(got[k]) == len(hist[k]) for k in hist)
    # (two environments never hold the same pool state in round one, so the sequences are unique)
    assert got == hist
    assert env.curriculum_counters()["next_unprocessed"] == nxt


def test_device_curriculum_round_two_invariants(miller_schupp):
    """After round one the next state is random: an unsolved one while nothing is solved, always a
    valid pool index, and the rollout runs inside a CUDA graph without host involvement."""
    import torch
    from ac_solver_b200.envs.vector_env import ACVectorEnv

    n, H = 32, 3
    pool = _initial_states(miller_schupp, 40)
    env = ACVectorEnv(pool[:n], horizon_length=H, clip_rewards=(-10, 1000))
    env.reset()
    env.enable_curriculum(pool, repeat_solved_prob=0.5, seed=11)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.randint(4, 12, (n,), device="cuda", generator=g, dtype=torch.int64)  # conjugations only: nothing gets solved
    for _ in range(30):
        env.step_device(acts)
    cs = env._curriculum["cur_state"].cpu().numpy()
    assert cs.min() >= 0 and cs.max() < len(pool)
    c = env.curriculum_counters()
    assert c["next_unprocessed"] == len(pool) and c["n_solved"] == 0 and c["random_draws"] > 0
    assert env.success_record()["solved"] == set()
    env.check_errors()
