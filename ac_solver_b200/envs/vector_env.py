"""GPU-resident vector environment -- batched stand-in for the ``gym.vector.SyncVectorEnv`` of
``ACEnv`` instances that the reference's PPO rollout steps one environment at a time on the host
(``ac_solver/agents/environment.py:60-127``, ``ac_solver/agents/training.py:154-228``).

All environment state (presentations, step counters, action logs) lives on the GPU; one call of
``step`` is one launch of the fused env-step kernel (``acs_env_step_batch``) plus a few small
torch ops for the auto-reset.  Semantics follow gymnasium 0.28.1's ``SyncVectorEnv`` as used by
the reference: when an environment terminates or truncates it is reset to its own initial state,
the returned observation is the reset observation, and ``infos["final_observation"]`` /
``infos["final_info"]`` (with ``{"actions": [...]}`` for solved episodes, ``ac_env.py:107-113``)
carry the last step of the finished episode.  The reference's own tests do not pin these
wrapper semantics; tests/test_vector_env_gpu.py pins them against a per-environment
loop of the oracle's ``ACEnv.step``.

Inputs decide the output container: numpy actions give numpy outputs (drop-in for
``training.py``), a CUDA tensor gives CUDA tensors (rollouts that never leave the device).
"""

from __future__ import annotations

import numpy as np

from .. import _lib
from .spaces import Box, Discrete
from .utils import is_array_valid_presentation


def _lens_of(states):
    """[n, 2*mrl] int8 -> uint8 [n, 2] relator lengths."""
    m = states.shape[1] // 2
    return np.stack([np.count_nonzero(states[:, :m], axis=1), np.count_nonzero(states[:, m:], axis=1)],
                    axis=1).astype(np.uint8)


def _normal_form_mask(states):
    """True for rows whose two words are freely AND cyclically reduced (what ACMove with
    cyclical=True leaves behind, and what ACS_FLAG_NORMALIZED promises)."""
    m = states.shape[1] // 2
    ok = np.ones(len(states), dtype=bool)
    rows = np.arange(len(states))
    for h in (states[:, :m].astype(np.int16), states[:, m:].astype(np.int16)):
        ln = np.count_nonzero(h, axis=1)
        if m > 1:
            ok &= ~(((h[:, :-1] + h[:, 1:]) == 0) & (h[:, :-1] != 0)).any(axis=1)
        last = h[rows, np.maximum(ln - 1, 0)]
        ok &= ~((ln >= 2) & (h[:, 0] == -last))
    return ok


class _EnvProxy:
    """``envs.envs[i]``: the two things the PPO loop touches (training.py:224,233)."""

    def __init__(self, owner, index):
        self._owner, self._index = owner, index

    @property
    def max_reward(self):
        return self._owner.max_reward

    @property
    def state(self):
        return self._owner.state[self._index].cpu().numpy()

    def reset(self, *, seed=None, options=None):
        """ac_env.py:115-131 for one environment (``options["starting_state"]`` honoured)."""
        o = self._owner
        src = options["starting_state"] if options and "starting_state" in options else o.initial_states_host[self._index]
        o.set_states([self._index], np.asarray(src)[None, :])
        return np.array(src, dtype=np.int8), {}


class ACVectorEnv:
    def __init__(self, initial_states, horizon_length=1000, device=None, clip_rewards=None, use_supermoves=False,
                 norm_rewards=False, gamma=0.99):
        """initial_states: [N, 2*mrl] presentations (one per environment), all valid
        (``ACEnvConfig.__post_init__``, ac_env.py:22-35).  clip_rewards: optional (min, max) as in
        ``TransformReward(np.clip)`` (environment.py:48-52).  norm_rewards / gamma: gymnasium's
        ``NormalizeReward`` per environment (environment.py:45-46), applied before the clip."""
        import torch

        if use_supermoves:
            raise NotImplementedError("ACEnv with supermoves is not yet implemented in this library.")
        init = np.asarray(initial_states)
        if init.ndim != 2 or init.shape[1] % 2:
            raise ValueError("initial_states must be [num_envs, 2*max_relator_length]")
        for row in init:
            if not is_array_valid_presentation(row):
                raise ValueError("initial state must be a valid presentation")
        if np.abs(init).max() > 2:
            raise ValueError("the GPU environment supports the two-generator alphabet {+-1, +-2} only")
        if not torch.cuda.is_available():
            raise _lib.AcsError("no CUDA device visible; ACVectorEnv has no CPU fallback")
        self.torch = torch
        self.L = _lib.lib()
        self.dev = torch.device("cuda", _lib.default_device() if device is None else device)
        self.num_envs, width = init.shape
        self.max_relator_length = width // 2
        self.horizon_length = int(horizon_length)
        self.max_reward = self.horizon_length * self.max_relator_length * 2  # ac_env.py:80
        self.clip_rewards = clip_rewards
        bound = np.full(width, 2, dtype=np.int8)
        self.single_observation_space = Box(-bound, bound, dtype=np.int8)
        self.single_action_space = Discrete(12)
        self.initial_states_host = init.astype(np.int8)
        self.initial_states = torch.from_numpy(self.initial_states_host).to(self.dev)
        n = self.num_envs
        self.state = self.initial_states.clone()
        self.step_count = torch.zeros(n, dtype=torch.int32, device=self.dev)
        self.reward = torch.zeros(n, dtype=torch.int32, device=self.dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        self.truncated = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        self.action_log = torch.zeros((n, max(self.horizon_length, 1)), dtype=torch.uint8, device=self.dev)
        self.err = torch.tensor([0, -1], dtype=torch.int64, device=self.dev)
        self.initial_normal_host = _normal_form_mask(self.initial_states_host)
        self.initial_lens = torch.from_numpy(_lens_of(self.initial_states_host)).to(self.dev)
        self.lens = self.initial_lens.clone()  # ACEnv.lengths (ac_env.py:84-92), kernel-maintained
        self._rows = torch.arange(n, device=self.dev)
        # caller-supplied states may be non-reduced: a step after planting such a state runs the
        # general kernel variant; states produced by the kernel are normal forms, and so are most
        # datasets (ACS_FLAG_NORMALIZED | ACS_FLAG_LENS_VALID is the steady state).
        self._normalized = bool(self.initial_normal_host.all())
        # the sync-free path runs ONE kernel variant per environment object: the steady-state variant when
        # every initial state is a normal form, else the general one (full simplification of both relators)
        self._device_flags = (_lib.FLAG_NORMALIZED | _lib.FLAG_LENS_VALID) if self._normalized else 0
        self.norm_rewards, self.gamma = bool(norm_rewards), float(gamma)
        self.reward_stats = torch.zeros((4, n), dtype=torch.float64, device=self.dev)  # returns, mean, var, count
        self.reward_stats[2] = 1.0
        self.reward_stats[3] = 1e-4
        self.reward_out = torch.zeros(n, dtype=torch.float32, device=self.dev)
        self._curriculum = None
        self.final_obs = torch.zeros_like(self.state)
        self.final_steps = torch.zeros(n, dtype=torch.int32, device=self.dev)
        self.envs = [_EnvProxy(self, i) for i in range(n)]

    # ---------------------------------------------------------------------------------------
    def set_states(self, indices, states):
        """Plant caller-supplied presentations into the given environments (per-env ``reset`` with
        ``starting_state``; the curriculum hook of training.py:223-224), resetting their counters."""
        t = self.torch
        s = np.ascontiguousarray(states, dtype=np.int8)
        for row in s:
            assert is_array_valid_presentation(row), f"{row} is not a valid presentation"
        idx = t.as_tensor(np.asarray(indices, dtype=np.int64), device=self.dev)
        self.state[idx] = t.from_numpy(s).to(self.dev)
        self.step_count[idx] = 0
        # normal forms keep the steady-state kernel variant valid; anything else forces one
        # general step (which re-simplifies and recounts every row)
        self.lens[idx] = t.from_numpy(_lens_of(s)).to(self.dev)
        if not _normal_form_mask(s).all():
            self._normalized = False
            self._device_flags = 0  # from now on the sync-free path runs the general kernel variant

    def reset(self, *, seed=None, options=None):
        """All environments back to their initial states -> (obs, {})."""
        self.state.copy_(self.initial_states)
        self.step_count.zero_()
        self.lens.copy_(self.initial_lens)
        self._normalized = bool(self.initial_normal_host.all())
        if self._curriculum is not None:  # environment i is on pool state i again (the solved record is kept)
            self._curriculum["cur_state"].copy_(self.torch.arange(self.num_envs, dtype=self.torch.int32, device=self.dev))
        return self.state.cpu().numpy(), {}

    def step(self, actions):
        t = self.torch
        as_numpy = not isinstance(actions, t.Tensor)
        act = t.as_tensor(np.ascontiguousarray(actions, dtype=np.uint8) if as_numpy else actions).to(
            device=self.dev, dtype=t.uint8).contiguous()
        if act.shape != (self.num_envs,):
            raise ValueError("actions must have shape (num_envs,)")
        # log the action at the pre-step counter (info["actions"] of a finished episode)
        pos = self.step_count.to(t.int64).clamp_(max=self.action_log.shape[1] - 1)
        self.action_log[self._rows, pos] = act
        self.err[0], self.err[1] = 0, -1
        # steady state: states are normal forms and self.lens is current (written by the last step)
        flags = (_lib.FLAG_NORMALIZED | _lib.FLAG_LENS_VALID) if self._normalized else 0
        stream = t.cuda.current_stream(self.dev).cuda_stream
        _lib.check(self.L.acs_env_step_batch(
            self.state.data_ptr(), act.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
            self.truncated.data_ptr(), self.step_count.data_ptr(), self.lens.data_ptr(), None, self.err.data_ptr(),
            self.num_envs, self.max_relator_length, self.horizon_length, flags, stream))
        self._normalized = True
        finished = (self.done | self.truncated).bool()
        n_bad, any_fin = (int(v) for v in t.stack([self.err[0], finished.any().to(t.int64)]).cpu())
        if n_bad:
            raise AssertionError(f"{n_bad} environments produced an invalid presentation (first: env {int(self.err[1])}); "
                                 "the reference raises AssertionError here (envs/utils.py:261-263)")
        if self.norm_rewards:
            reward = self.transformed_reward().to(t.float64)
        else:
            reward = self.reward.to(t.float64)
            if self.clip_rewards is not None:
                reward = reward.clamp(self.clip_rewards[0], self.clip_rewards[1])
        infos = {}
        obs = self.state
        if any_fin:
            fin_idx = finished.nonzero().flatten()
            final_obs = self.state[fin_idx].cpu().numpy()
            lens = self.step_count[fin_idx].cpu().numpy()
            logs = self.action_log[fin_idx].cpu().numpy()
            solved = self.done[fin_idx].cpu().numpy().astype(bool)
            fo = np.full(self.num_envs, None, dtype=object)
            fi = np.full(self.num_envs, None, dtype=object)
            mask = np.zeros(self.num_envs, dtype=bool)
            fin_host = fin_idx.cpu().numpy()
            for k, i in enumerate(fin_host):
                fo[i] = final_obs[k]
                fi[i] = {"actions": [int(a) for a in logs[k, : lens[k]]]} if solved[k] else {}
                mask[i] = True
            infos = {"final_observation": fo, "_final_observation": mask, "final_info": fi, "_final_info": mask.copy()}
            # auto-reset: the env's own initial state, counters zeroed (gymnasium SyncVectorEnv)
            self.state[fin_idx] = self.initial_states[fin_idx]
            self.step_count[fin_idx] = 0
            self.lens[fin_idx] = self.initial_lens[fin_idx]
            if not self.initial_normal_host[fin_host].all():
                self._normalized = False  # a non-normal initial state: next step re-simplifies
        if as_numpy:
            # one device-to-host copy for the four outputs: [obs | reward f64 | done | truncated] as bytes
            n, w = self.num_envs, self.state.shape[1]
            packed = t.cat([obs.reshape(-1).view(t.uint8), reward.contiguous().view(t.uint8), self.done, self.truncated]).cpu().numpy()
            o = packed[: n * w].view(np.int8).reshape(n, w).copy()
            r = packed[n * w : n * w + 8 * n].view(np.float64).copy()
            d = packed[n * w + 8 * n : n * w + 9 * n].astype(bool)
            tr = packed[n * w + 9 * n :].astype(bool)
            return o, r, d, tr, infos
        return obs.clone(), reward, self.done.bool(), self.truncated.bool(), infos

    # ---------------------------------------------------------------------------------------
    def step_device(self, actions):
        """GPU-resident step without any host synchronisation (CUDA-graph capturable): one
        ``acs_vecenv_step`` call = fused env-step kernel + auto-reset kernel.

        actions: CUDA tensor (any integer dtype).  Returns live views ``(obs int8 [N, 2*mrl],
        reward int32 [N], done uint8 [N], truncated uint8 [N])`` that the next call overwrites.
        Finished environments are already reset in ``obs``; their last observation is in
        ``self.final_obs``, the episode length in ``self.final_steps`` and -- for solved ones --
        the move sequence via ``final_actions(i)``.  Reward clipping (if configured) is applied by
        ``clipped_reward()`` / ``transformed_reward()``.  Errors (a move emptying a relator) accumulate in
        ``self.err`` and are raised by ``check_errors()``; the outputs of such a row are zero for that step.
        With ``enable_curriculum`` finished environments continue with another state of the pool instead of
        their own initial state.  Any max_relator_length <= 64 and non-normal-form initial states are served
        (by the general kernel variant)."""
        t = self.torch
        act = actions if actions.dtype == t.uint8 else actions.to(t.uint8)
        stream = t.cuda.current_stream(self.dev).cuda_stream
        if self._curriculum is not None:
            a = self._curriculum["args"]
            a.action = act.data_ptr()
            a.flags = self._device_flags
            _lib.check(self.L.acs_vecenv_curriculum_step(__import__("ctypes").byref(a), stream))
            return self.state, self.reward, self.done, self.truncated
        _lib.check(self.L.acs_vecenv_step(
            self.state.data_ptr(), self.initial_states.data_ptr(), act.data_ptr(), self.reward.data_ptr(),
            self.done.data_ptr(), self.truncated.data_ptr(), self.step_count.data_ptr(), self.lens.data_ptr(),
            self.initial_lens.data_ptr(), self.action_log.data_ptr(), self.action_log.shape[1],
            self.final_obs.data_ptr(), self.final_steps.data_ptr(), self.err.data_ptr(), self.num_envs,
            self.max_relator_length, self.horizon_length, self._device_flags, stream))
        return self.state, self.reward, self.done, self.truncated

    def transformed_reward(self):
        """Rewards of the last step as the PPO loop receives them: ``NormalizeReward`` (if configured, per
        environment, running statistics kept on the device) then the clip.  float32 CUDA tensor; sync-free."""
        t = self.torch
        clip = self.clip_rewards is not None
        lo, hi = (self.clip_rewards if clip else (0.0, 0.0))
        _lib.check(self.L.acs_reward_transform(
            self.reward.data_ptr(), self.done.data_ptr(), self.reward_stats.data_ptr(), self.reward_out.data_ptr(),
            self.num_envs, self.gamma, 1e-8, int(self.norm_rewards), int(clip), float(lo), float(hi),
            t.cuda.current_stream(self.dev).cuda_stream))
        return self.reward_out

    # ---- device-side curriculum (training.py:169-224) ------------------------------------------
    def enable_curriculum(self, pool_states, repeat_solved_prob=0.25, seed=0):
        """Let ``step_device`` move finished environments on to other initial states of ``pool_states``
        ([n_states, 2*mrl]; the first num_envs rows are the environments' current initial states) exactly
        as the reference's rollout loop does on the host, without leaving the device."""
        import ctypes as C

        t = self.torch
        pool = np.ascontiguousarray(pool_states, dtype=np.int8)
        if pool.ndim != 2 or pool.shape[1] != 2 * self.max_relator_length or len(pool) < self.num_envs:
            raise ValueError("pool_states must be [n_states >= num_envs, 2*max_relator_length]")
        for row in pool:
            assert is_array_valid_presentation(row), f"{row} is not a valid presentation"
        if not np.array_equal(pool[: self.num_envs], self.initial_states_host):
            raise ValueError("the first num_envs pool states must be the environments' initial states")
        if not _normal_form_mask(pool).all():
            self._device_flags = 0
        ns = len(pool)
        c = {
            "pool": t.from_numpy(pool).to(self.dev), "pool_lens": t.from_numpy(_lens_of(pool)).to(self.dev),
            "cur_state": t.arange(self.num_envs, dtype=t.int32, device=self.dev),
            "solved": t.zeros((ns + 3) // 4 * 4, dtype=t.uint8, device=self.dev),
            "solved_list": t.zeros(ns, dtype=t.int32, device=self.dev),
            "best": t.full((ns,), -1, dtype=t.int64, device=self.dev),
            "best_actions": t.zeros((ns, self.action_log.shape[1]), dtype=t.uint8, device=self.dev),
            "counters": t.tensor([self.num_envs, 0, 0, 0], dtype=t.int64, device=self.dev), "n_states": ns,
        }
        a = _lib.CurriculumArgs()
        a.state, a.pool, a.pool_lens, a.lens = (self.state.data_ptr(), c["pool"].data_ptr(), c["pool_lens"].data_ptr(),
                                                self.lens.data_ptr())
        a.reward, a.done, a.truncated, a.step_count = (self.reward.data_ptr(), self.done.data_ptr(),
                                                       self.truncated.data_ptr(), self.step_count.data_ptr())
        a.cur_state, a.solved, a.solved_list, a.best = (c["cur_state"].data_ptr(), c["solved"].data_ptr(),
                                                        c["solved_list"].data_ptr(), c["best"].data_ptr())
        a.best_actions, a.action_log = c["best_actions"].data_ptr(), self.action_log.data_ptr()
        a.final_obs, a.final_steps = self.final_obs.data_ptr(), self.final_steps.data_ptr()
        a.counters, a.err = c["counters"].data_ptr(), self.err.data_ptr()
        a.n, a.n_states, a.mrl, a.horizon = self.num_envs, ns, self.max_relator_length, self.horizon_length
        a.log_stride, a.flags = self.action_log.shape[1], self._device_flags
        a.repeat_solved_prob, a.seed = float(repeat_solved_prob), int(seed) & (2 ** 64 - 1)
        c["args"] = a
        self._curriculum = c
        del C

    def success_record(self):
        """``success_record`` of the reference's loop: {"solved": set, "unsolved": set} of pool indices."""
        c = self._curriculum
        solved = set(int(i) for i in np.flatnonzero(c["solved"][: c["n_states"]].cpu().numpy()))
        return {"solved": solved, "unsolved": set(range(c["n_states"])) - solved}

    def acmoves_hist(self):
        """``ACMoves_hist``: the shortest action sequence that solved each pool state so far."""
        c = self._curriculum
        best = c["best"].cpu().numpy()
        acts = c["best_actions"].cpu().numpy()
        return {int(s): [int(x) for x in acts[s, : int(best[s] >> 32)]] for s in np.flatnonzero(best >= 0)}

    def curriculum_counters(self):
        nxt, n_solved, draws, episodes = (int(v) for v in self._curriculum["counters"].cpu())
        return {"next_unprocessed": nxt, "n_solved": n_solved, "random_draws": draws, "episodes": episodes}

    def clipped_reward(self):
        r = self.reward.to(self.torch.float32)
        return r if self.clip_rewards is None else r.clamp(self.clip_rewards[0], self.clip_rewards[1])

    def final_actions(self, i):
        """``info["actions"]`` of the episode environment i finished in the last ``step_device``."""
        n = int(self.final_steps[i])
        return [int(a) for a in self.action_log[i, :n].cpu().numpy()]

    def check_errors(self):
        n_bad = int(self.err[0])
        if n_bad:
            raise AssertionError(f"{n_bad} environment steps produced an invalid presentation (first: env "
                                 f"{int(self.err[1])}); the reference raises AssertionError (envs/utils.py:261-263)")

    def close(self):
        pass
