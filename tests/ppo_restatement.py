"""TEST INFRASTRUCTURE: a torch (CPU, fp32) restatement of the PPO update path of the reference
(/root/reference/ac_solver/agents/training.py:118-352 -- rollout bookkeeping, GAE :230-250, losses :262-318,
clip + Adam :320-326, target-KL stop / beta schedule :328-336).  The reference's code is inline in
``ppo_training_loop`` (no callable pieces), so this restatement is pinned END TO END: oracle/gen_golden_ppo.py
runs the real reference loop on a deterministic fake vector environment and stores the trained parameters;
tests/test_ppo_cpu.py runs ``run_updates`` below on the same environment and compares.  The GPU kernels are then
checked against ``gae_ref`` / ``loss_ref``."""

import numpy as np
import torch
from torch.distributions import Categorical


class FakeVecEnv:
    """Deterministic stand-in for the SyncVectorEnv: observations / rewards from a seeded generator that also
    mixes in the actions; never terminates (keeps the run away from the curriculum branch)."""

    class _Space:
        def __init__(self, shape, n=None):
            self.shape, self.n = shape, n

    class _E:
        max_reward = 50.0
        supermoves = None

    def __init__(self, num_envs, width=8, n_actions=12, seed=0):
        self.num_envs, self.width = num_envs, width
        self.single_observation_space = self._Space((width,))
        self.single_action_space = self._Space((), n_actions)
        self.envs = [self._E() for _ in range(num_envs)]
        self.rng = np.random.default_rng(seed)

    def reset(self):
        self.rng = np.random.default_rng(0)
        return self.rng.integers(-2, 3, size=(self.num_envs, self.width)).astype(np.int8), {}

    def step(self, actions):
        a = np.asarray(actions).astype(np.int64)
        obs = ((self.rng.integers(-2, 3, size=(self.num_envs, self.width)) + a[:, None]) % 5 - 2).astype(np.int8)
        reward = -(self.rng.integers(2, 20, size=self.num_envs) + a % 3).astype(np.float64)
        z = np.zeros(self.num_envs, dtype=bool)
        return obs, reward, z, z.copy(), {}


def act_ref(agent, x, action=None):
    """ppo_agent.py:92-109 (Categorical sampling: consumes the torch RNG exactly like the reference)."""
    dist = Categorical(logits=agent.actor(x))
    if action is None:
        action = dist.sample()
    return action, dist.log_prob(action), dist.entropy(), agent.critic(x)


def gae_ref(rewards, values, dones, next_value, next_done, gamma, lam):
    """training.py:240-250 (tensors [T, N]; next_value [1, N])."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = 0
    for t in reversed(range(T)):
        nnt = 1.0 - (next_done if t == T - 1 else dones[t + 1])
        nv = next_value if t == T - 1 else values[t + 1]
        delta = rewards[t] + gamma * nv * nnt - values[t]
        adv[t] = last = delta + gamma * lam * nnt * last
    return adv, adv + values


def loss_ref(logits, newvalue, actions, old_logprob, adv, returns, old_values, a, beta=None):
    """training.py:262-318 for one minibatch; returns (loss, pg_loss, v_loss, entropy, approx_kl, clipfrac)."""
    dist = Categorical(logits=logits)
    newlogprob, entropy = dist.log_prob(actions), dist.entropy()
    logratio = newlogprob - old_logprob
    ratio = logratio.exp()
    kl_var = (ratio - 1) - logratio
    approx_kl = kl_var.mean().detach()
    clipfrac = ((ratio - 1.0).abs() > a.clip_coef).float().mean().detach()
    if a.norm_adv:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    pg1 = -adv * ratio
    if a.is_loss_clip:
        pg_loss = torch.max(pg1, -adv * torch.clamp(ratio, 1 - a.clip_coef, 1 + a.clip_coef)).mean()
    else:
        pg_loss = (pg1 + beta * kl_var).mean()
    nv = newvalue.view(-1)
    if a.clip_vloss:
        vc = old_values + torch.clamp(nv - old_values, -a.clip_coef, a.clip_coef)
        v_loss = 0.5 * torch.max((nv - returns) ** 2, (vc - returns) ** 2).mean()
    else:
        v_loss = 0.5 * ((nv - returns) ** 2).mean()
    ent = entropy.mean()
    return pg_loss - a.ent_coef * ent + v_loss * a.vf_coef, pg_loss, v_loss, ent, approx_kl, clipfrac


def run_updates(envs, a, agent, optimizer, lr_fn):
    """The reference's loop without logging / curriculum (the fake environment never finishes an episode)."""
    import random

    T, N = a.num_steps, a.num_envs
    obs = torch.zeros((T, N) + envs.single_observation_space.shape)
    actions, logprobs, rewards = torch.zeros((T, N)), torch.zeros((T, N)), torch.zeros((T, N))
    dones, values = torch.zeros((T, N)), torch.zeros((T, N))
    next_obs, next_done = torch.Tensor(envs.reset()[0]), torch.zeros(N)
    beta = None if a.is_loss_clip else a.beta
    for update in range(1, a.total_timesteps // a.batch_size + 1):
        random.seed(a.seed + update)
        np.random.seed(a.seed + update)
        torch.manual_seed(a.seed + update)
        if a.anneal_lr:
            optimizer.param_groups[0]["lr"] = lr_fn(update)
        for step in range(T):
            obs[step], dones[step] = next_obs, next_done
            with torch.no_grad():
                act, lp, _, val = act_ref(agent, next_obs)
            values[step], actions[step], logprobs[step] = val.flatten(), act, lp
            o, r, d, _tr, _ = envs.step(act.cpu().numpy())
            rewards[step] = torch.tensor(r).view(-1)
            next_obs, next_done = torch.Tensor(o), torch.Tensor(d)
        if not a.norm_rewards:
            rewards /= envs.envs[0].max_reward
        with torch.no_grad():
            adv, ret = gae_ref(rewards, values, dones, agent.get_value(next_obs).reshape(1, -1), next_done, a.gamma, a.gae_lambda)
        b = [x.reshape((-1,) + x.shape[2:]) for x in (obs, logprobs, actions, adv, ret, values)]
        inds = np.arange(a.batch_size)
        for _ in range(a.update_epochs):
            np.random.shuffle(inds)
            for s in range(0, a.batch_size, a.minibatch_size):
                mb = inds[s : s + a.minibatch_size]
                logits, nv = agent.actor(b[0][mb]), agent.critic(b[0][mb])
                loss, _, _, _, kl, _ = loss_ref(logits, nv, b[2].long()[mb], b[1][mb], b[3][mb], b[4][mb], b[5][mb], a, beta)
                optimizer.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(agent.parameters(), a.max_grad_norm)
                optimizer.step()
            if a.is_loss_clip:
                if a.target_kl is not None and kl > a.target_kl:
                    break
            else:
                beta = beta / 2 if kl < a.target_kl / 1.5 else (beta * 2 if kl > a.target_kl * 1.5 else beta)
    return agent
