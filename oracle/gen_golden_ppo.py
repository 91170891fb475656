#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates tests/golden/ppo.json from the REAL reference
(/root/reference/ac_solver/agents/training.py, ppo_agent.py), imported here with stubs for the packages this
container lacks (gymnasium, wandb): (i) get_curr_lr on a grid of schedules; (ii) the reference's
ppo_training_loop run on CPU on the deterministic fake environment of tests/ppo_restatement.py for a few
updates under several loss configurations -> sha-free summary of the trained parameters (per-tensor sums and
first values) that pins the torch restatement of the update path used by the GPU tests."""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
for name in ("gymnasium", "gymnasium.spaces", "wandb"):
    sys.modules.setdefault(name, types.ModuleType(name))
gym = sys.modules["gymnasium"]
gym.Env = object
gym.spaces = sys.modules["gymnasium.spaces"]
gym.spaces.Discrete = lambda n: None
gym.spaces.Box = lambda *a, **k: None
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402
from ac_solver.agents.ppo_agent import Agent  # noqa: E402
from ac_solver.agents.training import get_curr_lr, ppo_training_loop  # noqa: E402

from ppo_restatement import FakeVecEnv  # noqa: E402

CONFIGS = {
    "clip": dict(is_loss_clip=True, clip_vloss=True, norm_adv=True, update_epochs=2, target_kl=None, lr_decay="linear"),
    "clip_noclipv_nonorm": dict(is_loss_clip=True, clip_vloss=False, norm_adv=False, update_epochs=1, target_kl=0.01,
                                lr_decay="cosine"),
    "klpen": dict(is_loss_clip=False, clip_vloss=True, norm_adv=True, update_epochs=2, target_kl=0.01, lr_decay="linear"),
}


def make_args(**over):
    a = argparse.Namespace(
        exp_name="golden", seed=3, num_envs=4, num_steps=16, total_timesteps=4 * 16 * 3, num_minibatches=4, update_epochs=2,
        nodes_counts=[16, 16], anneal_lr=True, lr_decay="linear", warmup_period=0.0, learning_rate=1e-3, min_lr_frac=0.0,
        gamma=0.99, gae_lambda=0.95, norm_adv=True, norm_rewards=False, clip_coef=0.2, clip_vloss=True, ent_coef=0.01,
        vf_coef=0.5, max_grad_norm=0.5, target_kl=None, is_loss_clip=True, beta=0.9, wandb_log=False, horizon_length=100,
        epsilon=1e-5)
    for k, v in over.items():
        setattr(a, k, v)
    a.batch_size = a.num_envs * a.num_steps
    a.minibatch_size = a.batch_size // a.num_minibatches
    return a


def summary(agent):
    return {k: [float(v.double().sum()), float(v.double().abs().sum()), float(v.flatten()[0])] for k, v in agent.state_dict().items()}


def main():
    g = {"lr": []}
    for decay in ("linear", "cosine"):
        for warm in (0.0, 0.1, 0.5):
            for total in (2, 10, 97):
                for frac in (0.0, 0.1):
                    vals = [get_curr_lr(n, decay, warm, 2.5e-4, 2.5e-4 * frac, total) for n in range(1, total + 1)]
                    g["lr"].append({"decay": decay, "warmup": warm, "total": total, "min_frac": frac, "values": vals})
    g["train"] = {}
    cwd = os.getcwd()
    os.chdir("/tmp")  # the reference loop creates out/<run_name>
    for name, over in CONFIGS.items():
        a = make_args(**over)
        torch.manual_seed(a.seed)
        envs = FakeVecEnv(a.num_envs)
        agent = Agent(envs, a.nodes_counts)
        opt = torch.optim.Adam(agent.parameters(), lr=a.learning_rate, eps=a.epsilon)
        ppo_training_loop(envs, a, torch.device("cpu"), opt, agent, list(range(a.num_envs)), {"solved": set(), "unsolved": set()},
                          {}, set(), [])
        g["train"][name] = {"overrides": over, "params": summary(agent)}
    os.chdir(cwd)
    with open(os.path.join(ROOT, "tests", "golden", "ppo.json"), "w") as f:
        json.dump(g, f)
    print("wrote tests/golden/ppo.json", {k: len(v) for k, v in g.items()})


if __name__ == "__main__":
    main()
