#!/usr/bin/env python
"""BASELINE config 4: PPO rollout, environment side.  4096 GPU-resident environments, horizon
200, a torch MLP policy (the reference's 2x512 actor, ppo_agent.py:39-49) choosing actions on the
device; no host round trip in the step path.  Reports environment steps per second."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ac_solver_b200.envs.vector_env import ACVectorEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=200)
    ap.add_argument("--steps", type=int, default=400)
    args = ap.parse_args()
    ms = np.load(os.path.join(ROOT, "tests", "golden", "miller_schupp.npz"))
    init = ms["presentations36"][np.arange(args.envs) % 1190].astype(np.int8)
    env = ACVectorEnv(init, horizon_length=args.horizon, clip_rewards=(-10, 1000))
    torch.manual_seed(0)
    actor = torch.nn.Sequential(torch.nn.Linear(72, 512), torch.nn.Tanh(), torch.nn.Linear(512, 512), torch.nn.Tanh(),
                                torch.nn.Linear(512, 12)).cuda()
    obs = torch.from_numpy(env.reset()[0]).cuda()

    def rollout(n):
        nonlocal obs
        for _ in range(n):
            with torch.no_grad():
                logits = actor(obs.float())
                action = torch.distributions.Categorical(logits=logits).sample()
            obs, reward, done, trunc, infos = env.step(action)

    rollout(20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rollout(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"config": f"{args.envs} envs, horizon {args.horizon}, torch policy on device",
                      "env_steps_per_s": args.envs * args.steps / dt, "ms_per_vector_step": 1e3 * dt / args.steps}))


if __name__ == "__main__":
    main()
