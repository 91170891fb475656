#!/usr/bin/env python
"""BASELINE config 4: PPO rollout, environment side.  4096 GPU-resident environments, horizon
200, a torch MLP policy (the reference's 2x512 actor, ppo_agent.py:39-49) choosing actions on the
device.  Three ways of driving the same environment are timed:
  host-api   ACVectorEnv.step with CUDA tensors (gymnasium-style infos, one host sync per step)
  device     ACVectorEnv.step_device (no host sync)
  graph      policy forward + sampling + step_device captured in ONE CUDA graph and replayed
Reports environment steps per second for each."""
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
    ap.add_argument("--steps", type=int, default=1000)
    args = ap.parse_args()
    ms = np.load(os.path.join(ROOT, "tests", "golden", "miller_schupp.npz"))
    init = ms["presentations36"][np.arange(args.envs) % 1190].astype(np.int8)
    torch.manual_seed(0)
    actor = torch.nn.Sequential(torch.nn.Linear(72, 512), torch.nn.Tanh(), torch.nn.Linear(512, 512), torch.nn.Tanh(),
                                torch.nn.Linear(512, 12)).cuda()
    out = {"config": f"{args.envs} envs, horizon {args.horizon}, torch policy on device"}

    def policy(obs):
        with torch.no_grad():
            logits = actor(obs.float())
            # Gumbel-max sampling == Categorical(logits).sample(), graph-friendly
            g = -torch.log(-torch.log(torch.rand_like(logits).clamp_(1e-10, 1.0)))
            return (logits + g).argmax(dim=1)

    def timed(fn, n):
        fn(20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(n)
        torch.cuda.synchronize()
        return args.envs * n / (time.perf_counter() - t0)

    env = ACVectorEnv(init, horizon_length=args.horizon, clip_rewards=(-10, 1000))
    state = {"obs": torch.from_numpy(env.reset()[0]).cuda()}

    def run_host_api(n):
        for _ in range(n):
            state["obs"], reward, done, trunc, infos = env.step(policy(state["obs"]))

    out["host_api_env_steps_per_s"] = timed(run_host_api, min(args.steps, 300))

    env = ACVectorEnv(init, horizon_length=args.horizon, clip_rewards=(-10, 1000))
    env.reset()

    def run_device(n):
        for _ in range(n):
            env.step_device(policy(env.state))

    out["device_env_steps_per_s"] = timed(run_device, args.steps)
    env.check_errors()

    env = ACVectorEnv(init, horizon_length=args.horizon, clip_rewards=(-10, 1000))
    env.reset()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            env.step_device(policy(env.state))
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        env.step_device(policy(env.state))
        rew = env.clipped_reward()

    def run_graph(n):
        for _ in range(n):
            graph.replay()

    out["graph_env_steps_per_s"] = timed(run_graph, args.steps)
    env.check_errors()
    out["graph_us_per_vector_step"] = 1e6 * args.envs / out["graph_env_steps_per_s"]
    out["episodes_finished_flag_sum"] = int((env.done | env.truncated).sum())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
