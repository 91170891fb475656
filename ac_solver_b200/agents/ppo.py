"""PPO training on the GPU-resident AC environment -- the command-line entry point.

    python -m ac_solver_b200.agents.ppo --num-envs 4096 --num-steps 200 --nodes-counts 512 512

Flags are the reference's (agents/args.py mirrors ``ac_solver/agents/args.py``).  Differences from the
reference's ``ac_solver/agents/ppo.py``: a CUDA device is mandatory (the environments live on it), the
optimiser is a capturable Adam whose learning rate is a device scalar (the minibatch step is replayed
as a CUDA graph), and the trained bookkeeping is returned to the caller."""

from __future__ import annotations

import random

import numpy as np
import torch

from . import args as _args
from . import environment, ppo_agent, training


def seed_everything(seed, deterministic=True):
    for fn in (random.seed, np.random.seed, torch.manual_seed):
        fn(seed)
    torch.backends.cudnn.deterministic = bool(deterministic)


def make_optimizer(agent, learning_rate, eps, device):
    lr = torch.tensor(float(learning_rate), device=device)  # device-resident: annealed with fill_() between graph replays
    return torch.optim.Adam(agent.parameters(), lr=lr, eps=eps, capturable=True)


def train_ppo(argv=None):
    """Parse the flags, build environments / agent / optimiser and run ``ppo_training_loop``.
    Returns ``(last_log, success_record, ACMoves_hist)``."""
    cfg = _args.parse_args(argv)
    seed_everything(cfg.seed, cfg.torch_deterministic)
    if not (cfg.cuda and torch.cuda.is_available()):
        raise RuntimeError("ac_solver_b200 trains on a CUDA device only: the environment lives on the GPU")
    device = torch.device("cuda")
    envs, initial_states, curr_states, success_record, moves_hist, processed = environment.get_env(cfg)
    try:
        agent = ppo_agent.Agent(envs, cfg.nodes_counts).to(device)
        optimizer = make_optimizer(agent, cfg.learning_rate, cfg.epsilon, device)
        log = training.ppo_training_loop(envs, cfg, device, optimizer, agent, curr_states, success_record, moves_hist,
                                         processed, initial_states)
    finally:
        envs.close()
    return log, success_record, moves_hist


if __name__ == "__main__":
    train_ppo()
