"""Entry point of PPO training on the GPU-resident AC environment (reference: ``ac_solver/agents/ppo.py``):

    python -m ac_solver_b200.agents.ppo [--num-envs 4096 --num-steps 200 ...]      # flags: agents/args.py
"""

from __future__ import annotations

import random

import numpy as np
import torch
from torch.optim import Adam

from .args import parse_args
from .environment import get_env
from .ppo_agent import Agent
from .training import ppo_training_loop


def train_ppo(argv=None):
    args = parse_args(argv)
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.backends.cudnn.deterministic = args.torch_deterministic
    if not (torch.cuda.is_available() and args.cuda):
        raise RuntimeError("ac_solver_b200 trains on a CUDA device only: the environment lives on the GPU")
    device = torch.device("cuda")
    envs, initial_states, curr_states, success_record, ACMoves_hist, states_processed = get_env(args)
    agent = Agent(envs, args.nodes_counts).to(device)
    # capturable Adam with a device-resident learning rate: the whole minibatch step is one CUDA graph
    optimizer = Adam(agent.parameters(), lr=torch.tensor(args.learning_rate, device=device), eps=args.epsilon,
                     capturable=True)
    log = ppo_training_loop(envs, args, device, optimizer, agent, curr_states, success_record, ACMoves_hist,
                            states_processed, initial_states)
    envs.close()
    return log, success_record, ACMoves_hist


if __name__ == "__main__":
    train_ppo()
