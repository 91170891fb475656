"""Actor and critic networks of the PPO agent (reference: ``ac_solver/agents/ppo_agent.py``).

Same architecture, initialisation and public methods as the reference's ``Agent`` -- two independent
tanh MLPs, orthogonal weights (gain sqrt(2), last layer 0.01 for the actor and 1.0 for the critic), zero
biases -- so a reference checkpoint's ``actor`` / ``critic`` state dicts load unchanged.  Additions for the
device-resident loop: ``forward`` returns logits and values in one call (what the fused loss kernel
consumes), and ``sample`` draws actions with the Gumbel-max trick, which is ``Categorical.sample`` in
distribution but needs no host synchronisation and can be captured in a CUDA graph."""

from __future__ import annotations

import numpy as np
import torch
from torch import nn


def initialize_layer(layer, std=np.sqrt(2), bias_const=0.0):
    nn.init.orthogonal_(layer.weight, std)
    nn.init.constant_(layer.bias, bias_const)
    return layer


def build_network(nodes_counts, std=0.01):
    """[in, h1, ..., out] -> list of Linear / Tanh modules; the output layer is initialised with gain ``std``."""
    n = len(nodes_counts) - 1
    layers = []
    for k in range(n):
        last = k == n - 1
        lin = nn.Linear(nodes_counts[k], nodes_counts[k + 1])
        layers.append(initialize_layer(lin, std=std) if last else initialize_layer(lin))
        if not last:
            layers.append(nn.Tanh())
    return layers


class Agent(nn.Module):
    def __init__(self, envs, nodes_counts):
        super().__init__()
        input_dim = int(np.prod(envs.single_observation_space.shape))
        self.critic_nodes = [input_dim] + list(nodes_counts) + [1]
        self.actor_nodes = [input_dim] + list(nodes_counts) + [envs.single_action_space.n]
        self.critic = nn.Sequential(*build_network(self.critic_nodes, 1.0))
        self.actor = nn.Sequential(*build_network(self.actor_nodes, 0.01))

    # ---- reference API ---------------------------------------------------------------------
    def get_value(self, x):
        return self.critic(x)

    def get_action_and_value(self, x, action=None):
        """(action, log-probability, entropy, value) as in the reference (ppo_agent.py:92-109)."""
        logits = self.actor(x)
        value = self.critic(x)
        logp = torch.log_softmax(logits, dim=-1)
        if action is None:
            action = self._gumbel_argmax(logits)
        lp = logp.gather(-1, action.long().unsqueeze(-1)).squeeze(-1)
        entropy = -(logp.exp() * logp).sum(-1)
        return action, lp, entropy, value

    # ---- device-resident loop --------------------------------------------------------------
    def forward(self, x):
        """logits [B, n_actions], values [B]."""
        return self.actor(x), self.critic(x).squeeze(-1)

    @staticmethod
    def _gumbel_argmax(logits):
        u = torch.rand_like(logits).clamp_(1e-10, 1.0)
        return (logits - torch.log(-torch.log(u))).argmax(dim=-1)

    def sample(self, x):
        """action int64 [B], log-probability [B], value [B]; no host synchronisation."""
        logits, value = self.forward(x)
        action = self._gumbel_argmax(logits)
        lp = torch.log_softmax(logits, dim=-1).gather(-1, action.unsqueeze(-1)).squeeze(-1)
        return action, lp, value
