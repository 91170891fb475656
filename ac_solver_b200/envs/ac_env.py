"""Andrews-Curtis environment -- drop-in for the reference's ``ac_solver/envs/ac_env.py``.

``ACEnv`` keeps the reference's single-environment API (one GPU call per step); the batched,
GPU-resident equivalent used for rollouts is ``ac_solver_b200.envs.vector_env.ACVectorEnv``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Union

import numpy as np

from .ac_moves import ACMove
from .spaces import Box, Discrete, Env
from .utils import is_array_valid_presentation


@dataclass
class ACEnvConfig:
    """envs/ac_env.py:14-53"""

    initial_state: Union[np.ndarray, list] = field(default_factory=lambda: np.array([1, 0, 2, 0]))
    horizon_length: any = 1000
    use_supermoves: any = False

    def __post_init__(self):
        if isinstance(self.initial_state, list):
            self.initial_state = np.array(self.initial_state)
        if not isinstance(self.initial_state, np.ndarray):
            raise TypeError("initial_state must be a numpy array")
        if self.initial_state.ndim != 1:
            raise ValueError("initial_state must be a 1-dimensional array")
        if len(self.initial_state) % 2 != 0:
            raise ValueError("initial state must have even length")
        if not is_array_valid_presentation(self.initial_state):
            raise ValueError("initial state must be a valid presentation")

    @property
    def max_relator_length(self):
        return len(self.initial_state) // 2

    @classmethod
    def from_dict(cls, config_dict):
        return cls(
            initial_state=np.array(config_dict.get("initial_state", cls().initial_state)),
            horizon_length=config_dict.get("horizon_length", cls().horizon_length),
            use_supermoves=config_dict.get("use_supermoves", cls().use_supermoves),
        )


def _lengths(state, mrl):
    return [int(np.count_nonzero(state[k * mrl : (k + 1) * mrl])) for k in range(2)]


class ACEnv(Env):
    """envs/ac_env.py:56-134"""

    def __init__(self, config: ACEnvConfig = ACEnvConfig()):
        self.n_gen = 2
        self.max_relator_length = config.max_relator_length
        self.initial_state = config.initial_state
        self.horizon_length = config.horizon_length
        if config.use_supermoves:
            raise NotImplementedError("ACEnv with supermoves is not yet implemented in this library.")
        bound = np.full(self.max_relator_length * self.n_gen, self.n_gen, dtype=np.int8)
        self.observation_space = Box(-bound, bound, dtype=np.int8)
        self.action_space = Discrete(12)
        self.max_reward = self.horizon_length * self.max_relator_length * self.n_gen  # :80
        self.state = np.copy(self.initial_state)
        self.count_steps = 0
        self.lengths = _lengths(self.state, self.max_relator_length)
        self.actions = []

    def step(self, action):
        """envs/ac_env.py:95-113"""
        self.actions += [action]
        self.state, self.lengths = ACMove(action, self.state, self.max_relator_length, self.lengths)
        done = sum(self.lengths) == 2
        reward = self.max_reward * done - sum(self.lengths) * (1 - done)
        self.count_steps += 1
        truncated = self.count_steps >= self.horizon_length
        return self.state, reward, done, truncated, ({"actions": self.actions.copy()} if done else {})

    def reset(self, *, seed=None, options=None):
        """envs/ac_env.py:115-131"""
        src = options["starting_state"] if options and "starting_state" in options else self.initial_state
        self.state = np.copy(src)
        self.lengths = _lengths(self.state, self.max_relator_length)
        self.count_steps = 0
        self.actions = []
        return self.state, {}

    def render(self):
        pass
