"""Environment construction for PPO (reference: ``ac_solver/agents/environment.py``).

``get_env(args)`` returns the same six-tuple as the reference, but ``envs`` is ONE GPU-resident
``ACVectorEnv`` instead of a ``SyncVectorEnv`` of wrapped ``ACEnv`` objects: the ``NormalizeReward`` /
``TransformReward(clip)`` wrappers of ``make_env`` (environment.py:44-52) and the curriculum of the training
loop (training.py:169-224) run on the device inside it."""

from __future__ import annotations

import numpy as np

from ..envs.utils import change_max_relator_length_of_presentation, convert_relators_to_presentation
from ..envs.vector_env import ACVectorEnv
from .utils import load_initial_states_from_text_file


def get_env(args, device=None):
    if args.clip_rewards:
        assert args.min_rew < args.max_rew, "min_rew must be less than max_rew"
    clip = (args.min_rew, args.max_rew) if args.clip_rewards else None
    if args.fixed_init_state:
        presentation = convert_relators_to_presentation(args.relator1, args.relator2, args.max_relator_length)
        initial_states = [np.asarray(presentation, dtype=np.int8)]
        rows = np.stack([initial_states[0]] * args.num_envs)
    else:
        initial_states = load_initial_states_from_text_file(states_type=args.states_type)
        assert args.num_envs <= len(initial_states), \
            "Expect number of environments to be less than number of distinct initial states for now"
        args.max_relator_length = 36  # max(4n+2) for 1 <= n <= 7 (environment.py:85-87)
        initial_states = [np.asarray(change_max_relator_length_of_presentation(s, args.max_relator_length), dtype=np.int8)
                          for s in initial_states]
        rows = np.stack(initial_states[: args.num_envs])
    envs = ACVectorEnv(rows, horizon_length=args.horizon_length, device=device, clip_rewards=clip,
                       use_supermoves=args.use_supermoves, norm_rewards=args.norm_rewards, gamma=args.gamma)
    # (the reference leaves these four undefined for fixed_init_state=True and fails at return; here every
    # environment then restarts from the one fixed state)
    curr_states = list(range(args.num_envs)) if not args.fixed_init_state else [0] * args.num_envs
    states_processed = set(curr_states)
    success_record = {"solved": set(), "unsolved": set(range(len(initial_states)))}
    ACMoves_hist = {}
    if not args.fixed_init_state:
        envs.enable_curriculum(np.stack(initial_states), repeat_solved_prob=args.repeat_solved_prob, seed=args.seed)
    return envs, initial_states, curr_states, success_record, ACMoves_hist, states_processed
