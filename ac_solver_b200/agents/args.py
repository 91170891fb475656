"""Command-line arguments of the PPO trainer -- same flags, defaults and derived fields as the
reference's ``ac_solver/agents/args.py`` (``parse_args``), written as one table."""

from __future__ import annotations

import argparse


def _flag(x):
    if isinstance(x, bool):
        return x
    v = str(x).strip().lower()
    if v in ("y", "yes", "t", "true", "on", "1"):
        return True
    if v in ("n", "no", "f", "false", "off", "0"):
        return False
    raise argparse.ArgumentTypeError(f"invalid truth value {x!r}")  # distutils.util.strtobool raises ValueError here


# (flag, type, default); type "flag" = optional boolean value (``--x``, ``--x true``, ``--x false``)
_TABLE = [
    ("--exp-name", str, "args"),  # basename(args.py).rstrip(".py") in the reference
    ("--seed", int, 1),
    ("--torch-deterministic", "flag", True),
    ("--cuda", "flag", True),
    ("--wandb-log", "flag", False),
    ("--wandb-project-name", str, "AC-Solver-PPO"),
    ("--wandb-entity", str, None),
    # environment
    ("--fixed-init-state", "flag", False),
    ("--states-type", str, "all"),
    ("--repeat-solved-prob", float, 0.25),
    ("--max-relator-length", int, 7),
    ("--relator1", [int], [1, 1, -2, -2, -2]),
    ("--relator2", [int], [1, 2, 1, -2, -1, -2]),
    ("--horizon-length", int, 2000),
    ("--use_supermoves", "flag", False),
    # architecture
    ("--nodes-counts", [int], [256, 256]),
    # algorithm
    ("--is-loss-clip", "flag", True),
    ("--beta", float, 0.9),
    ("--total-timesteps", int, 200000),
    ("--learning-rate", float, 2.5e-4),
    ("--warmup-period", float, 0.0),
    ("--lr-decay", str, "linear"),
    ("--min-lr-frac", float, 0.0),
    ("--num-envs", int, 4),
    ("--num-steps", int, 2000),
    ("--anneal-lr", "flag", True),
    ("--gamma", float, 0.99),
    ("--gae-lambda", float, 0.95),
    ("--num-minibatches", int, 4),
    ("--update-epochs", int, 1),
    ("--norm-adv", "flag", True),
    ("--norm-rewards", "flag", False),
    ("--clip-rewards", "flag", True),
    ("--min-rew", int, -10),
    ("--max-rew", int, 1000),
    ("--clip-coef", float, 0.2),
    ("--clip-vloss", "flag", True),
    ("--ent-coef", float, 0.01),
    ("--vf-coef", float, 0.5),
    ("--max-grad-norm", float, 0.5),
    ("--target-kl", float, 0.01),
    ("--epsilon", float, 0.00001),
]


def build_parser():
    p = argparse.ArgumentParser(description="PPO on the AC environment (GPU-resident rollout)")
    for name, typ, default in _TABLE:
        if typ == "flag":
            p.add_argument(name, type=_flag, default=default, nargs="?", const=True)
        elif isinstance(typ, list):
            p.add_argument(name, type=typ[0], default=default, nargs="+")
        else:
            p.add_argument(name, type=typ, default=default)
    return p


def finalize(args):
    """Derived fields and the reference's sanity checks (args.py:283-296)."""
    args.batch_size = int(args.num_envs * args.num_steps)
    args.minibatch_size = int(args.batch_size // args.num_minibatches)
    assert 0.0 <= args.warmup_period <= 1.0, \
        "warmup period should be less than 1.0 as it is the fraction of total timesteps"
    assert args.lr_decay in ["linear", "cosine"], \
        f"lr-decay must be linear or cosine, not {args.lr_decay}. Other LR schedules not supported yet"
    assert 0.0 <= args.min_lr_frac <= 1.0, "min-lr-frac is the fraction of maximum lr to which we anneal."
    return args


def parse_args(argv=None):
    return finalize(build_parser().parse_args(argv))
