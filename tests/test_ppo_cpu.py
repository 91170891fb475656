"""CPU tests of the PPO package: the learning-rate schedule against values produced by the
reference's own get_curr_lr, the argument table against the reference's defaults, the agent's architecture,
and -- pinning the test restatement of the update path (tests/ppo_restatement.py) -- parameters trained by the
REAL reference loop on a fake environment (oracle/gen_golden_ppo.py) reproduced by the restatement."""

import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "ppo.json")) as f:
        return json.load(f)


def test_lr_schedule_equals_reference(golden):
    from ac_solver_b200.agents.training import get_curr_lr

    for case in golden["lr"]:
        mine = [get_curr_lr(n, case["decay"], case["warmup"], 2.5e-4, 2.5e-4 * case["min_frac"], case["total"])
                for n in range(1, case["total"] + 1)]
        assert mine == case["values"], case
    with pytest.raises(NotImplementedError):
        get_curr_lr(5, "exponential", 0.0, 1e-3, 0.0, 10)


def test_args_defaults_and_derived_fields():
    from ac_solver_b200.agents.args import parse_args

    a = parse_args([])
    assert (a.num_envs, a.num_steps, a.batch_size, a.minibatch_size) == (4, 2000, 8000, 2000)
    assert (a.gamma, a.gae_lambda, a.clip_coef, a.ent_coef, a.vf_coef, a.max_grad_norm, a.target_kl) == \
        (0.99, 0.95, 0.2, 0.01, 0.5, 0.5, 0.01)
    assert a.nodes_counts == [256, 256] and a.relator1 == [1, 1, -2, -2, -2] and a.horizon_length == 2000
    assert a.is_loss_clip and a.clip_vloss and a.norm_adv and not a.norm_rewards and a.clip_rewards
    b = parse_args(["--num-envs", "8", "--num-steps", "10", "--num-minibatches", "2", "--norm-rewards", "--clip-vloss", "false",
                    "--nodes-counts", "512", "512"])
    assert (b.batch_size, b.minibatch_size, b.norm_rewards, b.clip_vloss, b.nodes_counts) == (80, 40, True, False, [512, 512])
    with pytest.raises(AssertionError):
        parse_args(["--lr-decay", "step"])


def test_agent_architecture_matches_the_reference_layout():
    from ac_solver_b200.agents.ppo_agent import Agent
    from ppo_restatement import FakeVecEnv

    torch.manual_seed(0)
    ag = Agent(FakeVecEnv(4, width=72), [512, 512])
    assert [tuple(p.shape) for p in ag.critic.parameters()] == [(512, 72), (512,), (512, 512), (512,), (1, 512), (1,)]
    assert [tuple(p.shape) for p in ag.actor.parameters()] == [(512, 72), (512,), (512, 512), (512,), (12, 512), (12,)]
    assert list(ag.state_dict())[:2] == ["critic.0.weight", "critic.0.bias"]  # reference checkpoints load unchanged
    x = torch.randn(5, 72)
    a, lp, ent, v = ag.get_action_and_value(x)
    assert a.shape == (5,) and lp.shape == (5,) and ent.shape == (5,) and v.shape == (5, 1)
    d = torch.distributions.Categorical(logits=ag.actor(x))
    assert torch.allclose(lp, d.log_prob(a)) and torch.allclose(ent, d.entropy(), atol=1e-6)
    w = ag.actor[-1].weight
    assert abs(float((w @ w.T)[0, 0].detach()) - 1e-4) < 1e-6  # orthogonal rows with gain 0.01


@pytest.mark.parametrize("name", ["clip", "clip_noclipv_nonorm", "klpen"])
def test_restatement_reproduces_the_reference_training_run(golden, name):
    """tests/ppo_restatement.py == the reference's ppo_training_loop (run for real by oracle/gen_golden_ppo.py)."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    from ac_solver_b200.agents.ppo_agent import Agent
    from ac_solver_b200.agents.training import get_curr_lr
    from ppo_restatement import FakeVecEnv, run_updates
    import argparse

    over = golden["train"][name]["overrides"]
    a = argparse.Namespace(
        seed=3, num_envs=4, num_steps=16, total_timesteps=4 * 16 * 3, num_minibatches=4, update_epochs=2, nodes_counts=[16, 16],
        anneal_lr=True, lr_decay="linear", warmup_period=0.0, learning_rate=1e-3, min_lr_frac=0.0, gamma=0.99, gae_lambda=0.95,
        norm_adv=True, norm_rewards=False, clip_coef=0.2, clip_vloss=True, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5,
        target_kl=None, is_loss_clip=True, beta=0.9, epsilon=1e-5)
    for k, v in over.items():
        setattr(a, k, v)
    a.batch_size = a.num_envs * a.num_steps
    a.minibatch_size = a.batch_size // a.num_minibatches
    torch.manual_seed(a.seed)
    envs = FakeVecEnv(a.num_envs)
    agent = Agent(envs, a.nodes_counts)
    opt = torch.optim.Adam(agent.parameters(), lr=a.learning_rate, eps=a.epsilon)
    total = a.total_timesteps // a.batch_size
    run_updates(envs, a, agent, opt, lambda u: get_curr_lr(u, a.lr_decay, a.warmup_period, a.learning_rate,
                                                           a.learning_rate * a.min_lr_frac, total))
    for k, v in agent.state_dict().items():
        exp = golden["train"][name]["params"][k]
        got = [float(v.double().sum()), float(v.double().abs().sum()), float(v.flatten()[0])]
        np.testing.assert_allclose(got, exp, rtol=2e-5, atol=2e-6, err_msg=f"{name}: {k}")


def test_no_cpu_fallback_in_the_training_path():
    """The device-resident loop and its kernels refuse to run without CUDA instead of falling back."""
    import argparse

    from ac_solver_b200 import _lib
    from ac_solver_b200.agents.training import gae, ppo_training_loop

    with pytest.raises(_lib.AcsError):
        ppo_training_loop(None, argparse.Namespace(), "cpu", None, None, [], {}, {}, set(), [])
    x = torch.zeros(4, 3)
    with pytest.raises(AssertionError):
        gae(x, x, x, torch.zeros(3), torch.zeros(3), 0.99, 0.95)
    if not torch.cuda.is_available():
        from ac_solver_b200.agents.ppo import train_ppo

        with pytest.raises(RuntimeError):
            train_ppo(["--num-envs", "2", "--num-steps", "4"])
