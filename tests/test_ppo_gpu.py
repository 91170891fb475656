"""GPU tests of the PPO update path: acs_gae and acs_ppo_loss (csrc/ppo_kernels.cu) against the
torch restatement of the reference's update (tests/ppo_restatement.py, itself pinned against the real reference
loop by tests/test_ppo_cpu.py), and the device-resident training loop end to end."""

import argparse

import numpy as np
import pytest
import torch

from ppo_restatement import gae_ref, loss_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,N", [(1, 1), (16, 4), (200, 4096), (2000, 7)])
def test_gae_bit_identical(T, N):
    """Same fp32 operation order as training.py:240-250 -> equality, not closeness."""
    from ac_solver_b200.agents.training import gae

    g = torch.Generator().manual_seed(T * 131 + N)
    rewards = torch.randn(T, N, generator=g) * 3
    values = torch.randn(T, N, generator=g)
    dones = (torch.rand(T, N, generator=g) < 0.05).float()
    next_value, next_done = torch.randn(1, N, generator=g), (torch.rand(N, generator=g) < 0.2).float()
    adv_ref, ret_ref = gae_ref(rewards, values, dones, next_value, next_done, 0.99, 0.95)
    adv, ret = gae(rewards.cuda(), values.cuda(), dones.cuda(), next_value.reshape(-1).cuda(), next_done.cuda(), 0.99, 0.95)
    assert torch.equal(adv.cpu(), adv_ref) and torch.equal(ret.cpu(), ret_ref)


@pytest.mark.parametrize("is_loss_clip,clip_vloss,norm_adv", [(True, True, True), (True, False, False), (False, True, True),
                                                              (False, False, False)])
@pytest.mark.parametrize("B", [1, 37, 2048])
def test_fused_loss_and_gradient(B, is_loss_clip, clip_vloss, norm_adv):
    """Loss terms and d loss / d(logits, values) of the fused kernel vs torch autograd on the reference's formulas.
    Tolerance: 2e-5 relative + 1e-6 absolute on fp32 (fast-math exp/log in the kernel, different reduction order)."""
    from ac_solver_b200.agents.training import PPOLoss

    if B == 1 and norm_adv:
        pytest.skip("std of one sample is NaN in the reference as well")
    a = argparse.Namespace(is_loss_clip=is_loss_clip, clip_vloss=clip_vloss, norm_adv=norm_adv, clip_coef=0.2, ent_coef=0.01,
                           vf_coef=0.5)
    g = torch.Generator().manual_seed(B)
    logits = (torch.randn(B, 12, generator=g) * 2).requires_grad_()
    newvalue = torch.randn(B, generator=g).requires_grad_()
    actions = torch.randint(0, 12, (B,), generator=g)
    old_lp = torch.log_softmax(logits.detach() + 0.3 * torch.randn(B, 12, generator=g), -1).gather(1, actions[:, None]).squeeze(1)
    adv, returns = torch.randn(B, generator=g) * 2, torch.randn(B, generator=g)
    old_values = newvalue.detach() + 0.3 * torch.randn(B, generator=g)
    beta = 0.9
    ref = loss_ref(logits, newvalue, actions, old_lp, adv, returns, old_values, a, beta)
    ref[0].backward()
    lg, nv = logits.detach().cuda().requires_grad_(), newvalue.detach().cuda().requires_grad_()
    cfg = {k: getattr(a, k) for k in ("norm_adv", "is_loss_clip", "clip_vloss", "clip_coef", "ent_coef", "vf_coef")}
    loss, stats = PPOLoss.apply(lg, nv, actions.cuda(), old_lp.cuda(), adv.cuda(), returns.cuda(), old_values.cuda(),
                                None if is_loss_clip else torch.tensor([beta], device="cuda"), cfg)
    (loss * 1.0).backward()
    tol = dict(rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(stats[:6].cpu().numpy(), [float(x) for x in ref], **tol)
    np.testing.assert_allclose(float(loss), float(ref[0]), **tol)
    np.testing.assert_allclose(lg.grad.cpu().numpy(), logits.grad.numpy(), rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(nv.grad.cpu().numpy(), newvalue.grad.numpy(), rtol=2e-5, atol=1e-7)


def _tiny_args(**over):
    from ac_solver_b200.agents.args import parse_args

    a = parse_args(["--num-envs", "64", "--num-steps", "32", "--horizon-length", "24", "--total-timesteps", str(64 * 32 * 6),
                    "--nodes-counts", "64", "64", "--num-minibatches", "4", "--update-epochs", "2", "--learning-rate", "1e-3"])
    for k, v in over.items():
        setattr(a, k, v)
    return a


@pytest.mark.parametrize("use_graphs", [True, False])
def test_training_loop_runs_on_the_device(use_graphs, tmp_path, monkeypatch):
    """Six updates of the device-resident loop (curriculum, reward clip, graphs on and off): finite losses, the
    parameters move, episodes finish and the bookkeeping containers of the reference's signature are filled."""
    from ac_solver_b200.agents.environment import get_env
    from ac_solver_b200.agents.ppo_agent import Agent
    from ac_solver_b200.agents.training import ppo_training_loop

    monkeypatch.chdir(tmp_path)
    a = _tiny_args()
    torch.manual_seed(0)
    envs, initial_states, curr, rec, hist, processed = get_env(a)
    assert len(initial_states) == 1190 and a.max_relator_length == 36
    dev = torch.device("cuda")
    agent = Agent(envs, a.nodes_counts).to(dev)
    before = [p.detach().clone() for p in agent.parameters()]
    opt = torch.optim.Adam(agent.parameters(), lr=torch.tensor(a.learning_rate, device=dev), eps=a.epsilon, capturable=True)
    log = ppo_training_loop(envs, a, dev, opt, agent, curr, rec, hist, processed, initial_states, use_graphs=use_graphs,
                            checkpoint_every=3, progress=False)
    assert all(np.isfinite(log[k]) for k in ("losses/value_loss", "losses/policy_loss", "losses/entropy_loss", "losses/approx_kl"))
    assert log["charts/global_step"] == 64 * 32 * 6 and log["charts/episode"] >= 64  # horizon 24 < 32 steps
    assert 0 < log["losses/entropy_loss"] <= np.log(12) + 1e-4
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, agent.parameters()))
    assert len(processed) > 64 and len(curr) == 64 and len(rec["solved"]) + len(rec["unsolved"]) == 1190
    assert all(len(hist[s]) > 0 for s in rec["solved"])
    assert list(tmp_path.glob("out/*/ckpt.pt"))


def test_graph_and_eager_updates_agree():
    """One update from identical rollout data: the CUDA-graph minibatch step == the eager one (same kernels)."""
    from ac_solver_b200.agents.training import PPOLoss

    torch.manual_seed(1)
    dev = torch.device("cuda")
    nets = []
    for _ in range(2):
        torch.manual_seed(5)
        nets.append(torch.nn.Sequential(torch.nn.Linear(72, 64), torch.nn.Tanh(), torch.nn.Linear(64, 13)).to(dev))
    B = 512
    x = torch.randn(B, 72, device=dev)
    actions = torch.randint(0, 12, (B,), device=dev)
    old_lp, adv, ret, oldv = (torch.randn(B, device=dev) * 0.1 - 2.4, torch.randn(B, device=dev), torch.randn(B, device=dev),
                              torch.randn(B, device=dev))
    cfg = dict(norm_adv=True, is_loss_clip=True, clip_vloss=True, clip_coef=0.2, ent_coef=0.01, vf_coef=0.5)
    opts = [torch.optim.Adam(n.parameters(), lr=torch.tensor(1e-3, device=dev), capturable=True) for n in nets]

    def step(k):
        out = nets[k](x)
        loss, _ = PPOLoss.apply(out[:, :12], out[:, 12], actions, old_lp, adv, ret, oldv, None, cfg)
        opts[k].zero_grad(set_to_none=False)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(nets[k].parameters(), 0.5)
        opts[k].step()

    for _ in range(3):
        step(0)
        step(1)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step(1)
    for _ in range(4):
        step(0)
        graph.replay()
    torch.cuda.synchronize()
    for p, q in zip(nets[0].parameters(), nets[1].parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)


def test_fused_rollout_kernels():
    """acs_rollout_sample_record / acs_rollout_finish (csrc/ppo_kernels.cu) against the torch ops they replace:
    buffer writes at the device-resident time index, log-probabilities of the drawn actions, the action distribution
    (chi-square against softmax), episodic statistics and the ring of finished episodes."""
    from ac_solver_b200 import _lib

    L = _lib.lib()
    dev = torch.device("cuda")
    s = torch.cuda.current_stream().cuda_stream
    N, T, W, A = 5000, 4, 72, 12
    g = torch.Generator(device="cpu").manual_seed(0)
    state = torch.randint(-2, 3, (N, W), generator=g, dtype=torch.int8).to(dev)
    logits = (torch.randn(N, A, generator=g) * 1.5).to(dev)
    value = torch.randn(N, generator=g).to(dev)
    next_done = (torch.rand(N, generator=g) < 0.3).float().to(dev)
    ctr = torch.tensor([2, 77], dtype=torch.int64, device=dev)
    obs = torch.zeros((T, N, W), dtype=torch.int8, device=dev)
    dones, values, logprobs = (torch.zeros((T, N), device=dev) for _ in range(3))
    actions = torch.full((T, N), -1, dtype=torch.int64, device=dev)
    a8 = torch.zeros(N, dtype=torch.uint8, device=dev)
    _lib.check(L.acs_rollout_sample_record(state.data_ptr(), next_done.data_ptr(), logits.data_ptr(), value.data_ptr(), ctr.data_ptr(),
                                           obs.data_ptr(), dones.data_ptr(), values.data_ptr(), logprobs.data_ptr(), actions.data_ptr(),
                                           a8.data_ptr(), N, T, W, A, 1234, s))
    torch.cuda.synchronize()
    assert torch.equal(obs[2], state) and not obs[[0, 1, 3]].any()
    assert torch.equal(dones[2], next_done) and torch.equal(values[2], value)
    assert torch.equal(actions[2], a8.long()) and int(actions[2].min()) >= 0 and int(actions[2].max()) < A and (actions[[0, 1, 3]] == -1).all()
    ref_lp = torch.log_softmax(logits, -1).gather(1, actions[2][:, None]).squeeze(1)
    np.testing.assert_allclose(logprobs[2].cpu().numpy(), ref_lp.cpu().numpy(), rtol=1e-5, atol=1e-5)  # fast exp / log
    # same (seed, draw counter) -> same draws; another draw counter -> other draws
    act2 = torch.zeros_like(actions)
    _lib.check(L.acs_rollout_sample_record(state.data_ptr(), next_done.data_ptr(), logits.data_ptr(), value.data_ptr(), ctr.data_ptr(),
                                           obs.data_ptr(), dones.data_ptr(), values.data_ptr(), logprobs.data_ptr(), act2.data_ptr(),
                                           a8.data_ptr(), N, T, W, A, 1234, s))
    assert torch.equal(act2[2], actions[2])
    ctr[1] = 78
    _lib.check(L.acs_rollout_sample_record(state.data_ptr(), next_done.data_ptr(), logits.data_ptr(), value.data_ptr(), ctr.data_ptr(),
                                           obs.data_ptr(), dones.data_ptr(), values.data_ptr(), logprobs.data_ptr(), act2.data_ptr(),
                                           a8.data_ptr(), N, T, W, A, 1234, s))
    assert (act2[2] != actions[2]).float().mean() > 0.3
    # distribution: one logits row replicated 400 000 times, chi-square with 11 degrees of freedom (99.9 % quantile: 31.3)
    M = 400_000
    row = torch.tensor([0.3, -1.0, 2.0, 0.0, 0.5, -2.5, 1.0, 0.1, -0.4, 0.9, -3.0, 0.2])
    big = row.repeat(M, 1).to(dev)
    bobs = torch.zeros((1, M, 8), dtype=torch.int8, device=dev)
    bz = torch.zeros((1, M), device=dev)
    bact = torch.zeros((1, M), dtype=torch.int64, device=dev)
    c0 = torch.tensor([0, 5], dtype=torch.int64, device=dev)
    bd, bv, bl, b8 = bz.clone(), bz.clone(), bz.clone(), torch.zeros(M, dtype=torch.uint8, device=dev)
    bstate = torch.zeros((M, 8), dtype=torch.int8, device=dev)
    _lib.check(L.acs_rollout_sample_record(bstate.data_ptr(), bz.data_ptr(), big.data_ptr(), bz.data_ptr(), c0.data_ptr(), bobs.data_ptr(),
                                           bd.data_ptr(), bv.data_ptr(), bl.data_ptr(), bact.data_ptr(), b8.data_ptr(), M, 1, 8, A, 99, s))
    counts = torch.bincount(bact[0], minlength=A).double().cpu().numpy()
    p = torch.softmax(row.double(), 0).numpy()
    chi2 = float(((counts - M * p) ** 2 / (M * p)).sum())
    assert chi2 < 40.0, chi2

    # ---- finish: rewards[t], next_done, episodic statistics, ring of finished episodes ----
    r = torch.randn(N, generator=g).to(dev)
    done = (torch.rand(N, generator=g) < 0.005).to(torch.uint8).to(dev)
    trunc = (torch.rand(N, generator=g) < 0.005).to(torch.uint8).to(dev)
    rewards = torch.zeros((T, N), device=dev)
    nd = torch.zeros(N, device=dev)
    ep_r, ep_l = torch.randn(N, generator=g).to(dev), torch.randint(0, 50, (N,), generator=g).float().to(dev)
    ep_r0, ep_l0 = ep_r.clone(), ep_l.clone()
    ring_r, ring_l = torch.zeros(101, device=dev), torch.zeros(101, device=dev)
    ring_n = torch.ones(1, dtype=torch.int64, device=dev)
    ctr = torch.tensor([1, 0], dtype=torch.int64, device=dev)
    _lib.check(L.acs_rollout_finish(r.data_ptr(), done.data_ptr(), trunc.data_ptr(), ctr.data_ptr(), rewards.data_ptr(), nd.data_ptr(),
                                    ep_r.data_ptr(), ep_l.data_ptr(), ring_r.data_ptr(), ring_l.data_ptr(), ring_n.data_ptr(), N, T, 100, s))
    torch.cuda.synchronize()
    fin = (done | trunc).bool()
    assert torch.equal(rewards[1], r) and not rewards[[0, 2, 3]].any() and torch.equal(nd, done.float())
    assert torch.equal(ep_r[~fin], (ep_r0 + r)[~fin]) and torch.equal(ep_l[~fin], (ep_l0 + 1)[~fin])
    assert not ep_r[fin].any() and not ep_l[fin].any()
    nf = int(fin.sum())
    assert int(ring_n) == 1 + nf and 0 < nf < 100
    got = sorted(zip(ring_r[1 : 1 + nf].cpu().tolist(), ring_l[1 : 1 + nf].cpu().tolist()))
    exp = sorted(zip((ep_r0 + r)[fin].cpu().tolist(), (ep_l0 + 1)[fin].cpu().tolist()))
    assert got == exp and float(ring_r[0]) == 0.0


def test_fused_and_torch_rollout_paths_train():
    """The loop with the two bookkeeping kernels and the loop with the torch ops they replace both train (same
    algorithm, different random streams): finite losses, identical bookkeeping invariants."""
    from ac_solver_b200.agents.environment import get_env
    from ac_solver_b200.agents.ppo_agent import Agent
    from ac_solver_b200.agents.training import ppo_training_loop

    for fused in (True, False):
        a = _tiny_args()
        torch.manual_seed(0)
        envs, initial_states, curr, rec, hist, processed = get_env(a)
        dev = torch.device("cuda")
        agent = Agent(envs, a.nodes_counts).to(dev)
        opt = torch.optim.Adam(agent.parameters(), lr=torch.tensor(a.learning_rate, device=dev), eps=a.epsilon, capturable=True)
        log = ppo_training_loop(envs, a, dev, opt, agent, curr, rec, hist, processed, initial_states, checkpoint_every=0,
                                progress=False, fused_rollout=fused)
        assert np.isfinite(log["losses/value_loss"]) and log["charts/global_step"] == 64 * 32 * 6
        assert log["charts/episode"] >= 64 and 0 < log["losses/entropy_loss"] <= np.log(12) + 1e-4
        # every finished episode was truncated at the horizon (24) or solved earlier: mean length <= horizon
        assert 0 < log["charts/normalized_lengths_mean"] <= 1.0 + 1e-6
