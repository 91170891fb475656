"""PPO training loop on the GPU-resident environment (reference: ``ac_solver/agents/training.py``).

``ppo_training_loop`` keeps the reference's signature and algorithm (rollout of ``num_steps`` vector steps,
GAE, ``update_epochs`` x ``num_minibatches`` clipped / KL-penalised updates, target-KL early stop, linear /
cosine learning-rate schedule, checkpoint every 100 updates) but nothing in the rollout touches the host:

* rollout: policy forward, Gumbel-max sampling, fused env-step kernel, device-side reward wrappers and
  curriculum reset, and the writes into the [T, N] rollout buffers are ONE CUDA graph replayed ``num_steps``
  times (the time index is a device scalar, so the same graph serves every step);
* GAE: ``acs_gae`` (csrc/ppo_kernels.cu), bit-identical to the reference's reversed Python loop
  (training.py:230-250);
* update: minibatch gather, the two MLPs (torch / cuBLAS), the fused loss-and-gradient kernel
  ``acs_ppo_loss`` (training.py:262-318 in one launch), backward, gradient clipping and Adam are a second
  CUDA graph replayed once per minibatch.

The host sees one scalar per epoch (approx_kl for the early stop) and a handful per update (logging)."""

from __future__ import annotations

import math
import os
import random
import uuid

import numpy as np
import torch
from torch import nn

from .. import _lib


def get_curr_lr(n_update, lr_decay, warmup, max_lr, min_lr, total_updates):
    """Learning rate of update ``n_update`` (1-indexed): linear warm-up over the first ``warmup`` fraction of
    the updates, then linear or cosine decay from ``max_lr`` to ``min_lr`` (training.py:18-65)."""
    k, last = n_update - 1, total_updates - 1
    w_end = last * warmup
    if w_end > 0 and k <= w_end:
        return max_lr * k / w_end
    if lr_decay == "linear":
        slope = (max_lr - min_lr) / (w_end - last)
        return slope * k + (max_lr - slope * w_end)
    if lr_decay == "cosine":
        return min_lr + (max_lr - min_lr) * (1 + math.cos((k - w_end) / (last - w_end) * math.pi)) / 2
    raise NotImplementedError("Only 'linear' and 'cosine' lr-schedules are available.")


def gae(rewards, values, dones, next_value, next_done, gamma, gae_lambda, out=None):
    """Generalised advantage estimation on the device: [T, N] fp32 CUDA tensors in, (advantages, returns) out
    (written into ``out`` when given, so that a captured graph can keep reading the same buffers)."""
    T, N = rewards.shape
    for t in (rewards, values, dones, next_value, next_done):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    adv, ret = out if out is not None else (torch.empty_like(rewards), torch.empty_like(rewards))
    _lib.check(_lib.lib().acs_gae(rewards.data_ptr(), values.data_ptr(), dones.data_ptr(), next_value.data_ptr(),
                                  next_done.data_ptr(), adv.data_ptr(), ret.data_ptr(), int(T), int(N), float(gamma),
                                  float(gae_lambda), torch.cuda.current_stream(rewards.device).cuda_stream))
    return adv, ret


class PPOLoss(torch.autograd.Function):
    """loss = pg_loss - ent_coef * entropy + vf_coef * v_loss of one minibatch (training.py:262-318) from the
    actor's logits and the critic's values; forward and backward are the same kernel launch.  Returns
    ``(loss, stats)`` with stats = [loss, pg_loss, v_loss, entropy, approx_kl, clipfrac, 0, 0] (no gradient)."""

    @staticmethod
    def forward(ctx, logits, newvalue, actions, old_logprob, adv, returns, old_value, beta, cfg):
        L = _lib.lib()
        B, A = logits.shape
        logits, newvalue = logits.contiguous(), newvalue.contiguous()
        dlogits, dvalue = torch.empty_like(logits), torch.empty_like(newvalue)
        out = torch.zeros(8, dtype=torch.float32, device=logits.device)
        ws = torch.empty(L.acs_ppo_loss_workspace_bytes(), dtype=torch.uint8, device=logits.device)
        _lib.check(L.acs_ppo_loss(
            logits.data_ptr(), newvalue.data_ptr(), actions.data_ptr(), old_logprob.data_ptr(), adv.data_ptr(),
            returns.data_ptr(), old_value.data_ptr(), beta.data_ptr() if beta is not None else None, dlogits.data_ptr(),
            dvalue.data_ptr(), out.data_ptr(), ws.data_ptr(), int(B), int(A), int(cfg["norm_adv"]), int(cfg["is_loss_clip"]),
            int(cfg["clip_vloss"]), float(cfg["clip_coef"]), float(cfg["ent_coef"]), float(cfg["vf_coef"]),
            torch.cuda.current_stream(logits.device).cuda_stream))
        ctx.save_for_backward(dlogits, dvalue)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        dlogits, dvalue = ctx.saved_tensors
        return g_loss * dlogits, g_loss * dvalue, None, None, None, None, None, None, None


def _loss_cfg(args):
    return {"norm_adv": args.norm_adv, "is_loss_clip": args.is_loss_clip, "clip_vloss": args.clip_vloss,
            "clip_coef": args.clip_coef, "ent_coef": args.ent_coef, "vf_coef": args.vf_coef}


class _Graphed:
    """A step function that runs eagerly for its first ``warm`` calls and as a CUDA graph afterwards."""

    def __init__(self, fn, warm=3, enabled=True):
        self.fn, self.warm, self.enabled, self.calls, self.graph = fn, warm, enabled, 0, None

    def __call__(self):
        if not self.enabled or self.calls < self.warm:
            self.calls += 1
            return self.fn()
        if self.graph is None:
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):  # capture only records: the step is executed by the replay below
                self.fn()
        self.graph.replay()


def sync_curriculum(envs, curr_states, success_record, ACMoves_hist, states_processed):
    """Copy the device-side curriculum bookkeeping into the reference's host containers (training.py:169-222)."""
    if getattr(envs, "_curriculum", None) is None:
        return
    rec = envs.success_record()
    success_record["solved"].clear()
    success_record["solved"].update(rec["solved"])
    success_record["unsolved"].clear()
    success_record["unsolved"].update(rec["unsolved"])
    for s, acts in envs.acmoves_hist().items():
        ACMoves_hist[s] = acts
    cur = envs._curriculum["cur_state"].cpu().numpy()
    curr_states[:] = [int(c) for c in cur]
    nxt = envs.curriculum_counters()["next_unprocessed"]
    states_processed.update(range(min(nxt, envs._curriculum["n_states"])))
    states_processed.update(curr_states)


def ppo_training_loop(envs, args, device, optimizer, agent, curr_states, success_record, ACMoves_hist, states_processed,
                      initial_states, use_graphs=True, checkpoint_every=100, progress=True, fused_rollout=True):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.AcsError("ppo_training_loop runs on a CUDA device only (there is no CPU fallback)")
    T, N = args.num_steps, args.num_envs
    width = int(np.prod(envs.single_observation_space.shape))
    obs = torch.zeros((T, N, width), dtype=torch.int8, device=dev)  # int8 rollout store: a quarter of the reference's fp32
    actions = torch.zeros((T, N), dtype=torch.int64, device=dev)
    logprobs = torch.zeros((T, N), device=dev)
    rewards = torch.zeros((T, N), device=dev)
    dones = torch.zeros((T, N), device=dev)
    values = torch.zeros((T, N), device=dev)
    advantages = torch.zeros((T, N), device=dev)
    returns = torch.zeros((T, N), device=dev)
    ctr = torch.zeros(2, dtype=torch.int64, device=dev)  # {time index of the rollout, draw counter of the sampler}
    t_idx = ctr[0:1]
    next_done = torch.zeros(N, device=dev)
    ep_return = torch.zeros(N, device=dev)
    ep_length = torch.zeros(N, device=dev)
    # the reference's deque([0], maxlen=100) of episodic returns / lengths: slot 0 holds the initial 0, ring_n counts
    # the entries ever appended (initial one included); slot 100 is the dump slot of the torch-op path
    ring_ret = torch.zeros(101, device=dev)
    ring_len = torch.zeros(101, device=dev)
    ring_n = torch.ones(1, dtype=torch.int64, device=dev)
    action_u8 = torch.zeros(N, dtype=torch.uint8, device=dev)
    n_actions = int(envs.single_action_space.n)
    fused_rollout = bool(fused_rollout) and n_actions <= 16
    L = _lib.lib()
    transform = envs.norm_rewards or envs.clip_rewards is not None

    envs.reset()
    global_step = 0
    num_updates = args.total_timesteps // args.batch_size
    beta = None if args.is_loss_clip else torch.tensor([args.beta], dtype=torch.float32, device=dev)
    cfg = _loss_cfg(args)
    run_name = f"{args.exp_name}_ppo-ffn-nodes_{args.nodes_counts}_{uuid.uuid4()}"
    out_dir = os.path.join("out", run_name)
    wandb = None
    if getattr(args, "wandb_log", False):
        import wandb  # optional dependency, only when asked for

        wandb.init(project=args.wandb_project_name, name=run_name, config=vars(args), save_code=True)

    # ---- one vector step of the rollout, sync-free (graph) --------------------------------------
    def rollout_step_fused():
        # policy forward (cuBLAS) + TWO bookkeeping kernels around the environment step (csrc/ppo_kernels.cu)
        with torch.no_grad():
            state = envs.state
            logits, value = agent(state.float())
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(L.acs_rollout_sample_record(
                state.data_ptr(), next_done.data_ptr(), logits.data_ptr(), value.data_ptr(), ctr.data_ptr(), obs.data_ptr(),
                dones.data_ptr(), values.data_ptr(), logprobs.data_ptr(), actions.data_ptr(), action_u8.data_ptr(), N, T, width,
                n_actions, int(args.seed) & (2 ** 64 - 1), stream))
            envs.step_device(action_u8)
            r = envs.transformed_reward() if transform else envs.reward.float()
            _lib.check(L.acs_rollout_finish(
                r.data_ptr(), envs.done.data_ptr(), envs.truncated.data_ptr(), ctr.data_ptr(), rewards.data_ptr(),
                next_done.data_ptr(), ep_return.data_ptr(), ep_length.data_ptr(), ring_ret.data_ptr(), ring_len.data_ptr(),
                ring_n.data_ptr(), N, T, 100, stream))
            ctr.add_(1)

    def rollout_step():
        with torch.no_grad():
            state = envs.state
            obs.index_copy_(0, t_idx, state.unsqueeze(0))
            dones.index_copy_(0, t_idx, next_done.unsqueeze(0))
            action, lp, value = agent.sample(state.float())
            values.index_copy_(0, t_idx, value.unsqueeze(0))
            actions.index_copy_(0, t_idx, action.unsqueeze(0))
            logprobs.index_copy_(0, t_idx, lp.unsqueeze(0))
            envs.step_device(action)
            r = envs.transformed_reward() if transform else envs.reward.float()
            rewards.index_copy_(0, t_idx, r.unsqueeze(0))
            next_done.copy_(envs.done)
            # episodic statistics (training.py:162-163, 191-196)
            ep_return.add_(r)
            ep_length.add_(1.0)
            fin = (envs.done | envs.truncated).bool()
            pos = (ring_n + torch.cumsum(fin, 0) - 1) % 100
            slot = torch.where(fin, pos, torch.full_like(pos, 100))
            ring_ret.scatter_(0, slot, ep_return)
            ring_len.scatter_(0, slot, ep_length)
            ring_n.add_(fin.sum())
            ep_return.masked_fill_(fin, 0.0)
            ep_length.masked_fill_(fin, 0.0)
            ctr.add_(1)

    rollout = _Graphed(rollout_step_fused if fused_rollout else rollout_step, warm=3, enabled=use_graphs)

    # ---- one minibatch update (graph) -----------------------------------------------------------
    mb = args.minibatch_size
    mb_idx = torch.zeros(mb, dtype=torch.int64, device=dev)
    flat = dict(obs=obs.reshape(T * N, width), actions=actions.reshape(-1), logprobs=logprobs.reshape(-1),
                advantages=advantages.reshape(-1), returns=returns.reshape(-1), values=values.reshape(-1))
    stats_acc = torch.zeros(8, device=dev)   # running sums over the minibatches of an update (clipfrac mean)
    last_stats = torch.zeros(8, device=dev)
    capturable = all(g.get("capturable", False) for g in optimizer.param_groups)

    def minibatch_step(idx=None):
        idx = mb_idx if idx is None else idx
        x = flat["obs"].index_select(0, idx).float()
        logits, newvalue = agent(x)
        loss, stats = PPOLoss.apply(logits, newvalue, flat["actions"].index_select(0, idx),
                                    flat["logprobs"].index_select(0, idx), flat["advantages"].index_select(0, idx),
                                    flat["returns"].index_select(0, idx), flat["values"].index_select(0, idx), beta, cfg)
        optimizer.zero_grad(set_to_none=False)
        loss.backward()
        nn.utils.clip_grad_norm_(agent.parameters(), args.max_grad_norm)
        optimizer.step()
        last_stats.copy_(stats)
        stats_acc.add_(stats)

    update_step = _Graphed(minibatch_step, warm=3, enabled=use_graphs and capturable)

    def set_lr(lr):
        g = optimizer.param_groups[0]
        if torch.is_tensor(g["lr"]):
            g["lr"].fill_(lr)
        else:
            g["lr"] = lr

    print(f"total number of timesteps: {args.total_timesteps}, updates: {num_updates}")
    it = range(1, num_updates + 1)
    if progress:
        try:
            from tqdm import tqdm

            it = tqdm(it, desc="Training Progress", total=num_updates)
        except ImportError:
            pass
    log = {}
    update_seconds = []  # wall time of every whole iteration (rollout + GAE + updates + logging), for benchmarks
    import time as _time

    for update in it:
        _t0 = _time.perf_counter()
        random.seed(args.seed + update)
        np.random.seed(args.seed + update)
        if not (use_graphs):  # a captured graph owns its RNG offsets; reseeding would not reach it
            torch.manual_seed(args.seed + update)
        if args.anneal_lr:
            set_lr(get_curr_lr(update, args.lr_decay, args.warmup_period, args.learning_rate,
                               args.learning_rate * args.min_lr_frac, num_updates))

        # ---- rollout ----
        t_idx.zero_()
        for _ in range(T):
            rollout()
        global_step += T * N
        envs.check_errors()  # one host read per update

        if not args.norm_rewards:  # training.py:230-238: manual rescaling when there is no NormalizeReward wrapper
            rewards.div_(envs.max_reward)

        # ---- GAE (training.py:240-250) ----
        with torch.no_grad():
            next_value = agent.get_value(envs.state.float()).reshape(-1).contiguous()
        gae(rewards, values, dones, next_value, next_done, args.gamma, args.gae_lambda, out=(advantages, returns))

        # ---- optimisation (training.py:262-352) ----
        b_inds = np.arange(args.batch_size)
        stats_acc.zero_()
        n_mb = 0
        approx_kl = 0.0
        for _epoch in range(args.update_epochs):
            np.random.shuffle(b_inds)
            inds_dev = torch.from_numpy(b_inds).to(dev, non_blocking=False)
            for start in range(0, args.batch_size, mb):
                sel = inds_dev[start : start + mb]
                if sel.numel() != mb:  # ragged tail (batch_size not divisible by num_minibatches): eager, own size
                    minibatch_step(sel)
                else:
                    mb_idx.copy_(sel)
                    update_step()
                n_mb += 1
            approx_kl = float(last_stats[4])  # the epoch's one host read
            if args.is_loss_clip:
                if args.target_kl is not None and approx_kl > args.target_kl:
                    break
            else:
                b = float(beta)
                beta.fill_(b / 2 if approx_kl < args.target_kl / 1.5 else (b * 2 if approx_kl > args.target_kl * 1.5 else b))

        # ---- logging ----
        y_pred, y_true = flat["values"], flat["returns"]
        var_y = torch.var(y_true, unbiased=False)
        ev = float("nan") if float(var_y) == 0 else float(1 - torch.var(y_true - y_pred, unbiased=False) / var_y)
        n_episodes = int(ring_n) - 1
        k = min(n_episodes + 1, 100)
        rets, lens = ring_ret[:k].cpu().numpy(), ring_len[:k].cpu().numpy()
        if not args.norm_rewards:
            rets, lens = rets / envs.max_reward, lens / args.horizon_length
        ls = last_stats.cpu().numpy()
        lr_now = optimizer.param_groups[0]["lr"]
        log = {"charts/global_step": global_step, "charts/episode": n_episodes,
               "charts/normalized_returns_mean": float(rets.mean()), "charts/normalized_lengths_mean": float(lens.mean()),
               "charts/learning_rate": float(lr_now), "losses/value_loss": float(ls[2]), "losses/policy_loss": float(ls[1]),
               "losses/entropy_loss": float(ls[3]), "losses/approx_kl": float(ls[4]), "losses/explained_variance": ev,
               "losses/clipfrac": float(stats_acc[5]) / max(n_mb, 1), "debug/advantages_mean": float(flat["advantages"].mean()),
               "debug/advantages_std": float(flat["advantages"].std())}
        if getattr(envs, "_curriculum", None) is not None:
            c = envs.curriculum_counters()
            log["charts/solved"] = c["n_solved"]
            log["charts/unsolved"] = envs._curriculum["n_states"] - c["n_solved"]
        if wandb is not None:
            wandb.log(log)
        update_seconds.append(_time.perf_counter() - _t0)  # (the logging above read device scalars: the update is complete)
        log["perf/update_seconds"] = list(update_seconds)

        if checkpoint_every and update % checkpoint_every == 0:
            sync_curriculum(envs, curr_states, success_record, ACMoves_hist, states_processed)
            os.makedirs(out_dir, exist_ok=True)
            torch.save({"critic": agent.critic.state_dict(), "actor": agent.actor.state_dict(),
                        "optimizer": optimizer.state_dict(), "update": update, "episode": n_episodes,
                        "config": vars(args), "mean_return": float(rets.mean()), "success_record": success_record,
                        "value_loss": float(ls[2]), "policy_loss": float(ls[1]), "entropy_loss": float(ls[3]),
                        "approx_kl": float(ls[4]), "explained_var": ev, "clipfrac": log["losses/clipfrac"],
                        "global_step": global_step,
                        "round1_complete": (getattr(envs, "_curriculum", None) is not None
                                            and envs.curriculum_counters()["next_unprocessed"] >= envs._curriculum["n_states"]),
                        "curr_states": curr_states, "states_processed": states_processed,
                        "ACMoves_hist": ACMoves_hist, "supermoves": None}, os.path.join(out_dir, "ckpt.pt"))
            print(f"saving checkpoint to {out_dir}")

    sync_curriculum(envs, curr_states, success_record, ACMoves_hist, states_processed)
    return log
