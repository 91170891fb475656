#!/usr/bin/env python
"""bench.py -- headline benchmark of the AC-move hot path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the batched ``ACEnv.step`` kernel over one batch of 1 Mi synthetic
presentations (max_relator_length 36, uniform random AC' moves) per GPU.  Rows are independent,
so N GPUs run N shards with no data-path collective (weak scaling).  One JSON line is printed
by rank 0; see the task contract for the keys.  ``--impl reference`` times the CPU oracle (the
C restatement of the reference's pure-Python path, oracle/) on all host threads instead.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HORIZON = 200
# The synthetic states are freely and cyclically reduced, and ACEnv.step keeps them so: the
# steady-state (ACS_FLAG_NORMALIZED) variant of the kernel is the one an environment runs on
# every step but the first after a reset with caller-supplied states.  Like ACEnv.lengths, the
# relator lengths travel with the state (ACS_FLAG_LENS_VALID).  BENCH_FLAGS=0 times the general variant.
FLAGS = int(os.environ.get("BENCH_FLAGS", "6"))
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1 << 20, help="presentations per GPU per step")
    ap.add_argument("--mrl", type=int, default=36)
    ap.add_argument("--cpu-sample-rows", type=int, default=1 << 17)
    ap.add_argument("--skip-bfs", action="store_true")
    ap.add_argument("--bfs-budget", type=int, default=100_000_000)
    ap.add_argument("--bfs-timeout", type=float, default=90.0)
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []  # (phase, sm_mhz, reasons)
        self.phase = "warmup"
        self.stop_flag = False
        self.sm_max = None
        self.ok = False

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.ok = True
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((self.phase, sm, r))
                time.sleep(0.002)
        except Exception:
            self.ok = False

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        timed = [s for s in self.samples if s[0] == "timed"] or [s for s in self.samples if s[0] != "idle"]
        names = {
            0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
            0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
            0x100: "display_clock_setting",
        }
        mask = 0
        for s in timed:
            mask |= s[2]
        return {
            "sm_mhz": statistics.median(s[1] for s in timed),
            "sm_max_mhz": self.sm_max,
            "reasons": sorted(n for b, n in names.items() if mask & b),
            "samples": len(timed),
        }


def cpu_baseline(rows, mrl, seconds_target=12.0):
    """The CPU oracle (C restatement of the reference path) on all host threads, bounded sample."""
    from ac_solver_b200.synthetic import random_actions, random_presentations
    from oracle import oracle as O

    S = random_presentations(rows, mrl, seed=0)
    sc = np.zeros(rows, np.int32)
    A = random_actions(rows, seed=1)
    threads = host_threads()  # torchrun pins OMP_NUM_THREADS=1: ask for every core explicitly
    O.env_step_batch(S, A, sc, HORIZON, nthreads=threads)  # warm
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.env_step_batch(S, A, sc, HORIZON, nthreads=threads)
        reps += 1
        if time.perf_counter() - t0 > seconds_target or reps >= 2000:
            break
    dt = time.perf_counter() - t0
    return {
        "value": rows * reps / dt,
        "unit": "moves/s",
        "cores": threads,
        "kind": "port",
        "sample": f"{reps} passes of ACEnv.step over {rows} synthetic rows (mrl {mrl}), C oracle, {threads} OpenMP threads",
    }


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from ac_solver_b200.synthetic import random_actions, random_presentations
    from oracle import oracle as O

    rows, mrl = args.cpu_sample_rows, args.mrl
    S = random_presentations(rows, mrl, seed=0)
    A = random_actions(rows, seed=1)
    sc = np.zeros(rows, np.int32)
    threads = host_threads()  # torchrun pins OMP_NUM_THREADS=1: ask for every core explicitly
    for _ in range(max(args.warmup, 1)):
        O.env_step_batch(S, A, sc, HORIZON, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.env_step_batch(S, A, sc, HORIZON, nthreads=threads)
    dt = time.perf_counter() - t0
    v = rows * args.steps / dt
    sample = f"each step = ACEnv.step over a {rows}-row sample of the 1 Mi-row workload, C oracle on {threads} host threads"
    line = {
        "impl": "reference", "metric": "AC moves/sec (batched env steps)", "value": v, "unit": "moves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"batched ACEnv.step, {rows}-row sample of 1Mi rows, mrl {mrl}, uniform 12 moves, CPU"},
        "cpu_baseline": {"value": v, "unit": "moves/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the AC-move kernels have no CPU fallback")
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version
    # banner there) are sent to stderr until the line is printed.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))

    from ac_solver_b200 import _lib
    from ac_solver_b200.synthetic import random_actions, random_presentations

    L = _lib.lib()
    ctx = _lib.ctx(local_rank)
    rows, mrl = args.rows, args.mrl
    rowb = 2 * mrl
    nbuf = max(4, -(-3 * L2_BYTES // (rows * rowb)) + 1)  # rotated state buffers: working set > L2
    base = random_presentations(min(rows, 1 << 18), mrl, seed=rank)
    reps = -(-rows // len(base))
    host_states = np.tile(base, (reps, 1))[:rows]
    states = [torch.from_numpy(np.roll(host_states, 7919 * b, axis=0).copy()).cuda() for b in range(nbuf)]
    actions = [torch.from_numpy(random_actions(rows, seed=1 + 17 * b + rank)).cuda() for b in range(nbuf)]
    reward = torch.zeros(rows, dtype=torch.int32, device="cuda")
    done = torch.zeros(rows, dtype=torch.uint8, device="cuda")
    trunc = torch.zeros(rows, dtype=torch.uint8, device="cuda")
    stepc = [torch.zeros(rows, dtype=torch.int32, device="cuda") for _ in range(nbuf)]
    # ACEnv keeps `lengths` beside `state` (ac_env.py:84-92): so does the batched env (2 B per row)
    lens = [torch.stack([(s[:, :mrl] != 0).sum(1), (s[:, mrl:] != 0).sum(1)], dim=1).to(torch.uint8).contiguous()
            for s in states]
    err = torch.tensor([0, -1], dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step(i):
        b = i % nbuf
        rc = L.acs_env_step_batch(states[b].data_ptr(), actions[(i // nbuf + b) % nbuf].data_ptr(), reward.data_ptr(),
                                  done.data_ptr(), trunc.data_ptr(), stepc[b].data_ptr(), lens[b].data_ptr(), None,
                                  err.data_ptr(), rows, mrl, HORIZON, FLAGS, sptr)
        if rc != 0:
            _lib.check(rc)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.phase = "timed"
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record(stream)
    torch.cuda.synchronize()
    sampler.phase = "post"
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    n_bad = int(err[0].item())

    # ---- e2e: the vector-env call with HOST buffers (actions in; obs, reward, flags out) ----
    e2e_steps = max(3, min(args.steps, 30))
    h_act = [torch.from_numpy(random_actions(rows, seed=100 + b + rank)).pin_memory() for b in range(2)]
    h_obs = torch.empty((rows, rowb), dtype=torch.int8).pin_memory()
    h_rew = torch.empty(rows, dtype=torch.int32).pin_memory()
    h_done = torch.empty(rows, dtype=torch.uint8).pin_memory()
    h_tr = torch.empty(rows, dtype=torch.uint8).pin_memory()
    nbad = C.c_int64(0)

    def e2e_step(i):
        b = i % nbuf
        _lib.check(L.acs_env_step_host(ctx, states[b].data_ptr(), stepc[b].data_ptr(), h_act[i % 2].data_ptr(),
                                       h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr(), h_tr.data_ptr(),
                                       rows, mrl, HORIZON, FLAGS, C.byref(nbad)))

    for i in range(3):
        e2e_step(i)
    barrier()
    sampler.phase = "e2e"
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.phase = "post"
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_per_step = ms / args.steps
        moves_per_s = world * rows * args.steps / (ms * 1e-3)
        algo_bytes = 4 * mrl + 6  # 2*mrl in + 2*mrl out + action 1 + reward 4 + done 1 (150 at mrl 36)
        achieved = algo_bytes * rows / (ms_per_step * 1e-3) / 1e9
        # the CPU baseline is timed on rank 0 at N=1 only (it would stall the other ranks)
        cpu = cpu_baseline(args.cpu_sample_rows, mrl) if world == 1 else None
        line = {
            "metric": "AC moves/sec (batched env steps)",
            "value": moves_per_s,
            "unit": "moves/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "int8",
            "data": "synthetic",
            "config": {
                "workload": f"batched ACEnv.step microbench: {rows} synthetic random presentations per GPU, "
                            f"max_relator_length {mrl}, uniform random over the 12 AC' moves, horizon {HORIZON} "
                            "(BASELINE.json configs[1])",
                "rows_per_gpu": rows,
                "l2": f"rotating {nbuf} distinct in-place state buffers ({nbuf * rows * rowb >> 20} MiB > 126 MiB L2)",
                "parallelism": f"{world} independent shards, no collective",
                "rows_raising": n_bad,
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_bytes(), "peak_source": peak_src, "algorithmic_bytes_per_move": 4 * mrl + 6,
                "kernel": ("acs::ac_step_words_kernel<NW=9, TRUSTED=%s, LENS=%s, TR=128>"
                           % (bool(FLAGS & 2), bool(FLAGS & 4))) if mrl == 36 else "acs::ac_step_*_kernel",
                "kernel_flags": FLAGS,
                "frac_of_nominal_8TBs": achieved / 8000.0,
            },
            "cpu_baseline": cpu,
            "e2e": {
                "value": world * rows * e2e_steps / e2e_s,
                "unit": "moves/s",
                "h2d_bytes_per_step": rows * 1,
                "d2h_bytes_per_step": rows * (rowb + 4 + 1 + 1),
                "steps": e2e_steps,
                "api": "acs_env_step_host: pinned host actions in; host observations, rewards, done, truncated out",
            },
            "gpu_launches": args.steps,
            "clocks": sampler.summary(),
        }
    else:
        line = None

    def emit(extra=None):
        if rank == 0:
            if extra is not None:
                line["bfs"] = extra
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(line), flush=True)

    # Secondary metric (BFS nodes expanded/s).  At N > 1 it is a multi-rank collective program: a
    # watchdog guarantees that the headline line is printed and every rank exits even if it stalls.
    bfs_line = None
    if not args.skip_bfs and (rank == 0 or world > 1):
        def bail():
            emit({"error": f"sharded bfs did not finish within {args.bfs_timeout} s"})
            os._exit(0)

        dog = threading.Timer(args.bfs_timeout, bail)
        dog.daemon = True
        dog.start()
        try:
            bfs_line = bench_bfs(args, world, dist)
        except Exception as e:  # the search bench is auxiliary; never lose the headline line
            bfs_line = {"error": repr(e)}
            if world > 1:  # ranks may be out of step now: no further collectives
                dog.cancel()
                emit(bfs_line)
                os._exit(0)
        dog.cancel()
    emit(bfs_line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the latest committed `ncu --set full` capture of
    the step kernel (profiles/k1_step_r1_ncu_summary.json; cold-cache, the write-back of the last
    tiles is still in L2 when the capture ends, so it reads below the algorithmic bytes)."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_step_r1_ncu_summary.json")) as f:
            d = json.load(f)
        m = d[sorted(d)[-1]]
        return (float(m["dram__bytes_read.sum"]) + float(m["dram__bytes_write.sum"])) * 1e6
    except Exception:
        return None


def bench_bfs(args, world=1, dist=None):
    """Secondary metric: BFS nodes expanded / s on AK(3), mrl 24 (BASELINE.json configs[4]).
    One GPU: the single-device search (csrc/bfs.cu).  Several GPUs: the hash-partitioned search
    with an NCCL all-to-all per chunk (search/sharded.py), same total budget (strong scaling)."""
    ak3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18, np.int8)
    if world > 1:
        import contextlib
        import io

        import torch
        from ac_solver_b200.search.sharded import bfs_sharded

        with contextlib.redirect_stdout(io.StringIO()):
            bfs_sharded(ak3, 1_000_000)  # warm
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            solved, path, info = bfs_sharded(ak3, args.bfs_budget)
            torch.cuda.synchronize()
            dist.barrier()
        wall = time.perf_counter() - t0
        return {
            "metric": "BFS nodes expanded/sec", "nodes_expanded": info["n_expanded"], "visited": info["n_visited"],
            "workload": f"sharded bfs AK(3) mrl 24 budget {args.bfs_budget}, {world} GPUs, hash-partitioned, "
                        "NCCL all-to-all per chunk",
            "levels": info["n_levels"], "expanded_per_s_wall": info["n_expanded"] / wall, "seconds_wall": wall,
        }
    from ac_solver_b200.search.breadth_first import bfs_device

    bfs_device(ak3, 100000)  # warm
    t0 = time.perf_counter()
    solved, path, info = bfs_device(ak3, args.bfs_budget)
    wall = time.perf_counter() - t0
    # CPU baseline beside it: the C oracle's sequential bfs on one core,
    # bounded to a 2e6-node budget of the same search
    from oracle import oracle as O

    c0 = time.perf_counter()
    _, _, cinfo = O.bfs(ak3, 2_000_000)
    cpu_s = time.perf_counter() - c0
    return {
        "metric": "BFS nodes expanded/sec", "workload": f"bfs AK(3) mrl 24 budget {args.bfs_budget}, 1 GPU",
        "nodes_expanded": info["n_expanded"], "visited": info["n_visited"], "levels": info["n_levels"],
        "expanded_per_s_device": info["n_expanded"] / max(info["seconds_device"], 1e-9),
        "expanded_per_s_wall": info["n_expanded"] / wall, "seconds_wall": wall,
        "cpu_baseline": {"value": cinfo["n_expanded"] / cpu_s, "unit": "nodes expanded/s", "cores": 1, "kind": "port",
                         "sample": "C oracle bfs, same presentation, budget 2e6 (sequential algorithm, one core)"},
    }


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
