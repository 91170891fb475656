#!/usr/bin/env python
"""Multi-GPU check + timing of the native partitioned BFS (csrc/pbfs.cu), one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/run_partitioned.py [--budget B] [--skip-check]
Rank 0 compares result, counters and the visited ARRAY (insertion order) with the CPU oracle at
small budgets, then every rank times the big search; one JSON line per timed budget."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200.search.partitioned import PartitionedBfs, bfs_partitioned  # noqa: E402

AK2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0])
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=str, default="1e8")
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    if not args.skip_check:
        from oracle import oracle as O

        for pres, budget, cyc, chunk in ((AK2, 10, False, 0), (AK2, 5000, False, 0), (AK2, 1000000, False, 0),
                                         (AK2, 3000, True, 0), (AK3, 7777, False, 0), (AK3, 400000, False, 20000),
                                         (AK3, 3000000, False, 0)):
            solved, path, info = bfs_partitioned(pres, budget, cyc, want_visited=True, chunk_parents=chunk, verbose=None)
            if rank == 0:
                es, ep, ei = O.bfs(pres, budget, cyc, want_visited=True)
                ok = (solved, path) == (es, ep) and info["n_visited"] == ei["n_visited"] and \
                    info["n_expanded"] == ei["n_expanded"] and info["minlen_log"] == ei["minlen_log"] and \
                    np.array_equal(info["visited"], ei["visited"])
                print(f"check world={world} budget={budget} cyc={cyc}: {'OK' if ok else 'MISMATCH'} "
                      f"visited={info['n_visited']} chunks={info['chunks']}", flush=True)
                assert ok
    for b in args.budget.split(","):
        budget = int(float(b))
        with PartitionedBfs(24, budget, chunk_parents=args.chunk) as eng:
            eng.run(AK3)  # warm-up (first touch of the buffers, peer mappings)
            for _ in range(args.reps):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                solved, path, info = eng.run(AK3)
                dt = time.perf_counter() - t0
                t = torch.tensor([dt, info["seconds_device"]], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rank == 0:
                    print(json.dumps({"metric": "BFS nodes expanded/sec (partitioned, native driver)", "world": world,
                                      "budget": budget, "visited": info["n_visited"], "expanded": info["n_expanded"],
                                      "levels": info["n_levels"], "chunks": info["chunks"],
                                      "seconds_wall": float(t[0]), "seconds_device": float(t[1]),
                                      "expanded_per_s": info["n_expanded"] / float(t[1]),
                                      "records_recv_rank0": info["records_recv"][0],
                                      "n_local_rank0": info["n_local"][0]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
