#!/usr/bin/env python
"""Multi-GPU check + timing of the sharded BFS.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/run_sharded.py [--budget B]
Every rank runs the chunk loop; rank 0 compares with the CPU oracle at small budgets and prints
nodes expanded / s at the large one."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200.search.sharded import bfs_sharded  # noqa: E402

AK2 = np.array([1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0])
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=int, default=100_000_000)
    ap.add_argument("--skip-check", action="store_true")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    if not args.skip_check:
        from oracle import oracle as O

        for pres, budget, cyc in ((AK2, 10, False), (AK2, 5000, False), (AK2, 1000000, False), (AK2, 3000, True),
                                  (AK3, 7777, False), (AK3, 400000, False), (AK3, 3000000, False)):
            solved, path, info = bfs_sharded(pres, budget, cyc, want_visited=True, chunk_parents=60000)
            if rank == 0:
                es, ep, ei = O.bfs(pres, budget, cyc, want_visited=True)
                ok = (solved, path) == (es, ep) and info["n_visited"] == ei["n_visited"] and \
                    info["n_expanded"] == ei["n_expanded"] and np.array_equal(info["visited"], ei["visited"])
                print(f"check world={world} budget={budget} cyc={cyc}: {'OK' if ok else 'MISMATCH'} "
                      f"visited={info['n_visited']} local0={info['n_local']}", flush=True)
                assert ok
    bfs_sharded(AK3, 1_000_000)  # warm-up
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    solved, path, info = bfs_sharded(AK3, args.budget)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"metric": "BFS nodes expanded/sec (sharded)", "world": world, "budget": args.budget,
                          "visited": info["n_visited"], "expanded": info["n_expanded"], "levels": info["n_levels"],
                          "seconds_wall": dt, "expanded_per_s": info["n_expanded"] / dt,
                          "visited_per_s": info["n_visited"] / dt}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
