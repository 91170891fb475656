"""world_size-2 gloo test of the sharded BFS host logic (ac_solver_b200/search/sharded.py):
the real chunk loop with a numpy per-rank backend must reproduce the oracle bit for bit."""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

AK2 = [1, 1, -2, -2, -2, 0, 0, 1, 2, 1, -2, -1, -2, 0]
CASES = [(AK2, 1, False), (AK2, 13, False), (AK2, 700, False), (AK2, 3000, True), (AK2, 40000, False),
         ([1, -1, 2, 2, 1, 0, 0, 0, 2, 1, -1, -2, -2, 1, 0, 0], 300, False),
         ([1, 2, 0, 0, 1, 2, 0, 0], 50, False)]  # the last one raises AssertionError in the reference


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from ac_solver_b200.search.sharded import bfs_sharded
    from sharded_numpy_ops import NumpyShardOps

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for pres, budget, cyc in CASES:
        try:
            solved, path, info = bfs_sharded(np.array(pres), budget, cyc, ops_factory=NumpyShardOps, want_visited=True)
            out.append((solved, path, info["n_visited"], info["n_expanded"], info["budget_hit"], info["minlen_log"],
                        info.get("visited"), info["n_local"]))
        except AssertionError:
            out.append("AssertionError")
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 5])
def test_sharded_bfs_gloo(world):
    from oracle import oracle as O

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=800) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, (pres, budget, cyc) in enumerate(CASES):
        try:
            es, ep, ei = O.bfs(np.array(pres, np.int8), budget, cyc, want_visited=True)
        except AssertionError:
            assert all(results[r][k] == "AssertionError" for r in range(world))
            continue
        for rank in range(world):
            solved, path, nv, ne, bh, mins, vis, nloc = results[rank][k]
            assert (solved, path) == (es, ep), (k, rank)
            assert (nv, ne, bh, mins) == (ei["n_visited"], ei["n_expanded"], ei["budget_hit"], ei["minlen_log"]), (k, rank)
        assert np.array_equal(results[0][k][6], ei["visited"]), k  # rank 0 holds the gathered, ordered array
        if ei["n_visited"] > 100:  # both shards actually hold nodes
            assert all(results[r][k][7] > 0 for r in range(world))
            assert sum(results[r][k][7] for r in range(world)) == ei["n_visited"]
