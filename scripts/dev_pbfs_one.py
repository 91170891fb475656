#!/usr/bin/env python
"""One partitioned-BFS run for profiling: python scripts/dev_pbfs_one.py BUDGET [SIM_WORLD] [CHUNK]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200.search.partitioned import PartitionedBfs  # noqa: E402

AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)
budget = int(float(sys.argv[1]))
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
with PartitionedBfs(24, budget, sim_world=world, chunk_parents=chunk) as eng:
    t0 = time.perf_counter()
    solved, path, info = eng.run(AK3)
    dt = time.perf_counter() - t0
print(json.dumps({"budget": budget, "sim_world": world, "chunk": info["chunk_cap"], "visited": info["n_visited"],
                  "expanded": info["n_expanded"], "chunks": info["chunks"], "seconds_device": info["seconds_device"],
                  "seconds_wall": dt, "expanded_per_s_device": info["n_expanded"] / info["seconds_device"]}))
