#!/usr/bin/env python
"""Development timing of the native partitioned BFS on one GPU (simulated worlds included)."""
import json
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200.search.partitioned import PartitionedBfs  # noqa: E402

AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)

for budget, world, chunk in [(10**8, 1, 0), (10**8, 1, 1 << 20), (10**8, 2, 0), (10**9, 1, 0)]:
    if len(sys.argv) > 1 and budget > int(float(sys.argv[1])):
        continue
    with PartitionedBfs(24, budget, sim_world=world, chunk_parents=chunk) as eng:
        eng.run(AK3)
        t0 = time.perf_counter()
        solved, path, info = eng.run(AK3)
        dt = time.perf_counter() - t0
    print(json.dumps({"budget": budget, "sim_world": world, "chunk": info["chunk_cap"], "visited": info["n_visited"],
                      "expanded": info["n_expanded"], "chunks": info["chunks"], "seconds_device": info["seconds_device"],
                      "seconds_wall": dt, "expanded_per_s_device": info["n_expanded"] / info["seconds_device"],
                      "records_recv": sum(info["records_recv"])}), flush=True)
