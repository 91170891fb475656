"""Small driver for ncu: one BFS on AK(3) at the given budget (default 2e7)."""
import sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ac_solver_b200.search.breadth_first import bfs_device
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)
budget = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
s, p, i = bfs_device(AK3, budget)
print(i["n_visited"], i["n_expanded"], i["seconds_device"])
