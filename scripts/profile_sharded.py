"""world=1 run of the sharded BFS chunk loop with per-phase wall timing (cuda-synchronised)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200.search import sharded
AK3 = np.array([1, 1, 1, -2, -2, -2, -2] + [0] * 17 + [1, 2, 1, -2, -1, -2] + [0] * 18)
budget = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
T = {}
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(self, *a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(cls, name, g)
for n in ("__init__", "add_root", "begin_chunk", "expand_count", "expand_scatter", "insert_mark", "finish", "find_cut", "commit", "room"):
    wrap(sharded.GpuShardOps, n)
sharded.bfs_sharded(AK3, 1_000_000)
T.clear()
import io, contextlib
torch.cuda.synchronize(); t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    s, p, i = sharded.bfs_sharded(AK3, budget)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print("total", round(tot, 4), "visited", i["n_visited"], {k: round(v, 4) for k, v in T.items()}, "other", round(tot - sum(T.values()), 4))
