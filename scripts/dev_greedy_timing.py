"""Where the wall time of the config-3 greedy sweep goes: create / run / destroy per length group, run one after the
other (default) or concurrently from a thread pool as bench.py does (--threads)."""
import ctypes as C
import os
import sys
import time
from ast import literal_eval
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200 import _lib  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pres = [np.array(literal_eval(l), dtype=np.int8) for l in open(os.path.join(ROOT, "ac_solver_b200", "search", "miller_schupp", "data", "all_presentations.txt")) if l.strip()]
groups = {}
for k, p in enumerate(pres):
    groups.setdefault(p.size, []).append(k)
L = _lib.lib()
import torch  # noqa: E402

torch.cuda.init()
torch.zeros(1, device="cuda")
T0 = time.perf_counter()


def one(item):
    w, rows = item
    P8 = np.ascontiguousarray(np.stack([pres[k] for k in rows]))
    S = len(rows)
    h = C.c_void_p()
    t0 = time.perf_counter()
    _lib.check(L.acs_greedy_create(0, S, w // 2, 1000000, 0, 4096, C.byref(h)))
    t1 = time.perf_counter()
    paths = np.zeros((S, 4096, 2), np.int32)
    res = (_lib.SearchResult * S)()
    _lib.check(L.acs_greedy_run(h, P8.ctypes.data, paths.ctypes.data, res))
    t2 = time.perf_counter()
    L.acs_greedy_destroy(h)
    t3 = time.perf_counter()
    return (w // 2, S, t0 - T0, t1 - T0, t2 - T0, t3 - T0, res[0].seconds_device)


items = list(groups.items())
if "--threads" in sys.argv:
    with ThreadPoolExecutor(max_workers=8) as pool:
        out = list(pool.map(one, items))
else:
    out = [one(it) for it in items]
for mrl, S, a, b, c, d, dev in out:
    print(f"mrl {mrl}: {S} searches  create {a:.3f}->{b:.3f}  run ->{c:.3f} (device {dev:.3f})  destroy ->{d:.3f}")
print("wall %.3f" % (time.perf_counter() - T0))
