"""Where the wall time of the config-3 greedy sweep goes: create / run / destroy per length group (sequential)."""
import ctypes as C
import os
import sys
import time
from ast import literal_eval

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200 import _lib  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pres = [np.array(literal_eval(l), dtype=np.int8) for l in open(os.path.join(ROOT, "ac_solver_b200", "search", "miller_schupp", "data", "all_presentations.txt")) if l.strip()]
groups = {}
for k, p in enumerate(pres):
    groups.setdefault(p.size, []).append(k)
L = _lib.lib()
import torch
torch.cuda.init()
tot = [0, 0, 0]
for w, rows in groups.items():
    P8 = np.ascontiguousarray(np.stack([pres[k] for k in rows]))
    S = len(rows)
    h = C.c_void_p()
    t0 = time.perf_counter()
    _lib.check(L.acs_greedy_create(0, S, w // 2, 1000000, 0, 4096, C.byref(h)))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    paths = np.zeros((S, 4096, 2), np.int32)
    res = (_lib.SearchResult * S)()
    _lib.check(L.acs_greedy_run(h, P8.ctypes.data, paths.ctypes.data, res))
    t2 = time.perf_counter()
    L.acs_greedy_destroy(h)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"mrl {w // 2}: {S} searches  create {t1 - t0:.3f}  run {t2 - t1:.3f} (device {res[0].seconds_device:.3f})  destroy {t3 - t2:.3f}")
    tot[0] += t1 - t0
    tot[1] += t2 - t1
    tot[2] += t3 - t2
print("total create %.3f run %.3f destroy %.3f" % tuple(tot))
