"""Small driver for ncu: the bucket-batched greedy kernel on a slice of the Miller-Schupp rows.
    python scripts/profile_greedy.py [budget=100000] [rows=96] [mrl_group=36]"""
import os
import sys
from ast import literal_eval

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ac_solver_b200.search.greedy import greedy_search_batch  # noqa: E402

budget = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
nrows = int(sys.argv[2]) if len(sys.argv) > 2 else 96
width = 2 * (int(sys.argv[3]) if len(sys.argv) > 3 else 36)
data = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ac_solver_b200", "search", "miller_schupp", "data",
                    "all_presentations.txt")
rows = [np.array(literal_eval(l), dtype=np.int8) for l in open(data) if l.strip()]
rows = [r for r in rows[533:] if r.size == width][:nrows]  # unsolved rows of one max_relator_length group
res = greedy_search_batch(np.stack(rows), budget)
print(len(rows), "searches at budget", budget, "solved:", sum(1 for r in res if r[0]))
