#!/usr/bin/env python
"""BASELINE config 3: greedy_search over all 1190 Miller-Schupp presentations, batched across
presentations (one warp per search, one launch per max_relator_length group).

    python scripts/greedy_sweep.py [--budget 1000000] [--max-group 400]

Checks the outcome against the data shipped with the reference: rows 0..532 are solved with
exactly the stored paths (greedy_search_paths.txt, action+1 convention), rows 533.. fail."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ac_solver_b200.search.greedy import greedy_search_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=int, default=1_000_000)
    ap.add_argument("--max-group", type=int, default=400)
    args = ap.parse_args()
    ms = np.load(os.path.join(ROOT, "tests", "golden", "miller_schupp.npz"))
    offs, flat = ms["greedy_path_offsets"], ms["greedy_path_flat"]

    def row(k):
        m = int(ms["mrl"][k])
        p = ms["presentations36"][k]
        return np.concatenate([p[:m], p[36 : 36 + m]]).astype(np.int8)

    n_solved = n_path_ok = visited = expanded = 0
    wrong = []
    t0 = time.perf_counter()
    dev_s = 0.0
    for mrl in sorted(set(int(m) for m in ms["mrl"])):
        rows = [k for k in range(len(ms["mrl"])) if ms["mrl"][k] == mrl]
        for i in range(0, len(rows), args.max_group):
            part = rows[i : i + args.max_group]
            out = greedy_search_batch(np.stack([row(k) for k in part]), args.budget, path_cap=4096)
            dev_s += out[0][2]["seconds_device"]
            for k, (solved, path, info) in zip(part, out):
                visited += info["n_visited"]
                expanded += info["n_expanded"]
                n_solved += solved
                if k < 533:
                    exp = [(int(a) - 1, int(l)) for a, l in flat[offs[k] : offs[k + 1]]]
                    if solved and path == exp:
                        n_path_ok += 1
                    else:
                        wrong.append(k)
                elif solved:
                    wrong.append(k)
            print(f"mrl {mrl}: {len(part)} searches done, solved so far {n_solved}", file=sys.stderr, flush=True)
    wall = time.perf_counter() - t0
    # CPU baseline beside it: the C oracle on sampled unsolved rows, one core
    from oracle import oracle as O

    c0 = time.perf_counter()
    cpu_visited = sum(O.greedy_search(row(k), 100_000)[2]["n_visited"] for k in (533, 700, 900, 1189))
    cpu_rate = cpu_visited / (time.perf_counter() - c0)
    print(f"cpu baseline (C oracle, 1 core, 4 unsolved rows at budget 1e5): {cpu_rate:.3e} visited/s", file=sys.stderr)
    print(json.dumps({"config": "greedy_search over 1190 Miller-Schupp presentations", "budget": args.budget,
                      "solved": n_solved, "stored_paths_reproduced": n_path_ok, "mismatching_rows": wrong[:20],
                      "visited_total": visited, "expanded_total": expanded, "seconds_wall": wall,
                      "seconds_device": dev_s, "visited_per_s_device": visited / max(dev_s, 1e-9)}))


if __name__ == "__main__":
    main()
