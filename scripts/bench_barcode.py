#!/usr/bin/env python
"""Radius-r ball sizes of the Miller-Schupp presentations (barcode_analysis/5_steps_neibourhoods workload):
GPU engine (csrc/ball.cu) vs the reference's own C++ tool compiled into oracle/_ref, same inputs, results
compared.  One JSON line.
    python scripts/bench_barcode.py [--rows 1190] [--ref-rows 64] [--radius 5] [--classic]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from ast import literal_eval

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1190)
    ap.add_argument("--ref-rows", type=int, default=64)
    ap.add_argument("--radius", type=int, default=5)
    ap.add_argument("--classic", action="store_true")
    ap.add_argument("--simplex", type=int, default=0, help="also time the simplex dump of ac_bfs.cpp for this n")
    args = ap.parse_args()
    from ac_solver_b200.barcode import neighbourhood_sizes

    data = os.path.join(ROOT, "ac_solver_b200", "search", "miller_schupp", "data", "all_presentations.txt")
    lines = [l.strip() for l in open(data) if l.strip()][: args.rows]
    pres = [literal_eval(l) for l in lines]
    neighbourhood_sizes(pres[:2], radius=2)  # warm-up (context, module load)
    t0 = time.perf_counter()
    sizes = neighbourhood_sizes(pres, radius=args.radius, classic=args.classic)
    gpu_s = time.perf_counter() - t0
    out = {"metric": "radius-%d neighbourhood sizes, %s moves" % (args.radius, "classic" if args.classic else "prime"),
           "rows": len(pres), "gpu_seconds": gpu_s, "gpu_rows_per_s": len(pres) / gpu_s, "states_total": int(sum(sizes)),
           "gpu_states_per_s": sum(sizes) / gpu_s}
    exe = os.path.join(ROOT, "oracle", "_ref", "neibourhoods_ref")
    if os.path.exists(exe) and args.ref_rows > 0:
        step = max(1, len(lines) // args.ref_rows)
        idx = list(range(0, len(lines), step))[: args.ref_rows]
        with tempfile.TemporaryDirectory() as d:
            src, dst = os.path.join(d, "in.txt"), os.path.join(d, "out.txt")
            with open(src, "w") as f:
                f.write("\n".join(lines[i] for i in idx) + "\n")
            t0 = time.perf_counter()
            subprocess.run([exe, src, dst, str(args.radius), str(int(args.classic))], check=True, stdout=subprocess.DEVNULL)
            ref_s = time.perf_counter() - t0
            ref = [int(x) for x in open(dst).read().split()]
        mism = sum(1 for k, i in enumerate(idx) if ref[k] != sizes[i])
        ref_states = sum(ref)
        out.update({"reference": {"kind": "reference (oracle/_ref/neibourhoods_ref, the reference's C++ compiled here), 1 core",
                                  "rows": len(idx), "seconds": ref_s, "states_per_s": ref_states / ref_s,
                                  "extrapolated_seconds_all_rows": ref_s * sum(sizes) / max(ref_states, 1)},
                    "parity": {"rows_compared": len(idx), "mismatches": mism},
                    "speedup_vs_reference_1core": (sum(sizes) / gpu_s) / (ref_states / ref_s)})
    if args.simplex:
        import hashlib

        from ac_solver_b200.barcode import write_simplex_files

        n = args.simplex
        names = ("zero_simplices", "zero_filtrations", "one_simplices", "one_filtrations")
        with tempfile.TemporaryDirectory() as d:
            write_simplex_files(4, d, classic=args.classic)  # warm-up
            t0 = time.perf_counter()
            data = write_simplex_files(n, d, classic=args.classic)
            gpu_s = time.perf_counter() - t0
            mine = [hashlib.sha256(open(os.path.join(d, f"{x}_{n}"), "rb").read()).hexdigest() for x in names]
        sx = {"n": n, "vertices": int(data["n_vertices"]), "edges": int(len(data["one_filt"])), "gpu_seconds_incl_file_writes": gpu_s}
        ref_exe = os.path.join(ROOT, "oracle", "_ref", "ac_bfs_classic" if args.classic else "ac_bfs_prime")
        if os.path.exists(ref_exe):
            with tempfile.TemporaryDirectory() as d:
                t0 = time.perf_counter()
                subprocess.run([ref_exe, str(n)], cwd=d, check=True, stdout=subprocess.DEVNULL)
                sx["reference_seconds"] = time.perf_counter() - t0
                ref = [hashlib.sha256(open(os.path.join(d, f"{x}_{n}"), "rb").read()).hexdigest() for x in names]
            sx["files_byte_identical"] = mine == ref
        out["simplex"] = sx
    print(json.dumps(out))


if __name__ == "__main__":
    main()
