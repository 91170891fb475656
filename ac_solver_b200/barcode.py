"""The reference's `barcode_analysis` explorations on the GPU (csrc/ball.cu).

State model of the C++ tools (not of the Python package): an unordered pair of relators kept
sorted, no length cap, full free reduction after every move, 12 "prime" or 14 "classic" moves.

* ``neighbourhood_size`` -- barcode_analysis/5_steps_neibourhoods/neibourhoods.cpp:18-54: number of
  presentations within ``radius`` moves (known answers in its README.txt:24-43).
* ``simplex_data`` / ``write_simplex_files`` -- barcode_analysis/simplex_data_generation/*/ac_bfs.cpp:
  the graph of all presentations of total length <= n reachable from <a, b>: 0-simplices (vertices in
  discovery order), their filtration (total length), 1-simplices (cn < cc) and their filtration (the
  larger total length), in the files the reference writes.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


def _relators(presentation):
    p = np.asarray(presentation).astype(np.int64).ravel()
    if p.size % 2:
        raise ValueError("a presentation is a list of even length (two zero-padded relators)")
    h = p.size // 2
    r1, r2 = p[:h][p[:h] != 0], p[h:][p[h:] != 0]  # neibourhoods.cpp:80-90 drops the zeros of each half
    if len(r1) and np.abs(r1).max() > 2 or len(r2) and np.abs(r2).max() > 2:
        raise ValueError("letters must be +-1, +-2")
    return r1.astype(np.int8), r2.astype(np.int8)


def _explore(r1, r2, radius, size_cap, classic, max_nodes, want_nodes=False, want_edges=False, device=None):
    L = _lib.lib()
    dev = _lib.default_device() if device is None else device
    letters = np.ascontiguousarray(np.concatenate([r1, r2]), dtype=np.int8)
    if letters.size == 0:
        letters = np.zeros(1, np.int8)
    n_nodes, n_edges = C.c_int64(0), C.c_int64(0)
    sizes = np.zeros(max_nodes + 16, np.uint16) if want_nodes else None
    levels = np.zeros(max_nodes + 16, np.uint8) if want_nodes else None
    cap_edges = (max_nodes + 16) * (14 if classic else 12) if want_edges else 0
    edges = np.zeros((cap_edges, 3), np.uint32) if want_edges else None
    _lib.check(L.acs_ball_explore(dev, letters.ctypes.data, len(r1), len(r2), int(radius), int(size_cap), int(bool(classic)),
                                  int(max_nodes), C.byref(n_nodes), sizes.ctypes.data if want_nodes else None,
                                  levels.ctypes.data if want_nodes else None, len(sizes) if want_nodes else 0,
                                  edges.ctypes.data if want_edges else None, cap_edges, C.byref(n_edges)))
    out = {"n_nodes": int(n_nodes.value)}
    if want_nodes:
        out["sizes"], out["levels"] = sizes[: n_nodes.value], levels[: n_nodes.value]
    if want_edges:
        out["edges"] = edges[: n_edges.value]
    return out


def neighbourhood_size(presentation, radius=5, classic=False, max_nodes=4_000_000, device=None):
    """neibourhoods.cpp ``neibourhood(start, radius, classic)``."""
    r1, r2 = _relators(presentation)
    return _explore(r1, r2, radius, 0, classic, max_nodes, device=device)["n_nodes"]


def _ball_batch(rels, radius, classic, max_nodes, device):
    L = _lib.lib()
    letters = np.ascontiguousarray(np.concatenate([np.concatenate(r) for r in rels] + [np.zeros(1, np.int8)]), dtype=np.int8)
    off = np.zeros(2 * len(rels) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for r in rels for x in r])
    counts = np.zeros(len(rels), np.int64)
    rc = L.acs_ball_sizes(device, letters.ctypes.data, off.ctypes.data, len(rels), int(radius), int(bool(classic)),
                          int(max_nodes), counts.ctypes.data)
    return rc, counts


def neighbourhood_sizes(presentations, radius=5, classic=False, batch_roots=128, batch_bytes=12 << 30, device=None):
    """neibourhoods.cpp ``read_do_and_write`` (:58-103) for a list of presentations: the balls of a batch of
    presentations are explored TOGETHER (one breadth-first run over (root, state) pairs, csrc/ball.cu), batches
    sized so that the node store stays below ``batch_bytes``; a batch that outgrows its store is split."""
    rels = [_relators(p) for p in presentations]
    dev = _lib.default_device() if device is None else device
    out = np.zeros(len(rels), np.int64)

    def stride(batch):  # the engine's letter stride for this batch (ball.cu: a relator at most doubles per move)
        a = b = max(1, max(max(len(r[0]), len(r[1])) for r in batch))
        for _ in range(max(radius, 0)):
            a, b = max(a + b, max(a, b) + 2), max(a, b)
        return (min(max(max(a, b) + 2, 4), 512) + 3) // 4 * 4

    def run(lo, hi):
        batch = rels[lo:hi]
        max_nodes = max(int(batch_bytes // (2 * stride(batch) + 16)), len(batch) + 1024)
        rc, counts = _ball_batch(batch, radius, classic, max_nodes, dev)
        if rc == _lib.ACS_ERR_NOMEM and hi - lo > 1:  # more states than the store holds: halve the batch
            mid = (lo + hi) // 2
            run(lo, mid)
            run(mid, hi)
            return
        _lib.check(rc)
        out[lo:hi] = counts

    for lo in range(0, len(rels), batch_roots):
        run(lo, min(lo + batch_roots, len(rels)))
    return [int(x) for x in out]


def simplex_data(n, classic=False, max_nodes=50_000_000, device=None):
    """ac_bfs.cpp: everything it writes, as arrays: ``zero_filt`` [V], ``one_simplices`` [E,2], ``one_filt`` [E]."""
    r = _explore(np.array([1], np.int8), np.array([2], np.int8), -1, int(n), classic, max_nodes, want_nodes=True,
                 want_edges=True, device=device)
    return {"n_vertices": r["n_nodes"], "zero_filt": r["sizes"].astype(np.int64),
            "one_simplices": r["edges"][:, :2].astype(np.int64), "one_filt": r["edges"][:, 2].astype(np.int64)}


def write_simplex_files(n, outdir=".", classic=False, **kw):
    """The four files of ac_bfs.cpp:21-35,84-91, byte for byte in the reference's format."""
    d = simplex_data(n, classic, **kw)
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, f"zero_simplices_{n}"), "w") as f:
        f.write('{"0-simplices":[' + "".join(f"[{i}]," for i in range(d["n_vertices"])) + "[]]}")
    with open(os.path.join(outdir, f"zero_filtrations_{n}"), "w") as f:
        f.write('{"0-filt":[' + "".join(f"{int(v)}," for v in d["zero_filt"]) + "-5]}")
    with open(os.path.join(outdir, f"one_simplices_{n}"), "w") as f:
        f.write('{"1-simplices":[' + "".join(f"[{int(a)},{int(b)}]," for a, b in d["one_simplices"]) + "[]]}")
    with open(os.path.join(outdir, f"one_filtrations_{n}"), "w") as f:
        f.write('{"1-filt":[' + "".join(f"{int(v)}," for v in d["one_filt"]) + "-5]}")
    return d
