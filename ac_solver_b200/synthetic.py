"""Synthetic presentations for benchmarks and parity tests.

Rows are two independent words over {x, x^-1, y, y^-1}: length uniform in [min_len, max_len],
uniform over freely AND cyclically reduced words of that length, rows with r0 == r1^{+-1}
regenerated (those make the reference raise on a concatenation).  Pure numpy, vectorised."""

from __future__ import annotations

import numpy as np

_LETTER = np.array([2, 1, -2, -1], dtype=np.int8)  # 2-bit code -> letter; inverse = code ^ 2


def _random_codes(rng, n, mrl, min_len, max_len, cyclic):
    length = rng.integers(min_len, max_len + 1, size=n)
    codes = np.zeros((n, mrl), dtype=np.int8)
    codes[:, 0] = rng.integers(0, 4, size=n)
    for t in range(1, mrl):
        step = rng.integers(1, 4, size=n)  # any letter except the inverse of the previous one
        codes[:, t] = ((codes[:, t - 1] ^ 2) + step) % 4
    if cyclic:
        rows = np.arange(n)
        last = length - 1
        for _ in range(64):
            bad = (length > 1) & (codes[rows, last] == (codes[:, 0] ^ 2))
            if not bad.any():
                break
            idx = np.flatnonzero(bad)
            prev = codes[idx, np.maximum(last[idx] - 1, 0)]
            codes[idx, last[idx]] = ((prev ^ 2) + rng.integers(1, 4, size=idx.size)) % 4
        else:  # pragma: no cover
            raise RuntimeError("could not draw cyclically reduced words")
    return codes, length


def random_presentations(n, mrl, seed=0, min_len=1, max_len=None, cyclic=True):
    """-> int8 [n, 2*mrl] of right-padded reduced word pairs."""
    rng = np.random.default_rng(seed)
    max_len = mrl if max_len is None else max_len
    out = np.zeros((n, 2 * mrl), dtype=np.int8)
    pos = np.arange(mrl)[None, :]
    for h in range(2):
        codes, length = _random_codes(rng, n, mrl, min_len, max_len, cyclic)
        out[:, h * mrl : (h + 1) * mrl] = np.where(pos < length[:, None], _LETTER[codes], 0)
    # regenerate the rare rows with r0 == r1 or r0 == r1^-1
    for _ in range(16):
        r0, r1 = out[:, :mrl], out[:, mrl:]
        l1 = np.count_nonzero(r1, axis=1)
        inv = np.zeros_like(r1)
        for k in np.flatnonzero(np.count_nonzero(r0, axis=1) == l1):
            inv[k, : l1[k]] = -r1[k, : l1[k]][::-1]
        bad = (r0 == r1).all(axis=1) | ((r0 == inv).all(axis=1) & (l1 > 0))
        if not bad.any():
            break
        idx = np.flatnonzero(bad)
        codes, length = _random_codes(rng, idx.size, mrl, min_len, max_len, cyclic)
        out[idx, mrl:] = np.where(pos < length[:, None], _LETTER[codes], 0)
    return out


def random_actions(n, seed=1):
    return np.random.default_rng(seed).integers(0, 12, size=n).astype(np.uint8)
