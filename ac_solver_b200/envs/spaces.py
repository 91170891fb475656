"""Minimal stand-ins for ``gymnasium.spaces`` (gymnasium is optional; the reference uses only
``Box(low, high, dtype)`` and ``Discrete(n)``, envs/ac_env.py:68-77)."""

from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    from gymnasium import Env  # type: ignore
    from gymnasium.spaces import Box, Discrete  # type: ignore
except Exception:  # gymnasium absent: tiny local equivalents

    class Env:  # noqa: D401
        """Base class placeholder."""

    class Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class Box:
        def __init__(self, low, high, dtype=None):
            self.low = np.asarray(low)
            self.high = np.asarray(high)
            self.dtype = np.dtype(dtype) if dtype is not None else self.low.dtype
            self.shape = self.low.shape

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())
