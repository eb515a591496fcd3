"""`Discrete` / `Box` for the env's spaces: gym's or gymnasium's when importable, else minimal stand-ins
with the same attributes (tetris_env.py:377-392 only constructs them)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - neither package exists in the build image
    from gym.spaces import Box, Discrete  # type: ignore
except Exception:  # noqa: BLE001
    try:
        from gymnasium.spaces import Box, Discrete  # type: ignore
    except Exception:  # noqa: BLE001

        class Discrete:
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.int64
                self._rng = np.random.default_rng()

            def sample(self):
                return int(self._rng.integers(self.n))

            def contains(self, x):
                return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

            def __repr__(self):
                return f"Discrete({self.n})"

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

            def contains(self, x):
                return np.asarray(x).shape == self.shape

            def __repr__(self):
                return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
