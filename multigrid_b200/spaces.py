"""Observation / action space descriptions.

The reference describes its spaces with `gymnasium.spaces` (multigrid/core/agent.py:85-97,
multigrid/base.py:209-227). gymnasium is an optional dependency here: when it is importable its
classes are used, otherwise these minimal stand-ins with the same attribute names (`shape`,
`dtype`, `n`, `low`, `high`, dict access) are, so adapters that only read those keep working.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium.spaces import Box, Dict, Discrete, Text  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)

        def __repr__(self):
            return f"{type(self).__name__}(shape={self.shape}, dtype={self.dtype})"

    class Discrete(Space):
        def __init__(self, n: int):
            super().__init__((), np.int64)
            self.n = int(n)

        def contains(self, x) -> bool:
            return 0 <= int(x) < self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.int64):
            super().__init__(shape, dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

    class Text(Space):
        def __init__(self, max_length: int = 256):
            super().__init__((), None)
            self.max_length = max_length

    class Dict(dict):
        """A real dict subclass, like gymnasium's (the reference does `dict(space)`)."""

        def __init__(self, spaces=None, **kw):
            super().__init__(spaces or {}, **kw)

        @property
        def spaces(self):
            return self
