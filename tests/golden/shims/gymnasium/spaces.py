import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = None
        if seed is not None:
            self.seed(seed)

    @property
    def shape(self):
        return self._shape

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    def seed(self, seed=None):
        self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        return seed

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        super().__init__(shape, dtype, seed)
        self.low, self.high = low, high


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64, seed)
        self.n, self.start = int(n), int(start)

    def sample(self):
        return int(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)


class Dict(Space, dict):
    def __init__(self, spaces=None, seed=None, **kw):
        dict.__init__(self, spaces or {}, **kw)
        Space.__init__(self, None, None, seed)

    @property
    def spaces(self):
        return self

    def sample(self):
        return {k: s.sample() for k, s in self.items()}
