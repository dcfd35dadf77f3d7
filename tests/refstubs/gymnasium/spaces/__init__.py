import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(int(s) for s in shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(shape if shape is not None else np.shape(low), dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n, self.start = int(n), int(start)


class MultiBinary(Space):
    def __init__(self, n):
        super().__init__(tuple(np.atleast_1d(n)), np.int8)
        self.n = n


class Dict(Space, dict):
    def __init__(self, spaces=None, **kw):
        Space.__init__(self)
        dict.__init__(self, spaces or {}, **kw)

    @property
    def spaces(self):
        return self
