import numpy as np


class Array:
    def __init__(self, shape, dtype, name=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.name = name


class BoundedArray(Array):
    def __init__(self, shape, dtype, minimum, maximum, name=None):
        super().__init__(shape, dtype, name)
        self.minimum = np.array(minimum, dtype=self.dtype)
        self.maximum = np.array(maximum, dtype=self.dtype)


class DiscreteArray(BoundedArray):
    def __init__(self, num_values, dtype=np.int32, name=None):
        super().__init__((), dtype, 0, num_values - 1, name)
        self.num_values = int(num_values)
