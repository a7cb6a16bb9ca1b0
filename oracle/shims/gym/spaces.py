"""TEST INFRASTRUCTURE ONLY -- minimal gym.spaces (rad_search_env.py:355, 364 use them as class-level defaults)."""
import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape or ()), dtype
