"""cnld.util -- the helpers of /root/reference/cnld/util.py the hot path touches."""
import functools
from copy import deepcopy

import numpy as np


class _Memo:
    """Call cache behind `memoize`.  Same observable contract as the reference's decorator
    (util.py:549-588) -- a bounded `memo` dict on the wrapped function, argument keys built from
    `_memoize()` of abstract objects / array bytes / floats rounded to 18 digits, every hit returned as a
    deep copy -- written as a small class around one key function."""

    def __init__(self, func, maxsize):
        self.func, self.maxsize, self.memo = func, maxsize, {}
        func.memo = self.memo                      # the reference exposes the cache as func.memo

    @staticmethod
    def key_of(value):
        memo_key = getattr(value, '_memoize', None)
        if callable(memo_key):
            return memo_key()
        if isinstance(value, np.ndarray):
            return (value.shape, value.dtype.str, value.tobytes())
        if isinstance(value, float):
            return round(value, 18)
        try:
            hash(value)
            return value
        except TypeError:
            return str(value)

    def __call__(self, *args, **kwargs):
        key = (tuple(map(self.key_of, args)), tuple(sorted((k, self.key_of(v)) for k, v in kwargs.items())))
        hit = self.memo.get(key)
        if hit is None:
            hit = self.func(*args, **kwargs)
            if len(self.memo) <= self.maxsize:     # a full cache computes without storing
                self.memo[key] = hit
        return deepcopy(hit)


def memoize(func, maxsize=20):
    cache = _Memo(func, maxsize)

    @functools.wraps(func)
    def cached(*args, **kwargs):
        return cache(*args, **kwargs)
    return cached


def meshview(v1, v2, v3, mode='cartesian', as_list=True):
    """util.py:23-44, cartesian mode."""
    x, y, z = np.meshgrid(v1, v2, v3, indexing='ij')
    if as_list:
        return np.c_[x.ravel('F'), y.ravel('F'), z.ravel('F')]
    return x, y, z
