"""cnld.util -- the helpers of /root/reference/cnld/util.py the hot path touches."""
import functools
from copy import deepcopy

import numpy as np


def memoize(func, maxsize=20):
    """Mirror of util.memoize (util.py:549-588): objects with `_memoize()` contribute that key,
    ndarrays their bytes, floats are rounded to 18 digits; results are deep-copied."""
    def make_hashable(obj):
        if hasattr(obj, '_memoize'):
            return obj._memoize()
        try:
            hash(obj)
        except TypeError:
            if isinstance(obj, np.ndarray):
                return obj.tobytes()
            return str(obj)
        if isinstance(obj, float):
            return round(obj, 18)
        return obj

    func.memo = {}

    @functools.wraps(func)
    def decorator(*args, **kwargs):
        key = (tuple(make_hashable(a) for a in args),
               tuple((k, make_hashable(v)) for k, v in sorted(kwargs.items())))
        if key not in func.memo:
            if len(func.memo) > maxsize:
                return func(*args, **kwargs)
            func.memo[key] = func(*args, **kwargs)
        return deepcopy(func.memo[key])

    return decorator


def meshview(v1, v2, v3, mode='cartesian', as_list=True):
    """util.py:23-44, cartesian mode."""
    x, y, z = np.meshgrid(v1, v2, v3, indexing='ij')
    if as_list:
        return np.c_[x.ravel('F'), y.ravel('F'), z.ravel('F')]
    return x, y, z
