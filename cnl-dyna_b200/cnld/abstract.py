"""
cnld.abstract -- array description types (mirror of /root/reference/cnld/abstract.py:218-296).

The reference builds these with the third-party `namedlist` package (absent here); the same
field names, defaults, JSON layout (`__name__` tags, abstract.py:32-76,126-160) and helper
functions are provided on a small built-in record type, so array JSON files written by
cnl-dyna load unchanged.
"""
import json
from collections import OrderedDict
from collections.abc import Mapping
from copy import deepcopy

__all__ = ['SquareCmutMembrane', 'CircularCmutMembrane', 'Patch', 'Element', 'Array', 'dump', 'dumps',
           'load', 'loads', 'register_type', 'get_patch_count', 'get_membrane_count',
           'get_element_count', 'get_patches_from_array', 'get_membranes_from_array',
           'get_elements_from_array']

_TYPES = {}


class FACTORY:
    def __init__(self, f):
        self.f = f


def register_type(name, fields, excl=None):
    """Record type with defaults; `excl` fields are left out of the memoisation key
    (abstract.py:163-183)."""
    fields = OrderedDict(fields)
    excl = list(excl or [])

    def __init__(self, *args, **kwargs):
        for n, d in fields.items():
            setattr(self, n, d.f() if isinstance(d, FACTORY) else deepcopy(d))
        for n, a in zip(fields, args):
            setattr(self, n, a)
        for n, a in kwargs.items():
            if n not in fields:
                raise TypeError(f'{name}: unexpected field {n!r}')
            setattr(self, n, a)

    def _asdict(self):
        return OrderedDict((n, getattr(self, n)) for n in fields)

    def _memoize(self):
        d = self._asdict()
        for key in excl:
            d.pop(key, None)
        return json.dumps(_tagged(d))

    ns = dict(__init__=__init__, _asdict=_asdict, _fields=tuple(fields), _memoize=_memoize,
              _memoize_excl=excl, __getitem__=lambda self, k: getattr(self, k),
              keys=lambda self: list(fields), __contains__=lambda self, k: k in fields,
              __repr__=lambda self: f'{name}({", ".join(f"{k}={getattr(self, k)!r}" for k in fields)})',
              copy=lambda self: loads(dumps(self)), __copy__=lambda self: loads(dumps(self)))
    cls = type(name, (), ns)
    _TYPES[name] = cls
    return cls


def _is_abstract(obj):
    return hasattr(obj, '_asdict')


def _tagged(obj):
    if _is_abstract(obj) or isinstance(obj, Mapping):
        items = obj._asdict().items() if _is_abstract(obj) else obj.items()
        d = {'__name__': type(obj).__name__}
        for k, v in items:
            d[k] = _tagged(v)
        return d
    if isinstance(obj, (list, tuple)):
        return [_tagged(v) for v in obj]
    return obj


def _untagged(js):
    if isinstance(js, dict):
        js = dict(js)
        name = js.pop('__name__', None)
        d = {k: _untagged(v) for k, v in js.items()}
        if name in _TYPES:
            return _TYPES[name](**d)
        return d
    if isinstance(js, (list, tuple)):
        return [_untagged(v) for v in js]
    return js


def dump(obj, fp, indent=1, mode='w+', *args, **kwargs):
    with open(fp, mode) as f:
        json.dump(_tagged(obj), f, indent=indent, *args, **kwargs)


def dumps(obj, indent=1, *args, **kwargs):
    return json.dumps(_tagged(obj), indent=indent, *args, **kwargs)


def load(fp, *args, **kwargs):
    with open(fp, 'r') as f:
        return _untagged(json.load(f, *args, **kwargs))


def loads(s, *args, **kwargs):
    return _untagged(json.loads(s, *args, **kwargs))


_mem_common = [('thickness', (2e-6,)), ('density', (2040,)), ('y_modulus', (110e9,)), ('p_ratio', (0.22,)),
               ('isolation', 200e-9), ('permittivity', 6.3), ('gap', 50e-9), ('damping_freq_a', 0),
               ('damping_freq_b', 0), ('damping_mode_a', 0), ('damping_mode_b', 4),
               ('damping_ratio_a', 0.0), ('damping_ratio_b', 0.0), ('patches', FACTORY(list))]
SquareCmutMembrane = register_type(
    'SquareCmutMembrane',
    [('id', None), ('position', None), ('shape', 'square'), ('length_x', 40e-6), ('length_y', 40e-6),
     ('electrode_x', 40e-6), ('electrode_y', 40e-6)] + _mem_common, excl=['id', 'position', 'patches'])
CircularCmutMembrane = register_type(
    'CircularCmutMembrane',
    [('id', None), ('position', None), ('shape', 'circle'), ('radius', 40e-6 / 2),
     ('electrode_r', 40e-6 / 2)] + _mem_common, excl=['id', 'position', 'patches'])
Patch = register_type('Patch', [('id', None), ('position', None), ('length_x', None), ('length_y', None),
                                ('radius_min', None), ('radius_max', None), ('theta_min', None),
                                ('theta_max', None), ('area', None)])
Element = register_type('Element', [('id', None), ('position', None), ('kind', None), ('active', True),
                                    ('apodization', 1), ('delay', 0), ('dc_bias', 0),
                                    ('membranes', FACTORY(list))])
Array = register_type('Array', [('id', None), ('position', None), ('delay', 0),
                                ('elements', FACTORY(list))])


def get_elements_from_array(array, kind='both'):
    return list(array.elements)


def get_membranes_from_array(array):
    return [m for e in array.elements for m in e.membranes]


def get_patches_from_array(array):
    return [p for e in array.elements for m in e.membranes for p in m.patches]


def get_element_count(a, kind=None):
    return len(a.elements)


def get_membrane_count(a):
    return len(get_membranes_from_array(a))


def get_patch_count(a):
    return len(get_patches_from_array(a))
