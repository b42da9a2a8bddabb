"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/cnld_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
from conftest import ROOT


def _declared(header):
    src = open(os.path.join(ROOT, 'include', header)).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = set(re.findall(r'\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;', src))
    return sorted(names - {'void', 'real', 'uint', 'field', 'bool'})   # function-pointer members


def test_library_exports_every_declared_symbol():
    from cnld import _lib
    L = _lib.lib()
    names = []
    for h in sorted(os.listdir(os.path.join(ROOT, 'include'))):
        if h.endswith('.h'):
            names += _declared(h)
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    from cnld import _lib
    if _lib.device_count() > 0:
        pytest.skip('GPU present')
    with pytest.raises(_lib.Cnldb200Error, match='no CPU fallback'):
        _lib.Context(np.zeros((3, 3)), np.array([[0, 1, 2]], dtype=np.uint32))
    with pytest.raises(_lib.Cnldb200Error, match='no CPU fallback'):
        _lib.lrdecomp(np.eye(4))
