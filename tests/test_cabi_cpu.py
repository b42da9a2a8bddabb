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


def test_new_entry_points_refuse_without_a_gpu_and_host_parts_work():
    """Round-2 entry points: the compute ones fail loudly without a CUDA device; the host-only ones
    (pressure-plan chunking, world-size-1 communicator, coarse-space modes) run here."""
    from cnld import _lib, arrays, mesh
    am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(2, 1)), 4)
    info, order = _lib.pres_plan_stats(am.vertices, am.triangles)
    nt = len(am.triangles)
    assert sorted(order.tolist()) == list(range(nt))                    # every triangle in exactly one chunk
    assert 0 < info['chunks'] <= nt and info['max_nodes'] <= 24 and info['max_triangles'] <= 32
    assert info['node_entries'] >= len(am.vertices) and info['untouched_vertices'] == 0
    assert info['accumulating_entries'] == info['node_entries'] - len(am.vertices)   # first touch stores, later ones add
    c = _lib.Comm(0, 1, 0)                                              # single rank: no NCCL, no GPU needed
    a = np.arange(5.0)
    assert c.bcast(a) is a and c.gather(a).shape == (1, 5) and c.max(2.5) == 2.5 and c.bcast_object([1, 2]) == [1, 2]
    c.barrier()
    c.close()
    with pytest.raises(_lib.Cnldb200Error):
        _lib.Comm(2, 2, 0)                                              # rank outside the world
    if _lib.device_count() == 0:
        with pytest.raises(_lib.Cnldb200Error, match='no CPU fallback'):
            _lib.fft_to_fir(np.linspace(0, 1e6, 8), np.ones((2, 8), dtype=np.complex128))
        with pytest.raises(_lib.Cnldb200Error, match='no CPU fallback'):
            _lib.FirTable(np.ones((2, 2, 8)))


def test_membrane_modes_coarse_space():
    """Coarse space of the two-level GMRES preconditioner: per membrane the lowest K v = lambda M v modes on the
    free nodes, zero on clamped nodes and on the other membranes, identical membranes share the modes."""
    from cnld import _lib, arrays
    from cnld.sweep import SweepSetup
    s = SweepSetup.from_abstract(arrays.matrix_array(nelem=(2, 1)), 4)

    class Ctx:
        pass
    from scipy import sparse as sps
    c = Ctx()
    c.nv, c.triangles = s.nnodes, s.triangles
    pat = (abs(s.K) + abs(s.M) + abs(s.B)).tocsr()
    pat.sort_indices()
    rows = np.repeat(np.arange(c.nv), np.diff(pat.indptr))
    c._mbk = (pat.indptr.astype(np.int64), pat.indices.astype(np.uint32)) + tuple(
        np.asarray(a[rows, pat.indices]).ravel() for a in (s.K, s.M, s.B))
    c._loads = (None,) * 6 + (s.on_boundary,)
    W, f = _lib.membrane_modes(c, 6)
    n1 = s.nnodes // 2
    assert W.shape == (s.nnodes, 6) and np.all(np.diff(f) >= -1e-6) and 5e6 < f[0] < 50e6
    assert np.all(W[s.on_boundary.astype(bool)] == 0)
    assert np.allclose(np.abs(W[:n1]), np.abs(W[n1:]))                  # same modes on both membranes
    K, M = s.K.toarray()[:n1, :n1], s.M.toarray()[:n1, :n1]
    v = W[:n1, 0]
    lam = (v @ K @ v) / (v @ M @ v)
    assert abs(np.sqrt(lam) / (2 * np.pi) - f[0]) / f[0] < 1e-8         # Rayleigh quotient = reported frequency
