"""Host-side mirror (cnld.mesh / cnld.fem / cnld.impulse_response / cnld.abstract) against the
golden vectors generated from the REFERENCE's own modules, and against the oracle mesh."""
import os

import numpy as np
import pytest
from conftest import GOLDEN


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize('tag', ['1x1_r3', '2x1_r4', '1x1_r6'])
def test_fem_mirror_matches_reference_golden(tag):
    from cnld import arrays, fem, mesh
    g = np.load(os.path.join(GOLDEN, f'fem_{tag}.npz'))
    refn = int(g['refn'])
    arr = arrays.matrix_array(nelem=tuple(g['nelem']))
    am = mesh.Mesh.from_abstract(arr, refn)
    assert np.array_equal(am.triangles, g['triangles'])
    assert np.array_equal(am.vertices, g['vertices'])      # bit-exact node coordinates
    assert np.array_equal(am.on_boundary, g['on_boundary'])
    mem = arr.elements[0].membranes[0]
    K, M = fem.mem_k_matrix(mem, refn), fem.mem_m_matrix(mem, refn)
    assert _rel(K, g['K']) < 1e-12
    assert _rel(M, g['M']) < 1e-13
    assert _rel(fem.mem_b_matrix_eig(mem, refn, M, K), g['B']) < 1e-9
    assert _rel(fem.mem_eig(mem, refn)[0][:8], g['eigf']) < 1e-9
    assert _rel(fem.array_f_spmatrix(arr, refn).toarray(), g['F']) < 1e-13
    assert _rel(fem.array_avg_spmatrix(arr, refn).toarray(), g['AVG']) < 1e-13
    for i, f in enumerate(g['freqs']):
        G = fem.array_mbk_spmatrix(arr, refn, f).toarray()
        assert _rel(np.diag(G), g['mbk_diag'][i]) < 1e-9
        Ks, Ms, Bs = fem.array_kmb_spmatrices(arr, refn)
        w = 2 * np.pi * f
        assert _rel((Ks - w**2 * Ms - 1j * w * Bs).toarray(), G) < 1e-14


def test_mesh_mirror_equals_oracle_mesh(po):
    from cnld import arrays, mesh
    for nelem, refn in [((1, 1), 2), ((2, 2), 5), ((3, 1), 7)]:
        m = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=nelem), refn)
        mo = po.array_mesh(po.matrix_array(nelem=nelem), refn)
        for a in ('vertices', 'edges', 'triangles', 'triangle_edges', 'on_boundary', 'g', 'normals'):
            assert np.array_equal(getattr(m, a), getattr(mo, a)), a
        n = refn
        assert m.nvertices == (2 * n * n + 2 * n + 1) * nelem[0] * nelem[1]
        assert m.ntriangles == 4 * n * n * nelem[0] * nelem[1]
    c = mesh.circle(20e-6, 5)
    vo, eo, to, so = po.geometry_circle(20e-6, 4, 5)
    assert np.array_equal(c.vertices, vo) and np.array_equal(c.triangles, to)
    assert np.array_equal(c.triangle_edges, so)


def test_surface_utilities():
    from cnld.h2lib import (build_from_macrosurface3d_surface3d, check_surface3d, isclosed_surface3d,
                            isoriented_surface3d, new_sphere_macrosurface3d, refine_red_surface3d)
    sp = build_from_macrosurface3d_surface3d(new_sphere_macrosurface3d(), 6)
    assert (sp.vertices, sp.triangles) == (6 + 12 * 5 + 8 * 10, 8 * 36)
    assert check_surface3d(sp) == 0 and isclosed_surface3d(sp) and isoriented_surface3d(sp)
    assert abs(sp.g.sum() / 2 - 4 * np.pi) / (4 * np.pi) < 0.03
    assert np.allclose(np.linalg.norm(sp.x, axis=1), 1.0)
    assert np.all(np.einsum('ij,ij->i', sp.n, sp.x[sp.t[:, 0]]) > 0)     # outward normals
    r = refine_red_surface3d(sp)
    assert (r.vertices, r.triangles) == (sp.vertices + sp.edges, 4 * sp.triangles)
    assert check_surface3d(r) == 0 and isclosed_surface3d(r) and isoriented_surface3d(r)
    assert abs(r.g.sum() - sp.g.sum()) < 1e-12


def test_host_containers_roundtrip():
    from scipy import sparse as sps
    from cnld.h2lib import (AMatrix, AVector, SparseMatrix, add_amatrix, add_sparsematrix_amatrix,
                            addeval_amatrix_avector, addeval_sparsematrix_avector, clone_amatrix,
                            getsize_amatrix, scale_amatrix)
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    A = AMatrix.from_array(a)
    assert np.array_equal(A.a, a) and (A.rows, A.cols) == (5, 5)
    B = clone_amatrix(A)
    scale_amatrix(2 - 1j, B)
    assert np.allclose(B.a, (2 - 1j) * a)
    add_amatrix(1.0, False, A, B)
    assert np.allclose(B.a, (3 - 1j) * a)
    assert getsize_amatrix(A) >= 25 * 16
    x = rng.standard_normal(5) + 0j
    y = AVector(5)
    y.v[:] = 0
    addeval_amatrix_avector(1.0, A, AVector.from_array(x), y)
    assert np.allclose(y.v, a.T @ x)      # numpy sees the transpose of the stored matrix
    s = sps.random(5, 5, 0.4, random_state=1, format='csr') + sps.eye(5)
    S = SparseMatrix.from_array(s)
    assert S.nz == s.nnz
    y.v[:] = 0
    addeval_sparsematrix_avector(1.0, S, AVector.from_array(x), y)
    assert np.allclose(y.v, s @ x)
    Z = AMatrix(5, 5)
    add_sparsematrix_amatrix(1.0, False, S, Z)
    assert np.allclose(Z.a, s.toarray().T)


def test_abstract_json_roundtrip(tmp_path):
    from cnld import abstract, arrays
    arr = arrays.matrix_array(nelem=(2, 2))
    assert abstract.get_patch_count(arr) == 36 and abstract.get_membrane_count(arr) == 4
    p = tmp_path / 'array.json'
    abstract.dump(arr, str(p))
    back = abstract.load(str(p))
    assert abstract.dumps(back) == abstract.dumps(arr)
    assert '"__name__": "SquareCmutMembrane"' in abstract.dumps(arr)
    m0, m1 = arr.elements[0].membranes[0], arr.elements[3].membranes[0]
    assert m0._memoize() == m1._memoize()        # id / position / patches excluded
