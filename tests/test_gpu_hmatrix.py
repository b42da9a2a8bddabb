"""GPU parity of the H-matrix path: trees identical to the oracle's, dense leaves exact, ACA leaves
and the H-matvec within eps_aca of the dense operator, GMRES sweep against the dense-LU sweep."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.fixture(scope='module')
def L():
    from cnld import _lib
    assert _lib.device_count() > 0
    return _lib


@pytest.mark.parametrize('nelem,refn,clf,eta,strict', [((2, 2), 6, 16, 0.8, True), ((2, 1), 7, 8, 1.2, False)])
def test_trees_match_oracle(L, po, nelem, refn, clf, eta, strict):
    am = po.array_mesh(po.matrix_array(nelem=nelem), refn)
    ctx = L.Context(am.vertices, am.triangles)
    H = L.HMatrix(ctx, clf=clf, eta=eta, strict=strict)
    t = H.tree()
    ct = po.ClusterTree(am, clf)
    assert np.array_equal(t['perm'], ct.perm)
    assert np.array_equal(t['son0'], ct.son0) and np.array_equal(t['son1'], ct.son1)
    assert np.array_equal(t['start'], ct.start) and np.array_equal(t['size'], ct.size)
    assert np.array_equal(t['bmin'], ct.bmin) and np.array_equal(t['bmax'], ct.bmax)
    rc, cc, adm = po.build_block(ct, eta, strict)
    assert np.array_equal(t['leaf_rc'], rc) and np.array_equal(t['leaf_cc'], cc)
    assert np.array_equal(t['leaf_adm'].astype(bool), adm)
    # the leaves tile the matrix exactly once
    cover = np.zeros((am.nvertices, am.nvertices), dtype=np.int32)
    for r, c in zip(rc, cc):
        cover[np.ix_(ct.idx(r), ct.idx(c))] += 1
    assert cover.min() == 1 and cover.max() == 1
    H.close()
    ctx.close()


@pytest.mark.parametrize('eps', [1e-2, 1e-4, 1e-6])
def test_hmatrix_vs_dense_and_oracle(L, po, eps):
    am = po.array_mesh(po.matrix_array(nelem=(2, 2)), 6)
    k = 2 * np.pi * 5e6 / 1500
    ctx = L.Context(am.vertices, am.triangles)
    Z = ctx.assemble_z(k)
    H = L.HMatrix(ctx, clf=16, eta=0.8, strict=True)
    H.assemble(k, eps)
    Zh = H.todense()
    t = H.tree()
    ct = po.ClusterTree(am, 16)
    # inadmissible leaves reproduce the dense assembly (same rules, different summation order)
    for r, c, a in zip(t['leaf_rc'], t['leaf_cc'], t['leaf_adm']):
        if not a:
            ix = np.ix_(ct.idx(r), ct.idx(c))
            assert np.abs(Zh[ix] - Z[ix]).max() / np.abs(Z).max() < 1e-10
    err = _rel(Zh, Z)
    assert err < 0.2 * eps, err
    Ho = po.HMatrix(am, k, clf=16, eta=0.8, eps_aca=eps)
    # same pivoting rule -> same ranks and (up to rounding) the same low-rank factors
    ranks_o = np.array([d[0].shape[1] if isinstance(d, tuple) else 0 for _, _, d in Ho.leaves])
    assert np.array_equal(t['leaf_rank'], ranks_o)
    assert _rel(Zh, Ho.todense()) < 1e-9
    x = np.random.default_rng(0).standard_normal((am.nvertices, 3)) + 1j * np.random.default_rng(1).standard_normal((am.nvertices, 3))
    assert _rel(H.matvec(x), Zh @ x) < 1e-12
    assert _rel(H.matvec(x[:, 0], alpha=2 - 1j), (2 - 1j) * (Zh @ x[:, 0])) < 1e-12
    info = H.info()
    assert info['stored'] == Ho.size_entries
    H.close()
    ctx.close()


def test_gmres_sweep_vs_dense_lu(L, po):
    from scipy.linalg import block_diag
    arr = po.matrix_array(nelem=(2, 2))
    am = po.array_mesh(arr, 5)
    blocks = po.array_fem(arr, 5)
    ctx = L.Context(am.vertices, am.triangles)
    ctx.set_mbk(block_diag(*[b[0] for b in blocks]), block_diag(*[b[1] for b in blocks]), block_diag(*[b[2] for b in blocks]))
    ctx.set_loads(block_diag(*[b[3] for b in blocks]), block_diag(*[b[4] for b in blocks]), am.on_boundary)
    freqs = np.array([0.0, 2e6, 9e6, 30e6])
    xp, xn, _ = ctx.sweep(freqs)
    H = L.HMatrix(ctx, clf=16, eta=0.8, strict=True)
    # tight ACA + tight GMRES: the two solvers must agree to the north-star 1e-6
    xph, xnh, it, t = H.sweep(freqs, eps_aca=1e-8, tol=1e-10, restart=40, maxiter=400, batch=16)
    for i in range(len(freqs)):
        assert _rel(xnh[i], xn[i]) < 1e-6, (freqs[i], _rel(xnh[i], xn[i]))
        assert _rel(xph[i], xp[i]) < 1e-6
    assert it.max() < 400
    # the reference's default eps_aca = 1e-2: tolerance max(1e-6, eps_aca)
    xph2, xnh2, it2, _ = H.sweep(freqs[1:3], eps_aca=1e-2, tol=1e-6, batch=36)
    for i in range(2):
        assert _rel(xnh2[i], xn[i + 1]) < 1e-2
    H.close()
    ctx.close()


@pytest.mark.parametrize('nrhs', [4, 5, 31, 48, 70, 100])
def test_batched_matvec_equals_single_vector_and_dense(L, po, nrhs):
    """Batched H-matvec (nrhs >= 4: DMMA band kernels, t = V^T x then y|band) against the single-vector
    kernel and against the dense expansion of the same H-matrix, including a non-trivial alpha."""
    arr = po.matrix_array(nelem=(3, 2))
    am = po.array_mesh(arr, 7)
    ctx = L.Context(am.vertices, am.triangles)
    H = L.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=64)
    H.assemble(2 * np.pi * 6e6 / 1500, 1e-4)
    Zd = H.todense()
    n = am.nvertices
    rng = np.random.default_rng(nrhs)
    X = rng.standard_normal((n, nrhs)) + 1j * rng.standard_normal((n, nrhs))
    a = 0.5 - 2j
    Y = H.matvec(X, alpha=a)
    ref1 = np.stack([H.matvec(X[:, j], alpha=a) for j in (0, nrhs - 1)], axis=1)
    assert _rel(Y[:, [0, nrhs - 1]], ref1) < 1e-13
    assert _rel(Y, a * Zd @ X) < 1e-13
    # re-assembly at another wavenumber changes the ranks: the band lists must follow
    H.assemble(2 * np.pi * 19e6 / 1500, 1e-6)
    assert _rel(H.matvec(X), H.todense() @ X) < 1e-13
    H.close()
    ctx.close()


def test_run_distributed_hmatrix_method_single_rank():
    """The sweep driver's H-matrix method (cyclic frequency distribution, ACA + GMRES per frequency)
    against its dense-LU method on the same frequencies."""
    from cnld import arrays
    from cnld.sweep import SweepSetup, run_distributed
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(2, 2)), 5)
    freqs = np.array([0.0, 2e6, 9e6, 23e6, 41e6])
    xd, _ = run_distributed(setup, freqs, 0, 1)
    xh, _ = run_distributed(setup, freqs, 0, 1, method='hmatrix',
                            hmatrix=dict(eps_aca=1e-7, tol=1e-10, batch=16, restart=60, maxiter=400))
    assert xd.shape == xh.shape == (5, setup.npatch, setup.npatch)
    for i in range(len(freqs)):
        assert _rel(xh[i], xd[i]) < 1e-5, (freqs[i], _rel(xh[i], xd[i]))


def test_leaf_translation_classes(L, po):
    """Leaves that are translates of an earlier leaf share its data: the expanded matrix must equal the
    one assembled leaf by leaf, and a real fraction of the leaves must be shared on a membrane array."""
    arr = po.matrix_array(nelem=(4, 4))    # 2^k membranes per side: the bisection planes fall between membranes
    am = po.array_mesh(arr, 5)
    k = 2 * np.pi * 8e6 / 1500
    dense = {}
    for on in (False, True):
        ctx = L.Context(am.vertices, am.triangles)
        ctx.set_translation_classes(on)
        H = L.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=64)
        H.assemble(k, 1e-9)
        sh, info = H.sharing(), H.info()
        assert sh['assembled_admissible'] + sh['assembled_inadmissible'] + sh['shared'] == info['leaves']
        assert (sh['shared'] > info['leaves'] // 10) if on else (sh['shared'] == 0), (on, sh, info['leaves'])
        dense[on] = H.todense()
        if on:
            x = np.random.default_rng(0).standard_normal((am.nvertices, 7)) + 0j
            assert _rel(H.matvec(x), dense[on] @ x) < 1e-13
        H.close()
        ctx.close()
    assert _rel(dense[True], dense[False]) < 1e-8
    Z = po.assemble_z(am, k)
    assert _rel(dense[True], Z) < 1e-7
