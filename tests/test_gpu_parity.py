"""GPU parity: the CUDA path (through the C ABI, via cnld._lib) against the CPU oracle on the
same inputs.  Tolerances: Z entries 1e-10 relative to max|Z| (same quadrature rules, only
rounding / libm differences); displacements, patch averages and pressure 1e-6 relative l2
(the north-star tolerance; observed values are far smaller)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.fixture(scope='module')
def L():
    from cnld import _lib
    assert _lib.device_count() > 0, 'no CUDA device: libcnld_b200 has no CPU fallback'
    return _lib


@pytest.mark.parametrize('n,nrhs', [(37, 1), (64, 3), (200, 5), (517, 144)])
def test_lu_and_solve_vs_oracle(L, po, n, nrhs):
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + n * 0.7 * np.eye(n)
    B = rng.standard_normal((n, nrhs)) + 1j * rng.standard_normal((n, nrhs))
    LU = L.lrdecomp(G)
    LUo, st = po.lrdecomp(G)
    assert st == 0
    assert np.abs(LU - LUo).max() / np.abs(LUo).max() < 1e-12
    X = L.lrsolve(LU, B)
    Xo = np.stack([po.lrsolve(LUo, B[:, j]) for j in range(nrhs)], axis=1)
    assert _rel(X, Xo) < 1e-12
    assert _rel(G @ X, B) < 1e-11


@pytest.mark.parametrize('n', [37, 64, 200, 517, 1300])
def test_symmetric_lu_equals_general_lu(L, n):
    """lrdecomp_sym reads only the lower triangle and returns the same L and R as lrdecomp on the mirrored
    matrix (R = D L^T is formed by transposition instead of block-row products)."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = A + A.T + n * 1.5 * np.eye(n)
    LU = L.lrdecomp(A)
    LUs = L.lrdecomp_sym(np.tril(A) + 1e3 * np.triu(np.ones((n, n)), 1))   # garbage above the diagonal is ignored
    assert np.abs(LU - LUs).max() / np.abs(LU).max() < 1e-13


def test_lu_zero_pivot_raises(L):
    G = np.eye(70, dtype=np.complex128)
    G[5, 5] = 0.0
    with pytest.raises(RuntimeError, match='failed to calculate LU decomposition'):
        L.lrdecomp(G)


@pytest.mark.parametrize('nelem,refn,k', [((1, 1), 3, 2 * np.pi * 1e6 / 1500), ((2, 1), 4, 3.1e4),
                                          ((1, 1), 5, 0.0), ((2, 1), 3, 2.0e4 + 3.0e3j)])
def test_assemble_z_vs_oracle(L, po, nelem, refn, k):
    arr = po.matrix_array(nelem=nelem)
    am = po.array_mesh(arr, refn)
    ctx = L.Context(am.vertices, am.triangles)
    g, nrm = ctx.geometry()
    assert np.allclose(g, am.g, rtol=1e-14)
    assert np.allclose(nrm, am.normals, atol=1e-14)
    Z = ctx.assemble_z(k)
    Zo = po.assemble_z(am, k)
    assert np.abs(Z - Zo).max() / np.abs(Zo).max() < 1e-10
    ctx.close()


@pytest.mark.parametrize('q_reg,q_sing', [(1, 2), (3, 3), (4, 5)])
def test_assemble_z_other_orders(L, po, q_reg, q_sing):
    am = po.array_mesh(po.matrix_array(nelem=(1, 1)), 3)
    ctx = L.Context(am.vertices, am.triangles, q_reg=q_reg, q_sing=q_sing)
    Z = ctx.assemble_z(5.0e4)
    Zo = po.assemble_z(am, 5.0e4, q_reg, q_sing)
    assert np.abs(Z - Zo).max() / np.abs(Zo).max() < 1e-10
    ctx.close()


def _setup_sweep(L, po, nelem, refn):
    from scipy.linalg import block_diag
    arr = po.matrix_array(nelem=nelem)
    am = po.array_mesh(arr, refn)
    blocks = po.array_fem(arr, refn)
    ctx = L.Context(am.vertices, am.triangles)
    ctx.set_mbk(block_diag(*[b[0] for b in blocks]), block_diag(*[b[1] for b in blocks]),
                block_diag(*[b[2] for b in blocks]))
    ctx.set_loads(block_diag(*[b[3] for b in blocks]), block_diag(*[b[4] for b in blocks]),
                  am.on_boundary)
    return arr, am, blocks, ctx


@pytest.mark.parametrize('nstreams,symmetric', [(1, 0), (3, 0), (1, 1), (3, 1)])
def test_sweep_vs_oracle(L, po, nstreams, symmetric):
    arr, am, blocks, ctx = _setup_sweep(L, po, (2, 1), 4)
    ctx.set_streams(nstreams)
    ctx.set_symmetric(symmetric, 1)
    freqs = np.array([0.0, 1e6, 7.3e6, 20e6, 50e6])
    xp, xn, t = ctx.sweep(freqs)
    for i, f in enumerate(freqs):
        Xo, XPo = po.solve_frequency(am, blocks, f)
        assert _rel(xn[i].T, Xo) < 1e-6, (f, _rel(xn[i].T, Xo))
        assert _rel(xp[i].T, XPo) < 1e-6
        assert _rel(xn[i].T, Xo) < 1e-9     # what the same-rule comparison actually achieves
    ctx.sweep_device(freqs)
    assert np.array_equal(ctx.sweep_fetch(len(freqs)), xp) or _rel(ctx.sweep_fetch(len(freqs)), xp) < 1e-12
    ctx.close()


def test_symmetric_path_convergence_guard(L, po):
    """Symmetric-part elimination + refinement: 0 steps leaves the Sauter-Schwab asymmetry (~1e-7) in
    the result, 1 step removes it; an unreachable tolerance forces the general-path redo, which then
    reproduces the general result; an asymmetric K, M or B goes into the correction too."""
    arr, am, blocks, ctx = _setup_sweep(L, po, (2, 1), 4)
    freqs = np.array([2e6, 11e6, 37e6])
    xp0, xn0, _ = ctx.sweep(freqs)
    ctx.set_symmetric(1, 0)
    _, xn_raw, _ = ctx.sweep(freqs)
    assert 1e-12 < _rel(xn_raw, xn0) < 1e-4, _rel(xn_raw, xn0)
    ctx.set_symmetric(1, 1)
    xp1, xn1, _ = ctx.sweep(freqs)
    st = ctx.symmetric_stats()
    assert st['mode'] == 1 and st['redone'] == 0 and 0 < st['max_ratio'] < 1e-5, st
    assert _rel(xn1, xn0) < 1e-11 and _rel(xp1, xp0) < 1e-11, (_rel(xn1, xn0), _rel(xp1, xp0))
    ctx.set_symmetric(1, 1, tol=1e-30)
    xp2, xn2, _ = ctx.sweep(freqs)
    assert ctx.symmetric_stats()['redone'] == len(freqs)
    assert _rel(xn2, xn0) < 1e-13 and _rel(xp2, xp0) < 1e-13      # atomics: summation order differs between runs
    ctx.close()
    # asymmetric FEM operators
    from scipy.linalg import block_diag
    ctx = L.Context(am.vertices, am.triangles)
    K, M, B = (block_diag(*[b[i] for b in blocks]) for i in range(3))
    rng = np.random.default_rng(5)
    K = K * (1 + 1e-6 * rng.standard_normal(K.shape))
    B = B * (1 + 1e-6 * rng.standard_normal(B.shape))
    ctx.set_mbk(K, M, B)
    ctx.set_loads(block_diag(*[b[3] for b in blocks]), block_diag(*[b[4] for b in blocks]), am.on_boundary)
    _, xg, _ = ctx.sweep(freqs)
    ctx.set_symmetric(1, 2)
    _, xs, _ = ctx.sweep(freqs)
    st = ctx.symmetric_stats()
    assert st['mbk_asym_nnz'] > 0
    assert _rel(xs, xg) < 1e-10, (_rel(xs, xg), st)
    ctx.close()


def test_translation_classes(L, po):
    """A membrane array is one mesh translated to every membrane position, so Z's membrane blocks
    depend on the offset only: integrating one block per offset class and copying the others must
    reproduce integrating everything (to the rounding of the translated coordinates), and a mesh that
    is not an exact set of translated copies must not be treated as one."""
    from scipy.linalg import block_diag
    arr = po.matrix_array(nelem=(3, 2))
    am, blocks = po.array_mesh(arr, 3), po.array_fem(arr, 3)
    mbk = [block_diag(*[b[i] for b in blocks]) for i in range(3)]
    loads = (block_diag(*[b[3] for b in blocks]), block_diag(*[b[4] for b in blocks]), am.on_boundary)
    freqs = np.array([1.5e6, 9e6, 41e6])

    def run(vertices, classes):
        ctx = L.Context(vertices, am.triangles)
        ctx.set_mbk(*mbk)
        ctx.set_loads(*loads)
        ctx.set_symmetric(1, 1)
        ctx.set_translation_classes(classes)
        xp, xn, _ = ctx.sweep(freqs)
        info = ctx.translation_info()
        ctx.close()
        return xn, info

    x_all, info0 = run(am.vertices, False)
    x_cls, info1 = run(am.vertices, True)
    assert not info0['in_use'] and info1['in_use']
    assert info1['components'] == 6
    # 3 x 2 lattice: offsets (dx, dy) with dx in -2..2, dy in -1..1, up to sign -> 7 classes
    assert info1['offset_classes'] == 7
    assert _rel(x_cls, x_all) < 1e-11
    Xo, _ = po.solve_frequency(am, blocks, freqs[1])
    assert _rel(x_cls[1].T, Xo) < 1e-9
    # one vertex moved by 1e-3 of an edge: not translated copies any more
    V = am.vertices.copy()
    V[am.nvertices // 2 + 3, 0] += 1e-9
    x_ref, _ = run(V, False)
    x_new, info2 = run(V, True)
    assert not info2['in_use']
    assert _rel(x_new, x_ref) < 1e-12


def test_pressure_vs_oracle(L, po):
    arr, am, blocks, ctx = _setup_sweep(L, po, (2, 1), 3)
    rng = np.random.default_rng(1)
    obs = np.c_[rng.uniform(-1e-3, 1e-3, 300), rng.uniform(-1e-3, 1e-3, 300), rng.uniform(1e-4, 2.1e-3, 300)]
    ks = np.array([2 * np.pi * 1e6 / 1500, 2 * np.pi * 17e6 / 1500])
    for k in ks:
        assert _rel(ctx.pres_vector(obs[:7], k), po.pres_vector(am, obs[:7], k)) < 1e-12
    disp = rng.standard_normal((2, 5, am.nvertices)) + 1j * rng.standard_normal((2, 5, am.nvertices))
    p = ctx.pressure(obs, ks, disp)
    for ik, k in enumerate(ks):
        pv = po.pres_vector(am, obs, k)          # [nobs, nv]
        ref = disp[ik] @ pv.T                    # [nsrc, nobs]   (p_vector.dot(disp))
        assert _rel(p[ik], ref) < 1e-11
    ctx.close()


def test_pressure_many_sources_vs_oracle(L, po):
    """nsrc > 4 takes the node-space path (k_pvec_tiles + ZGEMM): ragged point counts (not a multiple of
    the 128-point tile), one point, source counts that are not a multiple of the ZGEMM tile."""
    arr, am, blocks, ctx = _setup_sweep(L, po, (2, 1), 4)
    rng = np.random.default_rng(2)
    for nobs, nsrc in [(1, 5), (300, 7), (1000, 18)]:
        obs = np.c_[rng.uniform(-1e-3, 1e-3, nobs), rng.uniform(-1e-3, 1e-3, nobs), rng.uniform(1e-4, 2.1e-3, nobs)]
        ks = np.array([0.0, 2 * np.pi * 3e6 / 1500, 2 * np.pi * 41e6 / 1500])
        disp = rng.standard_normal((3, nsrc, am.nvertices)) + 1j * rng.standard_normal((3, nsrc, am.nvertices))
        p = ctx.pressure(obs, ks, disp)
        for ik, k in enumerate(ks):
            ref = disp[ik] @ po.pres_vector(am, obs, k).T
            if k == 0:
                assert np.abs(p[ik]).max() == 0 and np.abs(ref).max() == 0      # scale -(k c)^2 2 rho = 0
            else:
                assert _rel(p[ik], ref) < 1e-11
    # the resident-points object: reusable page-locked output, device-resident rerun, sampled fetch
    obs = np.c_[rng.uniform(-1e-3, 1e-3, 777), rng.uniform(-1e-3, 1e-3, 777), rng.uniform(1e-4, 2.1e-3, 777)]
    pf = L.PressureField(ctx, obs, 9)
    disp = rng.standard_normal((9, am.nvertices)) + 1j * rng.standard_normal((9, am.nvertices))
    k = 2 * np.pi * 11e6 / 1500
    out = np.empty((9, 777), dtype=np.complex128)
    p = pf.run(k, disp, out=out)
    ref = disp @ po.pres_vector(am, obs, k).T
    assert p is out and _rel(p, ref) < 1e-11
    pf.run_device(k, 9)
    idx = np.array([0, 5, 776, 5], dtype=np.uint32)
    assert _rel(pf.fetch(9, idx), ref[:, idx]) < 1e-11
    assert _rel(pf.fetch(9), ref) < 1e-11
    # fewer sources than the capacity, and agreement with the few-source kernel (k_pressure)
    assert _rel(pf.run(k, disp[:3]), ctx.pressure(obs, [k], disp[None, :3])[0]) < 1e-12
    info = pf.info()
    assert info['chunks'] == 1 and info['tri_chunks'] >= 1
    pf.close()
    ctx.close()


def test_nearfield_subblock_vs_oracle_and_dense(L, po):
    am = po.array_mesh(po.matrix_array(nelem=(2, 1)), 5)
    k = 2 * np.pi * 9e6 / 1500
    ctx = L.Context(am.vertices, am.triangles)
    Z = ctx.assemble_z(k)
    rng = np.random.default_rng(5)
    for nr, nc in [(1, 1), (7, 122), (40, 3), (100, 100)]:
        ri = rng.choice(am.nvertices, nr, replace=False).astype(np.uint32)
        ci = rng.choice(am.nvertices, nc, replace=False).astype(np.uint32)
        N = ctx.nearfield(k, ri, ci)
        assert np.abs(N - Z[np.ix_(ri, ci)]).max() / np.abs(Z).max() < 1e-10
        assert np.abs(N - po.nearfield(am, k, ri, ci)).max() / np.abs(Z).max() < 1e-10
    assert ctx.nearfield(k, np.zeros(0, np.uint32), np.arange(4, dtype=np.uint32)).shape == (0, 4)     # empty input
    with pytest.raises(L.Cnldb200Error):
        ctx.nearfield(k, [1, 1], [2])          # repeated index
    ctx.close()


def test_compat_dense_entry_points(L):
    """The H2Lib-named dense entry points that used to loop on the host now run on the GPU behind a
    device-resident copy (densedev.cu): matvec, unpivoted LU + repeated solves with the cached factors,
    adjoint solve, triangular solves, Cholesky, CG, GMRES, spectral-norm estimates."""
    from cnld import h2lib as h
    rng = np.random.default_rng(11)
    n = 300
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + n * 0.8 * np.eye(n)
    A = h.AMatrix.from_array(np.ascontiguousarray(a.T))           # numpy view is the transpose of the stored matrix
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = h.AVector(n)
    y.v[:] = 1.0
    h.addeval_amatrix_avector(2.0 - 1j, A, h.AVector.from_array(x), y)
    assert _rel(y.v, 1.0 + (2.0 - 1j) * (a @ x)) < 1e-13
    LU = h.clone_amatrix(A)
    assert h.lrdecomp_amatrix(LU) == 0
    for _ in range(3):                                             # same factors, no re-upload
        b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        v = h.AVector.from_array(b.copy())
        h.lrsolve_amatrix_avector(False, LU, v)
        assert _rel(a @ v.v, b) < 1e-11
    v = h.AVector.from_array(b.copy())
    h.lrsolve_amatrix_avector(True, LU, v)                         # adjoint system
    assert _rel(a.conj().T @ v.v, b) < 1e-11
    assert h.norm2diff_lr_amatrix(A, LU) < 1e-10 * np.linalg.norm(a, 2)
    LU.a[7, 7] += 1.0                                              # host copy rewritten (a diagonal entry is always in the
                                                                   # fingerprint): the cached factors are stale
    v = h.AVector.from_array(b.copy())
    h.lrsolve_amatrix_avector(False, LU, v)
    assert _rel(a @ v.v, b) > 1e-6
    # triangular solves with the factors of a fresh factorisation
    LU = h.clone_amatrix(A)
    h.lrdecomp_amatrix(LU)
    lu = np.array(LU.a).T
    Lm, Um = np.tril(lu, -1) + np.eye(n), np.triu(lu)
    for lower, unit, trans, M in [(True, True, False, Lm), (False, False, False, Um), (True, True, True, Lm.conj().T),
                                  (False, False, True, Um.conj().T)]:
        v = h.AVector.from_array(b.copy())
        h.triangularsolve_amatrix_avector(lower, unit, trans, LU, v)
        assert _rel(M @ v.v, b) < 1e-10
    # Hermitian positive definite: Cholesky, CG
    hp = a.conj().T @ a
    Hm = h.AMatrix.from_array(np.ascontiguousarray(hp.T))
    CH = h.clone_amatrix(Hm)
    assert h.choldecomp_amatrix(CH) == 0
    lc = np.tril(np.array(CH.a).T)
    assert _rel(lc @ lc.conj().T, hp) < 1e-12
    v = h.AVector.from_array(b.copy())
    h.cholsolve_amatrix_avector(CH, v)
    assert _rel(hp @ v.v, b) < 1e-9
    bad = h.AMatrix.from_array(np.ascontiguousarray((hp - 2 * np.linalg.eigvalsh(hp)[5] * np.eye(n)).T))
    assert h.choldecomp_amatrix(bad) != 0                          # indefinite: reports the failing pivot
    w = h.AVector(n)
    w.v[:] = 0
    assert 0 < h.solve_cg_amatrix_avector(Hm, h.AVector.from_array(b.copy()), w, 1e-12, 2000) < 2000
    assert _rel(hp @ w.v, b) < 1e-9
    w.v[:] = 0
    assert h.solve_gmres_amatrix_avector(A, h.AVector.from_array(b.copy()), w, 1e-11, 500, 50) < 500
    assert _rel(a @ w.v, b) < 1e-9
    Bm = h.AMatrix.from_array(np.ascontiguousarray((a + 0.01 * rng.standard_normal((n, n))).T))
    est = h.norm2diff_lr_amatrix(Bm, LU)
    exact = np.linalg.norm(np.array(Bm.a).T - a, 2)
    assert 0.7 * exact < est <= exact * (1 + 1e-9)
