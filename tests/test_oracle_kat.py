"""Known-answer tests of the oracle itself.  The reference pins nothing for the H2Lib-backed
stages, so the oracle has to prove its own rules:
  * every rule family integrates polynomials over T x T exactly (the Sauter-Schwab regions tile the
    product of the two reference triangles once);
  * the singular families converge to semi-analytic Laplace (k = 0) integrals: closed-form inner
    integral of 1/|x-y| over a plane triangle + high-order outer quadrature;
  * convergence in q_sing, (near-)symmetry of Z, unpivoted-LU residuals, k -> 0 continuity."""
import numpy as np
import pytest


def _mono(a, b):       # int_{0<=s<=t<=1} t^a s^b = 1 / ((b+1)(a+b+2))
    return 1.0 / ((b + 1) * (a + b + 2))


@pytest.mark.parametrize('family,npts', [(0, 3**4), (1, 2 * 4**4), (2, 5 * 4**4), (3, 6 * 4**4)])
def test_rule_point_counts_and_moments(po, family, npts):
    xt, xs, yt, ys, w = po.singquad_rule(family, 3, 4)
    assert len(w) == npts                                  # q^4 | 2,5,6 * q2^4
    assert abs(w.sum() - 0.25) < 1e-15
    assert (xs >= -1e-15).all() and (xs <= xt + 1e-15).all() and (xt <= 1 + 1e-15).all()
    assert (ys >= -1e-15).all() and (ys <= yt + 1e-15).all() and (yt <= 1 + 1e-15).all()
    for a, b, c, d in [(1, 0, 0, 1), (0, 2, 1, 0), (1, 1, 1, 0), (2, 0, 0, 1)]:
        v = (w * xt**a * xs**b * yt**c * ys**d).sum()
        assert abs(v - _mono(a, b) * _mono(c, d)) < 1e-14


def test_gauss_legendre(po):
    for m in (1, 2, 4, 7):
        x, w = po.gauss(m)
        assert abs(w.sum() - 2) < 1e-14 and np.all(np.diff(x) > 0)
        for p in range(2 * m):
            assert abs((w * x**p).sum() - (0 if p % 2 else 2 / (p + 1))) < 1e-13


def _tri_potential(P, x):
    """int_T 1/|x - y| dy for x in the plane of the triangle P (3x2), edge-wise closed form."""
    tot = 0.0
    for a in range(3):
        p1, p2 = P[a], P[(a + 1) % 3]
        e = (p2 - p1) / np.linalg.norm(p2 - p1)
        d = np.dot(p1 - x, np.array([e[1], -e[0]]))
        if abs(d) < 1e-300:
            continue
        s1, s2 = np.dot(p1 - x, e), np.dot(p2 - x, e)
        tot += d * np.log((s2 + np.linalg.norm(p2 - x)) / (s1 + np.linalg.norm(p1 - x)))
    return tot


def _outer(po, P, Q, m=48):
    gx, gw = po.gauss(m)
    gx, gw = 0.5 * (gx + 1), 0.5 * gw
    A, B, C = P
    g = abs((B[0] - A[0]) * (C[1] - A[1]) - (B[1] - A[1]) * (C[0] - A[0]))
    tot = 0.0
    for i in range(m):
        for j in range(m):
            t, s = gx[i], gx[i] * gx[j]
            tot += gw[i] * gw[j] * gx[i] * _tri_potential(Q, A * (1 - t) + B * (t - s) + C * s)
    return tot * g


VERTS = [(0, 0), (1, 0), (0.3, 0.8), (1.2, 0.9), (-0.7, 0.5), (2.0, 0.1)]
TRIS = [(0, 1, 2), (1, 3, 2), (0, 2, 4), (1, 5, 3)]


def _mesh(po, tris):
    v = np.c_[np.array(VERTS, float), np.zeros(len(VERTS))]
    return po.Mesh(v, np.zeros((1, 2), np.uint32), np.array(tris, np.uint32), np.zeros((len(tris), 3), np.uint32))


@pytest.mark.parametrize('pair,name', [((0, 0), 'identical'), ((0, 1), 'edge'), ((0, 2), 'edge'), ((0, 3), 'vertex'),
                                       ((2, 3), 'distant')])
def test_laplace_pair_integrals_converge_to_semi_analytic(po, pair, name):
    V = np.array(VERTS, float)
    a, b = pair
    Pa, Pb = V[list(TRIS[a])], V[list(TRIS[b])]
    if a == b:
        ref = _outer(po, Pa, Pa)
        tris = [TRIS[a]]
    else:
        ref = _outer(po, Pa, Pa) + _outer(po, Pb, Pb) + _outer(po, Pa, Pb) + _outer(po, Pb, Pa)
        tris = [TRIS[a], TRIS[b]]
    m = _mesh(po, tris)
    errs = []
    for q in (4, 8):
        # the hats sum to one, so the sum of all entries is the constant-basis integral
        val = po.assemble_z(m, 0.0, q if name == 'distant' else 2, q).real.sum() * 4 * np.pi
        errs.append(abs(val - ref) / ref)
    assert errs[0] < 5e-5 and errs[1] < 5e-7 and errs[1] < errs[0]      # ref itself is good to ~5e-8


def test_z_properties(po):
    am = po.array_mesh(po.matrix_array(nelem=(2, 1)), 3)
    k = 2 * np.pi * 4e6 / 1500
    Z4, Z6 = po.assemble_z(am, k, 2, 4), po.assemble_z(am, k, 2, 6)
    assert np.abs(Z4 - Z4.T).max() / np.abs(Z4).max() < 1e-5        # symmetric up to quadrature error
    assert np.linalg.norm(Z6 - Z4) / np.linalg.norm(Z6) < 2e-4      # convergence in q_sing
    Z0, Ze = po.assemble_z(am, 0.0), po.assemble_z(am, 1e-3)
    assert np.abs(Z0.imag).max() == 0.0 and np.linalg.norm(Ze - Z0) / np.linalg.norm(Z0) < 1e-6
    # complex wavenumber: exp(-Im(k) r) damping
    Zc = po.assemble_z(am, k + 1e3j)
    assert np.abs(Zc).max() < np.abs(Z4).max() * 1.0001 and np.linalg.norm(Zc - Z4) > 0
    # sub-block entry == dense entry
    ri, ci = np.array([3, 17, 30, 8], np.uint32), np.array([40, 2, 17], np.uint32)
    assert np.abs(po.nearfield(am, k, ri, ci) - Z4[np.ix_(ri, ci)]).max() / np.abs(Z4).max() < 1e-13


def test_unpivoted_lu_and_solve(po):
    rng = np.random.default_rng(3)
    n = 60
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + 0.6 * n * np.eye(n)
    LU, st = po.lrdecomp(G)
    assert st == 0
    Lm, Um = np.tril(LU, -1) + np.eye(n), np.triu(LU)
    assert np.linalg.norm(Lm @ Um - G) / np.linalg.norm(G) < 1e-14
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.linalg.norm(G @ po.lrsolve(LU, b) - b) / np.linalg.norm(b) < 1e-13
    assert np.linalg.norm(po.lu_blocked(G, nb=16) - LU) / np.linalg.norm(LU) < 1e-13
    G[7, 7] = 0
    G[:7, 7] = 0
    G[7, :7] = 0
    assert po.lrdecomp(G)[1] == 8          # index of the failing pivot + 1, like lrdecomp_amatrix


def test_aca_block_accuracy(po):
    am = po.array_mesh(po.matrix_array(nelem=(2, 1)), 5)
    k = 2 * np.pi * 10e6 / 1500
    n = am.nvertices // 2
    ri, ci = np.arange(0, n, dtype=np.uint32), np.arange(n, 2 * n, dtype=np.uint32)    # the two membranes
    A = po.nearfield(am, k, ri, ci)
    prev = None
    for eps in (1e-2, 1e-4, 1e-7):
        U, V = po.aca_block(am, k, ri, ci, eps)
        err = np.linalg.norm(U @ V.T - A) / np.linalg.norm(A)
        assert err < 3 * eps and U.shape[1] < 25
        assert prev is None or U.shape[1] >= prev
        prev = U.shape[1]


def test_asymmetry_of_z_lives_on_touching_pairs_only(po):
    """The assumption behind the GPU's symmetric path, checked on the oracle alone:
    Z - Z^T is tiny and non-zero only at entries (i, j) whose vertices belong to touching triangles
    (the Sauter-Schwab rules are not symmetric under swapping the two triangles; the regular tensor
    rule is), and translated membranes give translated blocks."""
    arr = po.matrix_array(nelem=(2, 1))
    am = po.array_mesh(arr, 4)
    k = 2 * np.pi * 7e6 / 1500
    Z = po.assemble_z(am, k)
    n = am.nvertices
    A = Z - Z.T
    scale = np.abs(Z).max()
    assert 1e-12 < np.abs(A).max() / scale < 1e-5
    # vertex pairs fed by touching triangle pairs
    tri = np.asarray(am.triangles)
    v2t = [set() for _ in range(n)]
    for t, vs in enumerate(tri):
        for v in vs:
            v2t[v].add(t)
    touch = np.zeros((n, n), dtype=bool)
    for t, vs in enumerate(tri):
        nb = set().union(*[v2t[v] for v in vs])          # triangles sharing a vertex with t
        cols = np.unique(tri[sorted(nb)].ravel())
        touch[np.ix_(vs, cols)] = True
    assert np.abs(A[~touch]).max() <= 1e-15 * scale
    # translation: the two membranes' diagonal blocks agree to rounding, and the off-diagonal blocks are transposes
    h = n // 2
    assert np.abs(Z[:h, :h] - Z[h:, h:]).max() <= 1e-12 * scale
    assert np.abs(Z[h:, :h] - Z[:h, h:].T).max() <= 1e-12 * scale
