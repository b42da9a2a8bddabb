"""
cnld.fem -- finite-element inputs of the frequency-response path (mirror of
/root/reference/cnld/fem.py): plate stiffness K (rotation-free BPT triangles, fem.py:62-194),
lumped/consistent mass M (:779-860), Rayleigh damping B from two modal frequencies (:885-926),
patch load F and patch averaging AVG (:1004-1216) and the per-frequency block-diagonal
operator -(w^2) M - 1j w B + K (:1289-1315).

K, M, B, F, AVG do not depend on frequency.  They are computed once per membrane *type* --
the reference's memoisation keys exclude id / position / patches (abstract.py:241-243), so its
own code also reuses the first membrane's matrices -- and handed to the GPU once per sweep
(`cnld.sweep`); only the combination with w happens per frequency, on the device.
"""
import numpy as np
from scipy import linalg
from scipy import sparse as sps
from scipy.integrate import dblquad

from . import abstract, mesh, util

eps = np.finfo(np.float64).eps


def _membrane_mesh(mem, refn, center=(0, 0, 0)):
    if isinstance(mem, abstract.SquareCmutMembrane):
        return mesh.square(mem.length_x, mem.length_y, refn, center=center)
    return mesh.circle(mem.radius, refn, center=center)


def _edge_neighbours(tri_edges):
    """Triangle across each edge (edge j lies opposite vertex j); -1 on the rim."""
    nt = len(tri_edges)
    flat = tri_edges.ravel().astype(np.int64)
    order = np.argsort(flat, kind='stable')
    nb = -np.ones(3 * nt, dtype=np.int64)
    same = flat[order][1:] == flat[order][:-1]
    a, b = order[:-1][same], order[1:][same]
    nb[a] = b // 3
    nb[b] = a // 3
    return nb.reshape(nt, 3)


@util.memoize
def mem_k_matrix(mem, refn, type='bpt'):
    if type.lower() != 'bpt':
        raise NotImplementedError('only the BPT stiffness (the reference default, fem.py:49) is provided')
    return mem_k_matrix_bpt(mem, refn)


@util.memoize
def mem_k_matrix_bpt(mem, refn):
    """Basic plate triangle (Onate & Zarate 2000): the curvature of each control triangle is
    the boundary integral of the averaged deflection gradient of the triangle and its edge
    neighbour; 6-node patch stiffness Bp^T D Bp * area (fem.py:62-194)."""
    am = _membrane_mesh(mem, refn)
    xy = am.vertices[:, :2]
    tris = am.triangles.astype(np.int64)
    area = am.triangle_areas
    ob = am.on_boundary
    nb = _edge_neighbours(am.triangle_edges)
    h, E, nu = mem.thickness[0], mem.y_modulus[0], mem.p_ratio[0]
    D = np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]]) * E * h**3 / (12 * (1 - nu**2))
    # hat-function gradients per triangle: inv(J^T) [[-1,1,0],[-1,0,1]]
    P = xy[tris]
    J = np.stack([P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]], axis=2)
    grad = np.linalg.inv(np.transpose(J, (0, 2, 1))) @ np.array([[-1., 1., 0.], [-1., 0., 1.]])
    n = len(xy)
    K = np.zeros((n, n))
    for p in range(len(tris)):
        loc = list(tris[p])
        cols = [0, 1, 2]
        Bp = np.zeros((3, 6))
        for j in range(3):
            q = nb[p, j]
            if q < 0:
                continue
            far = [v for v in tris[q] if v not in tris[p]][0]
            loc.append(far)
            cols.append(3 + j)
            # edge j runs from vertex j+2 to vertex j+1 so that z x r points outward
            r = P[p, (j + 1) % 3] - P[p, (j + 2) % 3]
            ln = np.hypot(r[0], r[1])
            nx, ny = -r[1] / ln, r[0] / ln
            T = np.array([[-nx, 0.], [0., -ny], [-ny, -nx]])
            Bp[:, :3] += ln / 2 * T.dot(grad[p])
            where = [cols[loc.index(v)] for v in tris[q]]
            Bp[:, where] += ln / 2 * T.dot(grad[q])
        Bp /= area[p]
        Kp = Bp.T.dot(D).dot(Bp) * area[p]
        K[np.ix_(loc, loc)] += Kp[np.ix_(cols, cols)]
    K[ob, :] = 0
    K[:, ob] = 0
    K[ob, ob] = 1
    return K


def _clamp(M, ob):
    M[ob, :] = 0
    M[:, ob] = 0
    M[ob, ob] = 1
    return M


@util.memoize
def mem_cm_matrix(mem, refn):
    """Consistent mass with linear shape functions (fem.py:792-826)."""
    am = _membrane_mesh(mem, refn)
    mass = sum(x * y for x, y in zip(mem.density, mem.thickness))
    tris, area = am.triangles.astype(np.int64), am.triangle_areas
    Mt = np.array([[1, .5, .5], [.5, 1, .5], [.5, .5, 1]]) / 12
    M = np.zeros((am.nvertices, am.nvertices))
    for tt in range(len(tris)):
        M[np.ix_(tris[tt], tris[tt])] += 2 * Mt * mass * area[tt]
    return _clamp(M, am.on_boundary)


@util.memoize
def mem_dlm_matrix(mem, refn):
    """Diagonally lumped mass (fem.py:829-860)."""
    am = _membrane_mesh(mem, refn)
    mass = sum(x * y for x, y in zip(mem.density, mem.thickness))
    tris, area = am.triangles.astype(np.int64), am.g / 2
    M = np.zeros((am.nvertices, am.nvertices))
    for tt in range(len(tris)):
        M[tris[tt], tris[tt]] += 1 / 3 * mass * area[tt]
    return _clamp(M, am.on_boundary)


@util.memoize
def mem_m_matrix(mem, refn, mu=0.5):
    return mu * mem_dlm_matrix(mem, refn) + (1 - mu) * mem_cm_matrix(mem, refn)


@util.memoize
def mem_eig(mem, refn):
    """Eigenfrequencies (Hz) and modes of the clamped membrane (fem.py:885-903)."""
    ob = _membrane_mesh(mem, refn).on_boundary
    M = mem_m_matrix(mem, refn, mu=0.5)
    K = mem_k_matrix(mem, refn)
    w, v = linalg.eig(linalg.inv(M).dot(K)[np.ix_(~ob, ~ob)])
    idx = np.argsort(np.sqrt(np.abs(w)))
    return np.sqrt(np.abs(w))[idx] / (2 * np.pi), v[:, idx]


def mem_b_matrix_eig(mem, refn, M, K):
    """Rayleigh damping alpha M + beta K matching two modal damping ratios (fem.py:906-926)."""
    eigf, _ = mem_eig(mem, refn)
    oa, ob_ = eigf[mem.damping_mode_a] * 2 * np.pi, eigf[mem.damping_mode_b] * 2 * np.pi
    A = 1 / 2 * np.array([[1 / oa, oa], [1 / ob_, ob_]])
    alpha, beta = linalg.inv(A).dot([mem.damping_ratio_a, mem.damping_ratio_b])
    return alpha * M + beta * K


def _patch_fractions(am, mem, pat):
    """Covered fraction of every triangle by one patch: 1 / 0 from the three vertex tests,
    partially covered triangles by the same loose dblquad (epsrel = epsabs = 1e-1) the reference
    uses (fem.py:1020-1105)."""
    xy = am.vertices[:, :2]
    tris = am.triangles
    px, py, _ = pat.position
    if isinstance(mem, abstract.SquareCmutMembrane):
        plx, ply = pat.length_x, pat.length_y

        def load(x, y):
            return 1 if (x - (px - plx / 2) >= -2 * eps and x - (px + plx / 2) <= 2 * eps and
                         y - (py - ply / 2) >= -2 * eps and y - (py + ply / 2) <= 2 * eps) else 0
    else:
        def load(x, y):
            r = np.sqrt((x - px)**2 + (y - py)**2)
            th = np.arctan2(y - py, x - px)
            th1 = th - 2 * eps
            th1 = th1 + 2 * np.pi if th1 < -np.pi else th1
            th2 = th + 2 * eps
            th2 = th2 - 2 * np.pi if th2 > np.pi else th2
            if r - pat.radius_min >= -2 * eps and r - pat.radius_max <= 2 * eps:
                if pat.theta_min <= th1 <= pat.theta_max or pat.theta_min <= th2 <= pat.theta_max:
                    return 1
            return 0
    inside = np.array([[load(*xy[v]) for v in tri] for tri in tris])
    frac = np.where(inside.all(axis=1), 1.0, 0.0)
    for tt in np.nonzero(inside.any(axis=1) & ~inside.all(axis=1))[0]:
        (xi, yi), (xj, yj), (xk, yk) = xy[tris[tt]]
        integ, _ = dblquad(lambda psi, eta: load((xj - xi) * psi + (xk - xi) * eta + xi,
                                                 (yj - yi) * psi + (yk - yi) * eta + yi),
                           0, 1, 0, lambda x: 1 - x, epsrel=1e-1, epsabs=1e-1)
        frac[tt] = integ / (1 / 2)
    return frac


@util.memoize
def _mem_patch_fractions(mem, refn):
    am = _membrane_mesh(mem, refn, center=mem.position)
    frac = np.array([_patch_fractions(am, mem, pat) for pat in mem.patches])
    return (am.triangles.astype(np.int64), np.array(am.triangle_areas), np.array(am.on_boundary),
            am.nvertices, frac)


@util.memoize
def mem_patch_f_matrix(mem, refn):
    """Pressure-load vectors, one column per patch: covered triangle area shared equally by the
    triangle's non-boundary nodes (fem.py:1004-1105)."""
    tris, area, ob, nv, frac = _mem_patch_fractions(mem, refn)
    F = np.zeros((nv, len(mem.patches)))
    for ip in range(len(mem.patches)):
        for tt in np.nonzero(frac[ip])[0]:
            F[tris[tt], ip] += 1 / (1 * np.sum(~ob[tris[tt]])) * frac[ip, tt] * area[tt]
        F[ob, ip] = 0
    return F


@util.memoize
def mem_patch_avg_matrix(mem, refn):
    """Patch-averaging vectors: covered area / 3 per node, divided by the patch area
    (fem.py:1109-1216)."""
    tris, area, ob, nv, frac = _mem_patch_fractions(mem, refn)
    A = np.zeros((nv, len(mem.patches)))
    for ip, pat in enumerate(mem.patches):
        for tt in np.nonzero(frac[ip])[0]:
            A[tris[tt], ip] += 1 / 3 * frac[ip, tt] * area[tt]
        A[:, ip] /= pat.area
    return A


def _membranes(array):
    return [m for e in array.elements for m in e.membranes]


def array_f_spmatrix(array, refn):
    return sps.block_diag([mem_patch_f_matrix(m, refn) for m in _membranes(array)], format='csc')


def array_avg_spmatrix(array, refn):
    return sps.block_diag([mem_patch_avg_matrix(m, refn) for m in _membranes(array)], format='csc')


def array_kmb_spmatrices(array, refn):
    """Frequency-independent block-diagonal K, M, B of the whole array (CSR) -- what the sweep
    uploads once instead of rebuilding the combined operator for every frequency."""
    Ks, Ms, Bs = [], [], []
    for mem in _membranes(array):
        M = mem_m_matrix(mem, refn, mu=0.5)
        K = mem_k_matrix(mem, refn)
        Ks.append(K)
        Ms.append(M)
        Bs.append(mem_b_matrix_eig(mem, refn, M, K))
    return tuple(sps.block_diag(b, format='csr') for b in (Ks, Ms, Bs))


def array_mbk_spmatrix(array, refn, f, inv=False, format='csr'):
    """-(w^2) M - 1j w B + K, block diagonal (fem.py:1289-1315).  `format` is accepted because
    the reference's own driver passes it (generate_db.py:46)."""
    omg = 2 * np.pi * f
    blocks, inverses = [], []
    for mem in _membranes(array):
        M = mem_m_matrix(mem, refn, mu=0.5)
        K = mem_k_matrix(mem, refn)
        B = mem_b_matrix_eig(mem, refn, M, K)
        blocks.append(-(omg**2) * M - 1j * omg * B + K)
        if inv:
            inverses.append(np.linalg.inv(blocks[-1]))
    if inv:
        return sps.block_diag(blocks, format='csr'), sps.block_diag(inverses, format='csr')
    return sps.block_diag(blocks, format=format)
