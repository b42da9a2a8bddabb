"""Runs INSIDE the route-A package (PYTHONPATH = tests/_route_a only): the reference's own Cython
binding layer (cnld/h2lib/*.pyx, built unmodified against include/h2lib_compat and linked with
-lcnld_b200 by scripts/build_route_a.py) driven the way the reference drives it:

  mesh      cnld/mesh.py:306-384  (Macrosurface3d -> set_parametrization('square') -> build_from_macrosurface3d_surface3d)
  dense     cnld/compressed_formats.py:616-645 (ZFullMatrix), :220-223 (_smul), :209-212 (_add sparse),
            :247-262 (_lu / _lusolve), :254-267 (_chol / _cholsolve)
  H-matrix  cnld/compressed_formats.py:652-703 (ZHMatrix), :382-448 (HFormat _smul / _add / _lu / _lusolve)

Writes the results to an .npz for the parent test, which compares them with the oracle.
usage: python recipe.py out.npz [import-only]
"""
import sys

import numpy as np

import cnld.h2lib as h

if len(sys.argv) > 2 and sys.argv[2] == 'import-only':
    names = ['AMatrix', 'AVector', 'SparseMatrix', 'Cluster', 'Block', 'Surface3d', 'Macrosurface3d', 'Bem3d', 'RKMatrix',
             'HMatrix', 'Truncmode', 'new_slp_helmholtz_bem3d', 'new_dlp_helmholtz_bem3d', 'assemble_bem3d_amatrix',
             'assemble_bem3d_hmatrix', 'lrdecomp_amatrix', 'lrsolve_amatrix_avector', 'choldecomp_amatrix',
             'cholsolve_amatrix_avector', 'lrdecomp_hmatrix', 'lrsolve_hmatrix_avector', 'addmul_hmatrix', 'add_hmatrix',
             'add_amatrix_hmatrix', 'solve_cg_amatrix_avector', 'solve_cg_hmatrix_avector', 'solve_gmres_hmatrix_avector',
             'norm2diff_amatrix_hmatrix', 'norm2diff_lr_amatrix', 'norm2diff_lr_hmatrix']
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing
    print('route A import ok:', len([n for n in dir(h) if not n.startswith('_')]), 'public names')
    sys.exit(0)

LIN = h.basisfunctionbem3d.LINEAR
xl = yl = 40e-6
refn = 5
v = np.array([[-xl / 2, -yl / 2, 0.], [xl / 2, -yl / 2, 0.], [xl / 2, yl / 2, 0.], [-xl / 2, yl / 2, 0.], [0., 0., 0.]])
e = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [0, 4], [1, 4], [2, 4], [3, 4]], dtype=np.uint32)
t = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], dtype=np.uint32)
s = np.array([[5, 4, 0], [6, 5, 1], [7, 6, 2], [4, 7, 3]], dtype=np.uint32)
ms = h.Macrosurface3d(len(v), len(e), len(t))
ms.x[:] = v
ms.e[:] = e
ms.t[:] = t
ms.s[:] = s
ms.set_parametrization('square')
surf = h.build_from_macrosurface3d_surface3d(ms, refn)
h.prepare_surface3d(surf)
n = surf.vertices
X, T = np.array(surf.x, copy=True), np.array(surf.t, copy=True)

f = 7e6
k = 2 * np.pi * f / 1500.
alpha = -(2 * np.pi * f) ** 2 * 2 * 1000.
rng = np.random.default_rng(0)
# a sparse, diagonally dominant stand-in for the FEM operator (the real one is pinned elsewhere)
from scipy import sparse as sps
M = sps.random(n, n, density=0.05, random_state=1, dtype=np.float64)
M = (M + M.T).tocsr() * 1e9 + sps.eye(n) * (5e10 + 3e9j)
Msp = h.SparseMatrix.from_array(sps.csr_matrix(M.astype(np.complex128)))

# ---- dense recipe ---------------------------------------------------------------------------------------
bem = h.new_slp_helmholtz_bem3d(k, surf, 2, 4, LIN, LIN)
Z = h.AMatrix(n, n)
h.assemble_bem3d_amatrix(bem, Z)
Zn = np.array(Z.a).T.copy()                              # AMatrix.a is the transposed view (amatrix.pyx:34-37)
G = h.clone_amatrix(Z)
h.scale_amatrix(alpha, G)
h.add_sparsematrix_amatrix(1.0, False, Msp, G)
Gn = np.array(G.a).T.copy()
LU = h.clone_amatrix(G)
assert h.lrdecomp_amatrix(LU) == 0
B = rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3))
Xs = np.zeros_like(B)
for j in range(B.shape[1]):                               # one lrsolve per right-hand side, same LU (generate_db.py:82-91)
    x = h.AVector.from_array(B[:, j].copy())
    h.lrsolve_amatrix_avector(False, LU, x)
    Xs[:, j] = np.array(x.v)
lr_err = h.norm2diff_lr_amatrix(G, LU)
y = h.AVector(n)
h.clear_avector(y)
h.addeval_amatrix_avector(1.0, G, h.AVector.from_array(B[:, 0].copy()), y)
mv = np.array(y.v)
# Cholesky + CG on a Hermitian positive definite matrix
Hn = Zn.conj().T @ Zn + np.eye(n) * np.abs(Zn).max() ** 2
Hm = h.AMatrix.from_array(np.ascontiguousarray(Hn.T))     # from_array stores the transpose
CH = h.clone_amatrix(Hm)
h.choldecomp_amatrix(CH)
xc = h.AVector.from_array(B[:, 1].copy())
h.cholsolve_amatrix_avector(CH, xc)
xcg = h.AVector(n)
h.clear_avector(xcg)
it_cg = h.solve_cg_amatrix_avector(Hm, h.AVector.from_array(B[:, 1].copy()), xcg, 1e-12, 500)

# ---- H-matrix recipe ------------------------------------------------------------------------------------
root = h.build_bem3d_cluster(bem, 8, LIN)
broot = h.build_strict_block(root, root, 0.8, '2')
h.setup_hmatrix_aprx_paca_bem3d(bem, root, root, broot, 1e-6)
Zh = h.build_from_block_hmatrix(broot, 0)
h.assemble_bem3d_hmatrix(bem, broot, Zh)
zh_err = h.norm2diff_amatrix_hmatrix(Zh, Z)
tm = h.new_releucl_truncmode()
ident = h.clonestructure_hmatrix(Zh)
h.identity_hmatrix(ident)
Gbe = h.clonestructure_hmatrix(Zh)
h.clear_hmatrix(Gbe)
h.addmul_hmatrix(alpha, False, ident, False, Zh, tm, 1e-12, Gbe)          # HFormat._smul
Mh = h.clonestructure_hmatrix(Zh)
h.clear_hmatrix(Mh)
h.copy_sparsematrix_hmatrix(Msp, Mh)                                       # SparseFormat._as_hformat
Gh = h.clone_hmatrix(Gbe)
h.add_hmatrix(1, Mh, tm, 1e-12, Gh)                                        # HFormat._add
gh_err = h.norm2diff_amatrix_hmatrix(Gh, G)
LUh = h.clone_hmatrix(Gh)
h.lrdecomp_hmatrix(LUh, tm, 1e-12)                                         # HFormat._lu
xh = h.AVector.from_array(B[:, 0].copy())
h.lrsolve_hmatrix_avector(False, LUh, xh)                                  # HFormat._lusolve
lrh_err = h.norm2diff_lr_hmatrix(Gh, LUh)
xg = h.AVector(n)
h.clear_avector(xg)
it_gm = h.solve_gmres_hmatrix_avector(Gh, h.AVector.from_array(B[:, 0].copy()), xg, 1e-10, 400, 60)
dlp = h.new_dlp_helmholtz_bem3d(k, surf, 2, 4, LIN, LIN, 0.5) if len(sys.argv) > 2 and sys.argv[2] == 'dlp' else None

np.savez(sys.argv[1], X=X, T=T, k=k, alpha=alpha, Z=Zn, G=Gn, Mdense=M.toarray(), B=B, Xs=Xs, mv=mv, lr_err=lr_err,
         H=Hn, xchol=np.array(xc.v), xcg=np.array(xcg.v), it_cg=it_cg, zh_err=zh_err, gh_err=gh_err, xh=np.array(xh.v),
         lrh_err=lrh_err, xgmres=np.array(xg.v), it_gm=it_gm, hsize=h.getsize_hmatrix(Zh), n=n)
print('route A recipe done: n =', n)
