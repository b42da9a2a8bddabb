"""Batched band H-matvec vs the single-vector kernel and vs the dense expansion (GPU)."""
import sys, time; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib, arrays, mesh
rng = np.random.default_rng(0)
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
for nelem, refn, dense in (((2, 1), 6, True), ((3, 3), 9, True), ((8, 8), 15, False)):
    am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=nelem), refn)
    n = am.nvertices
    ctx = _lib.Context(np.array(am.vertices), np.array(am.triangles))
    H = _lib.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=64)
    H.assemble(2 * np.pi * 5e6 / 1500, 1e-3)
    Zd = H.todense() if dense else None
    for nrhs in (1, 4, 5, 31, 40, 70, 96, 100):
        X = rng.standard_normal((n, nrhs)) + 1j * rng.standard_normal((n, nrhs))
        Y = H.matvec(X, alpha=0.5 - 2j)
        ref = np.stack([H.matvec(X[:, j], alpha=0.5 - 2j) for j in range(min(nrhs, 6))], axis=1)
        t0 = time.perf_counter(); H.matvec(X); dt = time.perf_counter() - t0
        msg = 'vs single-vector %.2e' % rel(Y[:, :ref.shape[1]], ref)
        if dense:
            msg += '  vs dense %.2e' % rel(Y, (0.5 - 2j) * Zd @ X)
        print(nelem, n, 'nrhs', nrhs, msg, ' %.2f ms incl copies' % (dt * 1e3), flush=True)
    H.close(); ctx.close()
