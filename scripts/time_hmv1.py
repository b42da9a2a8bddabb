"""Single-vector H-matvec at the 16x16 array (N = 194 816) for ncu / bandwidth: python scripts/time_hmv1.py"""
import sys, time; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib, arrays, mesh
nelem = (16, 16) if len(sys.argv) < 2 else (int(sys.argv[1]),) * 2
am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=nelem), 19 if nelem[0] == 16 else 15)
n = am.nvertices
ctx = _lib.Context(np.array(am.vertices), np.array(am.triangles))
H = _lib.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=int(sys.argv[2]) if len(sys.argv) > 2 else 64)
H.assemble(2 * np.pi * 5e6 / 1500, 1e-2)
info = H.info()
rng = np.random.default_rng(0)
x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
for _ in range(3):
    t0 = time.perf_counter(); y = H.matvec(x); dt = time.perf_counter() - t0
rows = rng.choice(n, 32, replace=False).astype(np.uint32)
exact = ctx.nearfield(2 * np.pi * 5e6 / 1500, rows, np.arange(n, dtype=np.uint32)) @ x
print('n', n, 'stored entries', info['stored'], 'bytes %.3e' % (16 * info['stored']), 'matvec incl copies %.3f ms' % (dt * 1e3),
      'sampled-row error %.2e' % (np.linalg.norm(y[rows] - exact) / np.linalg.norm(exact)))
