"""One batched H-matvec at 8x8 / refn 15 for ncu: python scripts/time_hmv.py [nrhs]"""
import sys; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib, arrays, mesh
nrhs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(8, 8)), 15)
n = am.nvertices
ctx = _lib.Context(np.array(am.vertices), np.array(am.triangles))
H = _lib.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=int(sys.argv[2]) if len(sys.argv) > 2 else 64)
H.assemble(2 * np.pi * 5e6 / 1500, 1e-2)
rng = np.random.default_rng(0)
X = rng.standard_normal((n, nrhs)) + 1j * rng.standard_normal((n, nrhs))
for _ in range(3):
    Y = H.matvec(X)
print(np.linalg.norm(Y))
