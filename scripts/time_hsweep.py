"""One C4 frequency (16x16 array, 2304 right-hand sides) through the H-matrix + GMRES sweep with a fixed number of
coarse modes per membrane: python scripts/time_hsweep.py f_MHz m1,m2,...   (m = 0: block Jacobi only)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
import numpy as np
from cnld import _lib, arrays
from cnld.sweep import SweepSetup
f = float(sys.argv[1]) * 1e6
ms = [int(v) for v in sys.argv[2].split(',')]
ne = int(sys.argv[3]) if len(sys.argv) > 3 else 16
refn = int(sys.argv[4]) if len(sys.argv) > 4 else 19
setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(ne, ne)), refn)
ctx = _lib.Context(setup.vertices, setup.triangles)
ctx.set_mbk(setup.K, setup.M, setup.B)
ctx.set_loads(setup.F, setup.AVG, setup.on_boundary)
H = _lib.HMatrix(ctx)
fm = H.enable_coarse(16)
print('mode frequencies MHz', (fm / 1e6).round(1).tolist())
ref = None
for m in ms:
    _lib.lib().cnld_b200_set_hsweep_coarse(m if m else 0)
    xp, _, it, t = H.sweep(np.array([f]), eps_aca=1e-2, tol=1e-6, restart=50, maxiter=300, batch=96, want_nodes=False)
    if ref is None:
        ref = xp
    print(json.dumps(dict(f=f, coarse=m, used=H.coarse_used(), seconds=float(t[0]), iters_mean=float(it.mean()), iters_max=int(it.max()),
                          rel_vs_first=float(np.linalg.norm(xp - ref) / np.linalg.norm(ref)))), flush=True)
