"""One C4 frequency (16x16 array, 2304 right-hand sides) through the H-matrix + GMRES sweep for several GMRES
batch widths: python scripts/time_hsweep_batch.py f_MHz b1,b2,... [restart]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
import numpy as np
from cnld import _lib, arrays
from cnld.sweep import SweepSetup
f = float(sys.argv[1]) * 1e6
bs = [int(v) for v in sys.argv[2].split(',')]
restart = int(sys.argv[3]) if len(sys.argv) > 3 else 50
setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(16, 16)), 19)
ctx = _lib.Context(setup.vertices, setup.triangles)
ctx.set_mbk(setup.K, setup.M, setup.B)
ctx.set_loads(setup.F, setup.AVG, setup.on_boundary)
H = _lib.HMatrix(ctx)
H.enable_coarse(16)
ref = None
for b in bs:
    for rep in range(2):
        xp, _, it, t = H.sweep(np.array([f]), eps_aca=1e-2, tol=1e-6, restart=restart, maxiter=300, batch=b, want_nodes=False)
    if ref is None:
        ref = xp
    print(json.dumps(dict(f=f, batch=b, restart=restart, seconds=float(t[0]), iters_mean=float(it.mean()), iters_max=int(it.max()),
                          rel_vs_first=float(np.linalg.norm(xp - ref) / np.linalg.norm(ref)))), flush=True)
    for nr in (b,):
        s = H.bench_matvec(nr, 3)
        print(json.dumps(dict(matvec_nrhs=nr, seconds=s, per_vector_us=s / nr * 1e6)), flush=True)
