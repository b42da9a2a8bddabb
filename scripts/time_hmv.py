"""Device-resident H-matvec timing: python scripts/time_hmv.py [nelem] [refn] [f_MHz] [nrhs,...]"""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
import numpy as np
from cnld import _lib, arrays, mesh
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 8
refn = int(sys.argv[2]) if len(sys.argv) > 2 else 15
f = float(sys.argv[3]) * 1e6 if len(sys.argv) > 3 else 5e6
nrhs_list = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [64, 96]
am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(ne, ne)), refn)
n = len(am.vertices)
ctx = _lib.Context(np.array(am.vertices), np.array(am.triangles))
H = _lib.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=64)
t0 = time.perf_counter()
H.assemble(2 * np.pi * f / 1500, 1e-2)
ta = time.perf_counter() - t0
info = H.info()
for nrhs in [(v, m) for m in (2, 4) for v in nrhs_list]:
    nrhs, mode = nrhs
    _lib.lib().cnld_b200_set_hmv_ywarps(mode)
    s = H.bench_matvec(nrhs, 3)
    print(json.dumps(dict(ywarps=mode, n=n, f=f, nrhs=nrhs, seconds=s, stored=info['stored'], maxrank=info['maxrank'], assemble_s=ta,
                          tflops_8=8.0 * info['stored'] * nrhs / s / 1e12, gbs_leafdata=16.0 * info['stored'] / s / 1e9)), flush=True)
