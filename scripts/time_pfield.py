"""Time the many-source field-pressure path (k_pvec_tiles variants + ZGEMM contraction) on the C3 mesh."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
from cnld import _lib, arrays, mesh

nobs = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
variants = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else range(8)
am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(4, 4)), 21)
ctx = _lib.Context(am.vertices, am.triangles)
rng = np.random.default_rng(0)
obs = np.c_[rng.uniform(-1e-3, 1e-3, nobs), rng.uniform(-1e-3, 1e-3, nobs), rng.uniform(1e-4, 2.1e-3, nobs)]
disp = rng.standard_normal((144, len(am.vertices))) + 1j * rng.standard_normal((144, len(am.vertices)))
k = 2 * np.pi * 20e6 / 1500
ref = None
for v in variants:
    _lib.lib().cnld_b200_set_pvec_variant(v)
    pf = _lib.PressureField(ctx, obs, 144)
    pf.run(k, disp, fetch=False)
    pf.run_device(k, 144)
    pf.run_device(k, 144)
    inf = pf.info()
    got = pf.fetch(144, np.arange(0, nobs, 997, dtype=np.uint32))
    if ref is None:
        ref = got
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(json.dumps(dict(variant=v, nobs=nobs, evals_per_s=3 * len(am.triangles) * nobs / inf['pvec_seconds'], err_vs_first=err, **inf)), flush=True)
    pf.close()
