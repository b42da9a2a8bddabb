"""Field pressure of the 4x4 array on a grid of observation points (for ncu): python scripts/time_pressure.py [npts_per_side]"""
import sys, time; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib, arrays, mesh
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100
am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(4, 4)), 21)
ctx = _lib.Context(np.array(am.vertices), np.array(am.triangles))
g = np.linspace(-1e-3, 1e-3, m); z = np.linspace(0.1e-3, 2.1e-3, m)
X, Y, Z = np.meshgrid(g, g, z, indexing='ij')
obs = np.c_[X.ravel('F'), Y.ravel('F'), Z.ravel('F')]
rng = np.random.default_rng(0)
disp = rng.standard_normal((1, 1, am.nvertices)) + 1j * rng.standard_normal((1, 1, am.nvertices))
k = 2 * np.pi * 5e6 / 1500
for _ in range(2):
    t0 = time.perf_counter(); p = ctx.pressure(obs, [k], disp); dt = time.perf_counter() - t0
print('points', len(obs), 'seconds incl copies %.4f' % dt, 'kernel evaluations/s %.3e' % (3.0 * am.ntriangles * len(obs) / dt))
