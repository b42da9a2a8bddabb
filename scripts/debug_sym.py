"""Symmetric path vs general path (GPU): LU kernel check, then sweep parity and timing."""
import sys, time; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib as L
rng = np.random.default_rng(0)
for n in (37, 64, 100, 200, 700, 1500):
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = A + A.T + n * 1.5 * np.eye(n)
    LU = L.lrdecomp(A)
    Al = np.tril(A) + 1e3 * np.triu(np.ones((n, n)), 1)   # garbage above the diagonal must be ignored
    LUs = L.lrdecomp_sym(Al)
    print('lu_sym', n, np.abs(LU - LUs).max() / np.abs(LU).max(), flush=True)
from cnld import arrays
from cnld.sweep import FrequencySweep, SweepSetup
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
for nelem, refn, fi in (((1, 1), 15, (1e6, 20e6)), ((2, 2), 15, (0.0, 3e6, 32e6)), ((4, 4), 21, (10.4e6, 33e6))):
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn)
    sw = FrequencySweep(setup, nstreams=1)
    freqs = np.array(fi)
    xp, xn, _ = sw.run(freqs)
    t0 = sw.ctx.stage_times()
    for steps in (0, 1, 2, 3, 4):
        sw.ctx.set_symmetric(1, steps)
        xps, xns, _ = sw.run(freqs)
        print(nelem, 'refine', steps, 'x_node rel', [float('%.2e' % rel(xns[i], xn[i])) for i in range(len(freqs))],
              'x_patch rel', [float('%.2e' % rel(xps[i], xp[i])) for i in range(len(freqs))], flush=True)
    t1 = sw.ctx.stage_times()
    print(nelem, 'general stages', t0, '\n  symmetric(4 steps) stages', t1, flush=True)
    sw.ctx.set_symmetric(0, 0)
    sw.close()
