"""Outer-block-inverse solve vs substitution (GPU): parity + timing at C3, LU sizes around the block sizes."""
import sys; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib as L
from cnld import arrays
from cnld.sweep import FrequencySweep, SweepSetup
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
for nelem, refn, fi in (((2, 2), 15, (0.0, 3e6, 32e6)), ((4, 4), 21, (10.4e6, 33e6))):
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn)
    for sym in (True, False):
        sw = FrequencySweep(setup, nstreams=1, symmetric=sym)
        freqs = np.array(fi)
        res = []
        for inv in (0, 1):
            L.lib().cnld_b200_set_lu_outer_inverse(inv)
            xp, xn, _ = sw.run(freqs)
            res.append((xn, sw.ctx.stage_times()))
        print(nelem, 'sym' if sym else 'general', 'inverse-vs-substitution rel', rel(res[1][0], res[0][0]),
              'solve ms/freq: %.2f -> %.2f' % (res[0][1]['solve'] / len(freqs) * 1e3, res[1][1]['solve'] / len(freqs) * 1e3),
              'factor ms/freq: %.2f -> %.2f' % (res[0][1]['factor'] / len(freqs) * 1e3, res[1][1]['factor'] / len(freqs) * 1e3), flush=True)
        sw.close()
