"""Tune the symmetric path at C3 with many frequencies in flight (GPU): streams x outer block width x reserved SMs."""
import sys, time; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib as L
from cnld import arrays
from cnld.sweep import FrequencySweep, SweepSetup
setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(4, 4)), 21)
allf = np.linspace(0.025e6, 50e6, 2000)
rng = np.random.default_rng(1)
for ns in (6, 8, 4):
    sw = FrequencySweep(setup, nstreams=ns)
    sw.ctx.set_symmetric(1, 1)
    nf = 2 * ns
    freqs = allf[rng.choice(2000, nf, replace=False)]
    sw.run_device(freqs[:ns])
    for outer in (512, 768, 1024):
        for res in (0, 6, 12, 24):
            L.lib().cnld_b200_set_lu_outer(outer); L.lib().cnld_b200_set_lu_reserve(res)
            best = 0
            for rep in range(2):
                sw.run_device(freqs)
                t = sw.ctx.stage_times()
                best = max(best, nf / t['device_seconds'])
            print('streams', ns, 'outer', outer, 'reserve', res, 'asm %.1f fac %.1f sol %.1f ms/freq  -> %.3f freq/s' % (
                t['assemble'] / nf * 1e3, t['factor'] / nf * 1e3, t['solve'] / nf * 1e3, best), flush=True)
    L.lib().cnld_b200_set_lu_outer(512); L.lib().cnld_b200_set_lu_reserve(12)
    sw.close()
