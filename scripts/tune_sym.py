"""Tune the symmetric path at C3 (GPU): outer block width x reserved SMs, 1 and 2 streams."""
import sys, time; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import _lib as L
from cnld import arrays
from cnld.sweep import FrequencySweep, SweepSetup
setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(4, 4)), 21)
freqs = np.linspace(0.025e6, 50e6, 2000)[[417, 903, 1311, 1777, 222, 1505]]
for ns in (3,):
    sw = FrequencySweep(setup, nstreams=ns)
    sw.ctx.set_symmetric(1, 1)
    sw.ctx.run_device(freqs[:2]) if hasattr(sw.ctx, 'run_device') else sw.run(freqs[:2])
    for outer in (512, 1024):
        for res in (12, 24):
            L.lib().cnld_b200_set_lu_outer(outer); L.lib().cnld_b200_set_lu_reserve(res)
            sw.run(freqs)
            t = sw.ctx.stage_times()
            print('streams', ns, 'outer', outer, 'reserve', res, 'asm %.1f fac %.1f sol %.1f ms/freq  -> %.3f freq/s' % (
                t['assemble'] / 6 * 1e3, t['factor'] / 6 * 1e3, t['solve'] / 6 * 1e3, 6 / t['device_seconds']), flush=True)
    sw.close()
