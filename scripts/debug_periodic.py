"""Translation classes vs integrating every block (GPU)."""
import sys; sys.path.insert(0, 'cnl-dyna_b200')
import numpy as np
from cnld import arrays
from cnld.sweep import FrequencySweep, SweepSetup
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
for nelem, refn, fi in (((1, 1), 6, (1e6, 20e6)), ((2, 1), 5, (3e6, 17e6)), ((2, 2), 15, (0.0, 3e6, 32e6)), ((4, 4), 21, (10.4e6, 33e6))):
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn)
    sw = FrequencySweep(setup, nstreams=1, symmetric=False)
    freqs = np.array(fi)
    xp, xn, _ = sw.run(freqs)
    tg = sw.ctx.stage_times()
    sw.ctx.set_symmetric(1, 1)
    out = []
    for on in (0, 1):
        sw.ctx.set_translation_classes(on)
        xps, xns, _ = sw.run(freqs)
        out.append((on, sw.ctx.translation_info(), sw.ctx.symmetric_stats(), [float('%.2e' % rel(xns[i], xn[i])) for i in range(len(freqs))],
                    sw.ctx.stage_times()['assemble'] / len(freqs)))
    print(nelem, refn, 'general assemble %.4f' % (tg['assemble'] / len(freqs)))
    for o in out: print('   ', o, flush=True)
    sw.close()
