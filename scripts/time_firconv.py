"""HBM throughput of the time-domain FIR convolution at the C3 table size (P = 144 patches, 4000 taps)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
import numpy as np
from cnld import _lib
P, nfir = 144, 4000
rng = np.random.default_rng(0)
fir = rng.standard_normal((P, P, nfir))
tab = _lib.FirTable(fir)
for n in range(nfir + 10):
    tab.push(rng.standard_normal(P))
tab.base(2.5e-9)
reps = 50
t0 = time.perf_counter()
for _ in range(reps):
    x = tab.base(2.5e-9)
dt = (time.perf_counter() - t0) / reps
ref = 2.5e-9 * 0
print(json.dumps(dict(P=P, nfir=nfir, table_bytes=fir.nbytes, seconds_per_conv_incl_d2h=dt, gbs=fir.nbytes * (nfir - 1) / nfir / dt / 1e9)))
