"""End-to-end throughput of the drop-in driver: generate_db.main on the C3 array (4x4, refn 21) with node
displacements stored (2.13 M rows per frequency), dense path.  python scripts/time_generate_db.py [nfreq] [store_nodes]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
import numpy as np
from cnld import abstract, arrays, database
from cnld.scripts import generate_db

nfreq = int(sys.argv[1]) if len(sys.argv) > 1 else 12
store = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
tmp = tempfile.mkdtemp(dir='/dev/shm' if os.access('/dev/shm', os.W_OK) else None)
arr_file = os.path.join(tmp, 'array.json')
abstract.dump(arrays.matrix_array(nelem=(4, 4)), arr_file)
step = 200e3
cfg = generate_db.Config(freqs=(1e6, 1e6 + (nfreq - 1) * step, step), mesh_refn=21, array_config=arr_file, format='FullFormat')
db = os.path.join(tmp, 'resp.db')
t0 = time.perf_counter()
generate_db.main(cfg, db, block=6, store_nodes=store)
dt = time.perf_counter() - t0
f, ppfr = database.read_patch_to_patch_freq_resp(db)
print(json.dumps(dict(config='c3 generate_db end to end (setup + sweep + database + index build + impulse responses)',
                      frequencies=len(f), store_nodes=store, seconds=dt, frequencies_per_s=len(f) / dt,
                      db_bytes=os.path.getsize(db), cores=len(os.sched_getaffinity(0)))))
