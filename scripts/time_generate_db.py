"""End-to-end throughput of the drop-in driver: generate_db.main on the C3 array (4x4, refn 21) with node
displacements stored (2.13 M rows per frequency), dense path.
python scripts/time_generate_db.py [nfreq] [store_nodes] [block] [sql]"""
import json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
import numpy as np
from cnld import _lib, abstract, arrays, database
from cnld.scripts import generate_db

nfreq = int(sys.argv[1]) if len(sys.argv) > 1 else 48
store = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
block = int(sys.argv[3]) if len(sys.argv) > 3 else 12
if len(sys.argv) > 4 and sys.argv[4] == 'sql':
    os.environ['CNLD_B200_DB_MODE'] = 'sql'          # the round-2 SQL writer (staged prepared statements + CREATE INDEX)
need = (nfreq * 260e6 if store else nfreq * 12e6) + 1e9
base = None
for d in ('/dev/shm', tempfile.gettempdir()):
    if os.access(d, os.W_OK):
        st = os.statvfs(d)
        if st.f_bavail * st.f_frsize > need:
            base = d
            break
if base is None:
    raise SystemExit(f'no directory with {need / 1e9:.1f} GB free')
tmp = tempfile.mkdtemp(dir=base)
try:
    arr_file = os.path.join(tmp, 'array.json')
    abstract.dump(arrays.matrix_array(nelem=(4, 4)), arr_file)
    step = 200e3
    cfg = generate_db.Config(freqs=(1e6, 1e6 + (nfreq - 1) * step, step), mesh_refn=21, array_config=arr_file, format='FullFormat')
    db = os.path.join(tmp, 'resp.db')
    t0 = time.perf_counter()
    generate_db.main(cfg, db, block=block, store_nodes=store, nstreams=6)
    dt = time.perf_counter() - t0
    f, ppfr = database.read_patch_to_patch_freq_resp(db)
    import sqlite3
    con = sqlite3.connect(db)
    nrows = con.execute('SELECT max(rowid) FROM patch_to_node_freq_resp').fetchone()[0]
    quick = con.execute('PRAGMA quick_check').fetchall() if nfreq <= 24 else None
    con.close()
    print(json.dumps(dict(config='c3 generate_db end to end (setup + sweep + database + index build + impulse responses)',
                          frequencies=len(f), store_nodes=store, block=block, seconds=dt, frequencies_per_s=len(f) / dt,
                          stages=generate_db.LAST_TIMING, node_rows=nrows, db_bytes=os.path.getsize(db), db_dir=base,
                          writer=_lib.db_counters(), quick_check=quick, cores=len(os.sched_getaffinity(0)))))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
