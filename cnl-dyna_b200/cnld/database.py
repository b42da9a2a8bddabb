"""
cnld.database -- sqlite storage of frequency / impulse responses.  Same schema, table and column
names and reader return shapes as /root/reference/cnld/database.py (tables :22-92, create_db
:130-139, append_* :144-194, read_* :216-320) plus the progress helpers of
/root/reference/cnld/util.py:373-416, so databases are interchangeable.  Writers take whole arrays
and insert them in one transaction (the reference's row-by-row inserts under a
global lock would cap a GPU sweep).
"""
import os
import sqlite3 as sql
from contextlib import closing
from itertools import repeat

import numpy as np

for _t, _c in ((np.float64, float), (np.float32, float), (np.int64, int), (np.int32, int), (np.uint64, int),
               (np.uint32, int)):
    sql.register_adapter(_t, _c)

NODE = 'CREATE TABLE node (node_id integer primary key, x float, y float, z float)'
PATCH_TO_PATCH_FREQ_RESP = ('CREATE TABLE patch_to_patch_freq_resp (source_patch integer, dest_patch integer, '
                            'frequency float, wavenumber float, displacement_real float, displacement_imag float, '
                            'time_solve float)')
PATCH_TO_PATCH_FREQ_RESP_INDEX = ('CREATE INDEX patch_to_patch_freq_resp_index ON patch_to_patch_freq_resp '
                                  '(source_patch, dest_patch, frequency)')
PATCH_TO_NODE_FREQ_RESP = ('CREATE TABLE patch_to_node_freq_resp (source_patch integer, node_id integer, x float, '
                           'y float, z float, frequency float, wavenumber float, displacement_real float, '
                           'displacement_imag float, FOREIGN KEY (node_id, x, y, z) REFERENCES node (node_id, x, y, z))')
PATCH_TO_NODE_FREQ_RESP_INDEX = ('CREATE INDEX patch_to_node_freq_resp_index ON patch_to_node_freq_resp '
                                 '(source_patch, node_id, frequency)')
PATCH_TO_PATCH_IMP_RESP = ('CREATE TABLE patch_to_patch_imp_resp (source_patch integer, dest_patch integer, '
                           'time float, displacement float)')
PATCH_TO_PATCH_IMP_RESP_INDEX = ('CREATE INDEX patch_to_patch_imp_resp_index ON patch_to_patch_imp_resp '
                                 '(source_patch, dest_patch, time)')


def open_db(f):
    def decorator(first, *args, **kwargs):
        if isinstance(first, sql.Connection):
            return f(first, *args, **kwargs)
        with closing(sql.connect(first)) as con:
            return f(con, *args, **kwargs)
    return decorator


def read_db(f):
    def decorator(first, *args, **kwargs):
        if isinstance(first, sql.Connection):
            return f(first, *args, **kwargs)
        if not os.path.isfile(first):
            raise IOError('File does not exist')
        with closing(sql.connect(first)) as con:
            return f(con, *args, **kwargs)
    return decorator


@open_db
def create_table(con, table_defn, index_defn=None):
    with con:
        con.execute(table_defn)
        if index_defn is not None:
            con.execute(index_defn)


@open_db
def create_metadata_table(con, **kwargs):
    cols = list(kwargs.keys())
    with con:
        con.execute('DROP TABLE IF EXISTS metadata')
        con.execute('CREATE TABLE metadata (%s)' % ', '.join(f'"{c}" TEXT' for c in cols))
        con.execute('INSERT INTO metadata VALUES (%s)' % ','.join('?' * len(cols)), [str(v) for v in kwargs.values()])


@open_db
def create_db(con, **kwargs):
    create_metadata_table(con, **kwargs)
    create_table(con, NODE)
    create_table(con, PATCH_TO_PATCH_FREQ_RESP, PATCH_TO_PATCH_FREQ_RESP_INDEX)
    create_table(con, PATCH_TO_NODE_FREQ_RESP, PATCH_TO_NODE_FREQ_RESP_INDEX)
    create_table(con, PATCH_TO_PATCH_IMP_RESP, PATCH_TO_PATCH_IMP_RESP_INDEX)


def _rows(kwargs, keys):
    n = max(len(v) for v in kwargs.values() if hasattr(v, '__len__'))
    cols = [kwargs[k] if hasattr(kwargs[k], '__len__') else (kwargs[k] if isinstance(kwargs[k], repeat) else repeat(kwargs[k]))
            for k in keys]
    return zip(*[c if isinstance(c, repeat) else list(c)[:n] for c in cols])


@read_db
def append_node(con, **kwargs):
    with con:
        con.executemany('INSERT INTO node VALUES (?,?,?,?)', _rows(kwargs, ['node_id', 'x', 'y', 'z']))


@read_db
def append_patch_to_patch_freq_resp(con, **kwargs):
    keys = ['source_patch', 'dest_patch', 'frequency', 'wavenumber', 'displacement_real', 'displacement_imag',
            'time_solve']
    with con:
        con.executemany('INSERT INTO patch_to_patch_freq_resp VALUES (?,?,?,?,?,?,?)', _rows(kwargs, keys))


@read_db
def append_patch_to_node_freq_resp(con, **kwargs):
    keys = ['source_patch', 'node_id', 'x', 'y', 'z', 'frequency', 'wavenumber', 'displacement_real',
            'displacement_imag']
    with con:
        con.executemany('INSERT INTO patch_to_node_freq_resp VALUES (?,?,?,?,?,?,?,?,?)', _rows(kwargs, keys))


@read_db
def append_patch_to_patch_imp_resp(con, **kwargs):
    with con:
        con.executemany('INSERT INTO patch_to_patch_imp_resp VALUES (?,?,?,?)',
                        _rows(kwargs, ['source_patch', 'dest_patch', 'time', 'displacement']))


def append_patch_to_patch_imp_resp_block(file, times, ppir, threads=None):
    """All rows of patch_to_patch_imp_resp from ppir[P(src), P(dst), nt] -- the rows generate_db.py:143-164
    builds with meshgrid and appends through pandas -- by the C writer (csrc/dbwriter.cpp)."""
    from . import _lib
    if threads is None:
        threads = min(16, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    _lib.db_append_imp_resp(file, times, np.ascontiguousarray(ppir, dtype=np.float64), threads=max(1, int(threads)))


def append_frequency_block(file, freqs, wavenums, x_patch, x_node=None, nodes=None, times=None, job_ids=None,
                           threads=None):
    """One transaction for a block of frequencies: x_patch[nf, P(src), P(dst)] and optionally
    x_node[nf, P(src), N] -- the rows the reference writes one source patch at a time through Python
    (generate_db.py:101-127, database.py:144-194) -- plus the progress rows `job_ids`, written by the C
    bulk writer (prepared statements on libsqlite3, node rows staged by `threads` threads and appended by
    raw record copy; csrc/dbwriter.cpp).  Same tables, columns, values and row order."""
    from . import _lib
    if not os.path.isfile(file):
        raise IOError('File does not exist')
    if threads is None:
        threads = min(8, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    if index_exists(file, 'patch_to_node_freq_resp_index'):
        threads = 1          # the staged raw-copy append needs a target without indexes (begin_bulk)
    _lib.db_append_block(file, freqs, wavenums, x_patch, x_node, nodes, times, job_ids, fast=max(1, int(threads)))


@read_db
def append_frequency_block_py(con, freqs, wavenums, x_patch, x_node=None, nodes=None, times=None):
    """The same rows through sqlite3.executemany (kept as the cross-check of the C writer)."""
    nf, P, _ = x_patch.shape
    with con:
        for i in range(nf):
            src, dst = np.meshgrid(np.arange(P), np.arange(P), indexing='ij')
            t = float(times[i]) if times is not None else 0.0
            con.executemany('INSERT INTO patch_to_patch_freq_resp VALUES (?,?,?,?,?,?,?)',
                            zip(src.ravel().tolist(), dst.ravel().tolist(), repeat(float(freqs[i])),
                                repeat(float(wavenums[i])), x_patch[i].real.ravel().tolist(),
                                x_patch[i].imag.ravel().tolist(), repeat(t)))
            if x_node is not None:
                N = x_node.shape[2]
                ids = list(range(N))
                xs, ys, zs = (nodes[:, d].tolist() for d in range(3))
                for s in range(P):
                    con.executemany('INSERT INTO patch_to_node_freq_resp VALUES (?,?,?,?,?,?,?,?,?)',
                                    zip(repeat(s), ids, xs, ys, zs, repeat(float(freqs[i])), repeat(float(wavenums[i])),
                                        x_node[i, s].real.tolist(), x_node[i, s].imag.tolist()))


@open_db
def index_exists(con, name):
    return con.execute("SELECT count(*) FROM sqlite_master WHERE type='index' and name=?", (name,)).fetchone()[0] != 0


def begin_bulk(file):
    """Drop the two frequency-response indexes for the duration of a sweep (a bulk load maintains them
    row by row otherwise); finish_bulk rebuilds them, leaving the reference's schema (database.py:31-45)."""
    from . import _lib
    _lib.db_begin_bulk(file)


def finish_bulk(file):
    from . import _lib
    _lib.db_finish(file)


@read_db
def read_metadata(con, as_dict=False):
    cur = con.execute('SELECT * FROM metadata')
    cols = [d[0] for d in cur.description]
    row = cur.fetchone()
    if as_dict:
        return dict(zip(cols, row))
    import pandas as pd
    return pd.DataFrame([row], columns=cols)


@read_db
def read_node(con):
    t = np.array(con.execute('SELECT x, y, z FROM node ORDER BY node_id').fetchall(), dtype=np.float64)
    return t.T


@read_db
def read_patch_to_patch_freq_resp(con):
    t = np.array(con.execute('SELECT source_patch, dest_patch, frequency, displacement_real, displacement_imag '
                             'FROM patch_to_patch_freq_resp ORDER BY source_patch, dest_patch, frequency').fetchall())
    freqs = np.unique(t[:, 2])
    ns, nd = len(np.unique(t[:, 0])), len(np.unique(t[:, 1]))
    return freqs, (t[:, 3] + 1j * t[:, 4]).reshape((ns, nd, len(freqs)))


@read_db
def read_patch_to_node_freq_resp(con):
    t = np.array(con.execute('SELECT source_patch, node_id, x, y, z, frequency, displacement_real, displacement_imag '
                             'FROM patch_to_node_freq_resp ORDER BY source_patch, node_id, frequency').fetchall())
    freqs = np.unique(t[:, 5])
    sel = (t[:, 0] == 0) & (t[:, 5] == freqs[0])
    nodes = t[sel][:, 2:5]
    ns = len(np.unique(t[:, 0]))
    return freqs, (t[:, 6] + 1j * t[:, 7]).reshape((ns, len(nodes), len(freqs))), nodes


@read_db
def read_patch_to_patch_imp_resp(con):
    t = np.array(con.execute('SELECT source_patch, dest_patch, time, displacement FROM patch_to_patch_imp_resp '
                             'ORDER BY source_patch, dest_patch, time').fetchall())
    times = np.unique(t[:, 2])
    ns, nd = len(np.unique(t[:, 0])), len(np.unique(t[:, 1]))
    return times, t[:, 3].reshape((ns, nd, len(times)))


# progress table (util.py:373-416): checkpoint / resume of a sweep
@open_db
def table_exists(con, name):
    return con.execute("SELECT count(*) FROM sqlite_master WHERE type='table' and name=?", (name,)).fetchone()[0] != 0


@open_db
def create_progress_table(con, njobs):
    with con:
        con.execute('CREATE TABLE progress (job_id INTEGER PRIMARY KEY, is_complete boolean)')
        con.executemany('INSERT INTO progress (is_complete) VALUES (?)', repeat((False,), njobs))


@open_db
def get_progress(con):
    done = np.array([r[0] for r in con.execute('SELECT is_complete FROM progress ORDER BY job_id').fetchall()],
                    dtype=bool)
    return done, int(done.sum()) + 1


@open_db
def update_progress(con, job_id):
    with con:
        ids = job_id if hasattr(job_id, '__len__') else [job_id]
        con.executemany('UPDATE progress SET is_complete=1 WHERE job_id=?', [(int(j),) for j in ids])
