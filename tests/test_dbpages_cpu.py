"""The direct b-tree page writer of the node-response table (csrc/dbpages.cpp) against SQLite itself:
files written as pages must pass PRAGMA integrity_check and hold, row for row and rowid for rowid, what the
SQL path (prepared statements through libsqlite3) writes for the same arrays -- including SQLite's storage
rules for REAL columns (integral values as integers, NaN as NULL) -- and the index generated from the block
descriptors must be the one CREATE INDEX builds (same schema row, used by the planner, same query results).
Reference rows and schema: /root/reference/cnld/database.py:22-92, cnld/scripts/generate_db.py:101-130."""
import os
import sqlite3

import numpy as np
import pytest


def _make(path, page_size=None):
    from cnld import database
    if page_size:
        con = sqlite3.connect(path)
        con.execute(f'PRAGMA page_size={page_size}')
        con.execute('VACUUM')
        con.close()
    database.create_db(path, note='dbpages test')
    database.create_progress_table(path, 16)
    return path


def _block(nf, P, N, seed):
    r = np.random.default_rng(seed)
    freqs = 1e6 + 200e3 * np.arange(nf) + (0.5 if seed % 2 else 0) + 1e6 * seed   # integral and fractional values
    wn = 2 * np.pi * freqs / 1500
    xp = r.standard_normal((nf, P, P)) + 1j * r.standard_normal((nf, P, P))
    xn = r.standard_normal((nf, P, N)) + 1j * r.standard_normal((nf, P, N))
    xn[:, :, ::7] = 0                      # masked boundary nodes: exact zeros
    xn[:, :, min(3, N - 1)] = 1.0
    xn[0, 0, N - 1] = np.nan + 0j          # SQLite stores NaN as NULL
    nodes = r.standard_normal((N, 3))
    nodes[:, 2] = 0
    nodes[1 % N, 0] = 3.0
    nodes[2 % N, 1] = -2e15                # integral but beyond 48 bits: stays a float
    return freqs, wn, xp, xn, nodes


def _dump(path):
    con = sqlite3.connect(path)
    out = dict(
        rows=con.execute('SELECT rowid, * FROM patch_to_node_freq_resp ORDER BY rowid').fetchall(),
        check=con.execute('PRAGMA integrity_check').fetchall(),
        index=con.execute("SELECT name, tbl_name, sql FROM sqlite_master WHERE type='index' ORDER BY name").fetchall(),
        by_index=con.execute('SELECT rowid, frequency FROM patch_to_node_freq_resp WHERE source_patch=0 AND node_id=2 '
                             'ORDER BY frequency').fetchall(),
        plan=con.execute('EXPLAIN QUERY PLAN SELECT rowid FROM patch_to_node_freq_resp WHERE source_patch=0 AND '
                         'node_id=2 ORDER BY frequency').fetchall(),
        ordered=con.execute('SELECT rowid FROM patch_to_node_freq_resp ORDER BY source_patch, node_id, frequency').fetchall(),
        patch=con.execute('SELECT rowid, * FROM patch_to_patch_freq_resp ORDER BY rowid').fetchall(),
        patch_ordered=con.execute('SELECT rowid FROM patch_to_patch_freq_resp ORDER BY source_patch, dest_patch, frequency').fetchall(),
        patch_plan=con.execute('EXPLAIN QUERY PLAN SELECT displacement_real FROM patch_to_patch_freq_resp WHERE source_patch=0 AND '
                               'dest_patch=0 ORDER BY frequency').fetchall(),
        progress=con.execute('SELECT * FROM progress ORDER BY job_id').fetchall(),
        types=con.execute('SELECT typeof(frequency), typeof(z), typeof(displacement_real), count(*) FROM '
                          'patch_to_node_freq_resp GROUP BY 1, 2, 3 ORDER BY 1, 2, 3').fetchall())
    con.close()
    return out


def _same_rows(a, b):
    if len(a) != len(b):
        return False
    for ra, rb in zip(a, b):
        for x, y in zip(ra, rb):
            if x != y and not (x is None and y is None):
                return False
    return True


def _write(path, blocks, mode, monkeypatch, threads=4, bulk=True):
    from cnld import database
    monkeypatch.setenv('CNLD_B200_DB_MODE', mode)
    if bulk:
        database.begin_bulk(path)
    job = 1
    for fr, wn, xp, xn, nodes in blocks:
        database.append_frequency_block(path, fr, wn, xp, xn, nodes, None, job_ids=np.arange(len(fr)) + job, threads=threads)
        job += len(fr)
    if bulk:
        database.finish_bulk(path)


@pytest.mark.parametrize('page_size', [None, 512, 1024, 65536])
@pytest.mark.parametrize('shape', [(3, 50, 2), (1, 5, 1), (7, 1300, 3)])
def test_pages_equal_sql_path(tmp_path, monkeypatch, page_size, shape):
    P, N, nf = shape
    blocks = [_block(nf, P, N, b) for b in range(2)]
    from cnld import _lib
    out, used = {}, {}
    for mode in ('pages', 'sql'):
        f = _make(str(tmp_path / f'{mode}.db'), page_size)
        c0 = _lib.db_counters()
        _write(f, blocks, mode, monkeypatch)
        c1 = _lib.db_counters()
        used[mode] = {k: c1[k] - c0[k] for k in c0}
        out[mode] = _dump(f)
    rows = 2 * nf * P * N
    # the page path was really taken (no silent fallback), and the SQL path is really SQL
    prow = 2 * nf * P * P
    assert used['pages'] == dict(rows_paged=rows, rows_sql=0, index_direct=1, index_sql=0, imp_rows_paged=0, imp_rows_sql=0,
                                 patch_rows_paged=prow, patch_rows_sql=0)
    assert used['sql'] == dict(rows_paged=0, rows_sql=rows, index_direct=0, index_sql=1, imp_rows_paged=0, imp_rows_sql=0,
                               patch_rows_paged=0, patch_rows_sql=prow)
    a, b = out['pages'], out['sql']
    assert a['check'] == [('ok',)] and b['check'] == [('ok',)]
    assert len(a['rows']) == 2 * nf * P * N and _same_rows(a['rows'], b['rows'])
    assert a['index'] == b['index']
    assert any('patch_to_node_freq_resp_index' in r[0] for r in a['index'])
    assert a['by_index'] == b['by_index'] and len(a['by_index']) == 2 * nf
    assert 'patch_to_node_freq_resp_index' in a['plan'][0][-1]
    assert a['ordered'] == b['ordered']
    assert a['types'] == b['types']           # integral REALs read back as 'real' in both
    assert _same_rows(a['patch'], b['patch']) and a['progress'] == b['progress']
    assert a['patch_ordered'] == b['patch_ordered'] and 'patch_to_patch_freq_resp_index' in a['patch_plan'][0][-1]


def test_pages_after_sql_rows_and_resume(tmp_path, monkeypatch):
    """Rows that went in through SQL first (an interrupted sweep written by another build, or the SQL fallback):
    the page writer appends behind SQLite's own b-tree (moves a root leaf, reuses / frees interior pages) and
    the index falls back to CREATE INDEX because the descriptors do not cover the table."""
    P, N, nf = 4, 700, 2
    blocks = [_block(nf, P, N, b) for b in range(4)]
    fa, fb = _make(str(tmp_path / 'a.db')), _make(str(tmp_path / 'b.db'))
    from cnld import _lib
    _write(fa, blocks[:1], 'sql', monkeypatch)            # SQLite builds a multi-level tree
    c0 = _lib.db_counters()
    _write(fa, blocks[1:3], 'pages', monkeypatch)         # resumed: begin_bulk sees rows -> SQL index at finish
    _write(fa, blocks[3:], 'pages', monkeypatch)
    c1 = _lib.db_counters()
    assert {k: c1[k] - c0[k] for k in c0} == dict(rows_paged=3 * nf * P * N, rows_sql=0, index_direct=0, index_sql=2, imp_rows_paged=0,
                                                  imp_rows_sql=0, patch_rows_paged=3 * nf * P * P, patch_rows_sql=0)
    _write(fb, blocks, 'sql', monkeypatch)
    a, b = _dump(fa), _dump(fb)
    assert a['check'] == [('ok',)]
    assert _same_rows(a['rows'], b['rows']) and a['ordered'] == b['ordered'] and a['index'] == b['index']
    assert _same_rows(a['patch'], b['patch']) and a['patch_ordered'] == b['patch_ordered']
    # a root that is still a leaf with cells when the page writer arrives
    fc, fd = _make(str(tmp_path / 'c.db')), _make(str(tmp_path / 'd.db'))
    tiny = [_block(1, 1, 3, 9)]
    _write(fc, tiny, 'sql', monkeypatch)
    _write(fc, blocks[:1], 'pages', monkeypatch)
    _write(fd, tiny + blocks[:1], 'sql', monkeypatch)
    c, d = _dump(fc), _dump(fd)
    assert c['check'] == [('ok',)] and _same_rows(c['rows'], d['rows']) and c['ordered'] == d['ordered']


def test_pages_without_bulk_bracket_and_with_index_present(tmp_path, monkeypatch):
    """No begin_bulk: the node index exists, so the page path does not apply and the rows go through SQL."""
    P, N, nf = 2, 40, 2
    blocks = [_block(nf, P, N, 0)]
    fa, fb = _make(str(tmp_path / 'a.db')), _make(str(tmp_path / 'b.db'))
    from cnld import _lib
    c0 = _lib.db_counters()
    _write(fa, blocks, 'pages', monkeypatch, bulk=False)
    assert _lib.db_counters()['rows_paged'] == c0['rows_paged']
    _write(fb, blocks, 'sql', monkeypatch, bulk=False)
    a, b = _dump(fa), _dump(fb)
    assert a['check'] == [('ok',)] and _same_rows(a['rows'], b['rows']) and a['index'] == b['index']


def test_pages_refuse_wal_and_autovacuum(tmp_path, monkeypatch):
    """Files the page writer must not touch (WAL journal, auto-vacuum pointer maps) take the SQL path."""
    from cnld import database
    P, N, nf = 2, 300, 1
    blocks = [_block(nf, P, N, 0)]
    for pragma in ('PRAGMA journal_mode=WAL', 'PRAGMA auto_vacuum=FULL'):
        f = str(tmp_path / (pragma.split('=')[1] + '.db'))
        con = sqlite3.connect(f)
        con.execute(pragma)
        con.execute('VACUUM')
        con.close()
        database.create_db(f, note='x')
        database.create_progress_table(f, 4)
        from cnld import _lib
        c0 = _lib.db_counters()
        _write(f, blocks, 'pages', monkeypatch)
        assert _lib.db_counters()['rows_paged'] == c0['rows_paged']
        d = _dump(f)
        assert d['check'] == [('ok',)] and len(d['rows']) == nf * P * N
        assert any('patch_to_node_freq_resp_index' in r[0] for r in d['index'])


def test_pages_through_generate_db_readers(tmp_path, monkeypatch):
    """What the reference's readers return (database.py:262-288) from a page-written file."""
    from cnld import database
    P, N, nf = 3, 60, 4
    fr, wn, xp, xn, nodes = _block(nf, P, N, 2)
    xn[0, 0, N - 1] = 0
    f = _make(str(tmp_path / 'r.db'))
    _write(f, [(fr, wn, xp, xn, nodes)], 'pages', monkeypatch)
    freqs, resp, nd = database.read_patch_to_node_freq_resp(f)
    assert np.array_equal(freqs, fr) and np.array_equal(nd, nodes)
    assert np.array_equal(resp, np.transpose(xn, (1, 2, 0)))
    assert os.path.getsize(f) % 4096 == 0


@pytest.mark.parametrize('page_size', [None, 512, 65536])
def test_imp_resp_pages_equal_sql_and_python_writer(tmp_path, monkeypatch, page_size):
    """patch_to_patch_imp_resp (database.py:47-53): page writer == prepared statements == the reference-style
    executemany over meshgrid rows (generate_db.py:150-164), index included."""
    from cnld import database
    P, nt = 5, 331
    r = np.random.default_rng(3)
    t = np.arange(nt) / 100e6
    h = r.standard_normal((P, P, nt))
    h[1, 2, :7] = 0
    h[0, 0, 5] = 2.0
    out = {}
    for mode in ('pages', 'sql', 'python'):
        f = _make(str(tmp_path / f'{mode}.db'), page_size)
        if mode == 'python':
            src, dst, times = np.meshgrid(np.arange(P), np.arange(P), t, indexing='ij')
            database.append_patch_to_patch_imp_resp(f, source_patch=src.ravel(), dest_patch=dst.ravel(), time=times.ravel(),
                                                    displacement=h.ravel())
        else:
            from cnld import _lib
            monkeypatch.setenv('CNLD_B200_DB_MODE', mode)
            c0 = _lib.db_counters()
            database.append_patch_to_patch_imp_resp_block(f, t, h, threads=3)
            c1 = _lib.db_counters()
            assert c1['imp_rows_paged'] - c0['imp_rows_paged'] == (P * P * nt if mode == 'pages' else 0)
            assert c1['imp_rows_sql'] - c0['imp_rows_sql'] == (P * P * nt if mode == 'sql' else 0)
        con = sqlite3.connect(f)
        out[mode] = dict(
            rows=con.execute('SELECT rowid, * FROM patch_to_patch_imp_resp ORDER BY rowid').fetchall(),
            check=con.execute('PRAGMA integrity_check').fetchall(),
            index=con.execute("SELECT name, tbl_name, sql FROM sqlite_master WHERE type='index' ORDER BY name").fetchall(),
            ordered=con.execute('SELECT rowid FROM patch_to_patch_imp_resp ORDER BY source_patch, dest_patch, time').fetchall(),
            plan=con.execute('EXPLAIN QUERY PLAN SELECT displacement FROM patch_to_patch_imp_resp WHERE source_patch=1 AND '
                             'dest_patch=2 ORDER BY time').fetchall())
        con.close()
        tt, ppir = database.read_patch_to_patch_imp_resp(f)
        assert np.array_equal(tt, t) and np.array_equal(ppir, h)
    for mode in ('pages', 'sql'):
        assert out[mode]['check'] == [('ok',)]
        assert _same_rows(out[mode]['rows'], out['python']['rows'])
        assert out[mode]['index'] == out['python']['index'] and out[mode]['ordered'] == out['python']['ordered']
        assert 'patch_to_patch_imp_resp_index' in out[mode]['plan'][0][-1]
    # a second call finds rows: SQL path behind the existing pages, still consistent
    f = str(tmp_path / 'pages.db')
    monkeypatch.setenv('CNLD_B200_DB_MODE', 'pages')
    database.append_patch_to_patch_imp_resp_block(f, t, h, threads=3)
    con = sqlite3.connect(f)
    assert con.execute('PRAGMA integrity_check').fetchall() == [('ok',)]
    assert con.execute('SELECT count(*) FROM patch_to_patch_imp_resp').fetchone()[0] == 2 * P * P * nt
    con.close()
