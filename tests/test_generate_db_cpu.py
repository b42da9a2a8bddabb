"""Host logic of the sweep driver on CPU: the generate_db mirror with the GPU sweep replaced by an
oracle-backed stand-in -- block loop, background database writer, progress / resume, error propagation."""
import numpy as np
import pytest


class _FakeSweep:
    """FrequencySweep stand-in: the oracle solves this rank's frequencies on the host."""
    calls = 0

    def __init__(self, setup, device=0, nstreams=2, **kw):
        import pyoracle as po
        self.setup, self.po = setup, po
        self.am = po.Mesh(setup.vertices, np.zeros((1, 2), np.uint32), setup.triangles, np.zeros_like(setup.triangles))
        self.am.on_boundary = setup.on_boundary.astype(bool)

    def run(self, freqs, want_nodes=True, out_nodes=None):
        s, po = self.setup, self.po
        type(self).calls += 1
        K, M, B, F, A = (m.toarray() for m in (s.K, s.M, s.B, s.F, s.AVG))
        xp, xn = [], []
        for f in freqs:
            w = 2 * np.pi * f
            G = (K - w**2 * M - 1j * w * B).astype(np.complex128)
            if f != 0:
                G = G + (-w**2 * 2 * s.rho) * po.assemble_z(self.am, w / s.c, s.q_reg, s.q_sing)
            X = np.conj(np.linalg.solve(G, F.astype(np.complex128)))
            X[self.am.on_boundary] = 0
            xn.append(X.T)
            xp.append((A.T @ X).T)
        return np.array(xp), (np.array(xn) if want_nodes else None), np.zeros(len(freqs))

    def close(self):
        pass


def test_generate_db_block_loop_writer_thread_and_resume(tmp_path, po, monkeypatch):
    from cnld import abstract, arrays, database
    from cnld.scripts import generate_db
    monkeypatch.setattr(generate_db, 'FrequencySweep', _FakeSweep)
    arr = arrays.matrix_array(nelem=(1, 1))
    cfg_file = str(tmp_path / 'array.json')
    abstract.dump(arr, cfg_file)
    cfg = generate_db.Config(freqs=(1e6, 9e6, 2e6), mesh_refn=3, array_config=cfg_file, format='FullFormat')
    db = str(tmp_path / 'resp.db')
    _FakeSweep.calls = 0
    generate_db.main(cfg, db, block=2)
    assert _FakeSweep.calls == 3                       # 5 frequencies in blocks of 2
    freqs, ppfr = database.read_patch_to_patch_freq_resp(db)
    assert np.allclose(freqs, [1e6, 3e6, 5e6, 7e6, 9e6]) and ppfr.shape == (9, 9, 5)
    _, pnfr, nodes = database.read_patch_to_node_freq_resp(db)
    setup = generate_db.SweepSetup.from_abstract(arr, 3)
    xp, xn, _ = _FakeSweep(setup).run(freqs)
    assert np.allclose(nodes, setup.vertices)
    for i in range(len(freqs)):                        # rows land under the right frequency, in order
        assert np.allclose(ppfr[:, :, i], xp[i])
        assert np.allclose(pnfr[:, :, i], xn[i])
    done, _ = database.get_progress(db)
    assert done.all()
    before = _FakeSweep.calls
    generate_db.main(cfg, db)                          # finished database: nothing is recomputed
    assert _FakeSweep.calls == before
    # the finished file has the reference's indexes again (dropped only for the bulk load)
    assert database.index_exists(db, 'patch_to_node_freq_resp_index') and database.index_exists(db, 'patch_to_patch_freq_resp_index')


def test_generate_db_reference_entry_point_and_cli(tmp_path, po, monkeypatch):
    """main(cfg, args) with the reference's argparse namespace (generate_db.py:167-172), the -g / -c / -w
    command line of util.script_parser2 (util.py:466-506), the reference's Config defaults, and unknown
    formats refused instead of silently taking the dense path."""
    import argparse
    from cnld import abstract, arrays, database
    from cnld.scripts import generate_db
    monkeypatch.setattr(generate_db, 'FrequencySweep', _FakeSweep)
    assert generate_db.Config().format == 'HFormat' and generate_db.Config().eps_aca == 1e-2
    arr = arrays.matrix_array(nelem=(1, 1))
    arr_file = str(tmp_path / 'array.json')
    abstract.dump(arr, arr_file)
    parser = generate_db.script_parser()
    cfg_file = str(tmp_path / 'cfg.json')
    a = parser.parse_args(['-g', cfg_file])
    a.func(a)
    cfg = abstract.load(cfg_file)
    cfg = generate_db.Config(**(cfg._asdict() if hasattr(cfg, '_asdict') else dict(cfg)))
    assert cfg.format == 'HFormat'
    cfg.freqs, cfg.mesh_refn, cfg.array_config, cfg.format = (2e6, 4e6, 2e6), 3, arr_file, 'FullFormat'
    abstract.dump(cfg, cfg_file)
    db = str(tmp_path / 'r.db')
    a = parser.parse_args([db, '-c', cfg_file, '-w', '-t', '1'])
    assert (a.file, a.write_over, a.threads) == (db, True, 1)
    a.func(a)
    freqs, ppfr = database.read_patch_to_patch_freq_resp(db)
    assert np.allclose(freqs, [2e6, 4e6]) and ppfr.shape == (9, 9, 2)
    generate_db.main(cfg, argparse.Namespace(file=db, write_over=False, threads=None))     # resume: complete
    cfg.format = 'SparseFormat'
    with pytest.raises(ValueError, match='cfg.format'):
        generate_db.main(cfg, db, write_over=True)


def test_c_bulk_writer_equals_python_writer(tmp_path):
    """cnld_b200_db_append_block (single connection and staged with threads) writes exactly the rows of
    the executemany path, in the same order, and marks the progress rows."""
    import sqlite3
    from cnld import database
    rng = np.random.default_rng(0)
    nf, P, N = 3, 5, 37
    xp = rng.standard_normal((nf, P, P)) + 1j * rng.standard_normal((nf, P, P))
    xn = rng.standard_normal((nf, P, N)) + 1j * rng.standard_normal((nf, P, N))
    nodes, freqs, t = rng.standard_normal((N, 3)), np.array([1e6, 2.5e6, 4e6]), np.array([0.1, 0.2, 0.3])
    wn = 2 * np.pi * freqs / 1500.
    tables = {}
    for mode in ('py', 'c1', 'c4'):
        f = str(tmp_path / f'{mode}.db')
        database.create_db(f, a=1)
        database.create_progress_table(f, nf)
        if mode == 'py':
            database.append_frequency_block_py(f, freqs, wn, xp, xn, nodes, t)
        else:
            if mode == 'c4':
                database.begin_bulk(f)
            database.append_frequency_block(f, freqs, wn, xp, xn, nodes, t, job_ids=np.array([1, 3]), threads=int(mode[1]))
            done, _ = database.get_progress(f)
            assert done.tolist() == [True, False, True]
            database.finish_bulk(f)
        con = sqlite3.connect(f)
        tables[mode] = (con.execute('SELECT * FROM patch_to_patch_freq_resp').fetchall(),
                        con.execute('SELECT * FROM patch_to_node_freq_resp').fetchall())
        con.close()
    assert tables['py'] == tables['c1'] == tables['c4'] and len(tables['py'][1]) == nf * P * N
    # patch responses only (x_node = None), and a missing file
    f = str(tmp_path / 'p.db')
    database.create_db(f, a=1)
    database.append_frequency_block(f, freqs, wn, xp)
    assert len(database.read_patch_to_patch_freq_resp(f)[0]) == nf
    with pytest.raises(IOError):
        database.append_frequency_block(str(tmp_path / 'missing.db'), freqs, wn, xp)


def test_generate_db_writer_failure_reaches_the_caller(tmp_path, po, monkeypatch):
    from cnld import abstract, arrays, database
    from cnld.scripts import generate_db
    monkeypatch.setattr(generate_db, 'FrequencySweep', _FakeSweep)

    def boom(*a, **k):
        raise OSError('disk full')

    arr = arrays.matrix_array(nelem=(1, 1))
    cfg_file = str(tmp_path / 'array.json')
    abstract.dump(arr, cfg_file)
    cfg = generate_db.Config(freqs=(1e6, 9e6, 2e6), mesh_refn=3, array_config=cfg_file, format='FullFormat')
    monkeypatch.setattr(database, 'append_frequency_block', boom)
    with pytest.raises(OSError, match='disk full'):
        generate_db.main(cfg, str(tmp_path / 'resp.db'), block=1)
