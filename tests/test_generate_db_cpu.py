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
    cfg = generate_db.Config(freqs=(1e6, 9e6, 2e6), mesh_refn=3, array_config=cfg_file)
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


def test_generate_db_writer_failure_reaches_the_caller(tmp_path, po, monkeypatch):
    from cnld import abstract, arrays, database
    from cnld.scripts import generate_db
    monkeypatch.setattr(generate_db, 'FrequencySweep', _FakeSweep)

    def boom(*a, **k):
        raise OSError('disk full')

    arr = arrays.matrix_array(nelem=(1, 1))
    cfg_file = str(tmp_path / 'array.json')
    abstract.dump(arr, cfg_file)
    cfg = generate_db.Config(freqs=(1e6, 9e6, 2e6), mesh_refn=3, array_config=cfg_file)
    monkeypatch.setattr(database, 'append_frequency_block', boom)
    with pytest.raises(OSError, match='disk full'):
        generate_db.main(cfg, str(tmp_path / 'resp.db'), block=1)
