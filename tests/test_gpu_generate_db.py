"""End-to-end drop-in check on the GPU: the generate_db mirror writes the same database the
reference's driver would (schema, row order, resume), with values equal to the oracle's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('fmt,devices,tol', [('FullFormat', 1, 1e-6), ('FullFormat', [0, 0], 1e-6), ('HFormat', 1, 2e-2),
                                             ('HFormat', [0, 0], 2e-2)])
def test_generate_db_end_to_end_and_resume(tmp_path, po, fmt, devices, tol):
    """cfg.format picks the solver (dense LU / ACA H-matrix + GMRES, tolerance = the reference's eps_aca for the
    latter); devices = [0, 0] runs the multi-process driver (one worker per entry) with both workers on GPU 0."""
    from cnld import abstract, arrays, database
    from cnld.scripts import generate_db
    arr = arrays.matrix_array(nelem=(2, 1))
    cfg_file = str(tmp_path / 'array.json')
    abstract.dump(arr, cfg_file)
    cfg = generate_db.Config(freqs=(0, 8e6, 2e6), mesh_refn=3, array_config=cfg_file, format=fmt)
    db = str(tmp_path / 'out' / 'resp.db')
    generate_db.main(cfg, db, block=2, devices=devices)
    freqs, ppfr = database.read_patch_to_patch_freq_resp(db)
    assert np.allclose(freqs, [0, 2e6, 4e6, 6e6, 8e6]) and ppfr.shape == (18, 18, 5)
    f2, pnfr, nodes = database.read_patch_to_node_freq_resp(db)
    oarr = po.matrix_array(nelem=(2, 1))
    am = po.array_mesh(oarr, 3)
    blocks = po.array_fem(oarr, 3)
    assert np.allclose(nodes, am.vertices)
    for i, f in enumerate(freqs):
        Xo, XPo = po.solve_frequency(am, blocks, f)
        assert np.linalg.norm(pnfr[:, :, i].T - Xo) / np.linalg.norm(Xo) < tol
        assert np.linalg.norm(ppfr[:, :, i].T - XPo) / np.linalg.norm(XPo) < tol
    t, ppir = database.read_patch_to_patch_imp_resp(db)
    to, ho = po.fft_to_fir(freqs, ppfr, interp=cfg.freq_interp, use_kkr=True)
    assert ppir.shape == ho.shape and np.linalg.norm(ppir - ho) / np.linalg.norm(ho) < 1e-9
    done, _ = database.get_progress(db)
    assert done.all()
    # resume: a finished database is left alone
    generate_db.main(cfg, db)
    assert database.read_patch_to_patch_freq_resp(db)[1].shape == (18, 18, 5)


def test_compressed_formats_dense_flow(po):
    """The reference's process() recipe through the mirrored operator classes (FullFormat)."""
    from cnld import arrays, bem, fem
    from cnld.compressed_formats import MbkSparseMatrix
    from cnld.mesh import Mesh
    arr = arrays.matrix_array(nelem=(1, 1))
    refn, f, c, rho = 4, 5e6, 1500., 1000.
    k, omg = 2 * np.pi * f / c, 2 * np.pi * f
    Gfe = fem.array_mbk_spmatrix(arr, refn, f, format='csr')
    Z = bem.array_z_matrix(arr, refn, k, format='FullFormat', basis='linear', q_reg=2, q_sing=4)
    assert Z.shape == (41, 41) and Z.size > 0 and 'ZMatrix' in repr(Z)
    G = MbkSparseMatrix(Gfe) + (-omg**2 * 2 * rho * Z)
    Glu = G.lu()
    F = fem.array_f_spmatrix(arr, refn)
    mesh = Mesh.from_abstract(arr, refn)
    oarr = po.matrix_array(nelem=(1, 1))
    Xo, _ = po.solve_frequency(po.array_mesh(oarr, refn), po.array_fem(oarr, refn), f)
    for sid in (0, 4):
        b = np.array(F[:, sid].todense())
        x = np.conj(Glu.lusolve(b))
        x[mesh.on_boundary] = 0
        assert np.linalg.norm(x - Xo[:, sid]) / np.linalg.norm(Xo[:, sid]) < 1e-6
    y = Z * np.ones(41)
    assert np.linalg.norm(y - po.assemble_z(po.array_mesh(oarr, refn), k).sum(axis=1)) / np.linalg.norm(y) < 1e-9
