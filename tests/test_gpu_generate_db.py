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


def test_contact_simulation_from_database(tmp_path, po):
    """The whole chain of configs[4]: frequency sweep -> database -> impulse responses -> time-domain contact solver
    built by FixedStepSolver.from_array_and_db (simulation.pyx:338-365) -> samples, against the oracle's solver fed
    with the impulse responses the database holds; also with calc_fir (responses recomputed from the spectra)."""
    from cnld import abstract, arrays, database, simulation
    from cnld.scripts import generate_db
    arr = arrays.matrix_array(nelem=(2, 1))
    cfg_file = str(tmp_path / 'array.json')
    abstract.dump(arr, cfg_file)
    cfg = generate_db.Config(freqs=(0, 40e6, 1e6), mesh_refn=3, array_config=cfg_file, format='FullFormat')
    db = str(tmp_path / 'resp.db')
    generate_db.main(cfg, db, block=16)
    fir_t, fir = database.read_patch_to_patch_imp_resp(db)
    P, dt = fir.shape[0], fir_t[1] - fir_t[0]
    assert P == 18 and fir.shape[2] == 2 * (40 * cfg.freq_interp + 1) - 1
    nt = 260
    t_lim = (0.0, (nt - 1) * dt, dt)
    v_t = np.arange(0, nt * dt, 2 * dt)
    rng = np.random.default_rng(2)
    v = np.outer(1 / (1 + np.exp(-(v_t - 0.5e-6) / 0.08e-6)), 30 + 2 * rng.random(P)) + \
        2 * np.outer(np.sin(2 * np.pi * 5e6 * v_t) * (v_t > 1.2e-6), np.ones(P))
    kw = dict(k=3e14, n=1, x0=-12e-9, lmbd=3e13, atol=1e-12, maxiter=5)
    sol = simulation.FixedStepSolver.from_array_and_db(arr, db, (v_t, v), t_lim, **kw)
    assert np.allclose(sol.props.gap_eff, 50e-9 + 200e-9 / 6.3) and sol.npatch == P
    sol.solve()
    _, firi = po.td_interp_fir(fir_t, fir, dt)
    tv, vi = po.td_voltage(t_lim, (v_t, v), P)
    ref = po.td_solve(firi, vi, sol.props.gap_eff, dt, kw['k'], kw['n'], kw['x0'], kw['lmbd'], kw['atol'], kw['maxiter'],
                      nsteps=len(tv) - 2)
    n = sol.current_step
    assert n == len(tv) - 1 and (ref['x'] < kw['x0']).sum() > 300 and ref['iters'].max() == 10
    assert np.array_equal(sol.iters, ref['iters'])
    for key, view in (('x', sol.displacement), ('u', sol.velocity), ('p_tot', sol.pressure_total),
                      ('p_cont_spr', sol.pressure_contact_spring), ('p_cont_dmp', sol.pressure_contact_damper)):
        assert np.linalg.norm(view - ref[key][:n]) / np.linalg.norm(ref[key][:n]) < 1e-9, key
    sol2 = simulation.FixedStepSolver.from_array_and_db(arr, db, (v_t, v), t_lim, calc_fir=True, interp=cfg.freq_interp, **kw)
    assert np.linalg.norm(sol2._fir - sol._fir) / np.linalg.norm(sol._fir) < 1e-9
    sol.close()
    sol2.close()


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
