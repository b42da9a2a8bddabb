"""Host-side mirror (cnld.mesh / cnld.fem / cnld.impulse_response / cnld.abstract) against the
golden vectors generated from the REFERENCE's own modules, and against the oracle mesh."""
import os

import numpy as np
import pytest
from conftest import GOLDEN


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize('tag', ['1x1_r3', '2x1_r4', '1x1_r6'])
def test_fem_mirror_matches_reference_golden(tag):
    from cnld import arrays, fem, mesh
    g = np.load(os.path.join(GOLDEN, f'fem_{tag}.npz'))
    refn = int(g['refn'])
    arr = arrays.matrix_array(nelem=tuple(g['nelem']))
    am = mesh.Mesh.from_abstract(arr, refn)
    assert np.array_equal(am.triangles, g['triangles'])
    assert np.array_equal(am.vertices, g['vertices'])      # bit-exact node coordinates
    assert np.array_equal(am.on_boundary, g['on_boundary'])
    mem = arr.elements[0].membranes[0]
    K, M = fem.mem_k_matrix(mem, refn), fem.mem_m_matrix(mem, refn)
    assert _rel(K, g['K']) < 1e-12
    assert _rel(M, g['M']) < 1e-13
    assert _rel(fem.mem_b_matrix_eig(mem, refn, M, K), g['B']) < 1e-9
    assert _rel(fem.mem_eig(mem, refn)[0][:8], g['eigf']) < 1e-9
    assert _rel(fem.array_f_spmatrix(arr, refn).toarray(), g['F']) < 1e-13
    assert _rel(fem.array_avg_spmatrix(arr, refn).toarray(), g['AVG']) < 1e-13
    for i, f in enumerate(g['freqs']):
        G = fem.array_mbk_spmatrix(arr, refn, f).toarray()
        assert _rel(np.diag(G), g['mbk_diag'][i]) < 1e-9
        Ks, Ms, Bs = fem.array_kmb_spmatrices(arr, refn)
        w = 2 * np.pi * f
        assert _rel((Ks - w**2 * Ms - 1j * w * Bs).toarray(), G) < 1e-14


def test_mesh_mirror_equals_oracle_mesh(po):
    from cnld import arrays, mesh
    for nelem, refn in [((1, 1), 2), ((2, 2), 5), ((3, 1), 7)]:
        m = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=nelem), refn)
        mo = po.array_mesh(po.matrix_array(nelem=nelem), refn)
        for a in ('vertices', 'edges', 'triangles', 'triangle_edges', 'on_boundary', 'g', 'normals'):
            assert np.array_equal(getattr(m, a), getattr(mo, a)), a
        n = refn
        assert m.nvertices == (2 * n * n + 2 * n + 1) * nelem[0] * nelem[1]
        assert m.ntriangles == 4 * n * n * nelem[0] * nelem[1]
    c = mesh.circle(20e-6, 5)
    vo, eo, to, so = po.geometry_circle(20e-6, 4, 5)
    assert np.array_equal(c.vertices, vo) and np.array_equal(c.triangles, to)
    assert np.array_equal(c.triangle_edges, so)


def test_surface_utilities():
    from cnld.h2lib import (build_from_macrosurface3d_surface3d, check_surface3d, isclosed_surface3d,
                            isoriented_surface3d, new_sphere_macrosurface3d, refine_red_surface3d)
    sp = build_from_macrosurface3d_surface3d(new_sphere_macrosurface3d(), 6)
    assert (sp.vertices, sp.triangles) == (6 + 12 * 5 + 8 * 10, 8 * 36)
    assert check_surface3d(sp) == 0 and isclosed_surface3d(sp) and isoriented_surface3d(sp)
    assert abs(sp.g.sum() / 2 - 4 * np.pi) / (4 * np.pi) < 0.03
    assert np.allclose(np.linalg.norm(sp.x, axis=1), 1.0)
    assert np.all(np.einsum('ij,ij->i', sp.n, sp.x[sp.t[:, 0]]) > 0)     # outward normals
    r = refine_red_surface3d(sp)
    assert (r.vertices, r.triangles) == (sp.vertices + sp.edges, 4 * sp.triangles)
    assert check_surface3d(r) == 0 and isclosed_surface3d(r) and isoriented_surface3d(r)
    assert abs(r.g.sum() - sp.g.sum()) < 1e-12


def test_host_containers_roundtrip():
    from scipy import sparse as sps
    from cnld.h2lib import (AMatrix, AVector, SparseMatrix, add_amatrix, add_sparsematrix_amatrix,
                            addeval_sparsematrix_avector, clone_amatrix, getsize_amatrix, scale_amatrix)
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    A = AMatrix.from_array(a)
    assert np.array_equal(A.a, a) and (A.rows, A.cols) == (5, 5)
    B = clone_amatrix(A)
    scale_amatrix(2 - 1j, B)
    assert np.allclose(B.a, (2 - 1j) * a)
    add_amatrix(1.0, False, A, B)
    assert np.allclose(B.a, (3 - 1j) * a)
    assert getsize_amatrix(A) >= 25 * 16
    x = rng.standard_normal(5) + 0j
    y = AVector(5)
    y.v[:] = 0
    # (the dense matvec runs on the GPU: tests/test_gpu_parity.py::test_compat_dense_entry_points)
    s = sps.random(5, 5, 0.4, random_state=1, format='csr') + sps.eye(5)
    S = SparseMatrix.from_array(s)
    assert S.nz == s.nnz
    y.v[:] = 0
    addeval_sparsematrix_avector(1.0, S, AVector.from_array(x), y)
    assert np.allclose(y.v, s @ x)
    Z = AMatrix(5, 5)
    add_sparsematrix_amatrix(1.0, False, S, Z)
    assert np.allclose(Z.a, s.toarray().T)


def test_abstract_json_roundtrip(tmp_path):
    from cnld import abstract, arrays
    arr = arrays.matrix_array(nelem=(2, 2))
    assert abstract.get_patch_count(arr) == 36 and abstract.get_membrane_count(arr) == 4
    p = tmp_path / 'array.json'
    abstract.dump(arr, str(p))
    back = abstract.load(str(p))
    assert abstract.dumps(back) == abstract.dumps(arr)
    assert '"__name__": "SquareCmutMembrane"' in abstract.dumps(arr)
    m0, m1 = arr.elements[0].membranes[0], arr.elements[3].membranes[0]
    assert m0._memoize() == m1._memoize()        # id / position / patches excluded


def test_fir_mirror_matches_reference_golden():
    from cnld import impulse_response
    g = np.load(os.path.join(GOLDEN, 'fir.npz'))
    for interp in (1, 2, 3):
        for kkr in (1, 0):
            t, h = impulse_response.fft_to_fir(g['f'], g['s'], interp=interp, use_kkr=bool(kkr), axis=-1)
            assert np.allclose(t, g[f't_{interp}_{kkr}'], rtol=1e-13, atol=0)
            assert _rel(h, g[f'h_{interp}_{kkr}']) < 1e-11
    t, h = impulse_response.fft_to_sir(g['f'], g['s'][0], interp=2, use_kkr=False, axis=1)
    assert _rel(h, g['sir_h']) < 1e-11
    # axis handling: frequency on axis 0
    s0 = np.moveaxis(g['s'], -1, 0)
    t0, h0 = impulse_response.fft_to_fir(g['f'], s0, interp=2, use_kkr=True, axis=0)
    assert _rel(np.moveaxis(h0, 0, -1), g['h_2_1']) < 1e-11


def test_simulation_signals_and_statedb_match_reference_golden():
    """cnld.simulation's drive signals (simulation.pyx:640-712) and StateDB's time grid / drive interpolation
    (:108-126) against the vectors the reference's own module produced; without a GPU the solver itself must
    refuse to construct (no CPU fallback)."""
    from cnld import _lib, simulation as sim
    g = np.load(os.path.join(GOLDEN, 'td_solver.npz'))
    sigs = [sim.gaussian_pulse(5e6, 0.8, 100e6, td=1e-7), sim.gaussian_pulse(3e6, 0.5, 80e6, antisym=False),
            sim.logistic_ramp(2e-7, 1e-8), sim.logistic_ramp(2e-7, 1e-8, td=3e-7, tstop=1e-6),
            sim.linear_ramp(2e-7, 1e-8), sim.linear_ramp(2e-7, 1e-8, td=1e-7, tstop=6e-7), sim.winsin(5e6, 3, 1e-8, td=2e-7)]
    sigs.append(sim.sigadd(sigs[2], sigs[6], sigs[4]))
    for q, (t, v) in enumerate(sigs):
        assert t.shape == g[f'sig_t{q}'].shape and np.allclose(t, g[f'sig_t{q}'], rtol=1e-13, atol=1e-20), q
        assert np.allclose(v, g[f'sig_v{q}'], rtol=1e-12, atol=1e-15), q
    for c in range(int(g['ncases'])):
        db = sim.StateDB(tuple(g[f't_lim{c}']), (g[f'v_t{c}'], g[f'v{c}']), g[f'fir{c}'].shape[0])
        assert np.array_equal(db.t, g[f't{c}']) and np.array_equal(db.v, g[f'vi{c}'])
        assert db.x.shape == g[f'x{c}'].shape and not db.x.any()
        st = db.get_state_i(3)
        st.x = np.full(db.x.shape[1], 2.0)
        db.set_state_i(st)
        assert np.all(db.x[3] == 2.0) and not db.x[2].any() and db.get_state_t(db.t[3] * 1.01).i == 3
    x = np.array([-5e-8, 0.0, -3.5e-8])
    assert np.allclose(sim.make_p_cont_spr(2e14, 1, -3e-8)(x), [4e6, 0, 1e6])
    assert np.allclose(sim.make_p_cont_dmp(1e13, 1, -3e-8)(x, np.array([1.0, 1.0, -2.0])), [-2e5, 0, 1e5])
    assert np.allclose(sim.p_es_pp(10.0, -1e-8, 6e-8), -8.8541878188e-12 / 2 * 100 / 25e-16, rtol=1e-9)
    if _lib.device_count() == 0:
        a = g
        with pytest.raises(_lib.Cnldb200Error):
            sim.FixedStepSolver((a['fir_t2'], a['fir2']), (a['v_t2'], a['v2']), a['gap2'], a['gap_eff2'], tuple(a['t_lim2']),
                                float(a['k2']), 2, float(a['x02']), float(a['lmbd2']))
    with pytest.raises(NotImplementedError):
        sim.CompensationSolver()


def test_fixed_step_solver_host_logic_with_oracle_standin(po, monkeypatch):
    """The host side of cnld.simulation.FixedStepSolver (constructor resampling, look-ahead stepping with samples
    revealed only when step() reaches them, iterator, reset, solve, the error past the time grid) with the device
    object replaced by a stand-in that serves the oracle's solution; compared with the reference-generated golden."""
    from cnld import _lib, simulation as sim
    g = np.load(os.path.join(GOLDEN, 'td_solver.npz'))
    c = 1

    class StandIn:
        STATE = _lib.TimeDomainSolver.STATE

        def __init__(self, fir, v, gap_eff, dt, k, n, x0, lmbd, atol=1e-10, maxiter=5, e0=None, device=0):
            self.nt, self.npatch = v.shape
            self.sol = po.td_solve(fir, v, gap_eff, dt, k, n, x0, lmbd, atol, maxiter, nsteps=self.nt - 1)
            self.current_step, self.calls = 1, []

        def step(self, nsteps=1):
            if self.current_step + nsteps > self.nt:
                raise _lib.Cnldb200Error('td_step: past the time grid')
            self.calls.append(nsteps)
            self.current_step += nsteps

        def reset(self):
            self.current_step = 1

        def state(self, i0, i1, into=None):
            assert i1 <= self.current_step
            for key in self.STATE:
                into[key][i0:i1] = self.sol[key][i0:i1]

        def info(self, s0, s1):
            return self.sol['error'][s0 - 1:s1 - 1], self.sol['iters'][s0 - 1:s1 - 1]

        def close(self):
            pass

    monkeypatch.setattr(_lib, 'TimeDomainSolver', StandIn)
    nsteps = int(g[f'nsteps{c}'])
    s = sim.FixedStepSolver((g[f'fir_t{c}'], g[f'fir{c}']), (g[f'v_t{c}'], g[f'v{c}']), g[f'gap{c}'], g[f'gap_eff{c}'],
                            tuple(g[f't_lim{c}']), float(g[f'k{c}']), int(g[f'n{c}']), float(g[f'x0{c}']), float(g[f'lmbd{c}']),
                            atol=float(g[f'atol{c}']), maxiter=int(g[f'maxiter{c}']))
    assert np.array_equal(s._fir, g[f'firi{c}']) and s.current_step == 1 and len(s.error) == 0
    assert np.allclose(s.pressure_electrostatic[0], g[f'p_es{c}'][0], rtol=1e-12)
    for _ in range(5):
        s.step()
    assert s._dev.calls == [sim.FixedStepSolver.LOOKAHEAD]                     # one device call served five samples
    assert s.current_step == 6 and not s._db.x[6:].any() and np.allclose(s.displacement, g[f'x{c}'][:6], rtol=1e-9, atol=1e-20)
    assert np.array_equal(s.iters, g[f'iters{c}'][:5]) and s.get_previous_state().i == 5
    s.step(80)                                                                  # beyond the look-ahead: one more call
    assert s._dev.calls == [64, 64] and s.current_step == 86
    assert sum(1 for _ in s) == nsteps - 85 and s.current_step == nsteps + 1   # iterator = step until len(t) - 1
    assert np.allclose(s.pressure_total, g[f'p_tot{c}'][:nsteps + 1], rtol=1e-9, atol=1e-9)
    assert np.array_equal(s.iters, g[f'iters{c}']) and np.allclose(s.error, g[f'error{c}'], atol=1e-18)
    with pytest.raises(_lib.Cnldb200Error):
        s.step(2)
    s.reset()
    assert s.current_step == 1 and not s._db.x.any() and len(s.iters) == 0
    s.solve()
    assert s.current_step == nsteps + 1 and np.allclose(s.velocity, g[f'u{c}'][:nsteps + 1], rtol=1e-8, atol=1e-12)


def test_database_roundtrip(tmp_path):
    from cnld import database
    from cnld.scripts.generate_db import Config
    f = str(tmp_path / 'x.db')
    cfg = Config()
    database.create_db(f, **cfg._asdict())
    assert database.read_metadata(f, as_dict=True)['mesh_refn'] == '11'
    rng = np.random.default_rng(0)
    P, N, nf = 3, 7, 4
    nodes = rng.standard_normal((N, 3))
    database.append_node(f, node_id=range(N), x=nodes[:, 0], y=nodes[:, 1], z=nodes[:, 2])
    assert np.allclose(database.read_node(f), nodes.T)
    freqs = np.array([1e6, 2e6, 3e6, 4e6])
    xp = rng.standard_normal((nf, P, P)) + 1j * rng.standard_normal((nf, P, P))
    xn = rng.standard_normal((nf, P, N)) + 1j * rng.standard_normal((nf, P, N))
    database.create_progress_table(f, nf)
    database.append_frequency_block(f, freqs[:2], freqs[:2] / 1500, xp[:2], xn[:2], nodes)
    database.update_progress(f, [1, 2])
    done, ijob = database.get_progress(f)
    assert done.tolist() == [True, True, False, False] and ijob == 3
    database.append_frequency_block(f, freqs[2:], freqs[2:] / 1500, xp[2:], xn[2:], nodes)
    fr, pp = database.read_patch_to_patch_freq_resp(f)
    assert np.allclose(fr, freqs) and np.allclose(pp, np.transpose(xp, (1, 2, 0)))
    fr, pn, nd = database.read_patch_to_node_freq_resp(f)
    assert np.allclose(pn, np.transpose(xn, (1, 2, 0))) and np.allclose(nd, nodes)
    # the reference's one-source-patch-at-a-time writer signature still works
    from itertools import repeat
    database.append_patch_to_patch_freq_resp(f, frequency=repeat(5e6), wavenumber=repeat(1.0), source_patch=repeat(0),
                                             dest_patch=np.arange(P), displacement_real=np.ones(P),
                                             displacement_imag=np.zeros(P), time_solve=repeat(0.1))


def test_translation_plan_covers_every_pair_once():
    """Host logic of the symmetric-path assembly (csrc/periodic.cpp): the work list plus the block
    copies must account for every lower-triangle membrane block exactly once, offsets with the same
    class must really be equal, and a mesh that is not exact translated copies takes the full list."""
    from cnld import _lib, arrays, mesh
    am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(4, 3)), 3)
    V, T = np.array(am.vertices), np.array(am.triangles)
    nt = len(T)
    info, work, copy = _lib.translation_plan(V, T)
    assert info['periodic'] and info['components'] == 12
    # 4 x 3 lattice: offsets dx in -3..3, dy in -2..2, up to sign, minus zero -> (7*5 - 1) / 2 = 17 classes
    assert info['offset_classes'] == 17
    nvc, ntc = info['vertices_per_component'], nt // 12
    assert nvc * 12 == len(V)
    # integrated pairs: count (t, s) with s < t covered by the work items, per membrane block
    cover = np.zeros((12, 12), dtype=np.int64)
    for tb, te, sb, se in work.astype(np.int64):
        assert tb // ntc == (te - 1) // ntc and sb // ntc == (se - 1) // ntc      # an item never straddles blocks
        a, b = tb // ntc, sb // ntc
        t = np.arange(tb, te)[:, None]
        s = np.arange(sb, se)[None, :]
        cover[a, b] += int((s < t).sum())
    integrated = {(a, b) for a in range(12) for b in range(12) if cover[a, b]}
    assert (0, 0) in integrated and cover[0, 0] == ntc * (ntc - 1) // 2
    for (a, b) in integrated - {(0, 0)}:
        assert a > b and cover[a, b] == ntc * ntc
    assert len(integrated) == 1 + 17
    # copies: every other lower block exactly once, from an integrated block with the same (or opposite) offset
    org = V[::nvc][:12]                                  # first vertex of each membrane
    seen = set(integrated)
    for dr, dc, sr, sc, tr in copy.astype(np.int64):
        a, b, ra, rb = dr // nvc, dc // nvc, sr // nvc, sc // nvc
        assert (ra, rb) in integrated and (a, b) not in seen and a >= b
        seen.add((a, b))
        d, e = org[a] - org[b], org[ra] - org[rb]
        assert np.allclose(d, -e if tr else e, rtol=0, atol=1e-15)
    assert seen == {(a, b) for a in range(12) for b in range(a + 1)}
    # without detection, or with one vertex nudged: one self rectangle over everything
    for VV, allow in ((V, False), (V + 1e-9 * (np.arange(len(V)) == 7)[:, None], True)):
        info2, work2, copy2 = _lib.translation_plan(VV, T, allow=allow)
        assert not info2['periodic'] and len(copy2) == 0
        tot = sum(int((np.arange(sb, se)[None, :] < np.arange(tb, te)[:, None]).sum()) for tb, te, sb, se in work2.astype(np.int64))
        assert tot == nt * (nt - 1) // 2


def test_translation_plan_falls_back_on_anything_irregular():
    """Detection must be conservative: one component, permuted triangles, permuted vertices or a
    scaled copy all take the integrate-everything list (one self rectangle, no copies)."""
    from cnld import _lib, arrays, mesh
    am1 = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(1, 1)), 3)
    V1, T1 = np.array(am1.vertices), np.array(am1.triangles)
    nv, nt = len(V1), len(T1)
    shift = np.array([60e-6, 0.0, 0.0])

    def plan(V, T):
        info, work, copy = _lib.translation_plan(V, T)
        tot = sum(int((np.arange(sb, se)[None, :] < np.arange(tb, te)[:, None]).sum()) for tb, te, sb, se in work.astype(np.int64))
        return info, tot, len(copy)

    # a clean pair of translated copies is detected ...
    V2, T2 = np.vstack([V1, V1 + shift]), np.vstack([T1, T1 + nv]).astype(np.uint32)
    info, tot, ncopy = plan(V2, T2)
    assert info['periodic'] and info['components'] == 2 and info['offset_classes'] == 1 and ncopy == 1
    assert tot == nt * (nt - 1) // 2 + nt * nt
    full = (2 * nt) * (2 * nt - 1) // 2
    # ... a single component is not
    info, tot, ncopy = plan(V1, T1)
    assert not info['periodic'] and ncopy == 0 and tot == nt * (nt - 1) // 2
    # second copy with its triangles in another order
    info, tot, ncopy = plan(V2, np.vstack([T1, (T1 + nv)[::-1]]).astype(np.uint32))
    assert not info['periodic'] and ncopy == 0 and tot == full
    # second copy with two vertices swapped (connectivity renumbered consistently: same geometry, other numbering)
    perm = np.arange(nv); perm[[3, 5]] = perm[[5, 3]]
    Vb = V1[perm] + shift
    inv = np.argsort(perm)
    info, tot, ncopy = plan(np.vstack([V1, Vb]), np.vstack([T1, inv[T1] + nv]).astype(np.uint32))
    assert not info['periodic'] and ncopy == 0 and tot == full
    # second copy scaled by 1 + 1e-9
    info, tot, ncopy = plan(np.vstack([V1, V1 * (1 + 1e-9) + shift]), T2)
    assert not info['periodic'] and ncopy == 0 and tot == full
