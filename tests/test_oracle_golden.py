"""Oracle (numpy restatements) vs golden vectors produced by the REFERENCE's own
fem.py / impulse_response.py / bem_pressure.pyx (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
from conftest import GOLDEN


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize('tag', ['1x1_r3', '2x1_r4', '1x1_r6'])
def test_fem_matches_reference(po, tag):
    g = np.load(os.path.join(GOLDEN, f'fem_{tag}.npz'))
    refn = int(g['refn'])
    arr = po.matrix_array(nelem=tuple(g['nelem']))
    am = po.array_mesh(arr, refn)
    assert np.array_equal(am.triangles, g['triangles'])
    assert np.allclose(am.vertices, g['vertices'], rtol=0, atol=1e-20)
    assert np.array_equal(am.on_boundary, g['on_boundary'])
    blocks = po.array_fem(arr, refn)
    K, M, B, F, A = blocks[0]
    assert _rel(K, g['K']) < 1e-12
    assert _rel(M, g['M']) < 1e-13
    assert _rel(B, g['B']) < 1e-9      # through a dense non-symmetric eig
    from scipy.linalg import block_diag
    assert _rel(block_diag(*[b[3] for b in blocks]), g['F']) < 1e-13
    assert _rel(block_diag(*[b[4] for b in blocks]), g['AVG']) < 1e-13
    for i, f in enumerate(g['freqs']):
        G = po.mbk_dense(blocks, f)
        assert _rel(np.diag(G), g['mbk_diag'][i]) < 1e-9
        assert abs(G.sum() - g['mbk_sum'][i]) / abs(g['mbk_sum'][i]) < 1e-8


def test_fir_matches_reference(po):
    g = np.load(os.path.join(GOLDEN, 'fir.npz'))
    for interp in (1, 2, 3):
        for kkr in (1, 0):
            t, h = po.fft_to_fir(g['f'], g['s'], interp=interp, use_kkr=bool(kkr))
            assert np.allclose(t, g[f't_{interp}_{kkr}'], rtol=1e-13, atol=0)
            assert _rel(h, g[f'h_{interp}_{kkr}']) < 1e-11
    t, h = po.fft_to_fir(g['f'], g['s'][0], interp=2, use_kkr=False)
    assert _rel(h, g['sir_h']) < 1e-11


def test_fir_conv_matches_reference(po):
    """oracle fir_conv == the reference's fir_conv_cy loop (simulation.pyx:51-77) on the golden cases."""
    g = np.load(os.path.join(GOLDEN, 'fir_conv.npz'))
    for c in range(int(g['ncases'])):
        dt, offset = g[f'arg{c}']
        x = po.fir_conv(g[f'fir{c}'], g[f'p{c}'], dt, int(offset))
        assert _rel(x, g[f'x{c}']) < 1e-13


def test_td_solver_matches_reference(po):
    """oracle td_solve == the reference's FixedStepSolver (simulation.pyx:297-555) stepped through its own class:
    table resampling, drive interpolation, every state array, iteration counts and final fixed-point errors;
    case 1 / 2 include steps where patch 0 is out of contact (np.vectorize then truncates the contact pressures)."""
    g = np.load(os.path.join(GOLDEN, 'td_solver.npz'))
    for c in range(int(g['ncases'])):
        t_lim = tuple(g[f't_lim{c}'])
        _, firi = po.td_interp_fir(g[f'fir_t{c}'], g[f'fir{c}'], t_lim[2])
        assert np.array_equal(firi, g[f'firi{c}'])
        t, v = po.td_voltage(t_lim, (g[f'v_t{c}'], g[f'v{c}']), firi.shape[0])
        assert np.array_equal(t, g[f't{c}']) and np.array_equal(v, g[f'vi{c}'])
        r = po.td_solve(firi, v, g[f'gap_eff{c}'], t_lim[2], float(g[f'k{c}']), float(g[f'n{c}']), float(g[f'x0{c}']),
                        float(g[f'lmbd{c}']), float(g[f'atol{c}']), int(g[f'maxiter{c}']), nsteps=int(g[f'nsteps{c}']))
        assert np.array_equal(r['iters'], g[f'iters{c}'])
        for key in ('x', 'u', 'p_es', 'p_cont_spr', 'p_cont_dmp', 'p_tot'):
            assert _rel(r[key], g[f'{key}{c}']) < 1e-9, (c, key)
        assert np.abs(r['error'] - g[f'error{c}']).max() < 1e-9 * np.abs(g[f'x{c}']).max()
        trunc = g[f'p_cont_spr{c}'][g[f'x{c}'][:, 0] >= g[f'x0{c}']]
        assert np.array_equal(trunc, np.trunc(trunc))      # the quirk is really in the reference's output


def test_pressure_matches_reference(po):
    g = np.load(os.path.join(GOLDEN, 'pressure.npz'))
    am = po.Mesh(g['vertices'], np.zeros((1, 2), np.uint32), g['triangles'],
                 np.zeros_like(g['triangles']))
    assert np.allclose(am.triangle_areas, g['areas'], rtol=1e-13)
    for ik, k in enumerate(g['ks']):
        pv = po.pres_vector(am, g['obs'], k)
        assert _rel(pv, g['pvec'][ik]) < 1e-12
