"""Route A of INTEGRATION.md on the GPU: the reference's own Cython binding layer (built by
scripts/build_route_a.py from /root/reference, linked with -lcnld_b200; the built extension modules
travel to the GPU box) runs the reference's recipes -- ZFullMatrix -> scale -> add sparse -> lu -> lusolve
(compressed_formats.py:616-645, 209-262) and the HFormat recipe (:382-448, :652-703) -- in a separate
interpreter; the results are compared here with the oracle and with plain linear algebra."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', '_route_a')
sys.path.insert(0, os.path.join(ROOT, 'scripts'))


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def test_reference_cython_layer_runs_the_recipes(po, tmp_path):
    import build_route_a
    if not build_route_a.built():
        pytest.skip('route-A extensions were not built (scripts/build_route_a.py needs /root/reference)')
    out = str(tmp_path / 'route_a.npz')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'route_a', 'recipe.py'), out],
                       capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=OUT), cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    d = np.load(out)
    n = int(d['n'])
    # mesh: same refinement (node order included) as the oracle's restatement of the macrosurface contract
    v, e, t, s = po.geometry_square(40e-6, 40e-6, 5)
    assert np.array_equal(d['T'], t) and np.abs(d['X'] - v).max() < 1e-18
    # Z through assemble_bem3d_amatrix == oracle Z
    am = po.Mesh(v, e, t, s)
    Zo = po.assemble_z(am, float(d['k']))
    assert np.abs(d['Z'] - Zo).max() / np.abs(Zo).max() < 1e-10
    # G = alpha Z + M, LU solves, matvec
    G = complex(d['alpha']) * Zo + d['Mdense']
    assert _rel(d['G'], G) < 1e-12
    assert _rel(d['Xs'], np.linalg.solve(G, d['B'])) < 1e-9
    assert _rel(d['mv'], G @ d['B'][:, 0]) < 1e-12
    assert float(d['lr_err']) < 1e-10 * np.linalg.norm(G, 2)
    # Cholesky and CG on the Hermitian positive definite matrix
    xref = np.linalg.solve(d['H'], d['B'][:, 1])
    assert _rel(d['xchol'], xref) < 1e-9 and _rel(d['xcg'], xref) < 1e-8 and 0 < int(d['it_cg']) < 500
    # H-matrix recipe: ACA 1e-6 assembly, scale by addmul with the identity, add the sparse operator, LU, solve
    zn = np.linalg.norm(Zo, 2)
    assert float(d['zh_err']) < 1e-4 * zn
    assert float(d['gh_err']) < 1e-4 * np.linalg.norm(G, 2)
    assert _rel(d['xh'], np.linalg.solve(G, d['B'][:, 0])) < 1e-4
    assert float(d['lrh_err']) < 1e-9 * np.linalg.norm(G, 2)
    assert _rel(d['xgmres'], d['xh']) < 1e-7 and int(d['it_gm']) < 400
    assert int(d['hsize']) > 0
