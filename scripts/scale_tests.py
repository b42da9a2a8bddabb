"""Scale runs on one B200 (size-independent parity properties at BASELINE.json's full sizes):
  c5 : field pressure of the 4x4 array on 10^6 observation points (100^3 grid)
  h  : H-matrix assembly + matvec on larger arrays, checked on sampled rows against exact entries
Usage: python scripts/scale_tests.py [c5] [h8] [c4]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
from cnld import _lib, arrays, mesh  # noqa: E402


def c5():
    am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=(4, 4)), 21)
    ctx = _lib.Context(np.array(am.vertices), np.array(am.triangles))
    g = np.linspace(-1e-3, 1e-3, 100)
    z = np.linspace(0.1e-3, 2.1e-3, 100)
    X, Y, Z = np.meshgrid(g, g, z, indexing='ij')
    obs = np.c_[X.ravel('F'), Y.ravel('F'), Z.ravel('F')]
    rng = np.random.default_rng(0)
    out = {}
    for nsrc in (1, 4):
        disp = rng.standard_normal((1, nsrc, am.nvertices)) + 1j * rng.standard_normal((1, nsrc, am.nvertices))
        k = 2 * np.pi * 5e6 / 1500
        ctx.pressure(obs[:1000], [k], disp)                                     # warm-up
        t0 = time.perf_counter()
        p = ctx.pressure(obs, [k], disp)
        dt = time.perf_counter() - t0
        ke = 3.0 * am.ntriangles * len(obs)
        # parity properties: sampled points against the reference-shaped p_vec . disp, and linearity
        idx = rng.choice(len(obs), 64, replace=False)
        ref = disp[0] @ ctx.pres_vector(obs[idx], k).T
        err = np.linalg.norm(p[0][:, idx] - ref) / np.linalg.norm(ref)
        p2 = ctx.pressure(obs[idx], [k], 2.5 * disp)
        lin = np.linalg.norm(p2[0] - 2.5 * p[0][:, idx]) / np.linalg.norm(p2[0])
        out[f'nsrc{nsrc}'] = dict(seconds=dt, kernel_evals_per_s=ke / dt, points=len(obs), err_vs_pvec=err, linearity=lin)
    ctx.close()
    return out


def hscale(nelem, refn, eps=1e-2):
    am = mesh.Mesh.from_abstract(arrays.matrix_array(nelem=nelem), refn)
    n = am.nvertices
    ctx = _lib.Context(np.array(am.vertices), np.array(am.triangles))
    t0 = time.perf_counter()
    H = _lib.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=64)
    t_tree = time.perf_counter() - t0
    k = 2 * np.pi * 5e6 / 1500
    H.assemble(k, eps)                                                          # warm-up (JIT-free, but first touch)
    t0 = time.perf_counter()
    H.assemble(k, eps)
    t_asm = time.perf_counter() - t0
    info = H.info()
    info['sharing'] = H.sharing()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    H.matvec(x)
    t0 = time.perf_counter()
    y = H.matvec(x)
    t_mv = time.perf_counter() - t0
    rows = rng.choice(n, 48, replace=False).astype(np.uint32)
    exact = ctx.nearfield(k, rows, np.arange(n, dtype=np.uint32)) @ x
    err = np.linalg.norm(y[rows] - exact) / np.linalg.norm(exact)
    out = dict(n=n, tree_seconds=t_tree, assemble_seconds=t_asm, matvec_seconds_incl_copies=t_mv, info=info,
               compression=info['stored'] / float(n) ** 2, sampled_row_error=err, eps_aca=eps)
    H.close()
    ctx.close()
    return out


if __name__ == '__main__':
    what = sys.argv[1:] or ['c5', 'h8']
    res = {}
    if 'c5' in what:
        res['c5_pressure'] = c5()
    if 'h8' in what:
        res['h_8x8_refn15'] = hscale((8, 8), 15)
    if 'c4' in what:
        res['h_16x16_refn19'] = hscale((16, 16), 19)
    print(json.dumps(res, indent=1, default=float))


def hsweep(nelem, refn, nfreq=1, batch=32, eps=1e-2, tol=1e-6):
    """H-matrix + GMRES frequency sweep at scale (time per frequency, iteration counts)."""
    from cnld.sweep import SweepSetup
    t0 = time.perf_counter()
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn)
    t_setup = time.perf_counter() - t0
    ctx = _lib.Context(setup.vertices, setup.triangles)
    ctx.set_mbk(setup.K, setup.M, setup.B)
    ctx.set_loads(setup.F, setup.AVG, setup.on_boundary)
    H = _lib.HMatrix(ctx, clf=16, eta=0.8, strict=True, kmax=64)
    freqs = np.linspace(2e6, 20e6, nfreq) if nfreq > 1 else np.array([5e6])
    xp, _, it, t = H.sweep(freqs, eps_aca=eps, tol=tol, restart=50, maxiter=300, batch=batch, want_nodes=False)
    out = dict(n=setup.nnodes, P=setup.npatch, host_setup_seconds=t_setup, seconds_per_frequency=t.tolist(),
               iters_mean=float(it.mean()), iters_max=int(it.max()), finite=bool(np.isfinite(xp).all()),
               xp_norm=float(np.linalg.norm(xp)))
    H.close()
    ctx.close()
    return out


if __name__ == '__main__' and any(a.startswith('sweep') for a in sys.argv[1:]):
    r = {}
    if 'sweep8' in sys.argv:
        r['hsweep_8x8_refn15'] = hsweep((8, 8), 15, nfreq=2, batch=64)
    if 'sweep16' in sys.argv:      # configs[3]: 16x16 array, N = 194 816, P = 2 304 right-hand sides
        r['hsweep_16x16_refn19'] = hsweep((16, 16), 19, nfreq=1, batch=96)
    if 'sweep4' in sys.argv:
        r['hsweep_4x4_refn21'] = hsweep((4, 4), 21, nfreq=2, batch=48)
    print(json.dumps(r, indent=1, default=float))
