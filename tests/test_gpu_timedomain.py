"""GPU parity of the stages after the sweep: impulse responses (device DFT-GEMM vs the golden vectors the
reference's own impulse_response.py produced, 1e-9) and the time-domain FIR convolution (vs the golden
vectors of the reference's fir_conv_cy loop and the oracle at a larger size)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def test_device_impulse_response_vs_reference_golden(po):
    from cnld import _lib, impulse_response
    g = np.load(os.path.join(GOLDEN, 'fir.npz'))
    for interp in (1, 2, 3):
        for kkr in (1, 0):
            t, h = _lib.fft_to_fir(g['f'], g['s'], interp=interp, use_kkr=bool(kkr))
            assert np.allclose(t, g[f't_{interp}_{kkr}'], rtol=1e-13, atol=0)
            assert h.shape == g[f'h_{interp}_{kkr}'].shape and _rel(h, g[f'h_{interp}_{kkr}']) < 1e-9
    t, h = impulse_response.fft_to_sir(g['f'], g['s'][0], interp=2, use_kkr=False, axis=1, engine='device')
    assert _rel(h, g['sir_h']) < 1e-9
    # a sweep-sized batch (P x P = 324 spectra, 251 frequencies starting at 0 like the default sweep,
    # generate_db.py:257) against the host restatement, on a non-last axis as postprocess uses it
    rng = np.random.default_rng(4)
    f = np.arange(0, 50e6 + 1, 200e3)
    s = (rng.standard_normal((18, 18, len(f))) + 1j * rng.standard_normal((18, 18, len(f)))) * np.exp(-f / 2e7)
    for kkr in (True, False):
        th, hh = impulse_response.fft_to_fir(f, s, interp=2, use_kkr=kkr, engine='host')
        td, hd = impulse_response.fft_to_fir(f, s, interp=2, use_kkr=kkr, engine='device')
        assert np.allclose(td, th, rtol=1e-13, atol=0) and _rel(hd, hh) < 1e-9
        if kkr:                                          # Kramers-Kronig: causal response
            assert np.abs(hd[..., len(td) // 2 + 1:]).max() == 0
    with pytest.raises(_lib.Cnldb200Error):
        _lib.fft_to_fir(f[::-1].copy(), s)               # frequencies must increase


def test_fir_convolution_vs_reference_golden_and_oracle(po):
    from cnld import _lib
    g = np.load(os.path.join(GOLDEN, 'fir_conv.npz'))
    for c in range(int(g['ncases'])):
        dt, offset = g[f'arg{c}']
        tab = _lib.FirTable(g[f'fir{c}'])
        assert _rel(tab.conv(g[f'p{c}'], dt, int(offset)), g[f'x{c}']) < 1e-12
        tab.close()
    rng = np.random.default_rng(5)
    nsrc = ndest = 36
    nfir, dt = 700, 2.5e-9
    fir = rng.standard_normal((nsrc, ndest, nfir))
    p = rng.standard_normal((1500, nsrc))
    tab = _lib.FirTable(fir)
    for nsample, offset in [(0, 0), (1, 0), (1, 1), (333, 1), (700, 0), (701, 1), (1500, 1), (1500, 0), (5, 699), (5, 700)]:
        assert _rel(tab.conv(p[:nsample], dt, offset), po.fir_conv(fir, p[:nsample], dt, offset)) < 1e-12 or \
            (nsample == 0 or offset >= nfir) and np.all(tab.conv(p[:nsample], dt, offset) == 0)
    # the solver's stepping pattern (simulation.pyx:400-410, :459-503): base over the finished steps + add for the current one
    tab.reset()
    for n in range(1, 1201):
        tab.push(p[n - 1])
        if n in (1, 2, 699, 700, 1024, 1025, 1200):       # history growth, lag window shorter / longer than the table
            base = tab.base(dt)
            assert _rel(base, po.fir_conv(fir, p[:n], dt, 1)) < 1e-12
            assert _rel(base + tab.add(p[n], dt), po.fir_conv(fir, p[:n + 1], dt, 0)) < 1e-12
    # deterministic: the same call twice gives the same bits
    assert np.array_equal(tab.base(dt), tab.base(dt))
    tab.close()


def test_fixed_step_solver_vs_reference_golden(po):
    """FixedStepSolver on the device against the reference's own class (tests/golden/td_solver.npz: three drives
    into contact, linear and quadratic contact laws, maxiter 5 / 3, steps with patch 0 in and out of contact --
    np.vectorize's integer truncation included), stepping one sample at a time, many at a time and after reset."""
    from cnld import simulation as sim
    g = np.load(os.path.join(GOLDEN, 'td_solver.npz'))
    for c in range(int(g['ncases'])):
        nsteps = int(g[f'nsteps{c}'])
        n = int(g[f'n{c}'])
        s = sim.FixedStepSolver((g[f'fir_t{c}'], g[f'fir{c}']), (g[f'v_t{c}'], g[f'v{c}']), g[f'gap{c}'], g[f'gap_eff{c}'],
                                tuple(g[f't_lim{c}']), float(g[f'k{c}']), n, float(g[f'x0{c}']), float(g[f'lmbd{c}']),
                                atol=float(g[f'atol{c}']), maxiter=int(g[f'maxiter{c}']))
        assert np.array_equal(s._fir, g[f'firi{c}']) and s.current_step == 1
        assert np.array_equal(s.pressure_electrostatic[0], g[f'p_es{c}'][0])      # initial state, bit for bit

        def check():
            assert s.current_step == nsteps + 1 and np.array_equal(s.time, g[f't{c}'][:nsteps + 1])
            assert np.array_equal(s.iters, g[f'iters{c}'])
            for key, view in (('x', s.displacement), ('u', s.velocity), ('p_es', s.pressure_electrostatic),
                              ('p_cont_spr', s.pressure_contact_spring), ('p_cont_dmp', s.pressure_contact_damper),
                              ('p_tot', s.pressure_total)):
                assert view.shape == (nsteps + 1, s.npatch) and _rel(view, g[f'{key}{c}'][:nsteps + 1]) < 1e-9, (c, key)
            assert np.abs(s.error - g[f'error{c}']).max() < 1e-9 * np.abs(g[f'x{c}']).max()
            free = s.displacement[:, 0] >= float(g[f'x0{c}'])
            assert np.array_equal(s.pressure_contact_spring[free], np.trunc(s.pressure_contact_spring[free]))

        for _ in range(10):                      # one table pass per sample ...
            s.step()
        assert s.current_step == 11 and _rel(s.displacement, g[f'x{c}'][:11]) < 1e-9
        # the device runs ahead in whole blocks; the StateDB shows a sample only when step() has reached it
        assert s._done > 11 and not s._db.x[11:].any() and not s.get_current_state().x.any() and len(s.iters) == 10
        s.step(nsteps - 10)                      # ... then blocks of 8 samples per pass (history / in-block lags split)
        check()
        x_mixed = s.displacement.copy()
        s.reset()
        assert s.current_step == 1 and not s._db.x.any() and len(s.error) == 0
        assert sum(1 for _ in s) == nsteps                                          # iterator protocol, sample by sample
        check()
        assert _rel(s.displacement, x_mixed) < 1e-11                                # same sums, different association
        s.reset()
        s.solve()                                                                   # everything in one device call
        check()
        x_solve = s.displacement.copy()
        s.reset()
        s.solve()
        assert np.array_equal(x_solve, s.displacement)                              # deterministic, bit for bit
        with pytest.raises(Exception):
            s.step(2)                                                               # past the time grid
        s.close()


def test_fixed_step_solver_vs_oracle_larger(po):
    """36 patches, 700-tap table, 400 steps (history shorter and longer than the table), against the oracle's
    restatement of the solver; and the module-level fir_conv_cy / fir_conv_py against the oracle."""
    from cnld import _lib, simulation as sim
    rng = np.random.default_rng(11)
    P, nfir, dt, nt = 36, 700, 2e-9, 402
    tt = np.arange(nfir) * dt
    px = np.arange(P) % 6
    py = np.arange(P) // 6
    dist = np.hypot(px[:, None] - px[None, :], py[:, None] - py[None, :])
    amp = 7e-7 / (1 + 5 * dist)**2 * (1 + 0.1 * rng.standard_normal((P, P)))
    fir = amp[:, :, None] * np.exp(-tt / 70e-9) * np.sin(2 * np.pi * 5.5e6 * (1 + 0.03 * rng.standard_normal((P, P, 1))) * tt + 0.3)
    fir[:, :, 300:] = 0                         # causal (Kramers-Kronig) tables end in exact zeros: never streamed
    t = np.arange(nt) * dt
    v = np.outer(1 / (1 + np.exp(-(t - 60e-9) / 10e-9)), 19.5 + 1.5 * rng.random(P)) + 1.5 * np.outer(np.sin(2 * np.pi * 8e6 * t), np.ones(P))
    gap_eff = np.full(P, 50e-9 + 100e-9 / 6.3)
    # contact threshold inside the stable electrostatic range (|x0| < gap_eff / 3): a rounding-level perturbation of
    # the table stays at rounding level over the 400 steps (checked with the oracle; with x0 = -30 nm it grows to 10 %)
    kw = dict(k=1e14 * 15e-9**-0.5, n=1.5, x0=-15e-9, lmbd=2e13 * 15e-9**-0.5, atol=1e-13, maxiter=4)
    ref = po.td_solve(fir, v, gap_eff, dt, kw['k'], kw['n'], kw['x0'], kw['lmbd'], kw['atol'], kw['maxiter'], nsteps=nt - 1)
    dev = _lib.TimeDomainSolver(fir, v, gap_eff, dt, kw['k'], kw['n'], kw['x0'], kw['lmbd'], atol=kw['atol'], maxiter=kw['maxiter'])
    assert dev.taps == 300 and dev.nfir == nfir
    dev.step(nt - 1)
    got = dev.state(0, nt)
    err, it = dev.info(1, nt)
    assert (ref['x'] < kw['x0']).sum() > 500 and ref['iters'].max() > 4          # contact engaged, iterations vary
    assert ((ref['x'][:, 1:] < kw['x0']).any(axis=1) & (ref['x'][:, 0] >= kw['x0'])).sum() > 50   # truncation rule exercised
    assert np.array_equal(it, ref['iters'])
    for key in got:
        assert _rel(got[key], ref[key]) < 1e-9, key
    assert np.abs(err - ref['error']).max() < 1e-9 * np.abs(ref['x']).max()
    dev.reset()                                  # sample by sample (one table pass each) against the blocked passes
    for _ in range(nt - 1):
        dev.step(1)
    single = dev.state(0, nt)
    assert np.array_equal(dev.info(1, nt)[1], it)
    for key in got:
        assert _rel(single[key], got[key]) < 1e-11, key
    dev.close()
    p = rng.standard_normal((50, P))
    assert _rel(sim.fir_conv_cy(fir, p, dt, 1), po.fir_conv(fir, p, dt, 1)) < 1e-12
    assert _rel(sim.fir_conv_py(fir, p, 1 / dt, 0), po.fir_conv(fir, p, dt, 0)) < 1e-12


def test_field_spatial_impulse_responses_vs_reference_loop(po):
    """configs[4] in small: SIR of every patch at many field points through the many-source pressure path +
    device impulse response, against the reference's loop (bem_pressure.pyx:91-117: one point, one
    frequency at a time) restated with the oracle's pressure vector and fft_to_sir."""
    from cnld import arrays, bem_pressure
    from cnld.sweep import FrequencySweep, SweepSetup
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(2, 1)), 3)
    sw = FrequencySweep(setup, nstreams=2)
    freqs = np.arange(1e6, 12.1e6, 1e6)
    _, xn, _ = sw.run(freqs)                                   # [freq, patch, node]
    sw.close()
    rng = np.random.default_rng(9)
    pts = np.c_[rng.uniform(-3e-4, 3e-4, 7), rng.uniform(-3e-4, 3e-4, 7), rng.uniform(1e-4, 1e-3, 7)]
    t, sir = bem_pressure.field_patch_pres_imp_resp(setup.vertices, setup.triangles, freqs, xn, pts, setup.c, setup.rho)
    oarr = po.matrix_array(nelem=(2, 1))
    am = po.array_mesh(oarr, 3)
    sfr = np.empty((xn.shape[1], len(pts), len(freqs)), dtype=np.complex128)
    for i, f in enumerate(freqs):
        sfr[:, :, i] = xn[i] @ po.pres_vector(am, pts, 2 * np.pi * f / setup.c).T
    to, ho = po.fft_to_fir(freqs, sfr, interp=2, use_kkr=False)
    assert sir.shape == ho.shape and np.allclose(t, to, rtol=1e-13, atol=0) and _rel(sir, ho) < 1e-9
