"""Device-resident FixedStepSolver at the C3 table size (P = 144 patches, 4000 taps on the solver's step):
seconds per time step once the history is longer than the table, table bytes per second against the HBM peak,
and the oracle's restatement (numpy einsum, the host's BLAS threads) on a bounded sample of the same steps.
usage: python scripts/time_tdsolver.py [--steps N] [--no-cpu]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
from cnld import _lib

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=1000)
ap.add_argument('--patches', type=int, default=144)
ap.add_argument('--taps', type=int, default=4000)
ap.add_argument('--no-cpu', action='store_true')
a = ap.parse_args()
P, nfir, dt = a.patches, a.taps, 2e-9
side = int(round(P**0.5))
rng = np.random.default_rng(3)
tt = np.arange(nfir) * dt
px, py = np.arange(P) % side, np.arange(P) // side
dist = np.hypot(px[:, None] - px[None, :], py[:, None] - py[None, :])
amp = 7e-7 / (1 + 5 * dist)**2 * (1 + 0.1 * rng.standard_normal((P, P)))
fir = (amp[:, :, None] * np.exp(-tt / 70e-9) * np.sin(2 * np.pi * 5.5e6 * tt + 0.3)).astype(np.float64)
warm = nfir + 50
nt = warm + a.steps + 2
t = np.arange(nt) * dt
v = np.outer(1 / (1 + np.exp(-(t - 60e-9) / 10e-9)), 19.5 + 1.5 * rng.random(P)) + 1.5 * np.outer(np.sin(2 * np.pi * 8e6 * t), np.ones(P))
gap_eff = np.full(P, 50e-9 + 100e-9 / 6.3)
kw = dict(k=2.5e14, n=1, x0=-30e-9, lmbd=2e13, atol=1e-12, maxiter=5)
l0 = _lib.launch_count()
t0 = time.perf_counter()
s = _lib.TimeDomainSolver(fir, v, gap_eff, dt, kw['k'], kw['n'], kw['x0'], kw['lmbd'], atol=kw['atol'], maxiter=kw['maxiter'])
t_create = time.perf_counter() - t0
t0 = time.perf_counter()
s.step(warm)
t_warm = time.perf_counter() - t0
t0 = time.perf_counter()
s.step(a.steps)
t_run = time.perf_counter() - t0
err, it = s.info(1, s.current_step)
st = s.state(0, s.current_step)
per = t_run / a.steps
out = dict(config='FixedStepSolver on the device', P=P, nfir=nfir, steps=a.steps, table_bytes=fir.nbytes, create_seconds=t_create,
           warm_steps=warm, warm_seconds=t_warm, seconds_per_step=per, steps_per_s=1 / per, table_gbs=fir.nbytes * (nfir - 1) / nfir / per / 1e9,
           convolutions_per_step_mean=float(it[warm:].mean()), contact_samples=int((st['x'] < kw['x0']).sum()),
           gpu_launches=_lib.launch_count() - l0, xmin=float(st['x'].min()))
if os.environ.get('CNLD_B200_TD_TRACE'):
    tr = s.trace(warm + 16, warm + 16 + 24).astype(np.int64)
    rel = tr[:, :7] - tr[:, :1]
    out['trace_us_from_kernel_start'] = (rel / 1e3).round(1).tolist()
    out['trace_start_to_start_us'] = (np.diff(tr[:, 0]) / 1e3).round(1).tolist()
    out['trace_wait_return_minus_prev_end_us'] = ((tr[1:, 2] - tr[:-1, 6]) / 1e3).round(1).tolist()
if not a.no_cpu:
    import pyoracle as po
    n_cpu = 20                       # the oracle redoes the same samples from the device's history: ~10-30 s of host work
    s0 = warm + 1
    p_hist = st['p_tot']
    t0 = time.perf_counter()
    worst = 0.0
    for q in range(n_cpu):
        base = po.fir_conv(fir, p_hist[:s0 + q], dt, 1)
        worst = max(worst, float(np.abs(base - (st['x'][s0 + q] - dt * (p_hist[s0 + q] @ fir[:, :, 0]))).max()))
    t_cpu = (time.perf_counter() - t0) / n_cpu
    out.update(cpu_seconds_per_base_convolution=t_cpu, cpu_sample=f'{n_cpu} base convolutions (one per step) with the oracle (numpy einsum)',
               cpu_cores=os.cpu_count(), base_conv_vs_oracle_abs=worst, speedup_vs_cpu_per_step=t_cpu / per)
print(json.dumps(out))
