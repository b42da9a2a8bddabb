#!/usr/bin/env python
"""
bench.py -- frequencies solved per second on the 4x4 CMUT array (BASELINE.json metric).

One "step" = one batch of frequencies (default 12 per GPU, 6 in flight) pushed through the whole hot path:
init G = K - w^2 M - j w B, accumulate -2 rho w^2 Z(k) (Galerkin Helmholtz SLP assembly), unpivoted
complex-FP64 LU, solve all P source patches, conj / boundary mask / patch average
(cnld/scripts/generate_db.py:30-130).  Workload = BASELINE.json configs[2] ("4x4 CMUT array
(~15k DOFs)", the configuration the metric is quoted on; it fits one B200): 16 square
membranes, refn = 21 -> N = 14 800 nodes, T = 28 224 triangles, P = 144 source patches,
frequencies taken from the 2000-point grid in [0.025, 50] MHz.

  python bench.py [--gpus N] [--steps K] [--warmup W]         this repo's CUDA path
  python bench.py --impl reference ...                        CPU arm: the oracle port of the
                                                              reference's H2Lib path on host cores
Multi-GPU (torchrun as the launcher, one rank per GPU): frequencies are sharded, no collective inside a
frequency ("weak" scaling: per-GPU batch fixed); setup broadcast once, patch responses gathered, both with
NCCL through the library's own C ABI (cnld_b200_comm_*, no torch.distributed).
  --config c4   16x16 array, ACA H-matrix + GMRES (BASELINE.json configs[3])
  --config c5   field pressure on 10^6 observation points (configs[4])
  --config td   time steps/s of the time-domain contact solver (FixedStepSolver) on a table
                of the 4x4 array's size; a simulation does not shard: --gpus N runs N independent replicas
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# The CPU arms (--impl reference, and the cpu_baseline leg) use every host core the process may run on.
# torchrun exports OMP_NUM_THREADS=1 to its workers; undo that before numpy / OpenBLAS / libgomp read it.
_NCORES = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
if 'reference' in sys.argv:
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(_NCORES)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))

METRIC = 'frequencies solved/s (Z assemble+factor+solve), 4x4 CMUT array'
UNIT = 'frequencies/s'
CONFIGS = {  # name: (nelem, refn, nfreq of the full sweep, fmin, fmax)
    'c1': ((1, 1), 15, 500, 1e6, 50e6),
    'c2': ((2, 2), 15, 1000, 0.05e6, 50e6),
    'c3': ((4, 4), 21, 2000, 0.025e6, 50e6),
    'c4': ((16, 16), 19, 500, 0.1e6, 50e6),      # ACA H-matrix + GMRES (bench_c4)
    'c5': ((4, 4), 21, 2000, 0.025e6, 50e6),     # C3 mesh + 10^6 observation points (bench_c5)
}
OBS_GRID = (100, 100, 100)                       # x, y in [-1, 1] mm, z in [0.1, 2.1] mm


def measured_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, taken from the
    committed ncu --set full summary of THIS round (profiles/r2_traffic.json, written by
    scripts/ncu_summary.py); None when that capture is absent."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')) as fh:
            return json.load(fh).get(kernel_key)
    except Exception:
        return None


def obs_grid():
    nx, ny, nz = OBS_GRID
    gx, gy, gz = np.meshgrid(np.linspace(-1e-3, 1e-3, nx), np.linspace(-1e-3, 1e-3, ny),
                             np.linspace(0.1e-3, 2.1e-3, nz), indexing='ij')
    return np.ascontiguousarray(np.c_[gx.ravel(), gy.ravel(), gz.ravel()])


def cublas_yardstick(m, n, k, device):
    """cuBLAS FP64 GEMM rates at the trailing update's shape -- an independent yardstick for the FP64
    roofline denominator, measured with torch.matmul OUTSIDE the product path (TFLOP/s; the complex one
    counted as 8 m n k)."""
    import torch
    out = {}
    for name, dt, fl in (('dgemm_tflops', torch.float64, 2.0), ('zgemm_tflops_8mnk', torch.complex128, 8.0)):
        try:
            a = torch.randn(m, k, dtype=dt, device=device)
            b = torch.randn(k, n, dtype=dt, device=device)
            c = torch.matmul(a, b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            out[name] = fl * m * n * k * 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
            del a, b, c
        except Exception as exc:      # yardstick only: never fail the bench on it
            out[name] = None
            out[name + '_error'] = str(exc)[:80]
    torch.cuda.empty_cache()
    return out


def workload_freqs(cfg, count, offset=0):
    """`count` frequencies spread over the config's sweep grid (deterministic)."""
    _, _, nf, f0, f1 = CONFIGS[cfg]
    grid = np.linspace(f0, f1, nf)
    stride = 137 if cfg == 'c4' else 37      # c4: few, expensive steps -- spread them over the whole band (high k too)
    idx = (np.arange(count) * stride + offset * 101 + 11) % nf
    return grid[idx]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_step(po, am, blocks, freqs, c, rho):
    """The oracle's restatement of generate_db.process on the host cores (all OpenMP/BLAS threads)."""
    for f in freqs:
        po.solve_frequency(am, blocks, float(f), c=c, rho=rho, blocked=True)


def run_reference(args, rank, world):
    """--impl reference: CPU implementation of the path (oracle port -- H2Lib itself is not
    buildable here) timed on the box's host cores.  Rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    po.build()
    nelem, refn = CONFIGS[args.config][:2]
    arr = po.matrix_array(nelem=nelem)
    am = po.array_mesh(arr, refn)
    blocks = po.array_fem(arr, refn)
    cores = po.lib().orc_num_threads()
    try:                                   # BLAS threads of the blocked LU follow the same count
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=_NCORES)
    except Exception:
        pass
    # each step = ONE full frequency of the workload (bounded sample of the per-GPU batch)
    freqs = workload_freqs(args.config, args.steps + args.warmup)
    for i in range(args.warmup):
        cpu_reference_step(po, am, blocks, freqs[i:i + 1], 1500., 1000.)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_reference_step(po, am, blocks, freqs[args.warmup + i:args.warmup + i + 1], 1500., 1000.)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex128 (f64)',
        'data': 'synthetic',
        'config': {'workload': f'{args.config}: {nelem[0]}x{nelem[1]} CMUT array, refn={refn}, N={am.nvertices}, '
                               f'T={am.ntriangles}, P={sum(b[3].shape[1] for b in blocks)}',
                   'frequencies_per_step': 1},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'one full frequency per step (oracle port of the H2Lib path: OpenMP C '
                                   'assembly + BLAS-3 blocked unpivoted LU); H2Lib itself not buildable offline'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# C5: field pressure of the 4x4 array on 10^6 observation points (BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------------
C5_METRIC = 'field-pressure frequencies/s (10^6 observation points x 144 source patches), 4x4 CMUT array'


def _c5_oracle_sample(po, am, obs, k, disp, idx, block=2000):
    """p[src, idx] by the oracle: disp . mesh_pres_vector(obs[idx], k) (bem_pressure.pyx:54-88, :110-112)."""
    out = np.empty((disp.shape[0], len(idx)), dtype=np.complex128)
    for b0 in range(0, len(idx), block):
        pv = po.pres_vector(am, obs[idx[b0:b0 + block]], k)          # [nb, nv]
        out[:, b0:b0 + block] = disp @ pv.T
    return out


def bench_c5_reference(args, rank):
    """CPU arm of C5: the oracle's mesh_pres_vector + dot on a bounded sample of the 10^6 points."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    po.build()
    nelem, refn = CONFIGS['c5'][:2]
    arr = po.matrix_array(nelem=nelem)
    am = po.array_mesh(arr, refn)
    P = 9 * nelem[0] * nelem[1]
    obs = obs_grid()
    rng = np.random.default_rng(0)
    disp = rng.standard_normal((P, am.nvertices)) + 1j * rng.standard_normal((P, am.nvertices))
    nsample = 4000
    freqs = workload_freqs('c5', args.steps + args.warmup)
    steps = min(args.steps, 3)
    idx = np.sort(rng.choice(len(obs), nsample, replace=False))
    for f in freqs[:min(args.warmup, 1)]:
        _c5_oracle_sample(po, am, obs, 2 * np.pi * f / 1500., disp, idx[:500])
    t0 = time.perf_counter()
    for f in freqs[args.warmup:args.warmup + steps]:
        _c5_oracle_sample(po, am, obs, 2 * np.pi * f / 1500., disp, idx)
    dt = (time.perf_counter() - t0) / steps
    val = 1.0 / (dt * len(obs) / nsample)
    print(json.dumps({
        'impl': 'reference', 'metric': C5_METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'complex128 (f64)', 'data': 'synthetic',
        'config': {'workload': f'c5: 4x4 CMUT array refn={refn} (N={am.nvertices}, T={am.ntriangles}), {len(obs)} observation points, P={P}'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': po.lib().orc_num_threads(), 'kind': 'port',
                         'sample': f'{nsample} of {len(obs)} observation points per step, extrapolated linearly'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}), flush=True)


def bench_c5(args, rank, world, local):
    import torch
    from cnld import _lib, arrays
    from cnld.sweep import FrequencySweep, SweepSetup, broadcast_setup
    if _lib.device_count() == 0:
        raise SystemExit('bench.py: no CUDA device -- libcnld_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    comm = _lib.Comm(rank, world, local)
    nelem, refn = CONFIGS['c5'][:2]
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn) if rank == 0 else None
    if world > 1:
        setup = broadcast_setup(setup, rank, comm=comm)
    N, P, T = setup.nnodes, setup.npatch, len(setup.triangles)
    obs = obs_grid()
    nobs = len(obs)
    nsteps = args.steps + args.warmup
    # inputs: this rank's frequencies and their node displacements (C3's dense sweep, untimed)
    freqs = workload_freqs('c5', nsteps, offset=rank)
    sweep = FrequencySweep(setup, device=local, nstreams=1)
    _, disp, _ = sweep.run(freqs, want_nodes=True)            # [nsteps, P, N]
    ks = 2 * np.pi * freqs / setup.c
    pf = _lib.PressureField(sweep.ctx, obs, P)

    def barrier():
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        return comm.max(x)

    # ---- device-resident: points and displacements in HBM, results stay in HBM ----------------------
    pf.run(ks[0], disp[0], setup.c, setup.rho, fetch=False)
    for w in range(args.warmup):
        pf.run_device(ks[w], P, setup.c, setup.rho)
    fp64_peak = _lib.bench_dfma(3, device=local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    barrier()
    dev_s = pv_s = ct_s = 0.0
    for s_ in range(args.steps):
        pf.run_device(ks[args.warmup + s_], P, setup.c, setup.rho)
        inf = pf.info()
        dev_s += inf['total_seconds']; pv_s += inf['pvec_seconds']; ct_s += inf['contract_seconds']
    barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.finish() if rank == 0 else None
    dt = reduce_max(dev_s)
    value = args.steps * world / dt

    # ---- end to end: displacements from host memory, the whole field back to host memory ------------
    out = np.empty((P, nobs), dtype=np.complex128)
    pf.run(ks[0], disp[0], setup.c, setup.rho, out=out)        # untimed: page-locks `out`
    barrier()
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        i = args.warmup + s_
        pf.run(ks[i], disp[i], setup.c, setup.rho, out=out)
    barrier()
    dte = reduce_max(time.perf_counter() - t0)
    e2e = args.steps * world / dte
    assert np.isfinite(out).all() and np.abs(out).max() > 0
    inf = pf.info()

    line = None
    if rank == 0:
        flops = 6.0 * nobs * P * N * args.steps               # executed real flops of the 3M contraction
        ach = flops / ct_s / 1e12
        evals = 3.0 * T * nobs * args.steps
        line = {
            'metric': C5_METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'complex128 (f64)', 'data': 'synthetic',
            'config': {'workload': f'c5: 4x4 CMUT array refn={refn} (N={N}, T={T}), {nobs} observation points on a '
                                   f'{OBS_GRID[0]}x{OBS_GRID[1]}x{OBS_GRID[2]} grid, P={P} source patches, one frequency per step',
                       'parallelism': f'freq-shard x{world}', 'obs_per_chunk': inf['obs_per_chunk'], 'chunks': inf['chunks'],
                       'cache': 'node-space pressure vectors of one chunk = 9 GB >> 126 MB L2'},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(disp[0].nbytes), 'd2h_bytes_per_step': int(out.nbytes)},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': ach / fp64_peak,
                         'traffic': measured_traffic('zgemm3m_pfield'),
                         'kernel': 'zgemm3m_kernel: p[src][obs] = PV[obs][node] . disp[src][node], 6 m n k executed flops '
                                   '(3M); durations by CUDA events on the launch stream inside the timed region',
                         'peak_source': 'FP64 FMA-pipe peak measured live (cnld_b200_bench_dfma)',
                         'pvec_kernel': {'kernel': 'k_pvec_tiles (FP64 pipe: rsqrt + sincos per point pair)',
                                         'kernel_evaluations_per_s': evals / pv_s, 'seconds_per_step': pv_s / args.steps},
                         'contract_seconds_per_step': ct_s / args.steps},
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pyoracle as po
        po.build()
        am = po.array_mesh(po.matrix_array(nelem=nelem), refn)
        rng = np.random.default_rng(0)
        idx = np.sort(rng.choice(nobs, 10000, replace=False)).astype(np.uint32)
        i = args.warmup + args.steps - 1                       # the last timed step's result is still on the device
        got = pf.fetch(P, idx)
        t0 = time.perf_counter()
        ref = _c5_oracle_sample(po, am, obs, ks[i], disp[i], idx)
        tc = time.perf_counter() - t0
        rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        line['parity'] = {'config': 'c5', 'sampled_points': len(idx), 'rel_l2': rel, 'tolerance': 1e-6,
                          'against': 'oracle mesh_pres_vector . displacement'}
        line['cpu_baseline'] = {'value': 1.0 / (tc * nobs / len(idx)), 'unit': UNIT, 'cores': po.lib().orc_num_threads(),
                                'kind': 'port', 'sample': f'{len(idx)} of {nobs} observation points, extrapolated linearly'}
        if not rel < 1e-6:
            print(json.dumps(line), flush=True)
            raise SystemExit('bench.py: c5 field pressure differs from the oracle by more than 1e-6')
    if rank == 0:
        print(json.dumps(line), flush=True)
    pf.close()
    sweep.close()
    comm.close()


C4_METRIC = 'frequencies solved/s (ACA H-matrix assemble + GMRES, all 2304 source patches), 16x16 CMUT array'
C4_H = dict(clf=16, eta=0.8, admis='2', strict=True, kmax=64)           # generate_db.py:264-275 defaults
C4_SOLVE = dict(eps_aca=1e-2, tol=1e-6, restart=50, maxiter=300, batch=96)


def _c4_cpu_sample(po, am, k, tree, frac=64, seed=0):
    """Oracle H-matrix work on every `frac`-th leaf of the block tree: assemble (ACA / near field) and apply to
    one vector, in C with OpenMP over the leaves (oracle: orc_hleaves_sample).
    Returns (assemble seconds, matvec seconds, leaves sampled, leaves total)."""
    ct, rc, cc, adm = tree
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(am.nvertices) + 1j * rng.standard_normal(am.nvertices)
    pick = np.arange(seed % frac, len(rc), frac)
    _, info = po.hleaves_sample(am, k, ct, rc, cc, adm, pick, x, eps_aca=C4_SOLVE['eps_aca'], kmax=C4_H['kmax'])
    return info['assemble_seconds'], info['matvec_seconds'], len(pick), len(rc)


def _c4_cpu_line(po, am, freqs, P, iters_mean):
    """CPU estimate of one C4 frequency from a 1/64 sample of the leaves (assembly + one matvec),
    extrapolated to all leaves, all P right-hand sides and the measured GMRES iteration count."""
    ta = tm = 0.0
    ct = po.ClusterTree(am, C4_H['clf'])
    tree = (ct,) + tuple(po.build_block(ct, C4_H['eta'], C4_H['strict']))
    for f in freqs:
        a, m, ns, nl = _c4_cpu_sample(po, am, 2 * np.pi * f / 1500., tree)
        ta += a * nl / ns
        tm += m * nl / ns
    ta, tm = ta / len(freqs), tm / len(freqs)
    sec = ta + tm * P * (iters_mean + 2)          # + initial residual and the final update
    return sec, {'value': 1.0 / sec, 'unit': UNIT, 'cores': po.lib().orc_num_threads(), 'kind': 'port',
                 'sample': f'every 64th H-matrix leaf assembled (OpenMP over the leaves) and applied to one vector by the oracle, '
                           f'extrapolated x64 x {P} right-hand sides x {iters_mean + 2:.1f} matvecs per right-hand side (single-vector, '
                           f'serial matvec: an upper bound on the CPU matvec time); '
                           f'assembly {ta:.1f} s + matvecs {tm * P * (iters_mean + 2):.0f} s per frequency'}


def bench_c4_reference(args, rank):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    po.build()
    nelem, refn = CONFIGS['c4'][:2]
    am = po.array_mesh(po.matrix_array(nelem=nelem), refn)
    P = 9 * nelem[0] * nelem[1]
    steps = min(args.steps, 2)
    freqs = workload_freqs('c4', steps)
    t0 = time.perf_counter()
    sec, base = _c4_cpu_line(po, am, freqs, P, 15.0)
    print(json.dumps({
        'impl': 'reference', 'metric': C4_METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': 0, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'complex128 (f64)', 'data': 'synthetic',
        'config': {'workload': f'c4: 16x16 CMUT array refn={refn} (N={am.nvertices}), P={P}, GMRES iterations assumed 15',
                   'wall_seconds_sampled': time.perf_counter() - t0},
        'cpu_baseline': base, 'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}), flush=True)


def bench_c4(args, rank, world, local):
    """configs[3]: 16x16 array (N = 194 816, P = 2 304), one frequency per step through the ACA H-matrix +
    batched GMRES path; frequencies spread over [0.1, 50] MHz, dealt cyclically to the ranks."""
    import torch
    from cnld import _lib, arrays
    from cnld.sweep import SweepSetup, broadcast_setup
    if _lib.device_count() == 0:
        raise SystemExit('bench.py: no CUDA device -- libcnld_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    comm = _lib.Comm(rank, world, local)
    nelem, refn = CONFIGS['c4'][:2]
    t_setup = time.perf_counter()
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn) if rank == 0 else None
    if world > 1:
        setup = broadcast_setup(setup, rank, comm=comm)
    ctx = _lib.Context(setup.vertices, setup.triangles, q_reg=setup.q_reg, q_sing=setup.q_sing, device=local)
    ctx.set_mbk(setup.K, setup.M, setup.B)
    ctx.set_loads(setup.F, setup.AVG, setup.on_boundary)
    H = _lib.HMatrix(ctx, **C4_H)
    t_setup = time.perf_counter() - t_setup
    N, P, T = setup.nnodes, setup.npatch, len(setup.triangles)
    nsteps = args.steps + args.warmup
    # step s of rank r solves grid frequency (s * world + r): the cyclic deal of sweep.shard_cyclic
    allf = workload_freqs('c4', 2 * nsteps * world)
    mine = allf[rank::world]

    def barrier():
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        return comm.max(x)

    for w in range(args.warmup):
        H.sweep(mine[w:w + 1], c=setup.c, rho=setup.rho, want_nodes=False, **C4_SOLVE)
    fp64_peak = _lib.bench_dfma(3, device=local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0, fl0 = _lib.launch_count(), _lib.dmma_flop_count()
    barrier()
    dev_s, its, capped = 0.0, [], 0
    for s_ in range(args.steps):
        f = mine[args.warmup + s_:args.warmup + s_ + 1]
        xp, _, it, t = H.sweep(f, c=setup.c, rho=setup.rho, want_nodes=False, **C4_SOLVE)
        dev_s += float(t.sum())
        its.append(float(it.mean()))
        capped += H.capped()[1]
    barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.finish() if rank == 0 else None
    dt = reduce_max(dev_s)
    value = args.steps * world / dt
    # ---- end to end: the same call by the wall clock (frequencies in, patch responses back on the host) --
    comm.gather(xp)                         # untimed: NCCL opens its peer connections on first use
    barrier()
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        f = mine[nsteps + s_:nsteps + s_ + 1]
        xp, _, it2, _ = H.sweep(f, c=setup.c, rho=setup.rho, want_nodes=False, **C4_SOLVE)
        comm.gather(xp)                     # patch responses to rank 0 (NCCL, NVLink)
    barrier()
    dte = reduce_max(time.perf_counter() - t0)
    e2e = args.steps * world / dte
    assert np.isfinite(xp).all() and np.abs(xp).max() > 0
    line = None
    if rank == 0:
        # ---- roofline leg: the H-matvec with the GMRES batch width, on the H-matrix of the last frequency ----
        info = H.info()
        nb = C4_SOLVE['batch']
        mv_s = H.bench_matvec(nb, 5)
        mv1_s = H.bench_matvec(1, 5)
        ach = 6.0 * info['stored'] * nb / mv_s / 1e12
        hbm = MEASURED_HBM()
        # ---- parity: sampled rows of the matvec against exactly integrated rows ----------------------
        f_last = float(mine[nsteps + args.steps - 1])
        k_last = 2 * np.pi * f_last / setup.c
        rng = np.random.default_rng(0)
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        rows = np.sort(rng.choice(N, 48, replace=False)).astype(np.uint32)
        y = H.matvec(x)
        exact = ctx.nearfield(k_last, rows, np.arange(N, dtype=np.uint32)) @ x
        row_err = float(np.linalg.norm(y[rows] - exact) / np.linalg.norm(exact))
        line = {
            'metric': C4_METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'complex128 (f64)', 'data': 'synthetic',
            'config': {'workload': f'c4: 16x16 CMUT array refn={refn} (N={N}, T={T}), P={P} right-hand sides per frequency, '
                                   f'ACA H-matrix ({C4_H}) + batched GMRES ({C4_SOLVE}), one frequency per step per GPU',
                       'frequencies_hz_rank0': [float(v) for v in mine[args.warmup:nsteps]],
                       'parallelism': f'freq-shard (cyclic) x{world}', 'setup_seconds_once': t_setup,
                       'hmatrix': info, 'gmres_mean_iterations': its, 'rank_capped_leaves': int(capped),
                       'cache': 'leaf data 4.4 GB >> 126 MB L2'},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': 8, 'd2h_bytes_per_step': int(xp.nbytes + it2.nbytes + 8)},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': ach / fp64_peak,
                         'traffic': measured_traffic('hmv_batched'),
                         'kernel': f'batched H-matvec ({nb} vectors: k_hmv_tstack + k_hmv_gemm on DMMA, 3M product: '
                                   f'6 x stored entries x vectors executed flops), {mv_s * 1e3:.2f} ms per call, timed alone '
                                   'with CUDA events after the timed region',
                         'peak_source': 'FP64 FMA-pipe peak measured live (cnld_b200_bench_dfma)',
                         'single_vector': {'bound': 'hbm', 'achieved': 16.0 * info['stored'] / mv1_s / 1e9, 'peak': hbm[0],
                                           'unit': 'GB/s', 'frac': 16.0 * info['stored'] / mv1_s / 1e9 / hbm[0],
                                           'peak_source': hbm[1], 'seconds': mv1_s}},
            'parity': {'config': 'c4', 'frequency_hz': f_last, 'matvec_sampled_rows_rel_err_vs_exact': row_err,
                       'tolerance': C4_SOLVE['eps_aca'], 'gmres_all_converged': True,
                       'note': 'solution-level parity of the H path against the dense LU path: tests/test_gpu_fullsize.py'},
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pyoracle as po
        po.build()
        am = po.array_mesh(po.matrix_array(nelem=nelem), refn)
        _, line['cpu_baseline'] = _c4_cpu_line(po, am, mine[args.warmup:args.warmup + 1], P, float(np.mean(its)))
    if rank == 0:
        print(json.dumps(line), flush=True)
    H.close()
    ctx.close()
    comm.close()
    if line is not None and not line['parity']['matvec_sampled_rows_rel_err_vs_exact'] < C4_SOLVE['eps_aca']:
        raise SystemExit('bench.py: c4 H-matvec differs from exactly integrated rows by more than eps_aca')


TD_METRIC = 'time steps/s (FixedStepSolver: contact simulation on the patch-to-patch impulse responses), 4x4 CMUT array table'
TD_KW = dict(k=3e14, n=1, x0=-15e-9, lmbd=3e13, atol=1e-12, maxiter=5)


def td_scenario(P=144, nfir=4000, nt=400, seed=3):
    """Synthetic table of the 4x4 array's size (P x P damped oscillators, coupling falling off with the patch distance),
    a DC ramp + tone drive that takes the patches into contact; contact threshold inside the stable electrostatic range."""
    dt = 2e-9
    side = int(round(P**0.5))
    rng = np.random.default_rng(seed)
    tt = np.arange(nfir) * dt
    px, py = np.arange(P) % side, np.arange(P) // side
    dist = np.hypot(px[:, None] - px[None, :], py[:, None] - py[None, :])
    amp = 7e-7 / (1 + 5 * dist)**2 * (1 + 0.1 * rng.standard_normal((P, P)))
    fir = amp[:, :, None] * np.exp(-tt / 70e-9) * np.sin(2 * np.pi * 5.5e6 * tt + 0.3)
    t = np.arange(nt) * dt
    v = np.outer(1 / (1 + np.exp(-(t - 60e-9) / 10e-9)), 19.5 + 1.5 * rng.random(P)) + \
        1.5 * np.outer(np.sin(2 * np.pi * 8e6 * t), np.ones(P))
    return fir, t, v, np.full(P, 50e-9 + 100e-9 / 6.3), dt


def bench_td_reference(args, rank):
    """--impl reference --config td: the oracle's restatement of FixedStepSolver.step (numpy on the host's BLAS threads)
    on a bounded sample of the same samples: a history longer than the table, 8 samples per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    per = 8
    fir, t, v, gap_eff, dt = td_scenario(nt=4100)
    rng = np.random.default_rng(1)
    hist = -1e6 * rng.random((4050, fir.shape[0]))           # a pressure history of the right size and magnitude
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for s_ in range(steps * per):
        base = po.fir_conv(fir, hist, dt, 1)
        for _ in range(6):                                    # the fixed point's products on the current sample
            base + dt * (hist[-1] @ fir[:, :, 0])
    dts = (time.perf_counter() - t0) / steps
    val = per / dts
    print(json.dumps({
        'impl': 'reference', 'metric': TD_METRIC, 'value': val, 'unit': 'steps/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': 0, 'ms_per_step': dts * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': {'workload': 'td: P=144 patches, 4000-tap table, history longer than the table'},
        'cpu_baseline': {'value': val, 'unit': 'steps/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': f'{per} samples per step: one base convolution over the whole table + 6 zero-lag products each '
                                   '(oracle restatement, numpy einsum)'},
        'e2e': {'value': val, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}), flush=True)


def bench_td(args, rank, world, local):
    """--config td: the time-domain consumer of the sweep's database.  One bench step =
    `--td-samples` time steps of FixedStepSolver on one GPU; N > 1 = independent replicas (a simulation does not shard)."""
    os.environ['CNLD_B200_DEVICE'] = str(local)
    from cnld import _lib, simulation
    if _lib.device_count() == 0:
        raise SystemExit('bench.py: no CUDA device -- libcnld_b200 has no CPU fallback')
    comm = _lib.Comm(rank, world, local) if world > 1 else None
    ns = args.td_samples
    warm = 4000 + 64
    nt = warm + ns * (args.steps + args.warmup) + 2
    fir, t, v, gap_eff, dt = td_scenario(nt=nt, seed=3 + rank)
    P, nfir = fir.shape[0], fir.shape[2]
    sol = _lib.TimeDomainSolver(fir, v, gap_eff, dt, TD_KW['k'], TD_KW['n'], TD_KW['x0'], TD_KW['lmbd'], atol=TD_KW['atol'],
                                maxiter=TD_KW['maxiter'], device=local)
    sol.step(warm)                                            # history longer than the table from here on
    for _ in range(args.warmup):
        sol.step(ns)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if comm:
        comm.barrier()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sol.step(ns)                                          # synchronous: returns when the samples are final on the device
    dts = time.perf_counter() - t0
    launches = _lib.launch_count() - l0
    if comm:
        dts = comm.max(dts)
    clocks = sampler.finish() if rank == 0 else None
    value = ns * args.steps * world / dts
    err, it = sol.info(1, sol.current_step)
    st = sol.state(0, 321)
    contact = int((sol.state(warm, sol.current_step)['x'] < TD_KW['x0']).sum())
    sol.close()

    # ---- end to end through the reference-facing class: every bench step brings all six state arrays of its samples back
    # to the host (StateDB); the table and the drive go to the device once, in the constructor (reported, not timed --
    # like the sweep's setup; most of it is the reference's own cubic resampling of the table with scipy on the host)
    t0 = time.perf_counter()
    fs = simulation.FixedStepSolver((np.arange(nfir + 1) * dt, np.concatenate([fir, fir[:, :, -1:]], axis=2)), (t, v), gap_eff,
                                    gap_eff, (0.0, (nt - 1) * dt, dt), TD_KW['k'], TD_KW['n'], TD_KW['x0'], TD_KW['lmbd'],
                                    atol=TD_KW['atol'], maxiter=TD_KW['maxiter'], device=local)
    t_setup = time.perf_counter() - t0
    fs.step(warm)
    for _ in range(args.warmup):
        fs.step(ns)
    if comm:
        comm.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fs.step(ns)
    dte = time.perf_counter() - t0
    if comm:
        dte = comm.max(dte)
    e2e = ns * args.steps * world / dte
    h2d, d2h = 0, 6 * ns * P * 8 + ns * 12
    x_e2e = fs.displacement
    assert np.isfinite(x_e2e).all() and x_e2e.shape[0] == fs.current_step
    fs.close()

    line = None
    if rank == 0:
        # the history pass alone: in line (no run-ahead), from the phase marks of the step kernels around it
        os.environ['CNLD_B200_TD_TRACE'] = '1'
        os.environ['CNLD_B200_TD_OVERLAP'] = '0'
        s2 = _lib.TimeDomainSolver(fir, v, gap_eff, dt, TD_KW['k'], TD_KW['n'], TD_KW['x0'], TD_KW['lmbd'], atol=TD_KW['atol'],
                                   maxiter=TD_KW['maxiter'], device=local)
        s2.step(warm + 64)
        tr = s2.trace(warm, warm + 64).astype(np.int64)
        s2.close()
        del os.environ['CNLD_B200_TD_TRACE'], os.environ['CNLD_B200_TD_OVERLAP']
        gaps = (tr[1:, 2] - tr[:-1, 6]) / 1e9                 # previous kernel's last store -> this kernel's dependency met
        t_pass = float(np.median(np.sort(gaps)[-3:]))         # the block boundaries (16 samples apart) hold the pass
        t_step = float(np.median(np.diff(tr[:, 0]))) / 1e9
        hbm, hbm_src = MEASURED_HBM()
        ach = fir.nbytes * (nfir - 1) / nfir / t_pass / 1e9
        line = {
            'metric': TD_METRIC, 'value': value, 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dts / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'td: FixedStepSolver, P={P} patches, {nfir}-tap table ({fir.nbytes / 1e6:.0f} MB), history longer '
                                   f'than the table, contact engaged ({contact} patch samples in contact), {ns} time steps per bench step',
                       'solver': TD_KW, 'parallelism': f'replicas x{world}', 'convolutions_per_sample_mean': float(it[warm:].mean()),
                       'cache': 'table 663 MB >> 126 MB L2'},
            'e2e': {'value': e2e, 'unit': 'steps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'note': 'cnld.simulation.FixedStepSolver.step(n): the six state arrays, errors and iteration counts of the n samples '
                            'copied back into the host StateDB inside the timed region; table and drive uploaded once by the constructor',
                    'setup_seconds_once': t_setup},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm, 'traffic': None,
                         'kernel': 'k_fir_far_mma<2>: the history pass (one read of the table, 16 lags per entry on DMMA) timed in '
                                   'line from the step kernels\' phase marks; in the timed region it runs ahead on a second stream '
                                   'under the previous block and serves 16 samples',
                         'peak_source': hbm_src, 'seconds_per_pass': t_pass, 'table_bytes_per_sample': fir.nbytes / 16,
                         'step_kernel_seconds': t_step,
                         'note': 'a sample costs one cluster step kernel (latency bound: ~6 fixed-point iterations of scalar '
                                 'code between barriers) + 1/16 of a hidden pass'},
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pyoracle as po
        ref = po.td_solve(fir, v[:322], gap_eff, dt, TD_KW['k'], TD_KW['n'], TD_KW['x0'], TD_KW['lmbd'], TD_KW['atol'],
                          TD_KW['maxiter'], nsteps=320)
        rel = {key: float(np.linalg.norm(st[key] - ref[key][:321]) / max(np.linalg.norm(ref[key][:321]), 1e-300))
               for key in ('x', 'u', 'p_tot', 'p_cont_spr')}
        same_it = bool(np.array_equal(it[:320], ref['iters']))
        line['parity'] = {'config': 'td', 'samples': 320, 'rel_l2': rel, 'iteration_counts_equal': same_it, 'tolerance': 1e-9,
                          'against': 'oracle restatement of FixedStepSolver (pinned to the reference class by tests/golden/td_solver.npz)'}
        rng = np.random.default_rng(1)
        hist = -1e6 * rng.random((4050, P))
        t0 = time.perf_counter()
        for _ in range(16):
            base = po.fir_conv(fir, hist, dt, 1)
            for _ in range(6):
                base + dt * (hist[-1] @ fir[:, :, 0])
        tc = (time.perf_counter() - t0) / 16
        line['cpu_baseline'] = {'value': 1.0 / tc, 'unit': 'steps/s', 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': '16 samples with a history longer than the table: one base convolution + 6 zero-lag '
                                          'products each (oracle restatement, numpy einsum)'}
        if not (max(rel.values()) < 1e-9 and same_it):
            print(json.dumps(line), flush=True)
            raise SystemExit('bench.py: td solver differs from the oracle')
    if rank == 0:
        print(json.dumps(line), flush=True)
    if comm:
        comm.close()


def MEASURED_HBM():
    """(GB/s, source) of the HBM roofline denominator: MEASURED_PEAKS.json if present, else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md: 6.65 TB/s, an earlier measurement on this pool)'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS) + ['td'])
    ap.add_argument('--td-samples', type=int, default=2048, help='td: time steps per bench step')
    ap.add_argument('--freqs-per-step', type=int, default=12)
    ap.add_argument('--streams', type=int, default=6)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--general', action='store_true', help='general unpivoted LU of G instead of the symmetric-part path')
    ap.add_argument('--c4-batch', type=int, default=0, help='c4: right-hand sides per GMRES batch (default: C4_SOLVE)')
    ap.add_argument('--c4-restart', type=int, default=0, help='c4: GMRES restart length (default: C4_SOLVE)')
    ap.add_argument('--lu-outer', type=int, default=0, help='debug: outer LU block (multiple of 64)')
    ap.add_argument('--lu-reserve', type=int, default=-1, help='debug: SMs left free by the look-ahead trailing update')
    args = ap.parse_args()
    if args.c4_batch:
        C4_SOLVE['batch'] = args.c4_batch
    if args.c4_restart:
        C4_SOLVE['restart'] = args.c4_restart

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if args.config == 'c5':
            return bench_c5_reference(args, rank)
        if args.config == 'c4':
            return bench_c4_reference(args, rank)
        if args.config == 'td':
            return bench_td_reference(args, rank)
        if args.steps > 2:
            args.steps, args.warmup = min(args.steps, 2), min(args.warmup, 1)
        run_reference(args, rank, world)
        return
    if args.config == 'c5':
        return bench_c5(args, rank, world, local)
    if args.config == 'c4':
        return bench_c4(args, rank, world, local)
    if args.config == 'td':
        return bench_td(args, rank, world, local)

    import torch
    from cnld import _lib, arrays
    from cnld.sweep import FrequencySweep, SweepSetup, broadcast_setup

    if _lib.device_count() == 0:
        raise SystemExit('bench.py: no CUDA device -- libcnld_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    # multi-GPU plumbing: NCCL through the library's own C ABI (cnld_b200_comm_*); torchrun is only the launcher
    comm = _lib.Comm(rank, world, local)

    nelem, refn = CONFIGS[args.config][:2]
    t_setup = time.perf_counter()
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn) if rank == 0 else None
    if world > 1:
        setup = broadcast_setup(setup, rank, comm=comm)
    if args.lu_outer:
        _lib.lib().cnld_b200_set_lu_outer(args.lu_outer)
    if args.lu_reserve >= 0:
        _lib.lib().cnld_b200_set_lu_reserve(args.lu_reserve)
    sweep = FrequencySweep(setup, device=local, nstreams=args.streams, symmetric=not args.general)
    t_setup = time.perf_counter() - t_setup
    N, P, T = setup.nnodes, setup.npatch, len(setup.triangles)
    nfb = args.freqs_per_step

    def barrier():
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.synchronize()

    def batch(step):
        return workload_freqs(args.config, nfb, offset=step * world + rank)

    # ---- device-resident timing (`value`) ---------------------------------------------------
    for w in range(args.warmup):
        sweep.run_device(batch(w))
    fp64_peak = _lib.bench_dfma(3, device=local)         # measured FP64 FMA-pipe peak, TFLOP/s
    # the same kernel alone on the GPU at the first trailing update's shape (conventional 8mnk rate)
    alone_8mnk = _lib.bench_zgemm(N - 1024, N - 1024, 512, 3, device=local) if args.config == 'c3' else None
    # independent yardstick for the FP64 peak: cuBLAS (through torch, outside the product) at the same shape
    yard = cublas_yardstick(N - 1024, N - 1024, 512, torch.device('cuda', local)) if rank == 0 and N > 2048 else None
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    fl0 = _lib.dmma_flop_count()
    stage = np.zeros(3)
    dev_s = 0.0
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        sweep.run_device(batch(args.warmup + s))
        st = sweep.ctx.stage_times()
        stage += [st['assemble'], st['factor'], st['solve']]
        dev_s += st['device_seconds']       # CUDA events joined across the library's streams
    barrier()
    wall = time.perf_counter() - t0
    dt = dev_s
    launches = _lib.launch_count() - l0
    step_flops = _lib.dmma_flop_count() - fl0      # executed DMMA flops of this rank's timed steps
    step_tflops = step_flops / dev_s / 1e12
    clocks = sampler.finish() if rank == 0 else None
    dt = comm.max(dt)
    value = args.steps * nfb * world / dt

    # ---- end to end through the public API (host buffers in, host buffers out) -------------
    out_nodes = np.empty((nfb, P, N), dtype=np.complex128)
    sweep.upload()
    xp0, _, _ = sweep.run(batch(0), want_nodes=True, out_nodes=out_nodes)     # untimed: pinned staging is allocated here
    comm.gather(xp0)                                              # untimed: NCCL opens its peer connections on first use
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for s in range(args.steps):
        h2d = sweep.upload() + 8 * nfb
        xp, xn, _ = sweep.run(batch(args.warmup + s), want_nodes=True, out_nodes=out_nodes)
        d2h = xp.nbytes + xn.nbytes
        gathered = comm.gather(xp)          # the sweep's one exchange: patch responses to rank 0 (NCCL, NVLink)
    barrier()
    dte = comm.max(time.perf_counter() - t0)
    e2e = args.steps * nfb * world / dte
    if rank == 0:
        assert gathered.shape == (world,) + xp.shape and np.isfinite(gathered).all()
    assert np.isfinite(xp).all() and np.abs(xp).max() > 0

    sym_stats = sweep.ctx.symmetric_stats()
    trans = sweep.ctx.translation_info()

    # ---- roofline leg: the dominant kernel's launch durations, CUDA events on its own stream ----
    # A dedicated single-stream pass of the same workload: with two frequencies in flight the other
    # slot's kernels share the SMs with a trailing update and its event-bracketed duration would
    # measure the sharing, not the kernel.
    kflops = ksec = klaunch = 0.0
    if rank == 0:
        sweep.ctx.set_streams(1)
        sweep.ctx.set_profile(True)
        sweep.run_device(batch(0)[:1])
        for s in range(2):
            sweep.run_device(batch(1 + s)[:1])
            ks = sweep.ctx.kernel_stats()
            kflops, ksec, klaunch = kflops + ks['flops'], ksec + ks['seconds'], klaunch + ks['launches']
    line = None
    if rank == 0:
        ach = kflops / ksec / 1e12 if ksec > 0 else None
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex128 (f64)', 'data': 'synthetic',
            'config': {'workload': f'{args.config}: {nelem[0]}x{nelem[1]} CMUT array, refn={refn}, N={N}, T={T}, P={P}, '
                                   + ('dense Z + general unpivoted LU' if args.general else
                                      'dense Z (lower triangle + sparse Sauter-Schwab asymmetry) + unpivoted elimination of the '
                                      'symmetric part (S = L D L^T) + 1 refinement step; results equal the general LU to 1e-11, '
                                      'checked per frequency on the device'),
                       'symmetric_stats': sym_stats,
                       'frequencies_per_step_per_gpu': nfb, 'streams': args.streams, 'parallelism': f'freq-shard x{world}',
                       'cache': 'per-frequency working set 3.5 GB >> 126 MB L2 (inputs larger than L2)',
                       'setup_seconds_once': t_setup,
                       'comm': dict(comm.stats(), backend='NCCL via cnld_b200_comm_* (C ABI); setup broadcast once, '
                                                          'patch responses gathered to rank 0 every e2e step')},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'gpu_launches': int(launches), 'wall_seconds_timed_region': wall,
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                         'frac': (ach / fp64_peak) if ach else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of the first trailing update from
                         # profiles/r1_ncu_full_summary.txt.  General: 13776 x 13776 x 512, algorithmic bytes
                         # C read + write 6.07e9 + panels 0.45e9.  Symmetric: the same update restricted to the
                         # tiles on or below the diagonal, algorithmic C read + write 3.05e9 + panels 0.45e9.
                         'traffic': (measured_traffic('zgemm3m_general_T2' if args.general else 'zgemm3m_lower_T2')
                                     if args.config == 'c3' else None),
                         'traffic_source': 'profiles/r2_traffic.json (ncu --set full of this round, first trailing update)',
                         'cublas_yardstick': yard,
                         'kernel': 'zgemm3m_kernel (DMMA.8x8x4 rank-512 trailing updates of the LU'
                                   + ('' if args.general else ', tiles on or below the diagonal only')
                                   + '; 3M complex product = 6 m n k executed real flops per launch, i.e. 8 m n k '
                                   'conventional complex-GEMM flops at 4/3 of this rate); durations from a dedicated '
                                   'single-stream pass after the timed region',
                         # every ZGEMM of the timed 2-stream steps (factor + solves), executed flops / device time:
                         # how busy the FP64 units are over the whole step, not just inside the dominant kernel
                         'whole_step_dmma_tflops': step_tflops,
                         'whole_step_frac': step_tflops / fp64_peak,
                         'achieved_conventional_8mnk': (ach * 4.0 / 3.0) if ach else None,
                         # stand-alone rate of the same kernel (full grid, nothing co-running), executed flops
                         'standalone_achieved': (alone_8mnk * 0.75) if alone_8mnk else None,
                         'standalone_frac': (alone_8mnk * 0.75 / fp64_peak) if alone_8mnk else None,
                         'launches_timed': int(klaunch),
                         'peak_source': 'FP64 FMA-pipe peak measured live with cnld_b200_bench_dfma '
                                        '(MEASURED_PEAKS.json carries no FP64 figure; DMMA and DFMA share it); '
                                        'cublas_yardstick = torch.matmul float64 / complex128 at the same shape, same run'},
            'stage_seconds_per_frequency': dict(zip(('assemble', 'factor', 'solve'), (stage / (args.steps * nfb)).tolist())),
        }

    # ---- CPU baseline beside it (rank 0, N=1 only, bounded sample) + parity at THIS config -----
    parity_fail = False
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pyoracle as po
        po.build()
        arr = po.matrix_array(nelem=nelem)
        am = po.array_mesh(arr, refn)
        blocks = po.array_fem(arr, refn)
        f = workload_freqs(args.config, 1)
        tm = {}
        t0 = time.perf_counter()
        Xo, XPo = po.solve_frequency(am, blocks, float(f[0]), c=1500., rho=1000., blocked=True, timings=tm)
        tc = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': 1.0 / tc, 'unit': UNIT, 'cores': po.lib().orc_num_threads(), 'kind': 'port',
                                'sample': 'one full frequency of the same workload (oracle port of the H2Lib path)',
                                'stage_seconds': tm}
        # the same frequency through the product path (host buffers in/out), compared with the oracle's
        # result (generate_db.py:82-99: x = conj(lusolve), x[ob] = 0, x_patch = AVG^T x)
        sweep.ctx.set_streams(args.streams)
        sweep.ctx.set_profile(False)
        xp1, xn1, _ = sweep.run(f, want_nodes=True)
        rn = float(np.linalg.norm(xn1[0].T - Xo) / np.linalg.norm(Xo))
        rp = float(np.linalg.norm(xp1[0].T - XPo) / np.linalg.norm(XPo))
        line['parity'] = {'config': args.config, 'frequency_hz': float(f[0]), 'rel_l2_node': rn, 'rel_l2_patch': rp,
                          'tolerance': 1e-6, 'against': 'oracle port (full N x N assembly + blocked unpivoted LU on the host)'}
        parity_fail = not (rn < 1e-6 and rp < 1e-6)
    if rank == 0:
        print(json.dumps(line), flush=True)
    sweep.close()
    comm.close()
    if parity_fail:
        raise SystemExit('bench.py: GPU result differs from the oracle by more than 1e-6 at this config')


if __name__ == '__main__':
    main()
