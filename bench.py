#!/usr/bin/env python
"""
bench.py -- frequencies solved per second on the 4x4 CMUT array (BASELINE.json metric).

One "step" = one batch of frequencies (default 6 per GPU, 3 in flight) pushed through the whole hot path:
init G = K - w^2 M - j w B, accumulate -2 rho w^2 Z(k) (Galerkin Helmholtz SLP assembly), unpivoted
complex-FP64 LU, solve all P source patches, conj / boundary mask / patch average
(cnld/scripts/generate_db.py:30-130).  Workload = BASELINE.json configs[2] ("4x4 CMUT array
(~15k DOFs)", the configuration the metric is quoted on; it fits one B200): 16 square
membranes, refn = 21 -> N = 14 800 nodes, T = 28 224 triangles, P = 144 source patches,
frequencies taken from the 2000-point grid in [0.025, 50] MHz.

  python bench.py [--gpus N] [--steps K] [--warmup W]         this repo's CUDA path
  python bench.py --impl reference ...                        CPU arm: the oracle port of the
                                                              reference's H2Lib path on host cores
Multi-GPU (torchrun, one rank per GPU): frequencies are sharded, no collective inside a
frequency ("weak" scaling: per-GPU batch fixed); setup broadcast once, results gathered.
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

# dram bytes (read + write) of the first trailing update, ncu --set full: {general: ..., symmetric: ...}
TRAFFIC_T2 = {True: 8.28e9, False: 8.06e9}
METRIC = 'frequencies solved/s (Z assemble+factor+solve), 4x4 CMUT array'
UNIT = 'frequencies/s'
CONFIGS = {  # name: (nelem, refn, nfreq of the full sweep, fmin, fmax)
    'c1': ((1, 1), 15, 500, 1e6, 50e6),
    'c2': ((2, 2), 15, 1000, 0.05e6, 50e6),
    'c3': ((4, 4), 21, 2000, 0.025e6, 50e6),
}


def workload_freqs(cfg, count, offset=0):
    """`count` frequencies spread over the config's sweep grid (deterministic)."""
    _, _, nf, f0, f1 = CONFIGS[cfg]
    grid = np.linspace(f0, f1, nf)
    idx = (np.arange(count) * 37 + offset * 101 + 11) % nf
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS))
    ap.add_argument('--freqs-per-step', type=int, default=6)
    ap.add_argument('--streams', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--general', action='store_true', help='general unpivoted LU of G instead of the symmetric-part path')
    ap.add_argument('--lu-outer', type=int, default=0, help='debug: outer LU block (multiple of 64)')
    ap.add_argument('--lu-reserve', type=int, default=-1, help='debug: SMs left free by the look-ahead trailing update')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if args.steps > 2:
            args.steps, args.warmup = min(args.steps, 2), min(args.warmup, 1)
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from cnld import _lib, arrays
    from cnld.sweep import FrequencySweep, SweepSetup, broadcast_setup

    if _lib.device_count() == 0:
        raise SystemExit('bench.py: no CUDA device -- libcnld_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    nelem, refn = CONFIGS[args.config][:2]
    t_setup = time.perf_counter()
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=nelem), refn) if rank == 0 else None
    if world > 1:
        setup = broadcast_setup(setup, rank, torch.device('cuda', local))
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
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def batch(step):
        return workload_freqs(args.config, nfb, offset=step * world + rank)

    # ---- device-resident timing (`value`) ---------------------------------------------------
    for w in range(args.warmup):
        sweep.run_device(batch(w))
    fp64_peak = _lib.bench_dfma(3, device=local)         # measured FP64 FMA-pipe peak, TFLOP/s
    # the same kernel alone on the GPU at the first trailing update's shape (conventional 8mnk rate)
    alone_8mnk = _lib.bench_zgemm(N - 1024, N - 1024, 512, 3, device=local) if args.config == 'c3' else None
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
    tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    value = args.steps * nfb * world / dt

    # ---- end to end through the public API (host buffers in, host buffers out) -------------
    out_nodes = np.empty((nfb, P, N), dtype=np.complex128)
    sweep.upload()
    sweep.run(batch(0), want_nodes=True, out_nodes=out_nodes)     # untimed: pinned staging is allocated here
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for s in range(args.steps):
        h2d = sweep.upload() + 8 * nfb
        xp, xn, _ = sweep.run(batch(args.warmup + s), want_nodes=True, out_nodes=out_nodes)
        d2h = xp.nbytes + xn.nbytes
    barrier()
    dte = time.perf_counter() - t0
    tt = torch.tensor([dte], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dte = float(tt.item())
    e2e = args.steps * nfb * world / dte
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
                       'setup_seconds_once': t_setup},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'gpu_launches': int(launches), 'wall_seconds_timed_region': wall,
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                         'frac': (ach / fp64_peak) if ach else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of the first trailing update from
                         # profiles/r1_ncu_full_summary.txt.  General: 13776 x 13776 x 512, algorithmic bytes
                         # C read + write 6.07e9 + panels 0.45e9.  Symmetric: the same update restricted to the
                         # tiles on or below the diagonal, algorithmic C read + write 3.05e9 + panels 0.45e9.
                         'traffic': (TRAFFIC_T2[bool(args.general)] if args.config == 'c3' else None),
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
                                        '(MEASURED_PEAKS.json carries no FP64 figure; DMMA and DFMA share it)'},
            'stage_seconds_per_frequency': dict(zip(('assemble', 'factor', 'solve'), (stage / (args.steps * nfb)).tolist())),
        }

    # ---- CPU baseline beside it (rank 0, N=1 only, bounded sample) --------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import pyoracle as po
        po.build()
        arr = po.matrix_array(nelem=nelem)
        am = po.array_mesh(arr, refn)
        blocks = po.array_fem(arr, refn)
        f = workload_freqs(args.config, 1)
        t0 = time.perf_counter()
        cpu_reference_step(po, am, blocks, f, 1500., 1000.)
        tc = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': 1.0 / tc, 'unit': UNIT, 'cores': po.lib().orc_num_threads(), 'kind': 'port',
                                'sample': 'one full frequency of the same workload (oracle port of the H2Lib path)'}
    if rank == 0:
        print(json.dumps(line), flush=True)
    sweep.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
