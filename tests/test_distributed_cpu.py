"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sweep's shard / broadcast / gather
path with a CPU stand-in for the per-rank GPU solve (the oracle on a tiny mesh)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _solver(setup, freqs):
    # oracle-backed stand-in for FrequencySweep.run on this rank's shard
    import pyoracle as po
    am = po.Mesh(setup.vertices, np.zeros((1, 2), np.uint32), setup.triangles, np.zeros_like(setup.triangles))
    am.on_boundary = setup.on_boundary.astype(bool)
    out = []
    K, M, B, F, A = (m.toarray() for m in (setup.K, setup.M, setup.B, setup.F, setup.AVG))
    for f in freqs:
        w = 2 * np.pi * f
        G = (K - w**2 * M - 1j * w * B).astype(np.complex128)
        if f != 0:
            G = G + (-w**2 * 2 * setup.rho) * po.assemble_z(am, w / setup.c, setup.q_reg, setup.q_sing)
        X = np.conj(np.linalg.solve(G, F.astype(np.complex128)))
        X[am.on_boundary] = 0
        out.append((A.T @ X).T)          # [src, dst]
    return np.array(out).reshape(len(freqs), F.shape[1], F.shape[1])


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.join(here, '..', 'cnl-dyna_b200'), os.path.join(here, '..', 'oracle')):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from cnld import arrays
    from cnld.sweep import SweepSetup, run_distributed, shard
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    freqs = np.array([0.5e6, 1e6, 3e6, 7e6, 11e6])
    setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(1, 1)), 3) if rank == 0 else None
    xp, _ = run_distributed(setup, freqs, rank, world, solver=_solver)
    xc, _ = run_distributed(setup, freqs, rank, world, solver=_solver, distribution='cyclic')   # H-path default
    if rank == 0:
        ref = _solver(setup, freqs)
        err = max(np.abs(xp - ref).max(), np.abs(xc - ref).max()) / np.abs(ref).max()
        q.put((err, xp.shape, [shard(5, r, world) for r in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_is_a_block_partition():
    from cnld.sweep import shard
    for n in (0, 1, 5, 16, 2000):
        for w in (1, 2, 3, 8):
            parts = [shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in parts]
            assert max(sizes) - min(sizes) <= 1


def test_cyclic_shard_is_a_partition():
    from cnld.sweep import shard_cyclic
    for n in (0, 1, 5, 16, 501):
        for w in (1, 2, 3, 8):
            parts = [shard_cyclic(n, r, w) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_sweep_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, shape, parts = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shape == (5, 9, 9)
    assert parts == [(0, 3), (3, 5)]
    assert err < 1e-12       # same inputs after broadcast; the oracle's OpenMP sum order may differ


def test_setup_pack_roundtrip():
    from cnld import arrays
    from cnld.sweep import SweepSetup
    s = SweepSetup.from_abstract(arrays.matrix_array(nelem=(2, 1)), 3)
    t = SweepSetup.unpack({k: v.copy() for k, v in s.pack().items()})
    assert np.array_equal(s.vertices, t.vertices) and np.array_equal(s.triangles, t.triangles)
    for n in ('K', 'M', 'B', 'F', 'AVG'):
        assert (getattr(s, n) != getattr(t, n)).nnz == 0
    assert (s.c, s.rho, s.q_reg, s.q_sing) == (t.c, t.rho, t.q_reg, t.q_sing)
