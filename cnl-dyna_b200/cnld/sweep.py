"""
cnld.sweep -- frequency-sweep driver on top of libcnld_b200 (one process per GPU).

What generate_db.process does for ONE frequency in ONE forked process
(/root/reference/cnld/scripts/generate_db.py:30-130) -- re-mesh, re-assemble FEM, build Z, form
G, LU, loop over source patches -- is split here into a frequency-independent part done once
(mesh, quadrature tables, K/M/B, F, AVG, boundary mask -> GPU) and a per-frequency part that
runs entirely on the device for a whole shard of frequencies.

Multi-GPU: frequencies are independent jobs (generate_db.py:229-242), so they are
block-distributed over ranks with no collective inside a frequency; rank 0's setup is
broadcast once and the patch responses are gathered at the end (`run_distributed`).
"""
import numpy as np

from . import _lib, fem, mesh


class SweepSetup:
    """Frequency-independent host data of a sweep (plain arrays: cheap to broadcast)."""

    def __init__(self, vertices, triangles, on_boundary, K, M, B, F, AVG, c, rho, q_reg, q_sing):
        from scipy import sparse as sps
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.uint32)
        self.on_boundary = np.ascontiguousarray(on_boundary, dtype=np.uint8)
        self.K, self.M, self.B = (sps.csr_matrix(a) for a in (K, M, B))
        self.F, self.AVG = sps.csc_matrix(F), sps.csc_matrix(AVG)
        for m in (self.K, self.M, self.B, self.F, self.AVG):
            m.eliminate_zeros()       # block_diag of dense membrane blocks stores every zero: 16 x 925^2 entries at C3
            m.sort_indices()
        self.c, self.rho, self.q_reg, self.q_sing = float(c), float(rho), int(q_reg), int(q_sing)

    @classmethod
    def from_abstract(cls, array, refn, c=1500., rho=1000., q_reg=2, q_sing=4):
        am = mesh.Mesh.from_abstract(array, refn)
        K, M, B = fem.array_kmb_spmatrices(array, refn)
        return cls(np.array(am.vertices), np.array(am.triangles), np.array(am.on_boundary), K, M, B,
                   fem.array_f_spmatrix(array, refn), fem.array_avg_spmatrix(array, refn), c, rho,
                   q_reg, q_sing)

    @property
    def nnodes(self):
        return len(self.vertices)

    @property
    def npatch(self):
        return self.F.shape[1]

    # flat arrays <-> object, for torch.distributed broadcast
    def pack(self):
        out = {'vertices': self.vertices, 'triangles': self.triangles, 'on_boundary': self.on_boundary,
               'scalars': np.array([self.c, self.rho, self.q_reg, self.q_sing, self.nnodes, self.npatch],
                                   dtype=np.float64)}
        for name in ('K', 'M', 'B', 'F', 'AVG'):
            m = getattr(self, name)
            out[name + '_indptr'] = m.indptr.astype(np.int64)
            out[name + '_indices'] = m.indices.astype(np.int64)
            out[name + '_data'] = m.data.astype(np.float64)
        return out

    @classmethod
    def unpack(cls, d):
        from scipy import sparse as sps
        c, rho, q_reg, q_sing, n, P = d['scalars']
        n, P = int(n), int(P)
        mats = {}
        for name in ('K', 'M', 'B'):
            mats[name] = sps.csr_matrix((d[name + '_data'], d[name + '_indices'], d[name + '_indptr']), shape=(n, n))
        for name in ('F', 'AVG'):
            mats[name] = sps.csc_matrix((d[name + '_data'], d[name + '_indices'], d[name + '_indptr']), shape=(n, P))
        return cls(d['vertices'], d['triangles'], d['on_boundary'], mats['K'], mats['M'], mats['B'], mats['F'],
                   mats['AVG'], c, rho, int(q_reg), int(q_sing))


class FrequencySweep:
    """Device-resident sweep on one GPU.

    symmetric=True (default) eliminates the symmetric part of G at half the flops and recovers the
    reference's (non-symmetric) solution by `refine_steps` refinement steps with the sparse asymmetry;
    every frequency's convergence is checked on the device and a frequency that misses 1e-11 is redone
    with the general elimination.  symmetric=False is the reference's elimination order
    (compressed_formats.py:247-252, lrdecomp_amatrix)."""

    def __init__(self, setup, device=0, nstreams=3, symmetric=True, refine_steps=1):
        self.setup = setup
        self.ctx = _lib.Context(setup.vertices, setup.triangles, q_reg=setup.q_reg, q_sing=setup.q_sing,
                                device=device)
        self.ctx.set_streams(nstreams)
        self.ctx.set_symmetric(1 if symmetric else 0, refine_steps)
        s = self.setup
        self.ctx.set_mbk(s.K, s.M, s.B)
        self.ctx.set_loads(s.F, s.AVG, s.on_boundary)

    def upload(self):
        """Host -> device copy of the FEM operators, loads and boundary mask."""
        self.ctx.upload_mbk()
        self.ctx.upload_loads()
        return self.upload_bytes

    @property
    def upload_bytes(self):
        return int(sum(a.nbytes for a in self.ctx._mbk) + sum(a.nbytes for a in self.ctx._loads))

    def run(self, freqs, want_nodes=True, out_nodes=None):
        """-> x_patch[nf, P(src), P(dst)], x_node[nf, P(src), N] | None, seconds[nf]."""
        return self.ctx.sweep(freqs, self.setup.c, self.setup.rho, want_nodes=want_nodes, out_nodes=out_nodes)

    def run_device(self, freqs):
        self.ctx.sweep_device(freqs, self.setup.c, self.setup.rho)

    def close(self):
        self.ctx.close()


def shard(n, rank, world):
    """Contiguous block distribution of n jobs: [lo, hi) of `rank` (sizes differ by at most 1)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class TorchComm:
    """The communicator interface of `_lib.Comm` (bcast / bcast_object / gather / max / barrier) on top of an
    initialised torch.distributed process group -- what the CPU tests use (gloo, world_size 2); on GPUs the
    sweep uses `_lib.Comm` (NCCL through the C ABI, no torch)."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist, self.device = dist, device if device is not None else 'cpu'
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def bcast(self, arr, root=0):
        import torch
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1)).to(self.device)
        self.dist.broadcast(t, src=root)
        if self.rank != root:
            arr.view(np.uint8).reshape(-1)[:] = t.cpu().numpy()
        return arr

    def bcast_object(self, obj, root=0):
        box = [obj if self.rank == root else None]
        self.dist.broadcast_object_list(box, src=root)
        return box[0]

    def gather(self, arr, root=0):
        import torch
        mine = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).to(self.device)
        parts = [torch.empty_like(mine) for _ in range(self.world)] if self.rank == root else None
        self.dist.gather(mine, parts, dst=root)
        if self.rank != root:
            return None
        return np.stack([p.cpu().numpy().view(arr.dtype).reshape(arr.shape) for p in parts])

    def max(self, *vals):
        import torch
        t = torch.tensor(vals, dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist() if len(vals) != 1 else float(t[0])

    def barrier(self):
        self.dist.barrier()


def _as_comm(comm, device=None):
    return comm if comm is not None else TorchComm(device)


def broadcast_setup(setup, rank, device=None, comm=None):
    """Rank 0's SweepSetup to every rank: one collective per array, once per sweep.  `comm` is a
    `_lib.Comm` (NCCL over NVLink through the C ABI) or None for the torch.distributed adapter (gloo on
    CPU, nccl on GPUs) bound to `device`."""
    comm = _as_comm(comm, device)
    names = ['vertices', 'triangles', 'on_boundary', 'scalars'] + [
        f'{m}_{p}' for m in ('K', 'M', 'B', 'F', 'AVG') for p in ('indptr', 'indices', 'data')]
    packed = setup.pack() if rank == 0 else None
    meta = comm.bcast_object([(k, packed[k].shape, str(packed[k].dtype)) for k in names] if rank == 0 else None)
    out = {}
    for k, shape, dtype in meta:
        a = np.ascontiguousarray(packed[k]) if rank == 0 else np.empty(shape, dtype=np.dtype(dtype))
        out[k] = comm.bcast(a)
    return SweepSetup.unpack(out)


def shard_cyclic(n, rank, world):
    """Cyclic distribution of n jobs: indices rank, rank + world, ...  On the H-matrix path the cost
    of a frequency grows with k (ACA ranks, GMRES iterations), so contiguous blocks of the sorted
    frequency list would leave the rank with the highest block working long after the others."""
    return np.arange(rank, n, world)


def run_distributed(setup_or_none, freqs, rank, world, device=0, torch_device=None, nstreams=3,
                    solver=None, symmetric=True, method='dense', hmatrix=None, distribution=None, comm=None):
    """Distributed sweep: returns (x_patch[nf, P, P] on rank 0 else None, local seconds).
    method 'dense' (LU, default) or 'hmatrix' (ACA H-matrix + GMRES, `hmatrix` = keyword arguments
    of _lib.HMatrix / HMatrix.sweep).  distribution 'block' (default for dense) or 'cyclic' (default
    for hmatrix).  No collective inside a frequency: one broadcast of the setup, one gather of the
    patch responses.  `solver(setup, freqs) -> x_patch[nf_local, P, P]` replaces the GPU sweep; the
    tests inject a CPU stand-in to exercise the sharding / broadcast / gather logic under gloo.
    `comm`: `_lib.Comm` (NCCL through the C ABI) or None = torch.distributed adapter on `torch_device`."""
    freqs = np.asarray(freqs, dtype=np.float64)
    if world > 1:
        comm = _as_comm(comm, torch_device)
    setup = broadcast_setup(setup_or_none, rank, comm=comm) if world > 1 else setup_or_none
    distribution = distribution or ('cyclic' if method == 'hmatrix' else 'block')
    if distribution == 'cyclic':
        owned = [shard_cyclic(len(freqs), r, world) for r in range(world)]
    else:
        owned = [np.arange(*shard(len(freqs), r, world)) for r in range(world)]
    mine_idx = owned[rank]
    if solver is not None:
        xp = solver(setup, freqs[mine_idx])
        t = np.zeros(len(mine_idx))
    elif method == 'hmatrix':
        kw = dict(hmatrix or {})
        ctx = _lib.Context(setup.vertices, setup.triangles, q_reg=setup.q_reg, q_sing=setup.q_sing, device=device)
        ctx.set_mbk(setup.K, setup.M, setup.B)
        ctx.set_loads(setup.F, setup.AVG, setup.on_boundary)
        H = _lib.HMatrix(ctx, **{k: kw.pop(k) for k in ('clf', 'eta', 'admis', 'strict', 'kmax') if k in kw})
        xp, _, _, t = H.sweep(freqs[mine_idx], c=setup.c, rho=setup.rho, want_nodes=False, **kw)
        H.close()
        ctx.close()
    else:
        sw = FrequencySweep(setup, device=device, nstreams=nstreams, symmetric=symmetric)
        xp, _, t = sw.run(freqs[mine_idx], want_nodes=False)
        sw.close()
    if world == 1:
        out = np.empty_like(xp)
        out[mine_idx] = xp
        return out, t
    P = setup.npatch
    nmax = max(len(o) for o in owned)
    buf = np.zeros((nmax, P, P), dtype=np.complex128)
    buf[:len(mine_idx)] = xp
    gathered = comm.gather(buf)            # the one exchange of the sweep: patch responses to rank 0
    if rank != 0:
        return None, t
    out = np.empty((len(freqs), P, P), dtype=np.complex128)
    for g, o in zip(gathered, owned):
        out[o] = g[:len(o)]
    return out, t
