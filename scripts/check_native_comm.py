"""Multi-GPU check of the product driver WITHOUT torch.distributed (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/check_native_comm.py
cnld_b200_comm_* (NCCL via the C ABI): broadcast / gather / max; sweep.run_distributed with the dense LU and the
ACA H-matrix + GMRES methods on N ranks against the same sweep on one rank."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'cnl-dyna_b200'))
import numpy as np
from cnld import _lib, arrays
from cnld.sweep import FrequencySweep, SweepSetup, run_distributed

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
os.environ['CNLD_B200_DEVICE'] = str(local)
comm = _lib.Comm(rank, world, local)
a = np.arange(1000, dtype=np.float64) * (1 if rank == 0 else -1)
comm.bcast(a)
assert np.array_equal(a, np.arange(1000.0))
g = comm.gather(np.full((3, 2), rank, dtype=np.complex128))
if rank == 0:
    assert g.shape == (world, 3, 2) and all(np.all(g[r] == r) for r in range(world))
assert comm.max(float(rank)) == world - 1
assert comm.bcast_object({'n': 5} if rank == 0 else None) == {'n': 5}
setup = SweepSetup.from_abstract(arrays.matrix_array(nelem=(2, 2)), 5) if rank == 0 else None
freqs = np.linspace(1e6, 30e6, 7)
res = {}
for method, kw in (('dense', {}), ('hmatrix', dict(hmatrix=dict(eps_aca=1e-6, tol=1e-9, batch=36)))):
    xp, _ = run_distributed(setup, freqs, rank, world, device=local, comm=comm, method=method, **kw)
    if rank == 0:
        ref, _ = run_distributed(setup, freqs, 0, 1, device=local, method=method, **kw)
        res[method] = float(np.linalg.norm(xp - ref) / np.linalg.norm(ref))
        assert res[method] < 1e-9, (method, res[method])
if rank == 0:
    print(json.dumps(dict(world=world, rel_err_vs_single_rank=res, comm=comm.stats())))
comm.close()
