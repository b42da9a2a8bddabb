"""cnld.bem -- impedance matrices of membranes / arrays (mirror of /root/reference/cnld/bem.py)."""
import numpy as np
from scipy import sparse as sps

from . import abstract, mesh, util
from .compressed_formats import ZFullMatrix, ZHMatrix


@util.memoize
def mem_z_matrix(mem, refn, k, *args, **kwargs):
    """Impedance matrix in FullFormat for one membrane (bem.py:16-25)."""
    if isinstance(mem, abstract.SquareCmutMembrane):
        amesh = mesh.square(mem.length_x, mem.length_y, refn)
    else:
        amesh = mesh.circle(mem.radius, refn)
    return np.array(ZFullMatrix(amesh, k, *args, **kwargs).data)


def array_z_matrix(array, refn, k, format='HFormat', *args, **kwargs):
    """Impedance matrix in FullFormat or HFormat for an array (bem.py:28-37)."""
    amesh = mesh.Mesh.from_abstract(array, refn)
    if format.lower() in ['hformat', 'h']:
        return ZHMatrix(amesh, k, *args, **kwargs)
    return ZFullMatrix(amesh, k, *args, **kwargs)


class _MaskedFluidOperator:
    """x -> -w^2 2 rho M A M x with M the projector that zeroes the clamped boundary nodes: the fluid
    loading operator of the Krylov formulation (bem.py:40-72, scratch/freq_resp_db_krylov.py:38-79).
    `apply` is the H-/dense matvec on the GPU or the LU solve with the factored impedance matrix."""

    def __init__(self, apply, free, scale):
        self.apply, self.free, self.scale = apply, free, scale

    def __call__(self, x):
        v = np.where(self.free, np.asarray(x, dtype=np.complex128).ravel(), 0.0)
        return self.scale * np.where(self.free, self.apply(v), 0.0)


def array_z_linop(array, refn, f, c, rho, *args, **kwargs):
    """(Gbe, Gbe_inv) as scipy LinearOperators, like bem.array_z_linop (bem.py:40-72)."""
    from scipy.sparse.linalg import LinearOperator
    omega = 2 * np.pi * f
    Z = array_z_matrix(array, refn, omega / c, *args, **kwargs)
    Zlu = Z.lu()
    amesh = mesh.Mesh.from_abstract(array, refn)
    free = ~np.asarray(amesh.on_boundary, dtype=bool)
    n, scale = len(free), -omega**2 * rho * 2
    ops = [_MaskedFluidOperator(fn, free, scale) for fn in (Z._matvec, Zlu._triangularsolve)]
    return tuple(LinearOperator((n, n), dtype=np.complex128, matvec=op) for op in ops)
