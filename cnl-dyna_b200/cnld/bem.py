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


def array_z_linop(array, refn, f, c, rho, *args, **kwargs):
    """Linear operators x -> -w^2 2 rho Z x (boundary rows/cols masked) and its LU-based inverse
    (bem.py:40-72)."""
    from scipy.sparse.linalg import LinearOperator
    k, omg = 2 * np.pi * f / c, 2 * np.pi * f
    Z = array_z_matrix(array, refn, k, *args, **kwargs)
    Z_LU = Z.lu()
    amesh = mesh.Mesh.from_abstract(array, refn)
    ob, n = amesh.on_boundary, len(amesh.vertices)

    def mv(x):
        x = np.array(x, dtype=np.complex128).ravel()
        x[ob] = 0
        p = Z * x
        p[ob] = 0
        return -omg**2 * rho * 2 * p

    def inv_mv(x):
        x = np.array(x, dtype=np.complex128).ravel()
        x[ob] = 0
        p = Z_LU._triangularsolve(x)
        p[ob] = 0
        return -omg**2 * rho * 2 * p

    return (LinearOperator((n, n), dtype=np.complex128, matvec=mv),
            LinearOperator((n, n), dtype=np.complex128, matvec=inv_mv))
