"""
cnld.h2lib -- host-side mirror of cnl-dyna's Cython package `cnld.h2lib`
(/root/reference/cnld/h2lib/*.pyx), bound with ctypes to the H2Lib-named
symbols of libcnld_b200.so (include/cnld_b200_h2lib.h) instead of libh2.a.

Same class / function names, argument meaning and error behaviour as the
reference wrappers, including their quirks:
  * handles carry `ptr` + `owner`; owners call the matching `del_*` on GC
    (amatrix.pyx:30-32 etc.);
  * `AMatrix.a` is a C-order [rows, cols] view of the column-major buffer, i.e.
    numpy sees the transpose (amatrix.pyx:34-37) -- harmless for the symmetric
    matrices of this path, and kept so existing callers see identical arrays.
"""
import atexit
import ctypes as C
import enum

import numpy as np

from . import _lib

real = C.c_double
uint = C.c_uint


class field(C.Structure):
    _fields_ = [('re', real), ('im', real)]


def _f(z):
    z = complex(z)
    return field(z.real, z.imag)


class _avector(C.Structure):
    _fields_ = [('v', C.POINTER(field)), ('dim', uint), ('owner', C.c_void_p)]


class _amatrix(C.Structure):
    _fields_ = [('a', C.POINTER(field)), ('ld', uint), ('rows', uint), ('cols', uint),
                ('owner', C.c_void_p)]


class _sparsematrix(C.Structure):
    _fields_ = [('rows', uint), ('cols', uint), ('nz', uint), ('row', C.POINTER(uint)),
                ('col', C.POINTER(uint)), ('coeff', C.POINTER(field))]


class _surface3d(C.Structure):
    _fields_ = [('vertices', uint), ('edges', uint), ('triangles', uint),
                ('x', C.POINTER(real)), ('e', C.POINTER(uint)), ('t', C.POINTER(uint)),
                ('s', C.POINTER(uint)), ('n', C.POINTER(real)), ('g', C.POINTER(real)),
                ('hmin', real), ('hmax', real)]


_PHI = C.CFUNCTYPE(None, uint, real, real, C.c_void_p, C.POINTER(real))


class _macrosurface3d(C.Structure):
    _fields_ = [('vertices', uint), ('edges', uint), ('triangles', uint),
                ('x', C.POINTER(real)), ('e', C.POINTER(uint)), ('t', C.POINTER(uint)),
                ('s', C.POINTER(uint)), ('phi', C.c_void_p), ('phidata', C.c_void_p)]


class _bem3d(C.Structure):
    _fields_ = [('gr', C.POINTER(_surface3d)), ('sq', C.c_void_p), ('row_basis', C.c_int),
                ('col_basis', C.c_int), ('mass', C.POINTER(real)), ('alpha', field), ('k', field),
                ('kernel_const', field), ('priv', C.c_void_p), ('q_regular', uint),
                ('q_singular', uint), ('accur', real)]


class _cluster(C.Structure):
    pass


_cluster._fields_ = [('size', uint), ('idx', C.POINTER(uint)), ('sons', uint), ('son', C.POINTER(C.POINTER(_cluster))),
                     ('dim', uint), ('bmin', C.POINTER(real)), ('bmax', C.POINTER(real)), ('desc', uint), ('type', uint)]


class _block(C.Structure):
    pass


_block._fields_ = [('rc', C.POINTER(_cluster)), ('cc', C.POINTER(_cluster)), ('a', C.c_bool),
                   ('son', C.POINTER(C.POINTER(_block))), ('rsons', uint), ('csons', uint), ('desc', uint)]


class _rkmatrix(C.Structure):
    _fields_ = [('A', _amatrix), ('B', _amatrix), ('k', uint)]


class _hmatrix(C.Structure):
    pass


_hmatrix._fields_ = [('rc', C.POINTER(_cluster)), ('cc', C.POINTER(_cluster)), ('r', C.POINTER(_rkmatrix)),
                     ('f', C.POINTER(_amatrix)), ('son', C.POINTER(C.POINTER(_hmatrix))), ('rsons', uint),
                     ('csons', uint), ('refs', uint), ('desc', uint)]


class _truncmode(C.Structure):
    _fields_ = [('frobenius', C.c_bool), ('absolute', C.c_bool), ('blocks', C.c_bool), ('zeta_level', real),
                ('zeta_age', real)]


_ADMIS = C.CFUNCTYPE(C.c_bool, C.POINTER(_cluster), C.POINTER(_cluster), C.c_void_p)


class basisfunctionbem3d(enum.IntEnum):
    NONE = 0
    CONSTANT = ord('c')
    LINEAR = ord('l')


_L = None


def _lib_h2():
    global _L
    if _L is not None:
        return _L
    L = _lib.lib()
    P = C.POINTER
    sig = {
        'new_avector': (P(_avector), [uint]), 'new_zero_avector': (P(_avector), [uint]),
        'del_avector': (None, [P(_avector)]), 'clear_avector': (None, [P(_avector)]),
        'copy_avector': (None, [P(_avector), P(_avector)]), 'random_avector': (None, [P(_avector)]),
        'new_amatrix': (P(_amatrix), [uint, uint]), 'new_zero_amatrix': (P(_amatrix), [uint, uint]),
        'del_amatrix': (None, [P(_amatrix)]), 'clone_amatrix': (P(_amatrix), [P(_amatrix)]),
        'scale_amatrix': (None, [field, P(_amatrix)]), 'conjugate_amatrix': (None, [P(_amatrix)]),
        'addeval_amatrix_avector': (None, [field, P(_amatrix), P(_avector), P(_avector)]),
        'add_amatrix': (None, [field, C.c_bool, P(_amatrix), P(_amatrix)]),
        'getsize_amatrix': (C.c_size_t, [P(_amatrix)]),
        'addmul_amatrix': (None, [field, C.c_bool, P(_amatrix), C.c_bool, P(_amatrix), P(_amatrix)]),
        'new_raw_sparsematrix': (P(_sparsematrix), [uint, uint, uint]),
        'new_identity_sparsematrix': (P(_sparsematrix), [uint, uint]),
        'del_sparsematrix': (None, [P(_sparsematrix)]),
        'getsize_sparsematrix': (C.c_size_t, [P(_sparsematrix)]),
        'sort_sparsematrix': (None, [P(_sparsematrix)]), 'clear_sparsematrix': (None, [P(_sparsematrix)]),
        'add_sparsematrix_amatrix': (None, [field, C.c_bool, P(_sparsematrix), P(_amatrix)]),
        'addeval_sparsematrix_avector': (None, [field, P(_sparsematrix), P(_avector), P(_avector)]),
        'new_surface3d': (P(_surface3d), [uint, uint, uint]), 'del_surface3d': (None, [P(_surface3d)]),
        'prepare_surface3d': (None, [P(_surface3d)]),
        'merge_surface3d': (P(_surface3d), [P(_surface3d), P(_surface3d)]),
        'refine_red_surface3d': (P(_surface3d), [P(_surface3d)]),
        'translate_surface3d': (None, [P(_surface3d), P(real)]),
        'check_surface3d': (uint, [P(_surface3d)]), 'isclosed_surface3d': (C.c_bool, [P(_surface3d)]),
        'isoriented_surface3d': (C.c_bool, [P(_surface3d)]),
        'new_macrosurface3d': (P(_macrosurface3d), [uint, uint, uint]),
        'del_macrosurface3d': (None, [P(_macrosurface3d)]),
        'build_from_macrosurface3d_surface3d': (P(_surface3d), [P(_macrosurface3d), uint]),
        'new_sphere_macrosurface3d': (P(_macrosurface3d), []),
        'new_bem3d': (P(_bem3d), [P(_surface3d), C.c_int, C.c_int]), 'del_bem3d': (None, [P(_bem3d)]),
        'new_slp_helmholtz_bem3d': (P(_bem3d), [field, P(_surface3d), uint, uint, C.c_int, C.c_int]),
        'assemble_bem3d_amatrix': (None, [P(_bem3d), P(_amatrix)]),
        'lrdecomp_amatrix': (uint, [P(_amatrix)]),
        'lrsolve_amatrix_avector': (None, [C.c_bool, P(_amatrix), P(_avector)]),
        'triangularsolve_amatrix_avector': (None, [C.c_bool, C.c_bool, C.c_bool, P(_amatrix), P(_avector)]),
        'init_h2lib': (None, [C.c_void_p, C.c_void_p]), 'uninit_h2lib': (None, []),
        'freemem': (None, [C.c_void_p]),
        'new_cluster': (P(_cluster), [uint, P(uint), uint, uint]), 'del_cluster': (None, [P(_cluster)]),
        'build_bem3d_cluster': (P(_cluster), [P(_bem3d), uint, C.c_int]),
        'new_block': (P(_block), [P(_cluster), P(_cluster), C.c_bool, uint, uint]), 'del_block': (None, [P(_block)]),
        'build_strict_block': (P(_block), [P(_cluster), P(_cluster), C.c_void_p, C.c_void_p]),
        'build_nonstrict_block': (P(_block), [P(_cluster), P(_cluster), C.c_void_p, C.c_void_p]),
        'new_rkmatrix': (P(_rkmatrix), [uint, uint, uint]), 'del_rkmatrix': (None, [P(_rkmatrix)]),
        'new_hmatrix': (P(_hmatrix), [P(_cluster), P(_cluster)]), 'del_hmatrix': (None, [P(_hmatrix)]),
        'clear_hmatrix': (None, [P(_hmatrix)]), 'copy_hmatrix': (None, [P(_hmatrix), P(_hmatrix)]),
        'clone_hmatrix': (P(_hmatrix), [P(_hmatrix)]), 'clonestructure_hmatrix': (P(_hmatrix), [P(_hmatrix)]),
        'getsize_hmatrix': (C.c_size_t, [P(_hmatrix)]), 'getrows_hmatrix': (uint, [P(_hmatrix)]),
        'getcols_hmatrix': (uint, [P(_hmatrix)]), 'build_from_block_hmatrix': (P(_hmatrix), [P(_block), uint]),
        'addeval_hmatrix_avector': (None, [field, P(_hmatrix), P(_avector), P(_avector)]),
        'addevalsymm_hmatrix_avector': (None, [field, P(_hmatrix), P(_avector), P(_avector)]),
        'copy_sparsematrix_hmatrix': (None, [P(_sparsematrix), P(_hmatrix)]),
        'identity_hmatrix': (None, [P(_hmatrix)]), 'scale_hmatrix': (None, [field, P(_hmatrix)]),
        'add_hmatrix_amatrix': (None, [field, C.c_bool, P(_hmatrix), P(_amatrix)]),
        'setup_hmatrix_aprx_aca_bem3d': (None, [P(_bem3d), P(_cluster), P(_cluster), P(_block), real]),
        'setup_hmatrix_aprx_paca_bem3d': (None, [P(_bem3d), P(_cluster), P(_cluster), P(_block), real]),
        'setup_hmatrix_aprx_hca_bem3d': (None, [P(_bem3d), P(_cluster), P(_cluster), P(_block), uint, real]),
        'setup_hmatrix_aprx_inter_row_bem3d': (None, [P(_bem3d), P(_cluster), P(_cluster), P(_block), uint]),
        'assemble_bem3d_hmatrix': (None, [P(_bem3d), P(_block), P(_hmatrix)]),
        'new_truncmode': (P(_truncmode), []), 'del_truncmode': (None, [P(_truncmode)]),
        'new_releucl_truncmode': (P(_truncmode), []),
        'solve_gmres_amatrix_avector': (uint, [P(_amatrix), P(_avector), P(_avector), real, uint, uint]),
        'solve_gmres_hmatrix_avector': (uint, [P(_hmatrix), P(_avector), P(_avector), real, uint, uint]),
        'solve_cg_amatrix_avector': (uint, [P(_amatrix), P(_avector), P(_avector), real, uint]),
        'solve_cg_hmatrix_avector': (uint, [P(_hmatrix), P(_avector), P(_avector), real, uint]),
        'choldecomp_amatrix': (uint, [P(_amatrix)]), 'cholsolve_amatrix_avector': (None, [P(_amatrix), P(_avector)]),
        'new_dlp_helmholtz_bem3d': (P(_bem3d), [field, P(_surface3d), uint, uint, C.c_int, C.c_int, field]),
        'norm2diff_amatrix_sparsematrix': (real, [P(_sparsematrix), P(_amatrix)]),
        'norm2diff_amatrix_hmatrix': (real, [P(_hmatrix), P(_amatrix)]),
        'norm2diff_sparsematrix_hmatrix': (real, [P(_hmatrix), P(_sparsematrix)]),
        'norm2diff_lr_amatrix': (real, [P(_amatrix), P(_amatrix)]),
        'norm2diff_lr_hmatrix': (real, [P(_hmatrix), P(_hmatrix)]),
        'norm2diff_lr_amatrix_hmatrix': (real, [P(_amatrix), P(_hmatrix)]),
        'norm2diff_lr_hmatrix_amatrix': (real, [P(_hmatrix), P(_amatrix)]),
        'norm2diff_lr_sparsematrix_amatrix': (real, [P(_sparsematrix), P(_amatrix)]),
        'norm2diff_lr_sparsematrix_hmatrix': (real, [P(_sparsematrix), P(_hmatrix)]),
        'lrdecomp_hmatrix': (None, [P(_hmatrix), P(_truncmode), real]),
        'lrsolve_hmatrix_avector': (None, [C.c_bool, P(_hmatrix), P(_avector)]),
        'lreval_hmatrix_avector': (None, [C.c_bool, P(_hmatrix), P(_avector)]),
        'choldecomp_hmatrix': (None, [P(_hmatrix), P(_truncmode), real]),
        'cholsolve_hmatrix_avector': (None, [P(_hmatrix), P(_avector)]),
        'choleval_hmatrix_avector': (None, [P(_hmatrix), P(_avector)]),
        'addmul_hmatrix': (None, [field, C.c_bool, P(_hmatrix), C.c_bool, P(_hmatrix), P(_truncmode), real, P(_hmatrix)]),
        'add_hmatrix': (None, [field, P(_hmatrix), P(_truncmode), real, P(_hmatrix)]),
        'add_amatrix_hmatrix': (None, [field, C.c_bool, P(_amatrix), P(_truncmode), real, P(_hmatrix)]),
        'triangularsolve_hmatrix_avector': (None, [C.c_bool, C.c_bool, C.c_bool, P(_hmatrix), P(_avector)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _L = L
    return L


def _view(p, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    itemsize = np.dtype(dtype).itemsize
    buf = (C.c_char * (n * itemsize)).from_address(C.addressof(p.contents))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class _Handle:
    _del = None

    def __init__(self):
        self.ptr = None
        self.owner = False

    def _setup(self, ptr, owner):
        if not ptr:
            raise _lib.Cnldb200Error(_lib.lib().cnld_b200_last_error().decode())
        self.ptr = ptr
        self.owner = owner

    def __del__(self):
        if getattr(self, 'ptr', None) and self.owner and self._del and _L is not None:
            getattr(_L, self._del)(self.ptr)
            self.ptr = None

    @classmethod
    def wrap(cls, ptr, owner=True):
        obj = cls.__new__(cls)
        _Handle.__init__(obj)
        obj._setup(ptr, owner)
        return obj


class AVector(_Handle):
    _del = 'del_avector'

    def __init__(self, dim):
        super().__init__()
        self._setup(_lib_h2().new_avector(dim), True)

    @classmethod
    def from_array(cls, v):
        v = np.asarray(v).squeeze()
        if v.ndim == 0:
            v = v.reshape(1)
        assert v.ndim == 1
        obj = cls(v.size)
        np.copyto(obj.v, v.astype(np.complex128))
        return obj

    @property
    def dim(self):
        return self.ptr.contents.dim

    @property
    def v(self):
        return _view(self.ptr.contents.v, (self.dim,), np.complex128)


class AMatrix(_Handle):
    _del = 'del_amatrix'

    def __init__(self, rows, cols):
        super().__init__()
        self._setup(_lib_h2().new_zero_amatrix(rows, cols), True)

    @classmethod
    def from_array(cls, a):
        a = np.asarray(a).squeeze()
        assert a.ndim == 2
        obj = cls(a.shape[0], a.shape[1])
        np.copyto(obj.a, a.astype(np.complex128))
        return obj

    @property
    def rows(self):
        return self.ptr.contents.rows

    @property
    def cols(self):
        return self.ptr.contents.cols

    @property
    def a(self):
        # C-order view of the column-major buffer, as in amatrix.pyx:34-37
        return _view(self.ptr.contents.a, (self.rows, self.cols), np.complex128)


class SparseMatrix(_Handle):
    _del = 'del_sparsematrix'

    def __init__(self, rows, cols, nz):
        super().__init__()
        self._setup(_lib_h2().new_raw_sparsematrix(rows, cols, nz), True)

    @classmethod
    def from_array(cls, a):
        from scipy import sparse as sps
        a = a.astype(np.complex128)
        assert a.ndim == 2
        if not isinstance(a, sps.csr_matrix):
            a = sps.csr_matrix(a)
        obj = cls(a.shape[0], a.shape[1], a.nnz)
        obj.row[:] = a.indptr
        obj.col[:] = a.indices
        obj.coeff[:] = a.data
        return obj

    @property
    def rows(self):
        return self.ptr.contents.rows

    @property
    def cols(self):
        return self.ptr.contents.cols

    @property
    def nz(self):
        return self.ptr.contents.nz

    @property
    def row(self):
        return _view(self.ptr.contents.row, (self.rows + 1,), np.uint32)

    @property
    def col(self):
        return _view(self.ptr.contents.col, (self.nz,), np.uint32)

    @property
    def coeff(self):
        return _view(self.ptr.contents.coeff, (self.nz,), np.complex128)


class Surface3d(_Handle):
    _del = 'del_surface3d'

    def __init__(self, vertices, edges, triangles):
        super().__init__()
        self._setup(_lib_h2().new_surface3d(vertices, edges, triangles), True)

    vertices = property(lambda self: self.ptr.contents.vertices)
    edges = property(lambda self: self.ptr.contents.edges)
    triangles = property(lambda self: self.ptr.contents.triangles)
    hmin = property(lambda self: self.ptr.contents.hmin)
    hmax = property(lambda self: self.ptr.contents.hmax)
    x = property(lambda self: _view(self.ptr.contents.x, (self.vertices, 3), np.float64))
    e = property(lambda self: _view(self.ptr.contents.e, (self.edges, 2), np.uint32))
    t = property(lambda self: _view(self.ptr.contents.t, (self.triangles, 3), np.uint32))
    s = property(lambda self: _view(self.ptr.contents.s, (self.triangles, 3), np.uint32))
    n = property(lambda self: _view(self.ptr.contents.n, (self.triangles, 3), np.float64))
    g = property(lambda self: _view(self.ptr.contents.g, (self.triangles,), np.float64))


class Macrosurface3d(_Handle):
    _del = 'del_macrosurface3d'

    def __init__(self, vertices, edges, triangles):
        super().__init__()
        self._setup(_lib_h2().new_macrosurface3d(vertices, edges, triangles), True)

    def set_parametrization(self, type):
        name = {'square': 'affine', 'cube': 'affine', 'circle': 'circle', 'sphere': 'sphere'}.get(type.lower())
        if name is None:
            raise ValueError
        fn = getattr(_lib_h2(), f'cnld_b200_phi_{name}')
        self.ptr.contents.phi = C.cast(fn, C.c_void_p).value

    vertices = property(lambda self: self.ptr.contents.vertices)
    edges = property(lambda self: self.ptr.contents.edges)
    triangles = property(lambda self: self.ptr.contents.triangles)
    x = property(lambda self: _view(self.ptr.contents.x, (self.vertices, 3), np.float64))
    e = property(lambda self: _view(self.ptr.contents.e, (self.edges, 2), np.uint32))
    t = property(lambda self: _view(self.ptr.contents.t, (self.triangles, 3), np.uint32))
    s = property(lambda self: _view(self.ptr.contents.s, (self.triangles, 3), np.uint32))


class Bem3d(_Handle):
    _del = 'del_bem3d'

    def __init__(self, surf, row_basis, col_basis):
        super().__init__()
        self._surf = surf
        self._setup(_lib_h2().new_bem3d(surf.ptr, int(row_basis), int(col_basis)), True)

    @property
    def k(self):
        k = self.ptr.contents.k
        return complex(k.re, k.im)

    @property
    def kernel_const(self):
        k = self.ptr.contents.kernel_const
        return complex(k.re, k.im)


class Cluster(_Handle):
    """cluster.pyx: owner also frees the index array (cluster.pyx:19-22)."""
    _del = 'del_cluster'

    def __init__(self, size, idx, sons, dim):
        super().__init__()
        self._idx_keep = np.ascontiguousarray(idx, dtype=np.uint32)
        self._setup(_lib_h2().new_cluster(size, self._idx_keep.ctypes.data_as(C.POINTER(uint)), sons, dim), True)
        self._owns_idx = False

    def __del__(self):
        if getattr(self, 'ptr', None) and self.owner and _L is not None:
            if getattr(self, '_owns_idx', True):
                _L.freemem(C.cast(self.ptr.contents.idx, C.c_void_p))
            _L.del_cluster(self.ptr)
            self.ptr = None

    size = property(lambda self: self.ptr.contents.size)
    sons = property(lambda self: self.ptr.contents.sons)
    dim = property(lambda self: self.ptr.contents.dim)
    desc = property(lambda self: self.ptr.contents.desc)
    type = property(lambda self: self.ptr.contents.type)
    idx = property(lambda self: _view(self.ptr.contents.idx, (self.size,), np.uint32))
    bmin = property(lambda self: _view(self.ptr.contents.bmin, (self.dim,), np.float64))
    bmax = property(lambda self: _view(self.ptr.contents.bmax, (self.dim,), np.float64))

    @property
    def son(self):
        out = []
        for i in range(self.sons):
            c = Cluster.wrap(self.ptr.contents.son[i], False)
            c._parent = self
            out.append(c)
        return out


class Block(_Handle):
    _del = 'del_block'

    def __init__(self, rc, cc, a, rsons, csons):
        super().__init__()
        self._keep = (rc, cc)
        self._setup(_lib_h2().new_block(rc.ptr, cc.ptr, bool(a), rsons, csons), True)

    a = property(lambda self: bool(self.ptr.contents.a))
    rsons = property(lambda self: self.ptr.contents.rsons)
    csons = property(lambda self: self.ptr.contents.csons)
    desc = property(lambda self: self.ptr.contents.desc)
    rc = property(lambda self: Cluster.wrap(self.ptr.contents.rc, False))
    cc = property(lambda self: Cluster.wrap(self.ptr.contents.cc, False))

    @property
    def son(self):
        # the reference indexes rsons + csons sons (hmatrix.pyx:64-66, right only for 2x2); the
        # full rsons * csons list is returned here
        out = []
        for i in range(self.rsons * self.csons):
            b = Block.wrap(self.ptr.contents.son[i], False)
            b._parent = self
            out.append(b)
        return out


class RKMatrix(_Handle):
    _del = 'del_rkmatrix'

    def __init__(self, rows, cols, k):
        super().__init__()
        self._setup(_lib_h2().new_rkmatrix(rows, cols, k), True)

    k = property(lambda self: self.ptr.contents.k)

    @property
    def A(self):
        m = AMatrix.wrap(C.pointer(self.ptr.contents.A), False)
        m._parent = self
        return m

    @property
    def B(self):
        m = AMatrix.wrap(C.pointer(self.ptr.contents.B), False)
        m._parent = self
        return m


class HMatrix(_Handle):
    _del = 'del_hmatrix'

    def __init__(self, rc, cc):
        super().__init__()
        self._keep = (rc, cc)
        self._setup(_lib_h2().new_hmatrix(rc.ptr, cc.ptr), True)

    rsons = property(lambda self: self.ptr.contents.rsons)
    csons = property(lambda self: self.ptr.contents.csons)
    desc = property(lambda self: self.ptr.contents.desc)
    rc = property(lambda self: Cluster.wrap(self.ptr.contents.rc, False))
    cc = property(lambda self: Cluster.wrap(self.ptr.contents.cc, False))

    @property
    def r(self):
        if self.ptr.contents.r:
            m = RKMatrix.wrap(self.ptr.contents.r, False)
            m._parent = self
            return m

    @property
    def f(self):
        if self.ptr.contents.f:
            m = AMatrix.wrap(self.ptr.contents.f, False)
            m._parent = self
            return m

    @property
    def son(self):
        out = []
        for i in range(self.rsons * self.csons):
            h = HMatrix.wrap(self.ptr.contents.son[i], False)
            h._parent = self
            out.append(h)
        return out


class Truncmode(_Handle):
    _del = 'del_truncmode'

    def __init__(self):
        super().__init__()
        self._setup(_lib_h2().new_truncmode(), True)


# ---- function layer (same names as the cpdef wrappers) ---------------------------------
def init_h2lib():
    _lib_h2().init_h2lib(None, None)


def uninit_h2lib():
    if _L is not None:
        _L.uninit_h2lib()


def new_zero_avector(dim):
    return AVector.wrap(_lib_h2().new_zero_avector(dim), True)


def clear_avector(v):
    _lib_h2().clear_avector(v.ptr)


def copy_avector(v, w):
    _lib_h2().copy_avector(v.ptr, w.ptr)


def random_avector(v):
    _lib_h2().random_avector(v.ptr)


def new_zero_amatrix(rows, cols):
    return AMatrix.wrap(_lib_h2().new_zero_amatrix(rows, cols), True)


def clone_amatrix(src):
    return AMatrix.wrap(_lib_h2().clone_amatrix(src.ptr), True)


def scale_amatrix(alpha, a):
    _lib_h2().scale_amatrix(_f(alpha), a.ptr)


def conjugate_amatrix(a):
    _lib_h2().conjugate_amatrix(a.ptr)


def addeval_amatrix_avector(alpha, a, src, trg):
    _lib_h2().addeval_amatrix_avector(_f(alpha), a.ptr, src.ptr, trg.ptr)


def add_amatrix(alpha, atrans, a, b):
    _lib_h2().add_amatrix(_f(alpha), bool(atrans), a.ptr, b.ptr)


def getsize_amatrix(a):
    return _lib_h2().getsize_amatrix(a.ptr)


def addmul_amatrix(alpha, atrans, a, btrans, b, c):
    _lib_h2().addmul_amatrix(_f(alpha), bool(atrans), a.ptr, bool(btrans), b.ptr, c.ptr)


def new_raw_sparsematrix(rows, cols, nz):
    return SparseMatrix.wrap(_lib_h2().new_raw_sparsematrix(rows, cols, nz), True)


def new_identity_sparsematrix(rows, cols):
    return SparseMatrix.wrap(_lib_h2().new_identity_sparsematrix(rows, cols), True)


def getsize_sparsematrix(a):
    return _lib_h2().getsize_sparsematrix(a.ptr)


def sort_sparsematrix(a):
    _lib_h2().sort_sparsematrix(a.ptr)


def clear_sparsematrix(a):
    _lib_h2().clear_sparsematrix(a.ptr)


def add_sparsematrix_amatrix(alpha, atrans, a, b):
    _lib_h2().add_sparsematrix_amatrix(_f(alpha), bool(atrans), a.ptr, b.ptr)


def addeval_sparsematrix_avector(alpha, a, x, y):
    _lib_h2().addeval_sparsematrix_avector(_f(alpha), a.ptr, x.ptr, y.ptr)


def prepare_surface3d(gr):
    _lib_h2().prepare_surface3d(gr.ptr)


def merge_surface3d(gr1, gr2):
    return Surface3d.wrap(_lib_h2().merge_surface3d(gr1.ptr, gr2.ptr), True)


def refine_red_surface3d(gr):
    return Surface3d.wrap(_lib_h2().refine_red_surface3d(gr.ptr), True)


def translate_surface3d(gr, t):
    t = np.ascontiguousarray(t, dtype=np.float64)
    _lib_h2().translate_surface3d(gr.ptr, t.ctypes.data_as(C.POINTER(real)))


def check_surface3d(gr):
    return _lib_h2().check_surface3d(gr.ptr)


def isclosed_surface3d(gr):
    return bool(_lib_h2().isclosed_surface3d(gr.ptr))


def isoriented_surface3d(gr):
    return bool(_lib_h2().isoriented_surface3d(gr.ptr))


def build_from_macrosurface3d_surface3d(ms, refn):
    return Surface3d.wrap(_lib_h2().build_from_macrosurface3d_surface3d(ms.ptr, refn), True)


def new_sphere_macrosurface3d():
    return Macrosurface3d.wrap(_lib_h2().new_sphere_macrosurface3d(), True)


def new_slp_helmholtz_bem3d(k, surf, q_regular, q_singular, row_basis, col_basis):
    """Needs a CUDA device: the bem object owns the device-side mesh/quadrature state."""
    ptr = _lib_h2().new_slp_helmholtz_bem3d(_f(k), surf.ptr, q_regular, q_singular, int(row_basis),
                                            int(col_basis))
    obj = Bem3d.wrap(ptr, True)
    obj._surf = surf  # keep the surface alive: bem->gr borrows it
    return obj


def assemble_bem3d_amatrix(bem, G):
    _lib_h2().assemble_bem3d_amatrix(bem.ptr, G.ptr)


def build_bem3d_cluster(bem, clf, basis):
    c = Cluster.wrap(_lib_h2().build_bem3d_cluster(bem.ptr, clf, int(basis)), True)
    c._owns_idx = True
    return c


def _admis_fn(admis):
    name = {'2': 'admissible_2_cluster', 'max': 'admissible_max_cluster', 'sphere': 'admissible_sphere_cluster',
            '2min': 'admissible_2_min_cluster', '2_min': 'admissible_2_min_cluster'}.get(str(admis).lower())
    if name is None:
        raise TypeError
    return C.cast(getattr(_lib_h2(), name), C.c_void_p)


def _build_block(fn, rc, cc, eta, admis):
    e = real(eta)
    b = Block.wrap(getattr(_lib_h2(), fn)(rc.ptr, cc.ptr, C.cast(C.pointer(e), C.c_void_p), _admis_fn(admis)), True)
    b._keep = (rc, cc)
    return b


def build_strict_block(rc, cc, eta, admis):
    return _build_block('build_strict_block', rc, cc, eta, admis)


def build_nonstrict_block(rc, cc, eta, admis):
    return _build_block('build_nonstrict_block', rc, cc, eta, admis)


def setup_hmatrix_aprx_aca_bem3d(bem, rc, cc, tree, accur):
    _lib_h2().setup_hmatrix_aprx_aca_bem3d(bem.ptr, rc.ptr, cc.ptr, tree.ptr, accur)


def setup_hmatrix_aprx_paca_bem3d(bem, rc, cc, tree, accur):
    _lib_h2().setup_hmatrix_aprx_paca_bem3d(bem.ptr, rc.ptr, cc.ptr, tree.ptr, accur)


def setup_hmatrix_aprx_hca_bem3d(bem, rc, cc, tree, m, accur):
    _lib_h2().setup_hmatrix_aprx_hca_bem3d(bem.ptr, rc.ptr, cc.ptr, tree.ptr, m, accur)


def setup_hmatrix_aprx_inter_row_bem3d(bem, rc, cc, tree, m):
    _lib_h2().setup_hmatrix_aprx_inter_row_bem3d(bem.ptr, rc.ptr, cc.ptr, tree.ptr, m)


def build_from_block_hmatrix(b, k):
    h = HMatrix.wrap(_lib_h2().build_from_block_hmatrix(b.ptr, k), True)
    h._keep = b
    return h


def assemble_bem3d_hmatrix(bem, b, G):
    _lib_h2().assemble_bem3d_hmatrix(bem.ptr, b.ptr, G.ptr)


def clear_hmatrix(hm):
    _lib_h2().clear_hmatrix(hm.ptr)


def copy_hmatrix(src, trg):
    _lib_h2().copy_hmatrix(src.ptr, trg.ptr)


def _derived(fn, src):
    h = HMatrix.wrap(getattr(_lib_h2(), fn)(src.ptr), True)
    h._keep = src
    return h


def clone_hmatrix(src):
    return _derived('clone_hmatrix', src)


def clonestructure_hmatrix(src):
    return _derived('clonestructure_hmatrix', src)


def getsize_hmatrix(hm):
    return _lib_h2().getsize_hmatrix(hm.ptr)


def getrows_hmatrix(hm):
    return _lib_h2().getrows_hmatrix(hm.ptr)


def getcols_hmatrix(hm):
    return _lib_h2().getcols_hmatrix(hm.ptr)


def addeval_hmatrix_avector(alpha, hm, x, y):
    _lib_h2().addeval_hmatrix_avector(_f(alpha), hm.ptr, x.ptr, y.ptr)


def addevalsymm_hmatrix_avector(alpha, hm, x, y):
    _lib_h2().addevalsymm_hmatrix_avector(_f(alpha), hm.ptr, x.ptr, y.ptr)


def copy_sparsematrix_hmatrix(sp, hm):
    _lib_h2().copy_sparsematrix_hmatrix(sp.ptr, hm.ptr)


def identity_hmatrix(hm):
    _lib_h2().identity_hmatrix(hm.ptr)


def scale_hmatrix(alpha, hm):
    _lib_h2().scale_hmatrix(_f(alpha), hm.ptr)


def add_hmatrix_amatrix(alpha, atrans, a, b):
    _lib_h2().add_hmatrix_amatrix(_f(alpha), bool(atrans), a.ptr, b.ptr)


def new_releucl_truncmode():
    return Truncmode.wrap(_lib_h2().new_releucl_truncmode(), True)


def solve_gmres_hmatrix_avector(A, b, x, eps, maxiter, kmax):
    return _lib_h2().solve_gmres_hmatrix_avector(A.ptr, b.ptr, x.ptr, eps, maxiter, kmax)


def solve_gmres_amatrix_avector(A, b, x, eps, maxiter, kmax):
    return _lib_h2().solve_gmres_amatrix_avector(A.ptr, b.ptr, x.ptr, eps, maxiter, kmax)


def lrdecomp_amatrix(a):
    return _lib_h2().lrdecomp_amatrix(a.ptr)


def lrsolve_amatrix_avector(atrans, a, x):
    _lib_h2().lrsolve_amatrix_avector(bool(atrans), a.ptr, x.ptr)


def triangularsolve_amatrix_avector(alower, aunit, atrans, a, x):
    _lib_h2().triangularsolve_amatrix_avector(bool(alower), bool(aunit), bool(atrans), a.ptr, x.ptr)


# ---- the rest of the surface cnld/h2lib/*.pyx exposes (krylovsolvers / factorizations / matrixnorms / harith)
def solve_cg_amatrix_avector(A, b, x, eps, maxiter):
    return _lib_h2().solve_cg_amatrix_avector(A.ptr, b.ptr, x.ptr, eps, maxiter)


def solve_cg_hmatrix_avector(A, b, x, eps, maxiter):
    return _lib_h2().solve_cg_hmatrix_avector(A.ptr, b.ptr, x.ptr, eps, maxiter)


def choldecomp_amatrix(a):
    return _lib_h2().choldecomp_amatrix(a.ptr)


def cholsolve_amatrix_avector(a, x):
    _lib_h2().cholsolve_amatrix_avector(a.ptr, x.ptr)


def new_dlp_helmholtz_bem3d(k, surf, q_regular, q_singular, row_basis, col_basis, alpha):
    """Not on cnl-dyna's path and not implemented by libcnld_b200: raises with the library's message."""
    ptr = _lib_h2().new_dlp_helmholtz_bem3d(_f(k), surf.ptr, q_regular, q_singular, int(row_basis), int(col_basis), _f(alpha))
    if not ptr:
        raise NotImplementedError(_lib.lib().cnld_b200_last_error().decode())
    return Bem3d.wrap(ptr, True)


def _norm(name):
    def fn(a, b):
        return getattr(_lib_h2(), name)(a.ptr, b.ptr)
    fn.__name__ = name
    return fn


for _n in ('amatrix_sparsematrix', 'amatrix_hmatrix', 'sparsematrix_hmatrix', 'lr_amatrix', 'lr_hmatrix',
           'lr_amatrix_hmatrix', 'lr_hmatrix_amatrix', 'lr_sparsematrix_amatrix', 'lr_sparsematrix_hmatrix'):
    globals()['norm2diff_' + _n] = _norm('norm2diff_' + _n)


def lrdecomp_hmatrix(a, tm, eps):
    _lib_h2().lrdecomp_hmatrix(a.ptr, tm.ptr, eps)


def lrsolve_hmatrix_avector(atrans, a, x):
    _lib_h2().lrsolve_hmatrix_avector(bool(atrans), a.ptr, x.ptr)


def lreval_hmatrix_avector(atrans, a, x):
    _lib_h2().lreval_hmatrix_avector(bool(atrans), a.ptr, x.ptr)


def choldecomp_hmatrix(a, tm, eps):
    _lib_h2().choldecomp_hmatrix(a.ptr, tm.ptr, eps)


def cholsolve_hmatrix_avector(a, x):
    _lib_h2().cholsolve_hmatrix_avector(a.ptr, x.ptr)


def choleval_hmatrix_avector(a, x):
    _lib_h2().choleval_hmatrix_avector(a.ptr, x.ptr)


def addmul_hmatrix(alpha, xtrans, x, ytrans, y, tm, eps, z):
    _lib_h2().addmul_hmatrix(_f(alpha), bool(xtrans), x.ptr, bool(ytrans), y.ptr, tm.ptr, eps, z.ptr)


def add_hmatrix(alpha, a, tm, eps, b):
    _lib_h2().add_hmatrix(_f(alpha), a.ptr, tm.ptr, eps, b.ptr)


def add_amatrix_hmatrix(alpha, atrans, a, tm, eps, b):
    _lib_h2().add_amatrix_hmatrix(_f(alpha), bool(atrans), a.ptr, tm.ptr, eps, b.ptr)


def triangularsolve_hmatrix_avector(alower, aunit, atrans, a, x):
    _lib_h2().triangularsolve_hmatrix_avector(bool(alower), bool(aunit), bool(atrans), a.ptr, x.ptr)


__all__ = [n for n in dir() if not n.startswith('_') and n not in ('C', 'np', 'enum', 'atexit', 'real', 'uint')]

# same life-cycle as cnld/h2lib/__init__.py:20-29
init_h2lib()
atexit.register(uninit_h2lib)
