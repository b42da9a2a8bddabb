"""
ctypes binding of libcnld_b200.so (C ABI in include/cnld_b200.h).

This is the binding a cnl-dyna maintainer would drop next to cnld/h2lib/*.pyx (see
INTEGRATION.md).  The library is the product: if it is missing or no CUDA device is
present every compute call raises -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'libcnld_b200.so')

_lib = None

u32 = C.c_uint
f64 = C.c_double
vp = C.c_void_p


class cfield(C.Structure):
    """cnld_field by value (two doubles)."""
    _fields_ = [('re', C.c_double), ('im', C.c_double)]


class Cnldb200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Cnldb200Error(
                f'{LIB_PATH} not found: build it with `make -C cnl-dyna_b200` '
                '(or __graft_entry__.build()); libcnld_b200 has no CPU fallback')
        L = C.CDLL(LIB_PATH)
        L.cnld_b200_last_error.restype = C.c_char_p
        L.cnld_b200_version.restype = C.c_char_p
        L.cnld_b200_launch_count.restype = C.c_ulonglong
        L.cnld_b200_dmma_flop_count.restype = C.c_ulonglong
        L.cnld_b200_create.restype = vp
        L.cnld_b200_create.argtypes = [C.c_int, u32, u32, vp, vp, u32, u32]
        L.cnld_b200_destroy.argtypes = [vp]
        L.cnld_b200_set_streams.argtypes = [vp, C.c_int]
        L.cnld_b200_pin_host.argtypes = [vp, vp, C.c_size_t]
        L.cnld_b200_get_geometry.argtypes = [vp, vp, vp]
        L.cnld_b200_assemble_slp_ll.argtypes = [vp, f64, f64, vp, u32]
        L.cnld_b200_nearfield_slp_ll.argtypes = [vp, f64, f64, vp, u32, vp, u32, vp, u32]
        L.cnld_b200_lrdecomp.argtypes = [C.c_int, u32, vp, u32, vp]
        L.cnld_b200_lrdecomp_sym.argtypes = [C.c_int, u32, vp, u32, vp]
        L.cnld_b200_lrsolve.argtypes = [C.c_int, u32, vp, u32, vp, u32, u32]
        L.cnld_b200_zgemm.argtypes = [C.c_int, u32, u32, u32, f64, vp, u32, vp, u32, f64, vp, u32]
        L.cnld_b200_set_mbk.argtypes = [vp] * 6
        L.cnld_b200_set_loads.argtypes = [vp, u32] + [vp] * 7
        L.cnld_b200_sweep_run.argtypes = [vp, vp, u32, f64, f64, vp, vp, vp]
        L.cnld_b200_sweep_run_device.argtypes = [vp, vp, u32, f64, f64]
        L.cnld_b200_sweep_fetch.argtypes = [vp, u32, vp, vp]
        L.cnld_b200_sweep_stage_times.argtypes = [vp, vp]
        L.cnld_b200_set_profile.argtypes = [vp, C.c_int]
        L.cnld_b200_set_symmetric.argtypes = [vp, C.c_int, C.c_int]
        L.cnld_b200_set_symmetric_tol.argtypes = [vp, f64]
        L.cnld_b200_set_translation_classes.argtypes = [vp, C.c_int]
        L.cnld_b200_translation_info.argtypes = [vp, vp, vp]
        L.cnld_b200_symmetric_stats.argtypes = [vp, vp, vp, vp]
        L.cnld_b200_sweep_kernel_stats.argtypes = [vp, vp]
        L.cnld_b200_pres_vector.argtypes = [vp, vp, u32, f64, f64, f64, vp]
        L.cnld_b200_pressure.argtypes = [vp, vp, u32, vp, u32, f64, f64, vp, u32, vp]
        L.cnld_b200_pfield_create.restype = vp
        L.cnld_b200_pfield_create.argtypes = [vp, vp, u32, u32]
        L.cnld_b200_pfield_destroy.argtypes = [vp]
        L.cnld_b200_pfield_pin.argtypes = [vp, vp, C.c_size_t]
        L.cnld_b200_pfield_run.argtypes = [vp, f64, f64, f64, vp, u32, vp]
        L.cnld_b200_pfield_run_device.argtypes = [vp, f64, f64, f64, u32]
        L.cnld_b200_pfield_fetch.argtypes = [vp, u32, vp, u32, vp]
        L.cnld_b200_pfield_info.argtypes = [vp, vp]
        L.cnld_b200_pres_plan_stats.argtypes = [u32, u32, vp, vp, vp, vp]
        L.cnld_b200_fir_create.restype = vp
        L.cnld_b200_fir_create.argtypes = [C.c_int, u32, u32, u32, vp]
        L.cnld_b200_fir_destroy.argtypes = [vp]
        L.cnld_b200_fir_conv.argtypes = [vp, vp, u32, f64, C.c_int, vp]
        L.cnld_b200_fir_reset.argtypes = [vp]
        L.cnld_b200_fir_push.argtypes = [vp, vp]
        L.cnld_b200_fir_base.argtypes = [vp, f64, vp]
        L.cnld_b200_fir_add.argtypes = [vp, vp, f64, vp]
        L.cnld_b200_td_create.restype = vp
        L.cnld_b200_td_create.argtypes = [C.c_int, u32, u32, vp, u32, vp, vp] + [f64] * 7 + [C.c_int]
        L.cnld_b200_td_destroy.argtypes = [vp]
        L.cnld_b200_td_reset.argtypes = [vp]
        L.cnld_b200_td_step.argtypes = [vp, u32]
        L.cnld_b200_td_current_step.argtypes = [vp]
        L.cnld_b200_td_current_step.restype = u32
        L.cnld_b200_td_taps.argtypes = [vp]
        L.cnld_b200_td_taps.restype = u32
        L.cnld_b200_td_state.argtypes = [vp, u32, u32] + [vp] * 6
        L.cnld_b200_td_info.argtypes = [vp, u32, u32, vp, vp]
        L.cnld_b200_td_trace.argtypes = [vp, u32, u32, vp]
        L.cnld_b200_fft_to_fir.argtypes = [C.c_int, u32, vp, u32, vp, u32, C.c_int, vp, vp]
        L.cnld_b200_comm_create.restype = vp
        L.cnld_b200_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.cnld_b200_comm_destroy.argtypes = [vp]
        L.cnld_b200_comm_rank.argtypes = [vp]
        L.cnld_b200_comm_world.argtypes = [vp]
        L.cnld_b200_comm_bcast.argtypes = [vp, vp, C.c_size_t, C.c_int]
        L.cnld_b200_comm_gather.argtypes = [vp, vp, C.c_size_t, vp, C.c_int]
        L.cnld_b200_comm_allreduce_max.argtypes = [vp, vp, C.c_int]
        L.cnld_b200_comm_barrier.argtypes = [vp]
        L.cnld_b200_comm_stats.argtypes = [vp, vp]
        L.cnld_b200_db_append_block.argtypes = [C.c_char_p, u32, vp, vp, u32, vp, vp, u32, vp, vp, vp, C.c_int]
        L.cnld_b200_db_begin_bulk.argtypes = [C.c_char_p]
        L.cnld_b200_db_finish.argtypes = [C.c_char_p]
        L.cnld_b200_db_counters.argtypes = [vp]
        L.cnld_b200_db_append_imp_resp.argtypes = [C.c_char_p, u32, u32, vp, vp, C.c_int]
        L.cnld_b200_hmat_create.restype = vp
        L.cnld_b200_hmat_create.argtypes = [vp, u32, f64, C.c_int, C.c_int, u32]
        L.cnld_b200_hmat_destroy.argtypes = [vp]
        L.cnld_b200_hmat_info.argtypes = [vp, vp]
        L.cnld_b200_hmat_tree.argtypes = [vp] * 12
        L.cnld_b200_hmat_assemble.argtypes = [vp, f64, f64, f64]
        L.cnld_b200_hmat_matvec.argtypes = [vp, cfield, vp, vp, u32]
        L.cnld_b200_hmat_capped.argtypes = [vp, vp, vp]
        L.cnld_b200_hmat_set_coarse.argtypes = [vp, u32, vp, vp]
        L.cnld_b200_hmat_coarse_used.argtypes = [vp]
        L.cnld_b200_hmat_todense.argtypes = [vp, vp, u32]
        L.cnld_b200_hmat_bench_matvec.argtypes = [vp, u32, C.c_int, vp]
        L.cnld_b200_hsweep_run.argtypes = [vp, vp, vp, u32, f64, f64, f64, f64, u32, u32, u32, vp, vp, vp, vp]
        L.cnld_b200_bench_zgemm.argtypes = [C.c_int, u32, u32, u32, C.c_int, vp]
        L.cnld_b200_bench_dfma.argtypes = [C.c_int, C.c_int, vp]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise Cnldb200Error(lib().cnld_b200_last_error().decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def device_count():
    return lib().cnld_b200_device_count()


def launch_count():
    return int(lib().cnld_b200_launch_count())


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Context:
    """Mesh + quadrature + FEM data resident on one GPU (opaque cnld_ctx)."""

    def __init__(self, vertices, triangles, q_reg=2, q_sing=4, device=0):
        self.vertices = _c(vertices, np.float64)
        self.triangles = _c(triangles, np.uint32)
        self.nv, self.nt = len(self.vertices), len(self.triangles)
        self.device = device
        self.npatch = 0
        self._h = lib().cnld_b200_create(device, self.nv, self.nt, ptr(self.vertices),
                                         ptr(self.triangles), q_reg, q_sing)
        if not self._h:
            raise Cnldb200Error(lib().cnld_b200_last_error().decode())

    def close(self):
        if getattr(self, '_h', None):
            if getattr(self, '_pinned', None) is not None:
                lib().cnld_b200_pin_host(self._h, None, 0)
                self._pinned = None
            lib().cnld_b200_destroy(self._h)
            self._h = None

    __del__ = close

    def set_streams(self, n):
        check(lib().cnld_b200_set_streams(self._h, int(n)))

    def set_symmetric(self, mode, refine_steps=1, tol=None):
        """mode 1: factor the symmetric part at half the flops and refine with the sparse asymmetry."""
        check(lib().cnld_b200_set_symmetric(self._h, int(mode), int(refine_steps)))
        if tol is not None:
            check(lib().cnld_b200_set_symmetric_tol(self._h, float(tol)))

    def set_translation_classes(self, on=True):
        """Integrate one block of Z per membrane-offset class and copy the rest (symmetric path)."""
        check(lib().cnld_b200_set_translation_classes(self._h, int(bool(on))))

    def translation_info(self):
        nc, ncl = C.c_uint(0), C.c_uint(0)
        used = lib().cnld_b200_translation_info(self._h, C.byref(nc), C.byref(ncl))
        return dict(in_use=bool(used == 1), components=nc.value, offset_classes=ncl.value)

    def symmetric_stats(self):
        a, r, m = C.c_uint(0), C.c_uint(0), C.c_double(0.0)
        mode = lib().cnld_b200_symmetric_stats(self._h, C.byref(a), C.byref(r), C.byref(m))
        return dict(mode=mode, mbk_asym_nnz=a.value, redone=r.value, max_ratio=m.value)

    def geometry(self):
        g = np.zeros(self.nt)
        n = np.zeros((self.nt, 3))
        check(lib().cnld_b200_get_geometry(self._h, ptr(g), ptr(n)))
        return g, n

    def assemble_z(self, k):
        """Z[i, j] (mathematical indexing) of the Helmholtz SLP, linear basis."""
        k = complex(k)
        buf = np.zeros((self.nv, self.nv), dtype=np.complex128)  # column-major => buf = Z^T
        check(lib().cnld_b200_assemble_slp_ll(self._h, k.real, k.imag, ptr(buf), self.nv))
        return buf.T

    def nearfield(self, k, ridx, cidx):
        """Z[ridx, cidx] (bem->nearfield): Galerkin entries for arbitrary vertex index sets."""
        k = complex(k)
        ridx, cidx = _c(ridx, np.uint32), _c(cidx, np.uint32)
        buf = np.zeros((len(cidx), len(ridx)), dtype=np.complex128)
        check(lib().cnld_b200_nearfield_slp_ll(self._h, k.real, k.imag, ptr(ridx), len(ridx), ptr(cidx), len(cidx),
                                               ptr(buf), max(len(ridx), 1)))
        return buf.T

    def set_mbk(self, K, M, B):
        """K, M, B: scipy sparse (any format) or dense, real, N x N, same shape."""
        from scipy import sparse as sps
        K, M, B = (sps.csr_matrix(a) for a in (K, M, B))
        pat = (abs(K) + abs(M) + abs(B)).tocsr()
        pat.sort_indices()
        rows = np.repeat(np.arange(self.nv), np.diff(pat.indptr))
        cols = pat.indices
        vals = [np.asarray(a[rows, cols]).ravel().astype(np.float64) for a in (K, M, B)]
        self._mbk = (_c(pat.indptr, np.int64), _c(cols, np.uint32)) + tuple(_c(v, np.float64) for v in vals)
        self.upload_mbk()

    def upload_mbk(self):
        """(re-)copy the prepared K/M/B arrays host -> device"""
        check(lib().cnld_b200_set_mbk(self._h, *[ptr(a) for a in self._mbk]))

    def set_loads(self, F, AVG, on_boundary):
        from scipy import sparse as sps
        F, AVG = sps.csc_matrix(F), sps.csc_matrix(AVG)
        F.sort_indices()
        AVG.sort_indices()
        self.npatch = F.shape[1]
        self._loads = (_c(F.indptr, np.int64), _c(F.indices, np.uint32), _c(F.data, np.float64),
                       _c(AVG.indptr, np.int64), _c(AVG.indices, np.uint32), _c(AVG.data, np.float64),
                       _c(on_boundary, np.uint8))
        self.upload_loads()

    def upload_loads(self):
        check(lib().cnld_b200_set_loads(self._h, self.npatch, *[ptr(a) for a in self._loads]))

    def sweep(self, freqs, c=1500., rho=1000., want_nodes=True, out_nodes=None):
        """-> x_patch[nf, P(src), P(dst)], x_node[nf, P(src), N] or None, t[nf].
        `out_nodes` lets a caller reuse a preallocated [nf, P, N] complex128 buffer."""
        freqs = _c(freqs, np.float64)
        nf, P = len(freqs), self.npatch
        xp = np.empty((nf, P, P), dtype=np.complex128)
        xn = None
        if want_nodes:
            xn = out_nodes if out_nodes is not None else np.empty((nf, P, self.nv), dtype=np.complex128)
            assert xn.shape == (nf, P, self.nv) and xn.dtype == np.complex128 and xn.flags.c_contiguous
            if out_nodes is not None:
                self.pin(out_nodes)        # caller-owned, reused buffer: D2H lands in it directly
        t = np.zeros(nf)
        check(lib().cnld_b200_sweep_run(self._h, ptr(freqs), nf, c, rho, ptr(xp), ptr(xn), ptr(t)))
        return xp, xn, t

    def pin(self, arr):
        """Page-lock `arr` in place (kept referenced here until replaced or close())."""
        if arr is None:
            if getattr(self, '_pinned', None) is not None:
                lib().cnld_b200_pin_host(self._h, None, 0)
                self._pinned = None
            return
        base = arr if arr.base is None else arr.base
        cur = getattr(self, '_pinned', None)
        if cur is not None and cur is base:
            return
        if not isinstance(base, np.ndarray) or not base.flags.c_contiguous:
            return
        if lib().cnld_b200_pin_host(self._h, ptr(base), base.nbytes) == 0:
            self._pinned = base
        else:
            self._pinned = None

    def sweep_device(self, freqs, c=1500., rho=1000.):
        freqs = _c(freqs, np.float64)
        check(lib().cnld_b200_sweep_run_device(self._h, ptr(freqs), len(freqs), c, rho))

    def sweep_fetch(self, nf):
        xp = np.zeros((nf, self.npatch, self.npatch), dtype=np.complex128)
        check(lib().cnld_b200_sweep_fetch(self._h, nf, ptr(xp), None))
        return xp

    def stage_times(self):
        out = np.zeros(5)
        check(lib().cnld_b200_sweep_stage_times(self._h, ptr(out)))
        return dict(assemble=out[0], factor=out[1], solve=out[2], launches=int(out[3]), device_seconds=out[4])

    def set_profile(self, on=True):
        check(lib().cnld_b200_set_profile(self._h, int(on)))

    def kernel_stats(self):
        out = np.zeros(3)
        check(lib().cnld_b200_sweep_kernel_stats(self._h, ptr(out)))
        return dict(flops=out[0], seconds=out[1], launches=int(out[2]))

    def pres_vector(self, obs, k, c=1500., rho=1000.):
        obs = _c(np.atleast_2d(obs), np.float64)
        out = np.zeros((len(obs), self.nv), dtype=np.complex128)
        check(lib().cnld_b200_pres_vector(self._h, ptr(obs), len(obs), float(k), c, rho, ptr(out)))
        return out

    def pressure(self, obs, ks, disp, c=1500., rho=1000.):
        """disp[nk, nsrc, N] -> p[nk, nsrc, nobs]."""
        obs = _c(np.atleast_2d(obs), np.float64)
        ks = _c(np.atleast_1d(ks), np.float64)
        disp = _c(disp, np.complex128)
        nk, nsrc, nv = disp.shape
        assert nk == len(ks) and nv == self.nv
        out = np.zeros((nk, nsrc, len(obs)), dtype=np.complex128)
        check(lib().cnld_b200_pressure(self._h, ptr(obs), len(obs), ptr(ks), nk, c, rho, ptr(disp), nsrc,
                                       ptr(out)))
        return out


class PressureField:
    """Observation points resident on one GPU + workspace of the many-source field-pressure path
    (opaque cnld_pfield): p[src, obs] = mesh_pres_vector(obs, k) . disp[src] for all points at once
    (the loop of bem_pressure.pyx:103-112)."""

    def __init__(self, ctx, obs, nsrc_max):
        self.ctx = ctx
        self.obs = _c(np.atleast_2d(obs), np.float64)
        self.nobs, self.nsrc_max = len(self.obs), int(nsrc_max)
        self._h = lib().cnld_b200_pfield_create(ctx._h, ptr(self.obs), self.nobs, self.nsrc_max)
        if not self._h:
            raise Cnldb200Error(lib().cnld_b200_last_error().decode())
        self._pinned = None

    def close(self):
        if getattr(self, '_h', None):
            lib().cnld_b200_pfield_destroy(self._h)
            self._h = None
            self._pinned = None

    __del__ = close

    def run(self, k, disp, c=1500., rho=1000., out=None, fetch=True):
        """disp[nsrc, N] (host) -> p[nsrc, nobs]; `out` = reusable host buffer (page-locked on first use)."""
        disp = _c(disp, np.complex128)
        nsrc = disp.shape[0]
        assert disp.shape[1] == self.ctx.nv
        p = None
        if fetch:
            p = out if out is not None else np.empty((nsrc, self.nobs), dtype=np.complex128)
            assert p.shape == (nsrc, self.nobs) and p.dtype == np.complex128 and p.flags.c_contiguous
            if out is not None and self._pinned is not out:
                self._pinned = out if lib().cnld_b200_pfield_pin(self._h, ptr(out), out.nbytes) == 0 else None
        check(lib().cnld_b200_pfield_run(self._h, float(k), c, rho, ptr(disp), nsrc, ptr(p)))
        return p

    def run_device(self, k, nsrc, c=1500., rho=1000.):
        check(lib().cnld_b200_pfield_run_device(self._h, float(k), c, rho, int(nsrc)))

    def fetch(self, nsrc, idx=None):
        if idx is None:
            p = np.empty((nsrc, self.nobs), dtype=np.complex128)
            check(lib().cnld_b200_pfield_fetch(self._h, nsrc, None, 0, ptr(p)))
            return p
        idx = _c(idx, np.uint32)
        p = np.empty((nsrc, len(idx)), dtype=np.complex128)
        check(lib().cnld_b200_pfield_fetch(self._h, nsrc, ptr(idx), len(idx), ptr(p)))
        return p

    def info(self):
        out = np.zeros(7)
        check(lib().cnld_b200_pfield_info(self._h, ptr(out)))
        keys = ('pvec_seconds', 'contract_seconds', 'total_seconds', 'chunks', 'obs_per_chunk', 'tri_chunks',
                'node_entries')
        return {k: (float(v) if i < 3 else int(v)) for i, (k, v) in enumerate(zip(keys, out))}


def pres_plan_stats(vertices, triangles):
    """Host-only: chunking of the mesh used by the pressure-vector kernel."""
    x = _c(vertices, np.float64)
    t = _c(triangles, np.uint32)
    counts = np.zeros(6, dtype=np.uint32)
    order = np.zeros(len(t), dtype=np.uint32)
    check(lib().cnld_b200_pres_plan_stats(len(x), len(t), ptr(x), ptr(t), ptr(counts), ptr(order)))
    keys = ('chunks', 'node_entries', 'max_triangles', 'max_nodes', 'accumulating_entries', 'untouched_vertices')
    return dict(zip(keys, (int(v) for v in counts))), order


class FirTable:
    """Patch-to-patch impulse responses fir[nsrc, ndest, nfir] resident on one GPU (opaque cnld_fir): the
    convolution of the time-domain solver (simulation.pyx:51-77 fir_conv_cy, :400-410)."""

    def __init__(self, fir, device=0):
        fir = _c(fir, np.float64)
        assert fir.ndim == 3
        self.nsrc, self.ndest, self.nfir = fir.shape
        self._h = lib().cnld_b200_fir_create(device, self.nsrc, self.ndest, self.nfir, ptr(fir))
        if not self._h:
            raise Cnldb200Error(lib().cnld_b200_last_error().decode())

    def close(self):
        if getattr(self, '_h', None):
            lib().cnld_b200_fir_destroy(self._h)
            self._h = None

    __del__ = close

    def conv(self, p, dt, offset):
        """fir_conv_cy(fir, p, dt, offset) for p[nsample, nsrc] -> x[ndest]."""
        p = _c(np.asarray(p).reshape(-1, self.nsrc), np.float64)
        x = np.empty(self.ndest)
        check(lib().cnld_b200_fir_conv(self._h, ptr(p), len(p), float(dt), int(offset), ptr(x)))
        return x

    def reset(self):
        check(lib().cnld_b200_fir_reset(self._h))

    def push(self, p_row):
        check(lib().cnld_b200_fir_push(self._h, ptr(_c(p_row, np.float64))))

    def base(self, dt):
        x = np.empty(self.ndest)
        check(lib().cnld_b200_fir_base(self._h, float(dt), ptr(x)))
        return x

    def add(self, p_row, dt):
        x = np.empty(self.ndest)
        check(lib().cnld_b200_fir_add(self._h, ptr(_c(p_row, np.float64)), float(dt), ptr(x)))
        return x


class TimeDomainSolver:
    """FixedStepSolver's arithmetic on one GPU (opaque cnld_td; simulation.pyx:297-555): fir[P, P, nfir] on the
    step dt, v[nt, P] on the solver's time grid.  `step(n)` advances n samples without a host round trip."""
    STATE = ('x', 'u', 'p_es', 'p_cont_spr', 'p_cont_dmp', 'p_tot')

    def __init__(self, fir, v, gap_eff, dt, k, n, x0, lmbd, atol=1e-10, maxiter=5, e0=None, device=0):
        if e0 is None:
            from scipy.constants import epsilon_0 as e0
        fir = _c(fir, np.float64)
        v = _c(v, np.float64)
        gap_eff = _c(gap_eff, np.float64)
        assert fir.ndim == 3 and fir.shape[0] == fir.shape[1] and v.ndim == 2 and v.shape[1] == fir.shape[0] == len(gap_eff)
        self.npatch, self.nfir, self.nt = fir.shape[0], fir.shape[2], v.shape[0]
        self._h = lib().cnld_b200_td_create(device, self.npatch, self.nfir, ptr(fir), self.nt, ptr(v), ptr(gap_eff), float(dt),
                                            float(e0), float(k), float(n), float(x0), float(lmbd), float(atol), int(maxiter))
        if not self._h:
            raise Cnldb200Error(lib().cnld_b200_last_error().decode())

    def close(self):
        if getattr(self, '_h', None):
            lib().cnld_b200_td_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def current_step(self):
        return int(lib().cnld_b200_td_current_step(self._h))

    @property
    def taps(self):
        """Taps a step streams: trailing taps that are zero for every patch pair are skipped."""
        return int(lib().cnld_b200_td_taps(self._h))

    def reset(self):
        check(lib().cnld_b200_td_reset(self._h))

    def step(self, nsteps=1):
        check(lib().cnld_b200_td_step(self._h, int(nsteps)))

    def state(self, i0=0, i1=None, into=None):
        """Samples [i0, i1) of every state array -> dict of [i1 - i0, P] arrays (or written into `into`'s rows)."""
        i1 = self.current_step if i1 is None else i1
        out = {}
        for key in self.STATE:
            out[key] = into[key][i0:i1] if into is not None else np.empty((i1 - i0, self.npatch))
            assert out[key].flags.c_contiguous
        check(lib().cnld_b200_td_state(self._h, i0, i1, *[ptr(out[key]) for key in self.STATE]))
        return out

    def trace(self, s0, s1):
        """%globaltimer marks [ns] of the step kernels (solver created with CNLD_B200_TD_TRACE=1), [s1 - s0, 8]."""
        out = np.zeros((s1 - s0, 8), dtype=np.uint64)
        check(lib().cnld_b200_td_trace(self._h, s0, s1, ptr(out)))
        return out

    def info(self, s0=1, s1=None):
        s1 = self.current_step if s1 is None else s1
        err = np.empty(max(s1 - s0, 0))
        it = np.empty(max(s1 - s0, 0), dtype=np.int32)
        check(lib().cnld_b200_td_info(self._h, s0, s1, ptr(err), ptr(it)))
        return err, it


def fft_to_fir(freqs, s, interp=1, use_kkr=True, device=0):
    """One-sided spectra s[..., nf] -> (t[nfft], h[..., nfft]) on the GPU (impulse_response.py:191-224)."""
    freqs = _c(freqs, np.float64)
    s = np.asarray(s, dtype=np.complex128)
    lead = s.shape[:-1]
    flat = _c(s.reshape(-1, s.shape[-1]), np.complex128)
    nfft = 2 * ((len(freqs) - 1) * int(interp) + 1) - 1
    t = np.empty(nfft)
    h = np.empty((flat.shape[0], nfft))
    check(lib().cnld_b200_fft_to_fir(device, len(freqs), ptr(freqs), flat.shape[0], ptr(flat), int(interp), int(bool(use_kkr)),
                                     ptr(t), ptr(h)))
    return t, h.reshape(lead + (nfft,))


class Comm:
    """One rank of the sweep's communicator (opaque cnld_comm): NCCL over NVLink through the C ABI, no
    torch.distributed.  Rank / world / rendezvous default to the launcher's environment (torchrun: RANK,
    WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT; the id exchange uses CNLD_B200_COMM_PORT or
    MASTER_PORT + 17)."""

    def __init__(self, rank=None, world=None, device=None, addr=None, port=None):
        env = os.environ
        self.rank = int(env.get('RANK', 0)) if rank is None else int(rank)
        self.world = int(env.get('WORLD_SIZE', 1)) if world is None else int(world)
        self.device = int(env.get('LOCAL_RANK', 0)) if device is None else int(device)
        addr = addr or env.get('MASTER_ADDR', '127.0.0.1')
        port = int(port or env.get('CNLD_B200_COMM_PORT', 0) or int(env.get('MASTER_PORT', 29500)) + 17)
        self._h = lib().cnld_b200_comm_create(self.rank, self.world, self.device, addr.encode(), port)
        if not self._h:
            raise Cnldb200Error(lib().cnld_b200_last_error().decode())

    def close(self):
        if getattr(self, '_h', None):
            lib().cnld_b200_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def bcast(self, arr, root=0):
        """In place: every rank's `arr` (same shape / dtype, C-contiguous) becomes root's."""
        assert arr.flags.c_contiguous
        check(lib().cnld_b200_comm_bcast(self._h, ptr(arr), arr.nbytes, root))
        return arr

    def bcast_object(self, obj, root=0):
        """Small Python object (pickled) from root to every rank."""
        import pickle
        payload = np.frombuffer(pickle.dumps(obj), dtype=np.uint8).copy() if self.rank == root else None
        n = np.array([len(payload) if payload is not None else 0], dtype=np.int64)
        self.bcast(n, root)
        if payload is None:
            payload = np.empty(int(n[0]), dtype=np.uint8)
        self.bcast(payload, root)
        return pickle.loads(payload.tobytes())

    def gather(self, arr, root=0):
        """Equal-shaped arrays of all ranks stacked along a new first axis on root; None elsewhere."""
        arr = np.ascontiguousarray(arr)
        out = np.empty((self.world,) + arr.shape, dtype=arr.dtype) if self.rank == root else None
        check(lib().cnld_b200_comm_gather(self._h, ptr(arr), arr.nbytes, ptr(out), root))
        return out

    def max(self, *vals):
        v = np.array(vals, dtype=np.float64)
        check(lib().cnld_b200_comm_allreduce_max(self._h, ptr(v), len(v)))
        return v.tolist() if len(v) != 1 else float(v[0])

    def barrier(self):
        check(lib().cnld_b200_comm_barrier(self._h))

    def stats(self):
        out = np.zeros(3)
        check(lib().cnld_b200_comm_stats(self._h, ptr(out)))
        return dict(nccl_bytes=float(out[0]), nccl_seconds=float(out[1]), nccl_version=int(out[2]))


def db_append_block(file, freqs, wavenums, x_patch, x_node=None, nodes=None, times=None, job_ids=None, fast=True):
    """One transaction: the rows generate_db.process writes for a block of frequencies (C writer)."""
    freqs, wavenums = _c(freqs, np.float64), _c(wavenums, np.float64)
    x_patch = _c(x_patch, np.complex128)
    nf, P, _ = x_patch.shape
    N = 0
    if x_node is not None:
        x_node = _c(x_node, np.complex128)
        N = x_node.shape[2]
        nodes = _c(nodes, np.float64)
        assert x_node.shape[:2] == (nf, P) and nodes.shape == (N, 3)
    times = _c(times, np.float64) if times is not None else None
    job_ids = _c(job_ids, np.uint32) if job_ids is not None else None
    check(lib().cnld_b200_db_append_block(os.fsencode(file), nf, ptr(freqs), ptr(wavenums), P, ptr(x_patch), ptr(times), N,
                                          ptr(x_node), ptr(nodes), ptr(job_ids), int(fast)))


def db_begin_bulk(file):
    check(lib().cnld_b200_db_begin_bulk(os.fsencode(file)))


def db_finish(file):
    check(lib().cnld_b200_db_finish(os.fsencode(file)))


def db_append_imp_resp(file, times, h, threads=8):
    """patch_to_patch_imp_resp rows from h[P, P, nt] (C writer; pages into an empty table, SQL otherwise)."""
    times, h = _c(times, np.float64), _c(h, np.float64)
    P, P2, nt = h.shape
    assert P == P2 and times.shape == (nt,)
    check(lib().cnld_b200_db_append_imp_resp(os.fsencode(file), P, nt, ptr(times), ptr(h), int(threads)))


def db_counters():
    """Node rows written as pages / through SQL, node indexes generated directly / by CREATE INDEX (this process)."""
    out = np.zeros(8, np.uint64)
    check(lib().cnld_b200_db_counters(ptr(out)))
    return dict(zip(('rows_paged', 'rows_sql', 'index_direct', 'index_sql', 'imp_rows_paged', 'imp_rows_sql', 'patch_rows_paged',
                     'patch_rows_sql'), (int(v) for v in out)))


ADMIS = {'2': 0, 'max': 1, 'sphere': 2, '2min': 3, '2_min': 3}


class HMatrix:
    """Device H-matrix of the Helmholtz SLP operator bound to a Context (opaque cnld_hmat)."""

    def __init__(self, ctx, clf=16, eta=0.8, admis='2', strict=True, kmax=64):
        self.ctx = ctx
        self._h = lib().cnld_b200_hmat_create(ctx._h, clf, float(eta), ADMIS[str(admis).lower()], int(bool(strict)), kmax)
        if not self._h:
            raise Cnldb200Error(lib().cnld_b200_last_error().decode())

    def close(self):
        if getattr(self, '_h', None):
            lib().cnld_b200_hmat_destroy(self._h)
            self._h = None

    __del__ = close

    def info(self):
        out = np.zeros(8)
        check(lib().cnld_b200_hmat_info(self._h, ptr(out)))
        keys = ('clusters', 'leaves', 'admissible', 'inadmissible', 'capacity', 'stored', 'maxrank', 'n')
        return {k: int(v) for k, v in zip(keys, out)}

    def sharing(self):
        out = np.zeros(4)
        f = lib().cnld_b200_hmat_sharing
        f.argtypes = [vp, vp]
        check(f(self._h, ptr(out)))
        return dict(zip(('assembled_admissible', 'assembled_inadmissible', 'shared', 'pool_entries'), (int(v) for v in out)))

    def tree(self):
        i = self.info()
        ncl, nl, n = i['clusters'], i['leaves'], i['n']
        t = dict(perm=np.zeros(n, np.uint32), son0=np.zeros(ncl, np.int32), son1=np.zeros(ncl, np.int32),
                 start=np.zeros(ncl, np.uint32), size=np.zeros(ncl, np.uint32), bmin=np.zeros((ncl, 3)),
                 bmax=np.zeros((ncl, 3)), leaf_rc=np.zeros(nl, np.uint32), leaf_cc=np.zeros(nl, np.uint32),
                 leaf_adm=np.zeros(nl, np.uint8), leaf_rank=np.zeros(nl, np.uint32))
        check(lib().cnld_b200_hmat_tree(self._h, *[ptr(a) for a in t.values()]))
        return t

    def assemble(self, k, eps_aca=1e-2):
        k = complex(k)
        check(lib().cnld_b200_hmat_assemble(self._h, k.real, k.imag, float(eps_aca)))
        self._warn_capped(self.capped()[0])

    def capped(self):
        """(rank-capped low-rank leaves of the last assemble, summed over the last sweep)."""
        a, b = C.c_uint(0), C.c_uint(0)
        check(lib().cnld_b200_hmat_capped(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @staticmethod
    def _warn_capped(count):
        if count:
            import warnings
            warnings.warn(f'H-matrix: ACA stopped at the rank capacity kmax in {count} low-rank leaves before reaching '
                          'eps_aca; the approximation is less accurate than requested -- raise kmax', RuntimeWarning)

    def matvec(self, x, alpha=1.0):
        """alpha * Z @ x for x[n] or x[n, nrhs]."""
        x = np.asarray(x, dtype=np.complex128)
        X = np.ascontiguousarray(x.reshape(x.shape[0], -1).T)      # [nrhs][n]
        Y = np.zeros_like(X)
        a = complex(alpha)
        check(lib().cnld_b200_hmat_matvec(self._h, cfield(a.real, a.imag), ptr(X), ptr(Y), X.shape[0]))
        return Y.T.reshape(x.shape)

    def bench_matvec(self, nrhs, reps=5):
        """Device-resident seconds per matvec with nrhs vectors."""
        out = C.c_double(0)
        check(lib().cnld_b200_hmat_bench_matvec(self._h, int(nrhs), int(reps), C.byref(out)))
        return out.value

    def todense(self):
        n = self.ctx.nv
        buf = np.zeros((n, n), dtype=np.complex128)
        check(lib().cnld_b200_hmat_todense(self._h, ptr(buf), n))
        return buf.T

    def set_coarse(self, W, fmode):
        """Coarse space of the two-level GMRES preconditioner: W[n, m] (m <= 16), fmode[m] in Hz."""
        if W is None:
            check(lib().cnld_b200_hmat_set_coarse(self._h, 0, None, None))
            self._coarse = 0
            return
        W, fmode = _c(W, np.float64), _c(fmode, np.float64)
        assert W.shape == (self.ctx.nv, len(fmode))
        check(lib().cnld_b200_hmat_set_coarse(self._h, W.shape[1], ptr(W), ptr(fmode)))
        self._coarse = W.shape[1]

    def enable_coarse(self, m=16):
        """Build the coarse space from the context's FEM operators: the first m in-vacuo eigenmodes
        (K v = lambda M v on the free nodes) of every membrane = connected component of the mesh."""
        W, fmode = membrane_modes(self.ctx, m)
        self.set_coarse(W, fmode)
        return fmode

    def coarse_used(self):
        return int(lib().cnld_b200_hmat_coarse_used(self._h))

    def sweep(self, freqs, c=1500., rho=1000., eps_aca=1e-2, tol=1e-6, restart=50, maxiter=500, batch=32,
              want_nodes=True, allow_unconverged=False, coarse=16):
        """H-matrix + GMRES sweep -> x_patch[nf, P, P], x_node[nf, P, N] | None, iters[nf, P], t[nf].
        Raises if a right-hand side misses `tol` within `maxiter` steps (unless allow_unconverged)."""
        freqs = _c(freqs, np.float64)
        nf, P, n = len(freqs), self.ctx.npatch, self.ctx.nv
        if coarse and getattr(self, '_coarse', None) is None:
            self.enable_coarse(int(coarse))       # once per H-matrix (frequency independent)
        elif not coarse and getattr(self, '_coarse', 0):
            self.set_coarse(None, None)
            self._coarse = None
        xp = np.zeros((nf, P, P), dtype=np.complex128)
        xn = np.zeros((nf, P, n), dtype=np.complex128) if want_nodes else None
        it = np.zeros((nf, P), dtype=np.uint32)
        t = np.zeros(nf)
        rc = lib().cnld_b200_hsweep_run(self.ctx._h, self._h, ptr(freqs), nf, c, rho, eps_aca, tol, restart, maxiter,
                                        batch, ptr(xp), ptr(xn), ptr(it), ptr(t))
        if rc == 2 and allow_unconverged:
            import warnings
            warnings.warn(lib().cnld_b200_last_error().decode(), RuntimeWarning)
        else:
            check(rc)
        self._warn_capped(self.capped()[1])
        return xp, xn, it, t


def membrane_modes(ctx, m=16):
    """(W[n, m], fmode[m]): per membrane (connected component of the mesh) the first m eigenmodes of the
    in-vacuo plate problem K v = lambda M v on its unclamped nodes, zero elsewhere; identical membranes share
    one eigen-decomposition.  fmode = resonance frequencies (Hz) of the first membrane's modes."""
    from scipy import sparse as sps
    from scipy.linalg import eigh
    from scipy.sparse.csgraph import connected_components
    n = ctx.nv
    indptr, cols, Kv, Mv, _ = ctx._mbk
    K = sps.csr_matrix((Kv, cols, indptr), shape=(n, n))
    M = sps.csr_matrix((Mv, cols, indptr), shape=(n, n))
    free = ~np.asarray(ctx._loads[-1], dtype=bool)
    t = ctx.triangles.astype(np.int64)
    adj = sps.coo_matrix((np.ones(3 * len(t)), (np.r_[t[:, 0], t[:, 1], t[:, 2]], np.r_[t[:, 1], t[:, 2], t[:, 0]])), shape=(n, n))
    ncomp, lab = connected_components(adj, directed=False)
    W = np.zeros((n, m))
    fmode, cache = None, []
    for c in range(ncomp):
        idx = np.nonzero((lab == c) & free)[0]
        if len(idx) == 0:
            continue
        Kf = K[idx][:, idx].toarray()
        Mf = M[idx][:, idx].toarray()
        hit = next((v for kf, mf, v, _ in cache if kf.shape == Kf.shape and np.array_equal(kf, Kf) and np.array_equal(mf, Mf)), None)
        if hit is None:
            mm = min(m, len(idx))
            lam, V = eigh(0.5 * (Kf + Kf.T), 0.5 * (Mf + Mf.T), subset_by_index=[0, mm - 1])
            V = V / np.abs(V).max(axis=0)
            f = np.sqrt(np.maximum(lam, 0.0)) / (2 * np.pi)
            cache.append((Kf, Mf, V, f))
            hit = V
            if fmode is None:
                fmode = np.full(m, np.inf)
                fmode[:mm] = f
        W[idx, :hit.shape[1]] = hit
    if fmode is None:
        fmode = np.full(m, np.inf)
    return W, fmode


def lrdecomp(G, device=0):
    """Unpivoted LU of G (mathematical indexing); raises like FullFormat._lu
    (cnld/compressed_formats.py:247-252) on a zero pivot."""
    n = G.shape[0]
    a = np.array(np.asarray(G).T, dtype=np.complex128, order='C', copy=True)
    info = C.c_uint(0)
    check(lib().cnld_b200_lrdecomp(device, n, ptr(a), n, C.byref(info)))
    if info.value != 0:
        raise RuntimeError('failed to calculate LU decomposition')
    return a.T


def lrdecomp_sym(G, device=0):
    """Unpivoted LU of the complex symmetric matrix given by G's lower triangle (half the flops)."""
    n = G.shape[0]
    a = np.array(np.asarray(G).T, dtype=np.complex128, order='C', copy=True)
    info = C.c_uint(0)
    check(lib().cnld_b200_lrdecomp_sym(device, n, ptr(a), n, C.byref(info)))
    if info.value != 0:
        raise RuntimeError('failed to calculate LU decomposition')
    return a.T


def lrsolve(LU, b, device=0):
    n = LU.shape[0]
    a = np.ascontiguousarray(np.asarray(LU, dtype=np.complex128).T)
    b = np.asarray(b, dtype=np.complex128)
    x = np.array(b.reshape(n, -1).T, dtype=np.complex128, order='C', copy=True)  # [nrhs][n]
    check(lib().cnld_b200_lrsolve(device, n, ptr(a), n, ptr(x), x.shape[0], n))
    return x.T.reshape(b.shape)


def zgemm(A, B, Cm=None, alpha=1.0, device=0):
    """alpha*A@B (+ Cm) through the DMMA kernel; mathematical indexing."""
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.complex128)
    m, k = A.shape
    n = B.shape[1]
    a, b = np.ascontiguousarray(A.T), np.ascontiguousarray(B.T)
    c = np.zeros((n, m), dtype=np.complex128) if Cm is None else np.array(np.asarray(Cm).T, dtype=np.complex128, order='C', copy=True)
    check(lib().cnld_b200_zgemm(device, m, n, k, alpha, ptr(a), m, ptr(b), k, 0.0 if Cm is None else 1.0,
                                ptr(c), m))
    return c.T


def bench_zgemm(m, n, k, reps=5, device=0):
    out = C.c_double(0)
    check(lib().cnld_b200_bench_zgemm(device, m, n, k, reps, C.byref(out)))
    return out.value


def translation_plan(vertices, triangles, allow=True):
    """Host-only: (info dict, work[nw, 4], copy[nc, 5]) of the symmetric-path assembly plan."""
    x = np.ascontiguousarray(vertices, dtype=np.float64)
    t = np.ascontiguousarray(triangles, dtype=np.uint32)
    counts = np.zeros(6, dtype=np.uint32)
    f = lib().cnld_b200_translation_plan
    f.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint,
                  C.c_void_p, C.c_uint]
    check(f(len(x), len(t), ptr(x), ptr(t), int(allow), ptr(counts), None, 0, None, 0))
    work = np.zeros((int(counts[3]), 4), dtype=np.uint32)
    copy = np.zeros((int(counts[4]), 5), dtype=np.uint32)
    check(f(len(x), len(t), ptr(x), ptr(t), int(allow), ptr(counts), ptr(work), len(work), ptr(copy), len(copy)))
    info = dict(periodic=bool(counts[0]), components=int(counts[1]), offset_classes=int(counts[2]),
                vertices_per_component=int(counts[5]))
    return info, work, copy


def dmma_flop_count():
    return int(lib().cnld_b200_dmma_flop_count())


def bench_dfma(reps=5, device=0):
    out = C.c_double(0)
    check(lib().cnld_b200_bench_dfma(device, reps, C.byref(out)))
    return out.value
