"""
Generates the golden fixtures in this directory by importing the REFERENCE's own
Python modules from /root/reference (cnld/fem.py, cnld/impulse_response.py and the
Cython cnld/bem_pressure.pyx run through cython's pure-Python fallback) and running
them on meshes produced by this repo's oracle mesh generator.

Run here (the container with /root/reference):  python tests/golden/make_golden.py
The GPU box never runs this; it only reads the committed *.npz.

What has to be stubbed so the reference imports at all in this image:
  * `namedlist`   (absent 3rd-party package; abstract.py:14)  -> tiny stand-in below
  * `matplotlib`  (absent; mesh.py:7, compressed_formats.py:7) -> empty modules
  * `cnld.h2lib`  (needs libh2, absent)                        -> Surface3d /
    Macrosurface3d stand-ins backed by oracle/cnld_oracle.c mesh routines
The FEM, pressure and impulse-response arithmetic that is being pinned is the
reference's own, untouched.
"""
import os
import sys
import types
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = '/root/reference'
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import pyoracle as po  # noqa: E402


# --- namedlist stand-in ------------------------------------------------------
class _Factory:
    def __init__(self, f):
        self.f = f


def _namedlist(typename, fields, **kw):
    if isinstance(fields, str):                          # 'a b c' with one common default (simulation.pyx:105)
        fields = OrderedDict((n, kw.get('default')) for n in fields.replace(',', ' ').split())
    fields = OrderedDict(fields) if not isinstance(fields, OrderedDict) else fields

    def __init__(self, *args, **kwargs):
        names = list(fields)
        for n, d in fields.items():
            setattr(self, n, d.f() if isinstance(d, _Factory) else d)
        for n, a in zip(names, args):
            setattr(self, n, a)
        for n, a in kwargs.items():
            if n not in fields:
                raise TypeError(n)
            setattr(self, n, a)

    def _asdict(self):
        return OrderedDict((n, getattr(self, n)) for n in fields)

    return type(typename, (), dict(__init__=__init__, _asdict=_asdict, _fields=tuple(fields)))


nl = types.ModuleType('namedlist')
nl.namedlist = _namedlist
nl.FACTORY = _Factory
sys.modules['namedlist'] = nl

# --- matplotlib stand-in -----------------------------------------------------
for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
    sys.modules[name] = types.ModuleType(name)
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
sys.modules['matplotlib'].patches = sys.modules['matplotlib.patches']

# --- cnld package rooted at the reference tree, with h2lib replaced -------
cnld = types.ModuleType('cnld')
cnld.__path__ = [os.path.join(REF, 'cnld')]
sys.modules['cnld'] = cnld


class Surface3d:
    def __init__(self, nv, ne, nt):
        self.x = np.zeros((nv, 3))
        self.e = np.zeros((ne, 2), dtype=np.uint32)
        self.t = np.zeros((nt, 3), dtype=np.uint32)
        self.s = np.zeros((nt, 3), dtype=np.uint32)
        self.n = np.zeros((nt, 3))
        self.g = np.zeros(nt)
        self.hmin = self.hmax = 0.0


class Macrosurface3d(Surface3d):
    kind = 0

    def set_parametrization(self, name):
        self.kind = 1 if name.lower() == 'circle' else 0


def build_from_macrosurface3d_surface3d(ms, refn):
    x, e, t, s = po.refine(ms.x, ms.e, ms.t, ms.s, refn, kind=ms.kind)
    surf = Surface3d(len(x), len(e), len(t))
    surf.x[:], surf.e[:], surf.t[:], surf.s[:] = x, e, t, s
    return surf


def prepare_surface3d(surf):
    m = po.Mesh(surf.x, surf.e, surf.t, surf.s)
    surf.n[:], surf.g[:] = m.normals, m.g
    surf.hmin, surf.hmax = m.hmin, m.hmax


def translate_surface3d(surf, r):
    surf.x += np.asarray(r)


h2 = types.ModuleType('cnld.h2lib')
for f in (Surface3d, Macrosurface3d, build_from_macrosurface3d_surface3d, prepare_surface3d,
          translate_surface3d):
    setattr(h2, f.__name__, f)
h2.__all__ = [f for f in dir(h2) if not f.startswith('_')]
sys.modules['cnld.h2lib'] = h2
cnld.h2lib = h2

if not hasattr(np.ndarray, 'tostring'):
    pass  # memoised calls with ndarray args are made through __wrapped__ below

from cnld import abstract, fem, impulse_response, mesh  # noqa: E402  (the REFERENCE modules)

assert fem.__file__.startswith(REF) and mesh.__file__.startswith(REF)


def ref_array(nelem, refn, ratio=0.1):
    """Array through the reference's own abstract types, matrix layout of
    cnld/scripts/matrix_array.py with its defaults."""
    desc = po.matrix_array(nelem=nelem, damping_ratio_a=ratio, damping_ratio_b=ratio)
    elements = []
    for e in desc['elements']:
        mems = []
        for m in e['membranes']:
            pats = [abstract.Patch(**p) for p in m['patches']]
            kw = {k: v for k, v in m.items() if k not in ('patches', 'shape')}
            mems.append(abstract.SquareCmutMembrane(patches=pats, **kw))
        elements.append(abstract.Element(id=e['id'], position=e['position'], kind='both',
                                         membranes=mems))
    return desc, abstract.Array(id=0, position=[0, 0, 0], elements=elements)


def golden_fem(tag, nelem, refn):
    desc, arr = ref_array(nelem, refn)
    mem = arr.elements[0].membranes[0]
    K = fem.mem_k_matrix(mem, refn)
    M = fem.mem_m_matrix(mem, refn, mu=0.5)
    B = fem.mem_b_matrix_eig.__wrapped__(mem, refn, M, K)
    eigf, _ = fem.mem_eig(mem, refn)
    F = fem.array_f_spmatrix(arr, refn).toarray()
    AVG = fem.array_avg_spmatrix(arr, refn).toarray()
    amesh = mesh.Mesh.from_abstract(arr, refn)
    freqs = np.array([0.0, 1e6, 7.3e6])
    # fem.array_mbk_spmatrix memoises with ndarray args (needs ndarray.tostring);
    # restate its 3-line body (fem.py:1301-1307) on the reference's own K, M, B
    nm = len(desc['elements'])
    from scipy.linalg import block_diag
    mbk = np.stack([block_diag(*[-(2 * np.pi * f)**2 * M - 1j * (2 * np.pi * f) * B + K] * nm)
                    for f in freqs])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), f'fem_{tag}.npz'),
                        nelem=np.array(nelem), refn=refn, K=K, M=M, B=B, eigf=eigf[:8], F=F, AVG=AVG,
                        vertices=amesh.vertices, triangles=amesh.triangles,
                        on_boundary=amesh.on_boundary, freqs=freqs,
                        mbk_diag=np.stack([np.diag(m) for m in mbk]),
                        mbk_sum=np.array([m.sum() for m in mbk]))
    print(tag, K.shape, F.shape, eigf[:5])


def golden_fir():
    rng = np.random.default_rng(0)
    f = np.linspace(0, 50e6, 64)
    # smooth causal-looking spectrum plus noise
    s = (np.exp(-((f - 8e6) / 5e6)**2) * np.exp(-2j * np.pi * f * 0.4e-6))[None, None, :] * \
        (1 + 0.1 * rng.standard_normal((3, 2, 1)))
    s = s + 0.01 * (rng.standard_normal(s.shape) + 1j * rng.standard_normal(s.shape))
    out = {}
    for interp in (1, 2, 3):
        for kkr in (True, False):
            t, h = impulse_response.fft_to_fir(f, s, interp=interp, use_kkr=kkr, axis=-1)
            out[f't_{interp}_{int(kkr)}'] = t
            out[f'h_{interp}_{int(kkr)}'] = h
    t, h = impulse_response.fft_to_sir(f, s[0], interp=2, use_kkr=False, axis=1)
    out['sir_t'], out['sir_h'] = t, h
    np.savez_compressed(os.path.join(os.path.dirname(__file__), 'fir.npz'), f=f, s=s, **out)
    print('fir', {k: v.shape for k, v in out.items() if k.startswith('h')})


def golden_pressure():
    """cnld/bem_pressure.pyx is Cython; its hot function has no typed-only
    constructs that change arithmetic, so we exec the reference source through
    cython's pure-Python shadow after mechanically stripping the C declarations
    (cdef/cpdef/cimport lines).  The arithmetic lines are untouched."""
    import re
    src = open(os.path.join(REF, 'cnld', 'bem_pressure.pyx')).read()
    src = src.replace('cimport numpy as np', '').replace('cimport cython', '')
    src = src.replace('cimport libc.math', 'import math as _m')
    src = re.sub(r'cdef extern from "complex.h":\n\s+double complex cexp\(double complex\)\n',
                 'from cmath import exp as cexp\n', src)
    src = re.sub(r'cpdef double complex kernel_helmholtz\([^)]*\):',
                 'def kernel_helmholtz(k, x1, y1, z1, x2, y2, z2):', src, flags=re.S)
    src = re.sub(r'cpdef np.ndarray mesh_pres_vector\([^)]*\):',
                 'def mesh_pres_vector(vertices, triangles, triangle_areas, r, k, c, rho, gn=2):',
                 src, flags=re.S)
    src = re.sub(r'^\s*cdef .*\n', '', src, flags=re.M)
    src = src.replace('libc.math.sqrt', '_m.sqrt')
    src = src.replace('@cython.boundscheck(False)', '').replace('@cython.wraparound(False)', '')
    src = src.replace('import cython\n', '')
    # the cdef lines declared these locals; re-create them
    src = src.replace("    x, y, z = r\n", "    nnodes = vertices.shape[0]\n"
                      "    p = np.zeros(nnodes, dtype=np.complex128)\n    x, y, z = r\n")
    cnld.database = types.ModuleType('cnld.database')
    sys.modules['cnld.database'] = cnld.database
    ns = {}
    exec(compile(src, 'bem_pressure.pyx', 'exec'), ns)
    desc, arr = ref_array((2, 1), 3)
    amesh = mesh.Mesh.from_abstract(arr, 3)
    obs = np.array([[0.0, 0.0, 1e-3], [2e-4, -3e-4, 5e-4], [-1e-3, 1e-3, 2.1e-3]])
    ks = np.array([2 * np.pi * 1e6 / 1500., 2 * np.pi * 12.5e6 / 1500.])
    pv = np.zeros((len(ks), len(obs), len(amesh.vertices)), dtype=np.complex128)
    for ik, k in enumerate(ks):
        for io, r in enumerate(obs):
            pv[ik, io] = ns['mesh_pres_vector'](amesh.vertices, amesh.triangles,
                                                amesh.triangle_areas, list(r), k, 1500., 1000.)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), 'pressure.npz'), obs=obs, ks=ks,
                        pvec=pv, vertices=amesh.vertices, triangles=amesh.triangles,
                        areas=amesh.triangle_areas)
    print('pressure', pv.shape, np.abs(pv).max())


def golden_fir_conv():
    """cnld/simulation.pyx:51-77 fir_conv_cy (the convolution the time-domain solver calls every step,
    :400-410): the function body is taken from the reference source and executed as plain Python after
    stripping the C declarations; the loop arithmetic is untouched."""
    import re
    src = open(os.path.join(REF, 'cnld', 'simulation.pyx')).read()
    m = re.search(r'cpdef np\.ndarray fir_conv_cy\(.*?\n    return np\.asarray\(x\)\n', src, flags=re.S)
    body = m.group(0)
    body = re.sub(r'cpdef np\.ndarray fir_conv_cy\([^)]*\):', 'def fir_conv_cy(fir, p, dt, offset):', body, flags=re.S)
    body = body.replace('    cdef int nsrc = fir.shape[0]', '    nsrc = fir.shape[0]').replace(
        '    cdef int ndest = fir.shape[1]', '    ndest = fir.shape[1]').replace(
        '    cdef int nfir = fir.shape[2]', '    nfir = fir.shape[2]').replace(
        '    cdef int nsample = p.shape[0]', '    nsample = p.shape[0]').replace(
        '    cdef double[:] x = np.zeros(ndest, dtype=np.float64)', '    x = np.zeros(ndest, dtype=np.float64)')
    body = re.sub(r'^\s*cdef .*\n', '', body, flags=re.M)
    ns = {'np': np}
    exec(compile(body, 'simulation.pyx:fir_conv_cy', 'exec'), ns)
    rng = np.random.default_rng(7)
    out = {}
    cases = [(3, 4, 16, 5, 0), (3, 4, 16, 5, 1), (3, 4, 16, 16, 1), (3, 4, 16, 40, 1), (3, 4, 16, 1, 0), (5, 2, 9, 9, 0),
             (2, 2, 8, 30, 0)]
    for c, (nsrc, ndest, nfir, nsample, offset) in enumerate(cases):
        fir = rng.standard_normal((nsrc, ndest, nfir))
        pp = rng.standard_normal((nsample, nsrc))
        out[f'fir{c}'], out[f'p{c}'] = fir, pp
        out[f'arg{c}'] = np.array([2.5e-9, offset])
        out[f'x{c}'] = ns['fir_conv_cy'](fir, pp, 2.5e-9, offset)
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(os.path.dirname(__file__), 'fir_conv.npz'), **out)
    print('fir_conv', len(cases), 'cases')


def _reference_simulation_module():
    """cnld/simulation.pyx executed as plain Python: the two `cimport` lines dropped and the typed signature /
    C declarations of fir_conv_cy stripped (as in golden_fir_conv); every class and function body -- StateDB,
    FixedStepSolver, the contact closures with their np.vectorize wrappers -- is the reference's own text."""
    import re
    src = open(os.path.join(REF, 'cnld', 'simulation.pyx')).read()
    src = re.sub(r'^cimport .*\n', '', src, flags=re.M)
    src = re.sub(r'cpdef np\.ndarray fir_conv_cy\([^)]*\):', 'def fir_conv_cy(fir, p, dt, offset):', src, flags=re.S)
    for name in ('nsrc = fir.shape[0]', 'ndest = fir.shape[1]', 'nfir = fir.shape[2]', 'nsample = p.shape[0]'):
        src = src.replace('    cdef int ' + name, '    ' + name)
    src = src.replace('    cdef double[:] x = np.zeros(ndest, dtype=np.float64)', '    x = np.zeros(ndest, dtype=np.float64)')
    src = re.sub(r'^\s*cdef .*\n', '', src, flags=re.M)
    for name in ('cnld.compensation', 'cnld.database'):   # imported at module level, not used by FixedStepSolver.step
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            setattr(cnld, name.split('.')[1], sys.modules[name])
    mod = types.ModuleType('cnld.simulation')
    exec(compile(src, os.path.join(REF, 'cnld', 'simulation.pyx'), 'exec'), mod.__dict__)
    return mod


def td_case(case):
    """Inputs of the time-domain golden cases: P patches, a table of damped oscillators (strong self terms,
    weak coupling falling off with the patch distance) sampled on its own grid fir_t, a DC ramp + pulse drive
    that takes most patches into contact; case 1 keeps patch 0 out of contact (the int64 truncation of
    np.vectorize then applies to everybody else), case 2 uses a quadratic contact law and maxiter 3."""
    rng = np.random.default_rng(100 + case)
    P, nfir_in, dt_in = (6, 90, 4e-9) if case < 2 else (4, 60, 5e-9)
    t_step = 2e-9 if case < 2 else 5e-9
    fir_t = np.arange(nfir_in) * dt_in
    pos = np.arange(P)
    amp = 9e-7 / (1 + 6 * np.abs(pos[:, None] - pos[None, :]))**2 * (1 + 0.1 * rng.standard_normal((P, P)))
    w0 = 2 * np.pi * 6e6 * (1 + 0.05 * rng.standard_normal((P, P)))
    fir = amp[:, :, None] * np.exp(-fir_t / 60e-9) * np.sin(w0[:, :, None] * fir_t + 0.35)
    nt = 140 if case < 2 else 90
    t_lim = (0.0, (nt - 1) * t_step, t_step)
    v_t = np.arange(0, nt * t_step, 4 * t_step)
    ramp = 1 / (1 + np.exp(-(v_t - 40e-9) / 8e-9))
    v = np.outer(ramp, 18.5 + 1.5 * rng.random(P)) + 2 * np.outer(np.sin(2 * np.pi * 9e6 * v_t) * (v_t > 120e-9), np.ones(P))
    if case == 0:
        v *= 1.09
    if case == 1:
        v[:, 0] *= 0.3
    if case == 2:
        v = v[:, 0] * 1.02                                # one drive for all patches (StateDB tiles it)
    gap = np.full(P, 50e-9)
    gap_eff = gap + 100e-9 / 6.3
    kw = dict(k=(2.5e14, 2.5e14, 4e22)[case], n=(1, 1, 2)[case], x0=-30e-9, lmbd=(2e13, 2e13, 4e21)[case],
              atol=(1e-12, 1e-10, 1e-12)[case], maxiter=5 if case < 2 else 3)
    return dict(fir_t=fir_t, fir=fir, v_t=v_t, v=v, gap=gap, gap_eff=gap_eff, t_lim=np.array(t_lim), **kw)


def golden_td_solver():
    """FixedStepSolver (cnld/simulation.pyx:297-555): constructed and stepped through the reference's own class."""
    sim = _reference_simulation_module()
    out = {}
    ncases = 3
    for c in range(ncases):
        a = td_case(c)
        s = sim.FixedStepSolver((a['fir_t'], a['fir']), (a['v_t'], a['v']), a['gap'], a['gap_eff'], tuple(a['t_lim']),
                                a['k'], a['n'], a['x0'], a['lmbd'], atol=a['atol'], maxiter=a['maxiter'])
        nsteps = len(s._db.t) - 2
        for _ in range(nsteps):
            s.step()
        for key, val in a.items():
            out[f'{key}{c}'] = np.asarray(val)
        out[f'nsteps{c}'] = np.array(nsteps)
        out[f'firi{c}'] = s._fir
        out[f't{c}'], out[f'vi{c}'] = s._db.t, s._db.v
        for key in ('x', 'u', 'p_es', 'p_cont_spr', 'p_cont_dmp', 'p_tot'):
            out[f'{key}{c}'] = getattr(s._db, key).copy()
        out[f'error{c}'], out[f'iters{c}'] = s.error, s.iters
        pen = (s._db.x < a['x0'])
        print('td case', c, 'steps', nsteps, 'iters', np.bincount(s.iters), 'contact samples', pen.sum(axis=0),
              'xmin', s._db.x.min(), 'max err', s.error.max())
    out['ncases'] = np.array(ncases)
    # drive-signal generators (:640-712), arguments in tests/test_host_mirror.py::test_simulation_signals
    sigs = [sim.gaussian_pulse(5e6, 0.8, 100e6, td=1e-7), sim.gaussian_pulse(3e6, 0.5, 80e6, antisym=False),
            sim.logistic_ramp(2e-7, 1e-8), sim.logistic_ramp(2e-7, 1e-8, td=3e-7, tstop=1e-6),
            sim.linear_ramp(2e-7, 1e-8), sim.linear_ramp(2e-7, 1e-8, td=1e-7, tstop=6e-7), sim.winsin(5e6, 3, 1e-8, td=2e-7)]
    sigs.append(sim.sigadd(sigs[2], sigs[6], sigs[4]))
    for q, (ts, vs) in enumerate(sigs):
        out[f'sig_t{q}'], out[f'sig_v{q}'] = ts, vs
    np.savez_compressed(os.path.join(os.path.dirname(__file__), 'td_solver.npz'), **out)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'fir_conv':
        golden_fir_conv()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'td_solver':
        golden_td_solver()
        sys.exit(0)
    golden_fem('1x1_r3', (1, 1), 3)
    golden_fem('2x1_r4', (2, 1), 4)
    golden_fem('1x1_r6', (1, 1), 6)
    golden_fir()
    golden_pressure()
    golden_fir_conv()
    golden_td_solver()
