"""Route A of INTEGRATION.md on the CPU: the REFERENCE's own Cython binding layer
(/root/reference/cnld/h2lib/*.pyx) cythonized against include/h2lib_compat/, linked with -lcnld_b200,
imported as a package (star imports + init_h2lib, like cnld/h2lib/__init__.py), and every C symbol it
binds resolved by libcnld_b200.so.  No compute calls here (no GPU); the recipe runs in
tests/test_gpu_route_a.py."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', '_route_a')
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import build_route_a  # noqa: E402


@pytest.fixture(scope='module')
def route_a():
    if os.path.isdir('/root/reference/cnld/h2lib'):
        assert build_route_a.build('/root/reference'), 'route A extensions failed to build'
    if not build_route_a.built():
        pytest.skip('reference sources absent and no prebuilt route-A extensions')
    return OUT


def test_every_bound_symbol_is_exported(route_a):
    need = set()
    for so in glob.glob(os.path.join(route_a, 'cnld', 'h2lib', '*.so')):
        out = subprocess.run(['nm', '-D', '--undefined-only', so], capture_output=True, text=True, check=True).stdout
        need |= {ln.split()[-1].split('@')[0] for ln in out.splitlines() if ln.strip()}
    have_out = subprocess.run(['nm', '-D', '--defined-only', os.path.join(ROOT, 'cnl-dyna_b200', 'libcnld_b200.so')],
                              capture_output=True, text=True, check=True).stdout
    have = {ln.split()[-1] for ln in have_out.splitlines() if ln.strip()}
    h2 = {s for s in need if not s.startswith(('Py', '_Py', '__', '_ITM')) and s not in
          {'memcpy', 'memset', 'memcmp', 'memmove', 'strlen', 'strcmp', 'strncmp', 'strrchr', 'malloc', 'free', 'realloc',
           'calloc', 'cabs', 'carg', 'cexp', 'conj', 'creal', 'cimag', 'sqrt', 'fabs', 'cos', 'sin', 'acos', 'atan2',
           'abort', 'vsnprintf', 'snprintf'}}
    missing = sorted(s for s in h2 if s not in have)
    assert not missing, f'symbols bound by the reference Cython layer but not exported: {missing}'
    for s in ('lrdecomp_hmatrix', 'add_hmatrix', 'addmul_hmatrix', 'choldecomp_amatrix', 'solve_cg_amatrix_avector',
              'norm2diff_amatrix_hmatrix', 'new_dlp_helmholtz_bem3d', 'assemble_bem3d_amatrix', 'lrdecomp_amatrix'):
        assert s in h2 and s in have


def test_package_imports(route_a):
    env = dict(os.environ, PYTHONPATH=route_a)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'route_a', 'recipe.py'), '/dev/null', 'import-only'],
                       capture_output=True, text=True, env=env, cwd='/tmp')
    assert r.returncode == 0, r.stderr[-2000:]
    assert 'route A import ok' in r.stdout
