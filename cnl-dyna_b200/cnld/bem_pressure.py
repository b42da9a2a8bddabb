"""
cnld.bem_pressure -- field pressure from the surface mesh (mirror of
/root/reference/cnld/bem_pressure.pyx: mesh_pres_vector :54-88, array_patch_pres_imp_resp
:91-117, press_resp :120-134), evaluated by libcnld_b200's pressure kernels.  Besides the
reference's one-point entry there is `field_pressure` for 10^5..10^6 observation points at once.
"""
import numpy as np

from . import _lib, impulse_response, mesh


def gauss_quadrature(n, type=1):
    """Triangle rules of bem_pressure.pyx:19-36 (n = 1, 2; the reference's n = 3 branch returns a
    1-tuple by mistake and is unusable, so it is not offered)."""
    if n == 1:
        return [[1 / 3, 1 / 3]], [1, ]
    if n == 2:
        if type == 1:
            return [[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]], [1 / 3, 1 / 3, 1 / 3]
        return [[0, 1 / 2], [1 / 2, 0], [1 / 2, 1 / 2]], [1 / 3, 1 / 3, 1 / 3]
    raise ValueError('gauss_quadrature: n must be 1 or 2')


def kernel_helmholtz(k, x1, y1, z1, x2, y2, z2):
    r = np.sqrt((x2 - x1)**2 + (y2 - y1)**2 + (z2 - z1)**2)
    return np.exp(-1j * k * r) / (4 * np.pi * r)


_ctx_cache = {}


def _context(vertices, triangles):
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    t = np.ascontiguousarray(triangles, dtype=np.uint32)
    key = (v.tobytes(), t.tobytes())
    if key not in _ctx_cache:
        if len(_ctx_cache) > 4:
            _ctx_cache.clear()
        _ctx_cache[key] = _lib.Context(v, t)
    return _ctx_cache[key]


def mesh_pres_vector(vertices, triangles, triangle_areas, r, k, c, rho, gn=2):
    """p_vec[n] with p(r) = p_vec . displacement  (bem_pressure.pyx:54-88); `triangle_areas` is
    accepted for call compatibility -- the areas come from the same vertices/triangles."""
    if gn != 2:
        raise ValueError('mesh_pres_vector: only gn=2 (the reference default) is supported')
    return _context(vertices, triangles).pres_vector(np.asarray(r, dtype=np.float64)[None, :], k, c, rho)[0]


def field_pressure(vertices, triangles, obs, ks, disp, c, rho):
    """p[ik, isrc, iobs] = p_vec(obs, k_ik) . disp[ik, isrc, :] for all observation points."""
    return _context(vertices, triangles).pressure(obs, ks, disp, c, rho)


def array_patch_pres_imp_resp(array, refn, db_file, r, c, rho, use_kkr=False, interp=2):
    """Pressure spatial impulse response of every patch at field point r (bem_pressure.pyx:91-117)."""
    from . import database
    freqs, pnfr, nodes = database.read_patch_to_node_freq_resp(db_file)
    amesh = mesh.Mesh.from_abstract(array, refn)
    ks = 2 * np.pi * np.asarray(freqs) / c
    disp = np.ascontiguousarray(np.transpose(pnfr, (2, 0, 1)))          # [freq, patch, node]
    sfr = field_pressure(amesh.vertices, amesh.triangles, np.asarray(r, dtype=np.float64)[None, :], ks, disp,
                         c, rho)[:, :, 0].T                               # [patch, freq]
    return impulse_response.fft_to_sir(freqs, sfr, interp=interp, axis=1, use_kkr=use_kkr)


def field_patch_pres_imp_resp(vertices, triangles, freqs, disp, points, c, rho, use_kkr=False, interp=2):
    """array_patch_pres_imp_resp for MANY field points at once (BASELINE.json configs[4]): the spatial impulse
    response of every patch at every point, sir[patch, point, nfft].  disp[freq, patch, node] are the node
    displacements of the sweep; per frequency the whole field is one pass of the many-source pressure path
    (observation points resident on the GPU), the spectra then go through the device impulse response."""
    from . import _lib
    ctx = _context(vertices, triangles)
    freqs = np.asarray(freqs, dtype=np.float64)
    disp = np.asarray(disp, dtype=np.complex128)
    nf, P, _ = disp.shape
    points = np.atleast_2d(np.asarray(points, dtype=np.float64))
    pf = _lib.PressureField(ctx, points, P)
    try:
        sfr = np.empty((P, len(points), nf), dtype=np.complex128)
        buf = np.empty((P, len(points)), dtype=np.complex128)
        for i, f in enumerate(freqs):
            sfr[:, :, i] = pf.run(2 * np.pi * f / c, disp[i], c, rho, out=buf)
    finally:
        pf.close()
    return impulse_response.fft_to_sir(freqs, sfr, interp=interp, axis=-1, use_kkr=use_kkr)


def press_resp(array, refn, db_file, pes, r, c, rho, use_kkr=False, interp=2):
    """Field pressure for patch pressures pes[:, patch] (bem_pressure.pyx:120-134)."""
    sir_t, sir = array_patch_pres_imp_resp(array, refn, db_file, r, c, rho, interp=interp, use_kkr=use_kkr)
    sir = sir.T
    dt = sir_t[1] - sir_t[0]
    return np.sum([np.convolve(pes[:, i], sir[:, i]) * dt for i in range(pes.shape[1])], axis=0)
