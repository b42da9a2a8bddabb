"""
cnld.mesh -- triangular surface meshes (mirror of /root/reference/cnld/mesh.py).

Same entry points (`Mesh.from_geometry / from_abstract`, `geometry_square`, `geometry_circle`,
`square`, `circle`, `matrix_array`) and attributes; the H2Lib containers behind them are
libcnld_b200's (cnld.h2lib).  Node ordering is defined by
build_from_macrosurface3d_surface3d and by concatenating membranes in element/membrane order
(mesh.py:182-222), which fixes the DOF ordering of every matrix on the path.
"""
import numpy as np

from . import abstract, util
from .h2lib import (Macrosurface3d, Surface3d, build_from_macrosurface3d_surface3d, merge_surface3d,
                    prepare_surface3d, refine_red_surface3d, translate_surface3d)

eps = np.finfo(np.float64).eps


class Mesh:
    """2D triangular mesh on H2Lib-style containers (mesh.py:13-179)."""

    def __init__(self):
        self._surface = None
        self.on_boundary = None

    @classmethod
    def from_surface3d(cls, surf):
        obj = cls()
        obj._surface = surf
        obj._update_properties()
        return obj

    @classmethod
    def from_macrosurface3d(cls, ms, center=(0, 0, 0), refn=2):
        assert refn > 1
        obj = cls.from_surface3d(build_from_macrosurface3d_surface3d(ms, refn))
        obj.translate(center)
        return obj

    @classmethod
    def from_geometry(cls, vertices, edges, triangles, triangle_edges):
        surf = Surface3d(len(vertices), len(edges), len(triangles))
        surf.x[:] = vertices
        surf.e[:] = edges
        surf.t[:] = triangles
        surf.s[:] = triangle_edges
        return cls.from_surface3d(surf)

    @classmethod
    def from_abstract(cls, array, refn=1, **kwargs):
        return _from_abstract(cls, array, refn, **kwargs)

    def __add__(self, other):
        a, b = self._surface, other._surface
        if a is None and b is None:
            return Mesh()
        if a is None:
            return Mesh.from_surface3d(b)
        if b is None:
            return Mesh.from_surface3d(a)
        return Mesh.from_surface3d(merge_surface3d(a, b))

    vertices = property(lambda self: np.asarray(self._surface.x))
    edges = property(lambda self: np.asarray(self._surface.e))
    triangles = property(lambda self: np.asarray(self._surface.t))
    triangle_edges = property(lambda self: np.asarray(self._surface.s))
    normals = property(lambda self: np.asarray(self._surface.n))
    g = property(lambda self: np.asarray(self._surface.g))
    triangle_areas = property(lambda self: np.asarray(self._surface.g) / 2)
    hmin = property(lambda self: self._surface.hmin)
    hmax = property(lambda self: self._surface.hmax)
    nvertices = property(lambda self: len(self.vertices))
    nedges = property(lambda self: len(self.edges))
    ntriangles = property(lambda self: len(self.triangles))
    surface3d = property(lambda self: self._surface)

    def _update_properties(self):
        prepare_surface3d(self._surface)

    def refine(self, n=1):
        for _ in range(n):
            self._surface = refine_red_surface3d(self._surface)
            self._update_properties()

    def translate(self, r):
        translate_surface3d(self._surface, np.array(r, dtype=np.float64))


def _refine(v, e, t, s, refn, parametrization):
    ms = Macrosurface3d(len(v), len(e), len(t))
    ms.x[:] = v
    ms.e[:] = e
    ms.t[:] = t
    ms.s[:] = s
    ms.set_parametrization(parametrization)
    surf = build_from_macrosurface3d_surface3d(ms, refn)
    return (np.array(surf.x, copy=True), np.array(surf.e, copy=True), np.array(surf.t, copy=True),
            np.array(surf.s, copy=True))


@util.memoize
def geometry_square(xl, yl, refn=1, type=1):
    """Square membrane: 5 vertices / 8 edges / 4 macro triangles around the centre (type 1) or
    4 / 5 / 2 (type 2), refined by the macro-surface splitter (mesh.py:306-384)."""
    hx, hy = xl / 2, yl / 2
    if type == 1:
        v = np.array([[-hx, -hy, 0.], [hx, -hy, 0.], [hx, hy, 0.], [-hx, hy, 0.], [0., 0., 0.]])
        e = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [0, 4], [1, 4], [2, 4], [3, 4]], dtype=np.uint32)
        t = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], dtype=np.uint32)
        s = np.array([[5, 4, 0], [6, 5, 1], [7, 6, 2], [4, 7, 3]], dtype=np.uint32)
    elif type == 2:
        v = np.array([[-hx, -hy, 0.], [hx, -hy, 0.], [hx, hy, 0.], [-hx, hy, 0.]])
        e = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [1, 3]], dtype=np.uint32)
        t = np.array([[0, 1, 3], [1, 2, 3]], dtype=np.uint32)
        s = np.array([[4, 3, 0], [2, 4, 1]], dtype=np.uint32)
    else:
        raise ValueError('incorrect type')
    if refn > 1:
        v, e, t, s = _refine(v, e, t, s, refn, 'square')
    return v, e, t, s


@util.memoize
def geometry_circle(rl, n=4, refn=1):
    """Disc from n macro triangles around the centre mapped with the circle parametrization
    (mesh.py:444-502)."""
    ang = 2 * np.pi / n * np.arange(n) - np.pi
    v = np.zeros((n + 1, 3))
    v[:n, 0], v[:n, 1] = rl * np.cos(ang), rl * np.sin(ang)
    v[np.isclose(v, 0)] = 0.0
    idx = np.arange(n)
    e = np.zeros((2 * n, 2), dtype=np.uint32)
    e[:n] = np.c_[idx, (idx + 1) % n]
    e[n:] = np.c_[idx, np.full(n, n)]
    t = np.c_[idx, (idx + 1) % n, np.full(n, n)].astype(np.uint32)
    s = np.c_[(idx + 1) % n + n, idx + n, idx].astype(np.uint32)
    if refn > 1:
        v, e, t, s = _refine(v, e, t, s, refn, 'circle')
    return v, e, t, s


def _square_boundary(x, y, cx, cy, xl, yl):
    return ((np.abs(x - cx + xl / 2) <= 2 * eps) | (np.abs(x - cx - xl / 2) <= 2 * eps) |
            (np.abs(y - cy + yl / 2) <= 2 * eps) | (np.abs(y - cy - yl / 2) <= 2 * eps))


def square(xl, yl, refn=1, type=1, center=(0, 0, 0)):
    v, e, t, s = geometry_square(xl, yl, refn=refn, type=type)
    v = v + np.array(center)
    m = Mesh.from_geometry(v, e, t, s)
    m.on_boundary = _square_boundary(m.vertices[:, 0], m.vertices[:, 1], center[0], center[1], xl, yl)
    return m


def circle(rl, refn=1, center=(0, 0, 0)):
    v, e, t, s = geometry_circle(rl, n=4, refn=refn)
    v = v + np.array(center)
    m = Mesh.from_geometry(v, e, t, s)
    r = np.sqrt(((m.vertices - np.array(center))**2).sum(axis=1))
    m.on_boundary = np.abs(r - rl) <= 2 * eps
    return m


def _from_abstract(cls, array, refn=1, **kwargs):
    """Whole-array mesh: membranes concatenated in element/membrane order with contiguous
    per-membrane numbering; membrane-boundary nodes flagged (mesh.py:182-302)."""
    parts = []
    vidx = eidx = 0
    for elem in array.elements:
        for mem in elem.membranes:
            if isinstance(mem, abstract.SquareCmutMembrane):
                v, e, t, s = geometry_square(mem.length_x, mem.length_y, refn=refn)
            elif isinstance(mem, abstract.CircularCmutMembrane):
                v, e, t, s = geometry_circle(mem.radius, n=4, refn=refn)
            else:
                raise TypeError
            parts.append((v + np.array(mem.position), e + np.uint32(vidx), t + np.uint32(vidx),
                          s + np.uint32(eidx)))
            vidx += len(v)
            eidx += len(e)
    mesh = cls.from_geometry(*[np.concatenate([p[i] for p in parts], axis=0) for i in range(4)])
    nverts = len(mesh.vertices)
    mesh.membrane_ids = np.full(nverts, np.nan)
    mesh.element_ids = np.full(nverts, np.nan)
    mesh.on_boundary = np.zeros(nverts, dtype=bool)
    x, y, _ = mesh.vertices.T
    for elem in array.elements:
        for mem in elem.membranes:
            mx, my, _ = mem.position
            if isinstance(mem, abstract.SquareCmutMembrane):
                lx, ly = mem.length_x, mem.length_y
                inside = ((x >= mx - lx / 2 - 2 * eps) & (x <= mx + lx / 2 + 2 * eps) &
                          (y >= my - ly / 2 - 2 * eps) & (y <= my + ly / 2 + 2 * eps))
                mesh.on_boundary[inside] = _square_boundary(x[inside], y[inside], mx, my, lx, ly)
            else:
                r = np.sqrt((x - mx)**2 + (y - my)**2)
                inside = r <= mem.radius + 2 * eps
                mesh.on_boundary[inside] = (r[inside] <= mem.radius + 2 * eps) & (r[inside] >= mem.radius - 2 * eps)
            mesh.membrane_ids[inside] = mem.id
            mesh.element_ids[inside] = elem.id
    return mesh


def matrix_array(nx, ny, pitchx, pitchy, xl, yl, refn=1, **kwargs):
    """Convenience mesh of an nx x ny grid of square membranes (mesh.py:722-775)."""
    from .arrays import matrix_array as _abstract_matrix
    arr = _abstract_matrix(nelem=(nx, ny), elempitch=(pitchx, pitchy), length=(xl, yl))
    return Mesh.from_abstract(arr, refn)
