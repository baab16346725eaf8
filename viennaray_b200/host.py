"""Host-side trace set-up in float32, mirroring the reference's host code that
surrounds the hot path: disk radius (rayUtil.hpp:99-101), bounding-box
adjustment (rayUtil.hpp:104-143), trace settings (rayUtil.hpp:145-202), the
random source's state (raySourceRandom.hpp:14-23) and its orthonormal basis
(rayUtil.hpp:287-321)."""
import numpy as np

from . import capi

POS_X, NEG_X, POS_Y, NEG_Y, POS_Z, NEG_Z = range(6)
F = np.float32

# sourceDir, boundaryDir1, boundaryDir2, minMax, posNeg
_TRACE_SETTINGS = {
    POS_X: (0, 1, 2, 1, -1), NEG_X: (0, 1, 2, 0, 1), POS_Y: (1, 0, 2, 1, -1),
    NEG_Y: (1, 0, 2, 0, 1), POS_Z: (2, 0, 1, 1, -1), NEG_Z: (2, 0, 1, 0, 1),
}


def disk_factor(D):
    return 0.5 * (1.7320508 if D == 3 else 1.41421356237) * (1 + 1e-5)


def disk_radius(grid_delta, D):
    return F(F(grid_delta) * disk_factor(D))


def trace_settings(source_dir):
    return _TRACE_SETTINGS[source_dir]


def geometry_bbox(points, D):
    """min/max over the first D axes; unused axes 0 (rayGeometryDisk.hpp:131-159)."""
    lo = np.zeros(3, F)
    hi = np.zeros(3, F)
    lo[:D] = points[:, :D].min(0)
    hi[:D] = points[:, :D].max(0)
    return lo, hi


def adjust_bbox(lo, hi, source_dir, offset, D):
    lo, hi = lo.astype(F).copy(), hi.astype(F).copy()
    offset = F(offset)
    if D == 2:
        lo[2] = F(lo[2] - offset)
        hi[2] = F(hi[2] + offset)
        if source_dir in (POS_Z, NEG_Z):
            raise ValueError("Ray source is set in z-direction for 2D geometry")
    axis, _, _, min_max, _ = trace_settings(source_dir)
    if min_max:
        hi[axis] = F(hi[axis] + F(2) * offset)
    else:
        lo[axis] = F(lo[axis] - F(2) * offset)
    return lo, hi


def orthonormal_basis(vec):
    u = np.asarray(vec, F)
    u = u * (F(1) / np.sqrt(F(u[0] * u[0] + u[1] * u[1]) + F(u[2] * u[2]), dtype=F))
    if abs(u[0]) > abs(u[2]):
        h = np.array([-u[1], u[0], 0], F)
    else:
        h = np.array([0, -u[2], u[1]], F)
    h = h * (F(1) / np.sqrt(F(h[0] * h[0] + h[1] * h[1]) + F(h[2] * h[2]), dtype=F))
    w = np.array([u[1] * h[2] - u[2] * h[1], u[2] * h[0] - u[0] * h[2],
                  u[0] * h[1] - u[1] * h[0]], F)
    return np.stack([u, h, w]).astype(F)


def create_source_grid(lo, hi, num_points, grid_delta, source_dir, D=3):
    """Origins of rayInternal::createSourceGrid (rayUtil.hpp:566-611): a regular grid of about
    num_points points on the source plane, 1e-4 inside the lateral box."""
    axis, first, second, min_max, _ = trace_settings(source_dir)
    eps = 1e-4
    len1, len2 = F(hi[first] - lo[first]), F(hi[second] - lo[second])
    n1, n2 = int(round(float(len1 / F(grid_delta)))), int(round(float(len2 / F(grid_delta))))
    ratio = n1 // n2
    if ratio == 0:  # the reference divides by this integer ratio (rayUtil.hpp:584-586)
        raise ValueError("createSourceGrid needs the first lateral extent >= the second")
    n1, n2 = int(np.sqrt(num_points * ratio)), int(np.sqrt(num_points / ratio))
    d1 = F((len1 - 2 * eps) / F(n1 - 1))
    d2 = F((len2 - 2 * eps) / F(n2 - 1))
    pts = []
    uu = F(lo[second] + eps)
    while uu <= hi[second] - eps:
        vv = F(lo[first] + eps)
        while vv <= hi[first] - eps:
            p = [F(0)] * 3
            p[axis] = F((hi if min_max else lo)[axis])
            p[second] = F(0) if D == 2 else uu
            p[first] = vv
            pts.append(p)
            vv = F(vv + d1)
        uu = F(uu + d2)
    return np.asarray(pts, F)


def source_desc(lo, hi, source_dir, primary_dir=None, use_grid=False):
    s = capi.SourceDesc()
    s.useGrid = 1 if use_grid else 0
    s.bboxMin[:] = [float(x) for x in lo]
    s.bboxMax[:] = [float(x) for x in hi]
    s.rayDir, s.firstDir, s.secondDir, s.minMax, pn = trace_settings(source_dir)
    s.posNeg = float(pn)
    s.useBasis = 0
    if primary_dir is not None:
        s.useBasis = 1
        s.basis[:] = [float(x) for x in orthonormal_basis(primary_dir).ravel()]
    return s


def config(num_rays, seed, idx_begin=0, idx_end=None, max_reflections=0xFFFFFFFF,
           max_boundary_hits=1000, wdist=False):
    return capi.Config(num_rays, idx_begin, num_rays if idx_end is None else idx_end, seed,
                       max_reflections, max_boundary_hits, capi.FLAG_WDIST if wdist else 0)


def triangle_normals(verts, tris):
    """Unit normals (v1-v0)x(v2-v0) * (1/|.|) in float32 (rayMesh.hpp:104-118)."""
    v = np.asarray(verts, F)
    t = np.asarray(tris, np.int64)
    a = v[t[:, 1]] - v[t[:, 0]]
    b = v[t[:, 2]] - v[t[:, 0]]
    c = np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                  a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1).astype(F)
    n2 = (c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]).astype(F) + (c[:, 2] * c[:, 2]).astype(F)
    inv = (F(1) / np.sqrt(n2, dtype=F)).astype(F)
    return (c * inv[:, None]).astype(F)
