"""TEST INFRASTRUCTURE (oracle/): ctypes bindings for

  * ``libvr_oracle.so``  -- the plain-C restatement (oracle/vr_oracle.c), and
  * ``libvr_ref.so``     -- the reference's unmodified TraceKernel behind
                            oracle/ref_driver.cpp (+ substitute intersector).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(`viennaray_b200`) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
FLUX_SCALE = float(2**30)

_vp = C.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def build(reference_root="/root/reference"):
    """Compile the oracle (and oracle/_ref when the reference tree exists)."""
    targets = ["oracle"]
    if os.path.exists(os.path.join(reference_root, "include/viennaray/rayTraceKernel.hpp")):
        targets.append("ref")
    subprocess.check_call(["make", "-C", _HERE, "REF=" + reference_root] + targets,
                          stdout=subprocess.DEVNULL)


class Particle(C.Structure):
    _fields_ = [("kind", C.c_int), ("sticking", C.c_float), ("sourcePower", C.c_float),
                ("coneMinAngle", C.c_float), ("meanFreePath", C.c_float),
                ("stickingByMaterial", _vp), ("numMaterials", C.c_int)]

    def set_sticking_by_material(self, table):
        """sticking[materialId] of the hit primitive; None returns to the constant."""
        if table is None:
            self._table = None
            self.stickingByMaterial, self.numMaterials = None, 0
            return self
        self._table = np.ascontiguousarray(table, np.float32)  # kept alive with the struct
        self.stickingByMaterial = self._table.ctypes.data
        self.numMaterials = len(self._table)
        return self


class Config(C.Structure):
    _fields_ = [("numRays", C.c_uint64), ("seed", C.c_uint32), ("stream", C.c_uint32),
                ("maxReflections", C.c_uint32), ("maxBoundaryHits", C.c_uint32),
                ("usePrimaryDir", C.c_int), ("primaryDir", C.c_float * 3),
                ("useWdist", C.c_int)]


class Info(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("numRays", "totalTraces", "nonGeoHits", "geoHits",
                                          "particleHits", "boundaryHits", "reflections",
                                          "raysTerminated")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


DIFFUSE, SPECULAR, CONED_COSINE = 0, 1, 2
REFLECTIVE, PERIODIC, IGNORE = 0, 1, 2
POS_X, NEG_X, POS_Y, NEG_Y, POS_Z, NEG_Z = range(6)

_oracle = None
_ref = {}


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(_REF, "libvr_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.vro_scene_create.restype = _vp
        L.vro_scene_create.argtypes = [C.c_int]
        L.vro_scene_destroy.argtypes = [_vp]
        L.vro_scene_set_disks.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_float]
        L.vro_scene_set_triangles.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_uint32]
        L.vro_scene_setup.argtypes = [_vp, C.c_int, _vp, C.c_float]
        L.vro_scene_set_material_ids.argtypes = [_vp, _vp]
        L.vro_scene_set_source_grid.argtypes = [_vp, _vp, C.c_uint32]
        L.vro_scene_bbox.argtypes = [_vp, _vp]
        L.vro_scene_num_prims.restype = C.c_uint32
        L.vro_scene_num_prims.argtypes = [_vp]
        L.vro_scene_neighbors.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp)]
        L.vro_scene_normals.restype = _vp
        L.vro_scene_normals.argtypes = [_vp]
        L.vro_trace.argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp]
        L.vro_source_rays.argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint32, _vp]
        L.vro_intersect.argtypes = [_vp, _vp, C.c_uint32, _vp, _vp, _vp, _vp]
        L.vro_neighbor_hits.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp]
        L.vro_boundary_process_hit.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_uint32, C.c_float]
        L.vro_normalize_flux_source.argtypes = [_vp, _vp, C.c_uint64, _vp]
        L.vro_smooth_flux.argtypes = [_vp, _vp]
        L.vro_normalize_flux_max.argtypes = [_vp, _vp, _vp]
        L.vro_smooth_flux_k.argtypes = [_vp, C.c_int, _vp]
        L.vro_philox4x32.argtypes = [C.c_uint32] * 6 + [_vp]
        L.vro_math_sincos2pi.argtypes = [_vp, C.c_uint32, _vp, _vp]
        L.vro_math_pow.argtypes = [_vp, C.c_float, C.c_uint32, _vp]
        L.vro_math_acos.argtypes = [_vp, C.c_uint32, _vp]
        L.vro_reflect.argtypes = [C.c_int, C.c_int, _vp, _vp, C.c_float, C.c_uint32, C.c_uint64,
                                  C.c_uint32, _vp]
        _oracle = L
    return _oracle


def oracle_set_threads(n):
    """OpenMP threads of the C oracle's traces (n <= 0: query); returns the count in effect."""
    L = oracle_lib()
    L.vro_set_threads.argtypes = [C.c_int]
    L.vro_set_threads.restype = C.c_int
    return int(L.vro_set_threads(int(n)))


def have_ref():
    return os.path.exists(os.path.join(_REF, "libvr_ref.so"))


def have_ref_wdist():
    return os.path.exists(os.path.join(_REF, "libvr_ref_wdist.so"))


def ref_lib(wdist=False):
    """The reference's unmodified kernel; wdist=True: the build with -DVIENNARAY_USE_WDIST."""
    if wdist not in _ref:
        L = C.CDLL(os.path.join(_REF, "libvr_ref_wdist.so" if wdist else "libvr_ref.so"))
        L.ref_trace_disk.argtypes = [C.c_int, _vp, _vp, C.c_uint32, C.c_float, _vp, C.c_int,
                                     C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint64,
                                     C.c_uint64, C.c_uint, C.c_uint, _vp, C.c_int, C.c_int, _vp,
                                     _vp, _vp]
        L.ref_trace_triangle.argtypes = [_vp, C.c_uint32, _vp, C.c_uint32, C.c_float, _vp,
                                         C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_uint64, C.c_uint64, C.c_uint, C.c_uint, C.c_int,
                                         _vp, _vp, _vp]
        L.ref_neighbors.restype = C.c_uint32
        L.ref_neighbors.argtypes = [C.c_int, _vp, C.c_uint32, C.c_float, C.c_uint32, _vp, _vp]
        L.ref_disk_areas.argtypes = [C.c_int, _vp, _vp, C.c_uint32, C.c_float, _vp, C.c_int, _vp]
        L.ref_intersect_disks.argtypes = [_vp, _vp, C.c_uint32, C.c_float, _vp, _vp, C.c_int,
                                          _vp, C.c_uint32, _vp, _vp, _vp, _vp]
        L.ref_boundary_process_hit.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp,
                                               C.c_uint32, C.c_float]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_set_options.restype = None
        L.ref_set_options.argtypes = [C.c_float, _vp, C.c_int, _vp, C.c_uint32]
        assert L.ref_uses_wdist() == (1 if wdist else 0)
        _ref[wdist] = L
    return _ref[wdist]


class _RefOptions:
    """Mean free path / sticking table / material IDs for the reference calls inside a
    ``with`` block (ref_driver.cpp: ref_set_options)."""

    def __init__(self, L, mean_free_path, sticking_by_material, material_ids):
        self.L = L
        self.mfp = float(mean_free_path or 0.0)
        self.tab = None if sticking_by_material is None else \
            np.ascontiguousarray(sticking_by_material, np.float32)
        self.ids = None if material_ids is None else np.ascontiguousarray(material_ids, np.int32)

    def __enter__(self):
        self.L.ref_set_options(self.mfp, _p(self.tab), 0 if self.tab is None else len(self.tab),
                               _p(self.ids), 0 if self.ids is None else len(self.ids))

    def __exit__(self, *a):
        self.L.ref_set_options(0.0, None, 0, None, 0)


def disk_factor(D):
    # rayUtil.hpp:99-101
    return 0.5 * (1.7320508 if D == 3 else 1.41421356237) * (1 + 1e-5)


class OracleScene:
    """A scene + trace set-up held by the C oracle."""

    def __init__(self, D):
        self.L = oracle_lib()
        self.D = D
        self.h = self.L.vro_scene_create(D)
        self.n = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.vro_scene_destroy(self.h)
            self.h = None

    def set_disks(self, points, normals, radius):
        points = np.ascontiguousarray(points, np.float32)
        normals = np.ascontiguousarray(normals, np.float32)
        self.n = len(points)
        self.geo = "disk"
        self.L.vro_scene_set_disks(self.h, _p(points), _p(normals), self.n, np.float32(radius))

    def set_triangles(self, verts, tris):
        verts = np.ascontiguousarray(verts, np.float32)
        tris = np.ascontiguousarray(tris, np.uint32)
        self.n = len(tris)
        self.geo = "triangle"
        self.L.vro_scene_set_triangles(self.h, _p(verts), len(verts), _p(tris), self.n)

    def set_material_ids(self, ids):
        if ids is None:
            self.L.vro_scene_set_material_ids(self.h, None)
            return
        ids = np.ascontiguousarray(ids, np.int32)
        assert len(ids) == self.n
        self.L.vro_scene_set_material_ids(self.h, _p(ids))

    def setup(self, source_dir, bc, source_offset):
        bc = (C.c_int * 3)(*(list(bc) + [IGNORE] * 3)[:3])
        rc = self.L.vro_scene_setup(self.h, source_dir, bc, np.float32(source_offset))
        if rc:
            raise ValueError("invalid trace set-up")

    def set_source_grid(self, points):
        """Grid source (raySourceGrid.hpp); None / empty returns to the random source."""
        if points is None or len(points) == 0:
            self.L.vro_scene_set_source_grid(self.h, None, 0)
            return
        points = np.ascontiguousarray(points, np.float32)
        self.L.vro_scene_set_source_grid(self.h, _p(points), len(points))

    def bbox(self):
        out = np.zeros(6, np.float32)
        self.L.vro_scene_bbox(self.h, _p(out))
        return out.reshape(2, 3)

    def normals(self):
        ptr = self.L.vro_scene_normals(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (self.n, 3)).copy()

    def neighbors(self):
        off, idx = _vp(), _vp()
        self.L.vro_scene_neighbors(self.h, C.byref(off), C.byref(idx))
        offsets = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint32)), (self.n + 1,)).copy()
        total = int(offsets[-1])
        if total == 0:
            return offsets, np.zeros(0, np.uint32)
        indices = np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_uint32)), (total,)).copy()
        return offsets, indices

    @staticmethod
    def config(num_rays, seed, stream=0, max_reflections=0xFFFFFFFF, max_boundary_hits=1000,
               primary_dir=None, wdist=False):
        c = Config(num_rays, seed, stream, max_reflections, max_boundary_hits,
                   0 if primary_dir is None else 1, (C.c_float * 3)(0, 0, 0), 1 if wdist else 0)
        if primary_dir is not None:
            c.primaryDir[:] = [float(x) for x in primary_dir]
        return c

    def trace(self, particle, cfg, idx_begin=0, idx_end=None, flux=None):
        if idx_end is None:
            idx_end = cfg.numRays
        if flux is None:
            flux = np.zeros(self.n, np.uint64)
        info = Info()
        self.L.vro_trace(self.h, C.byref(particle), C.byref(cfg), idx_begin, idx_end, _p(flux),
                         C.byref(info))
        return flux, info

    def source_rays(self, particle, cfg, idx_begin, m):
        rays = np.zeros((m, 6), np.float32)
        self.L.vro_source_rays(self.h, C.byref(particle), C.byref(cfg), idx_begin, m, _p(rays))
        return rays

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32)
        m = len(rays)
        geom = np.zeros(m, np.uint32)
        prim = np.zeros(m, np.uint32)
        t = np.zeros(m, np.float32)
        ng = np.zeros((m, 3), np.float32)
        self.L.vro_intersect(self.h, _p(rays), m, _p(geom), _p(prim), _p(t), _p(ng))
        return geom, prim, t, ng

    def neighbor_hits(self, rays, prim, cap=16):
        rays = np.ascontiguousarray(rays, np.float32)
        prim = np.ascontiguousarray(prim, np.uint32)
        m = len(rays)
        count = np.zeros(m, np.uint32)
        out = np.full((m, cap), 0xFFFFFFFF, np.uint32)
        self.L.vro_neighbor_hits(self.h, _p(rays), _p(prim), m, cap, _p(count), _p(out))
        return count, out

    def boundary_process_hit(self, org, ray_dir3, direction, ng, prim_id, t):
        org = np.array(org, np.float32)
        rd = np.array(ray_dir3, np.float32)
        d = np.array(direction, np.float32)
        ng = np.array(ng, np.float32)
        reflect = self.L.vro_boundary_process_hit(self.h, _p(org), _p(rd), _p(d), _p(ng), prim_id,
                                                  np.float32(t))
        return reflect, org, rd, d

    def normalize_flux_source(self, flux, areas, num_rays):
        flux = np.ascontiguousarray(flux, np.float32).copy()
        areas = np.ascontiguousarray(areas, np.float32)
        self.L.vro_normalize_flux_source(self.h, _p(areas), num_rays, _p(flux))
        return flux

    def smooth_flux(self, flux, k=1):
        flux = np.ascontiguousarray(flux, np.float32).copy()
        self.L.vro_smooth_flux_k(self.h, int(k), _p(flux))
        return flux

    def normalize_flux_max(self, flux, areas):
        flux = np.ascontiguousarray(flux, np.float32).copy()
        areas = np.ascontiguousarray(areas, np.float32)
        self.L.vro_normalize_flux_max(self.h, _p(areas), _p(flux))
        return flux


def ref_trace_disk(D, points, normals, grid_delta, bc, source_dir, kind, sticking, source_power=1.0,
                   cone_min_angle=0.0, rays_per_point=0, rays_fixed=0, seed=12345, runs=1,
                   primary_dir=None, normalize=False, smooth=0, threads=None, wdist=False,
                   mean_free_path=0.0, sticking_by_material=None, material_ids=None):
    L = ref_lib(wdist)
    if threads:
        L.ref_set_threads(threads)
    points = np.ascontiguousarray(points, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    n = len(points)
    flux = np.zeros((runs, n), np.float32)
    info = np.zeros(8, np.uint64)
    sec = C.c_double()
    bc = (C.c_int * 3)(*(list(bc) + [IGNORE] * 3)[:3])
    pd = None if primary_dir is None else np.array(primary_dir, np.float32)
    with _RefOptions(L, mean_free_path, sticking_by_material, material_ids):
        rc = L.ref_trace_disk(D, _p(points), _p(normals), n, grid_delta, bc, source_dir, kind,
                              sticking, source_power, cone_min_angle, rays_per_point, rays_fixed,
                              seed, runs, _p(pd), int(normalize), smooth, _p(flux), _p(info),
                              C.byref(sec))
    assert rc == 0
    return flux, info, sec.value


def ref_trace_triangle(verts, tris, grid_delta, bc, source_dir, kind, sticking, source_power=1.0,
                       cone_min_angle=0.0, rays_per_point=0, rays_fixed=0, seed=12345, runs=1,
                       normalize=False, threads=None, mean_free_path=0.0,
                       sticking_by_material=None, material_ids=None):
    L = ref_lib()
    if threads:
        L.ref_set_threads(threads)
    verts = np.ascontiguousarray(verts, np.float32)
    tris = np.ascontiguousarray(tris, np.uint32)
    n = len(tris)
    flux = np.zeros((runs, n), np.float32)
    info = np.zeros(8, np.uint64)
    sec = C.c_double()
    bc = (C.c_int * 3)(*(list(bc) + [IGNORE] * 3)[:3])
    with _RefOptions(L, mean_free_path, sticking_by_material, material_ids):
        rc = L.ref_trace_triangle(_p(verts), len(verts), _p(tris), n, grid_delta, bc, source_dir,
                                  kind, sticking, source_power, cone_min_angle, rays_per_point,
                                  rays_fixed, seed, runs, int(normalize), _p(flux), _p(info),
                                  C.byref(sec))
    assert rc == 0
    return flux, info, sec.value


def ref_post_disk(D, points, normals, grid_delta, bc, source_dir, flux, norm=0, smooth=0,
                  rays_fixed=1000):
    """The reference's normalizeFlux (norm: 1 SOURCE, 2 MAX) / smoothFlux(smooth) of `flux`."""
    L = ref_lib()
    points = np.ascontiguousarray(points, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    out = np.ascontiguousarray(flux, np.float32).copy()
    bc = (C.c_int * 3)(*(list(bc) + [IGNORE] * 3)[:3])
    L.ref_post_disk.argtypes = [C.c_int, _vp, _vp, C.c_uint32, C.c_float, C.POINTER(C.c_int),
                                C.c_int, C.c_uint64, C.c_int, C.c_int, _vp]
    rc = L.ref_post_disk(D, _p(points), _p(normals), len(points), grid_delta, bc, source_dir,
                         rays_fixed, norm, smooth, _p(out))
    assert rc == 0
    return out


def ref_neighbors(D, points, distance, cap=64):
    L = ref_lib()
    points = np.ascontiguousarray(points, np.float32)
    n = len(points)
    counts = np.zeros(n, np.uint32)
    idx = np.zeros((n, cap), np.uint32)
    mx = L.ref_neighbors(D, _p(points), n, distance, cap, _p(counts), _p(idx))
    assert mx <= cap
    return counts, idx


def ref_disk_areas(D, points, normals, grid_delta, bc, source_dir):
    L = ref_lib()
    points = np.ascontiguousarray(points, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    out = np.zeros(len(points), np.float32)
    bc = (C.c_int * 3)(*(list(bc) + [IGNORE] * 3)[:3])
    L.ref_disk_areas(D, _p(points), _p(normals), len(points), grid_delta, bc, source_dir, _p(out))
    return out


def ref_intersect_disks(points, normals, radius, bbox, source_dir, rays):
    L = ref_lib()
    points = np.ascontiguousarray(points, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    rays = np.ascontiguousarray(rays, np.float32)
    bbox = np.ascontiguousarray(bbox, np.float32)
    m = len(rays)
    geom = np.zeros(m, np.uint32)
    prim = np.zeros(m, np.uint32)
    t = np.zeros(m, np.float32)
    ng = np.zeros((m, 3), np.float32)
    L.ref_intersect_disks(_p(points), _p(normals), len(points), radius, _p(bbox[0].copy()),
                          _p(bbox[1].copy()), source_dir, _p(rays), m, _p(geom), _p(prim), _p(t),
                          _p(ng))
    return geom, prim, t, ng


def ref_boundary_process_hit(D, bbox, bc, source_dir, org, direction, ng, prim_id, t):
    L = ref_lib()
    bbox = np.ascontiguousarray(bbox, np.float32)
    org = np.array(org, np.float32)
    d = np.array(direction, np.float32)
    ng = np.array(ng, np.float32)
    bc = (C.c_int * 3)(*(list(bc) + [IGNORE] * 3)[:3])
    reflect = L.ref_boundary_process_hit(D, _p(bbox[0].copy()), _p(bbox[1].copy()), bc, source_dir,
                                         _p(org), _p(d), _p(ng), prim_id, t)
    return reflect, org, d
