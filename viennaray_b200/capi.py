"""ctypes binding of the C ABI in include/viennaray_b200.h (the same calls a
cgo / JNI / C++ host would make).  Loading fails loudly when the CUDA library
has not been built; there is no fallback of any kind."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VR_LIB_PATH") or os.path.join(_HERE, "libviennaray_b200.so")

_vp = C.c_void_p
FLUX_FIXED_SCALE = float(2**30)

PARTICLE_DIFFUSE, PARTICLE_SPECULAR, PARTICLE_CONED_COSINE = 0, 1, 2
BOUNDARY_REFLECTIVE, BOUNDARY_PERIODIC, BOUNDARY_IGNORE = 0, 1, 2
FLAG_WDIST = 1

EXPORTS = [
    "vr_ctx_create", "vr_ctx_create_multi", "vr_ctx_num_devices", "vr_ctx_destroy",
    "vr_last_error", "vr_scene_set_disks",
    "vr_scene_set_triangles", "vr_scene_build_neighbors", "vr_scene_get_neighbors",
    "vr_scene_set_boundary", "vr_source_set_grid", "vr_scene_commit", "vr_trace",
    "vr_trace_device", "vr_flux_device", "vr_flux_download", "vr_flux_download_fixed",
    "vr_flux_postprocess", "vr_flux_postprocess_ex",
    "vr_ctx_stream", "vr_ctx_synchronize", "vr_last_kernel_ms", "vr_last_launch_count", "vr_build_neighbors", "vr_free",
    "vr_debug_intersect", "vr_debug_source_rays", "vr_debug_math", "vr_debug_philox",
    "vr_debug_reflect", "vr_debug_bvh_stats", "vr_debug_work_counters", "vr_debug_l2_read_bandwidth", "vr_debug_phase_timing",
    "vr_debug_phase_ms",
]


NORM_NONE, NORM_SOURCE, NORM_MAX = 0, 1, 2  # vr_flux_postprocess_ex


class SourceDesc(C.Structure):
    _fields_ = [("bboxMin", C.c_float * 3), ("bboxMax", C.c_float * 3), ("rayDir", C.c_int32),
                ("firstDir", C.c_int32), ("secondDir", C.c_int32), ("minMax", C.c_int32),
                ("posNeg", C.c_float), ("useBasis", C.c_int32), ("basis", C.c_float * 9),
                ("useGrid", C.c_int32)]


class ParticleDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("sticking", C.c_float), ("sourcePower", C.c_float),
                ("coneMinAngle", C.c_float), ("meanFreePath", C.c_float),
                ("stickingByMaterial", C.c_void_p), ("numMaterials", C.c_int32)]

    def set_sticking_by_material(self, table):
        """sticking[materialId of the hit primitive]; None returns to the constant."""
        if table is None:
            self._table = None
            self.stickingByMaterial, self.numMaterials = None, 0
            return self
        self._table = np.ascontiguousarray(table, np.float32)  # kept alive with the struct
        self.stickingByMaterial = self._table.ctypes.data
        self.numMaterials = len(self._table)
        return self


class Config(C.Structure):
    _fields_ = [("numRays", C.c_uint64), ("rayIdxBegin", C.c_uint64), ("rayIdxEnd", C.c_uint64),
                ("seed", C.c_uint32), ("maxReflections", C.c_uint32),
                ("maxBoundaryHits", C.c_uint32), ("flags", C.c_uint32)]


class TraceInfo(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("numRays", "totalRaysTraced", "nonGeometryHits",
                                          "geometryHits", "particleHits", "boundaryHits",
                                          "reflections", "raysTerminated")] + [("time", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class VrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("viennaray_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "viennaray_b200: %s is missing -- build it with "
                "`python -m viennaray_b200.build` (nvcc, sm_100a); there is no CPU fallback"
                % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.vr_last_error.restype = C.c_char_p
        L.vr_last_error.argtypes = [_vp]
        L.vr_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
        L.vr_ctx_create_multi.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)]
        L.vr_ctx_num_devices.argtypes = [_vp]
        L.vr_ctx_destroy.restype = None
        L.vr_ctx_destroy.argtypes = [_vp]
        L.vr_scene_set_disks.argtypes = [_vp, _vp, _vp, C.c_uint32, _vp, _vp, _vp]
        L.vr_scene_set_triangles.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp]
        L.vr_scene_set_boundary.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int]
        L.vr_scene_commit.argtypes = [_vp]
        L.vr_source_set_grid.argtypes = [_vp, _vp, C.c_uint32]
        L.vr_scene_build_neighbors.argtypes = [_vp, C.c_int, _vp, C.c_float]
        L.vr_scene_get_neighbors.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp)]
        L.vr_trace.argtypes = [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp]
        L.vr_trace_device.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int]
        L.vr_flux_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]
        L.vr_flux_download.argtypes = [_vp, _vp, _vp]
        L.vr_flux_download_fixed.argtypes = [_vp, _vp]
        L.vr_flux_postprocess.argtypes = [_vp, C.c_int, _vp, C.c_float, C.c_int, _vp]
        L.vr_flux_postprocess_ex.argtypes = [_vp, C.c_int, _vp, C.c_int, C.c_double, C.c_int,
                                             C.c_float, _vp]
        L.vr_ctx_stream.restype = _vp
        L.vr_ctx_stream.argtypes = [_vp]
        L.vr_ctx_synchronize.argtypes = [_vp]
        L.vr_last_kernel_ms.restype = C.c_float
        L.vr_last_kernel_ms.argtypes = [_vp]
        L.vr_last_launch_count.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.vr_build_neighbors.argtypes = [C.c_int, _vp, C.c_uint32, C.c_float, C.POINTER(_vp),
                                         C.POINTER(_vp)]
        L.vr_free.restype = None
        L.vr_free.argtypes = [_vp]
        L.vr_debug_intersect.argtypes = [_vp, _vp, C.c_uint32, _vp, _vp, _vp, C.c_uint32, _vp, _vp]
        L.vr_debug_source_rays.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint32, _vp]
        L.vr_debug_math.argtypes = [_vp, C.c_int, _vp, C.c_uint32, C.c_float, _vp]
        L.vr_debug_philox.argtypes = [_vp] + [C.c_uint32] * 6 + [_vp]
        L.vr_debug_reflect.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, C.c_float, C.c_uint32,
                                       C.c_uint64, C.c_uint32, _vp]
        L.vr_debug_bvh_stats.argtypes = [_vp, _vp]
        L.vr_debug_work_counters.argtypes = [_vp, _vp]
        L.vr_debug_l2_read_bandwidth.argtypes = [_vp, C.c_uint64, C.c_int, _vp]
        L.vr_debug_phase_timing.argtypes = [_vp, C.c_int]
        L.vr_debug_phase_ms.argtypes = [_vp, _vp, _vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def build_neighbors(D, points, distance):
    """Neighbour CSR (offsets, indices) of PointNeighborhood::init semantics."""
    L = lib()
    points = np.ascontiguousarray(points, np.float32)
    n = len(points)
    off, idx = _vp(), _vp()
    rc = L.vr_build_neighbors(D, _p(points), n, np.float32(distance), C.byref(off), C.byref(idx))
    if rc:
        raise VrError(rc, "vr_build_neighbors")
    offsets = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint32)), (n + 1,)).copy()
    total = int(offsets[-1])
    indices = (np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_uint32)), (total,)).copy()
               if total else np.zeros(0, np.uint32))
    L.vr_free(off)
    L.vr_free(idx)
    return offsets, indices


class Context:
    """One vr_ctx: one scene on one GPU -- or, with a list of device ordinals, on several
    GPUs of the node behind the same calls (vr_ctx_create_multi: rays sharded over the
    devices, one NCCL all-reduce of the result words inside vr_trace*)."""

    def __init__(self, device=0):
        self.L = lib()
        h = _vp()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self.L.vr_ctx_create_multi(len(device), ids, C.byref(h))
        else:
            rc = self.L.vr_ctx_create(int(device), C.byref(h))
        if rc:
            raise VrError(rc, self.L.vr_last_error(None).decode())
        self.h = h
        self.n = 0
        self.num_particles = 0

    def num_devices(self):
        return int(self.L.vr_ctx_num_devices(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.vr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc:
            raise VrError(rc, self.L.vr_last_error(self.h).decode())

    def set_disks(self, xyzr, normals, nb_offsets=None, nb_indices=None, material_ids=None):
        xyzr = np.ascontiguousarray(xyzr, np.float32)
        normals = np.ascontiguousarray(normals, np.float32)
        assert xyzr.ndim == 2 and xyzr.shape[1] == 4 and normals.shape == (len(xyzr), 3)
        if nb_offsets is not None:
            nb_offsets = np.ascontiguousarray(nb_offsets, np.uint32)
            nb_indices = np.ascontiguousarray(nb_indices, np.uint32)
            if len(nb_indices) == 0:
                nb_indices = np.zeros(1, np.uint32)
        if material_ids is not None:
            material_ids = np.ascontiguousarray(material_ids, np.int32)
        self.n = len(xyzr)
        self._ck(self.L.vr_scene_set_disks(self.h, _p(xyzr), _p(normals), self.n,
                                           _p(material_ids), _p(nb_offsets), _p(nb_indices)))

    def build_neighbors_device(self, D, points, distance):
        """Neighbour lists of the disks set before, built on the device."""
        points = np.ascontiguousarray(points, np.float32)
        assert points.shape == (self.n, 3)
        self._ck(self.L.vr_scene_build_neighbors(self.h, D, _p(points), np.float32(distance)))

    def get_neighbors(self):
        off, idx = _vp(), _vp()
        self._ck(self.L.vr_scene_get_neighbors(self.h, C.byref(off), C.byref(idx)))
        offsets = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint32)), (self.n + 1,)).copy()
        total = int(offsets[-1])
        indices = (np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_uint32)), (total,)).copy()
                   if total else np.zeros(0, np.uint32))
        self.L.vr_free(off)
        self.L.vr_free(idx)
        return offsets, indices

    def set_triangles(self, verts, tris, normals, material_ids=None):
        verts = np.ascontiguousarray(verts, np.float32)
        tris = np.ascontiguousarray(tris, np.uint32)
        normals = np.ascontiguousarray(normals, np.float32)
        if material_ids is not None:
            material_ids = np.ascontiguousarray(material_ids, np.int32)
        self.n = len(tris)
        self._ck(self.L.vr_scene_set_triangles(self.h, _p(verts), len(verts), _p(tris), self.n,
                                               _p(normals), _p(material_ids)))

    def set_boundary(self, bbox_min, bbox_max, first_dir, second_dir, cond_first, cond_second, D):
        lo = np.ascontiguousarray(bbox_min, np.float32)
        hi = np.ascontiguousarray(bbox_max, np.float32)
        self._ck(self.L.vr_scene_set_boundary(self.h, _p(lo), _p(hi), first_dir, second_dir,
                                              cond_first, cond_second, D))

    def set_source_grid(self, points):
        """Origins of the grid source (raySourceGrid.hpp); None releases them."""
        if points is None or len(points) == 0:
            self._ck(self.L.vr_source_set_grid(self.h, None, 0))
            return
        points = np.ascontiguousarray(points, np.float32)
        self._ck(self.L.vr_source_set_grid(self.h, _p(points), len(points)))

    def commit(self):
        self._ck(self.L.vr_scene_commit(self.h))

    @staticmethod
    def _particles(particles):
        arr = (ParticleDesc * len(particles))()
        for i, p in enumerate(particles):
            arr[i] = p
        return arr

    def trace(self, source, particles, config):
        """vr_trace: host in, host out.  Returns (flux[np, n] float64, infos)."""
        np_ = len(particles)
        flux = np.zeros((np_, self.n), np.float64)
        infos = (TraceInfo * np_)()
        arr = self._particles(particles)
        self._ck(self.L.vr_trace(self.h, C.byref(source), arr, np_, C.byref(config), _p(flux),
                                 infos))
        self.num_particles = np_
        return flux, list(infos)

    def trace_device(self, source, particles, config, sync=False):
        np_ = len(particles)
        arr = self._particles(particles)
        self._ck(self.L.vr_trace_device(self.h, C.byref(source), arr, np_, C.byref(config),
                                        1 if sync else 0))
        self.num_particles = np_

    def flux_device(self):
        ptr, words = _vp(), C.c_size_t()
        self._ck(self.L.vr_flux_device(self.h, C.byref(ptr), C.byref(words)))
        return ptr.value, words.value

    def flux_download(self):
        flux = np.zeros((self.num_particles, self.n), np.float64)
        infos = (TraceInfo * self.num_particles)()
        self._ck(self.L.vr_flux_download(self.h, _p(flux), infos))
        return flux, list(infos)

    def flux_download_fixed(self):
        flux = np.zeros((self.num_particles, self.n), np.uint64)
        self._ck(self.L.vr_flux_download_fixed(self.h, _p(flux)))
        return flux

    def flux_postprocess(self, particle=0, areas=None, norm_factor=1.0, smooth=False):
        """Normalised / smoothed float flux of one particle, computed on the device."""
        if areas is not None:
            areas = np.ascontiguousarray(areas, np.float32)
        out = np.zeros(self.n, np.float32)
        self._ck(self.L.vr_flux_postprocess(self.h, particle, _p(areas), np.float32(norm_factor),
                                            1 if smooth else 0, _p(out)))
        return out

    def flux_postprocess_ex(self, particle=0, areas=None, normalization=0, norm_factor=1.0,
                            smooth_neighbors=0, disk_radius=0.0):
        """normalizeFlux(SOURCE = 1 | MAX = 2) + smoothFlux(smooth_neighbors) on the device."""
        if areas is not None:
            areas = np.ascontiguousarray(areas, np.float32)
        out = np.zeros(self.n, np.float32)
        self._ck(self.L.vr_flux_postprocess_ex(self.h, particle, _p(areas), int(normalization),
                                               float(norm_factor), int(smooth_neighbors),
                                               np.float32(disk_radius), _p(out)))
        return out

    def stream(self):
        return self.L.vr_ctx_stream(self.h)

    def synchronize(self):
        self._ck(self.L.vr_ctx_synchronize(self.h))

    def last_kernel_ms(self):
        return float(self.L.vr_last_kernel_ms(self.h))

    def last_launch_count(self):
        k, it = C.c_int(), C.c_int()
        self._ck(self.L.vr_last_launch_count(self.h, C.byref(k), C.byref(it)))
        return k.value, it.value

    def debug_intersect(self, rays, nb_cap=16):
        rays = np.ascontiguousarray(rays, np.float32)
        m = len(rays)
        geom = np.zeros(m, np.uint32)
        prim = np.zeros(m, np.uint32)
        t = np.zeros(m, np.float32)
        cnt = np.zeros(m, np.uint32)
        nb = np.full((m, nb_cap), 0xFFFFFFFF, np.uint32)
        self._ck(self.L.vr_debug_intersect(self.h, _p(rays), m, _p(geom), _p(prim), _p(t), nb_cap,
                                           _p(cnt), _p(nb)))
        return geom, prim, t, cnt, nb

    def debug_source_rays(self, source, particle, config, idx_begin, m):
        rays = np.zeros((m, 6), np.float32)
        self._ck(self.L.vr_debug_source_rays(self.h, C.byref(source), C.byref(particle),
                                             C.byref(config), idx_begin, m, _p(rays)))
        return rays

    def debug_math(self, which, x, param=0.0):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(2 * len(x) if which == 0 else len(x), np.float32)
        self._ck(self.L.vr_debug_math(self.h, which, _p(x), len(x), np.float32(param), _p(out)))
        return out

    def debug_philox(self, k0, k1, c0, c1, c2, c3):
        out = np.zeros(4, np.uint32)
        self._ck(self.L.vr_debug_philox(self.h, k0, k1, c0, c1, c2, c3, _p(out)))
        return out

    def debug_reflect(self, kind, D, ray_dir, normal, cone_min_angle, seed, idx, m):
        rd = np.ascontiguousarray(ray_dir, np.float32)
        nn = np.ascontiguousarray(normal, np.float32)
        out = np.zeros((m, 3), np.float32)
        self._ck(self.L.vr_debug_reflect(self.h, kind, D, _p(rd), _p(nn),
                                         np.float32(cone_min_angle), seed, idx, m, _p(out)))
        return out

    def bvh_stats(self):
        out = np.zeros(8, np.uint64)
        self._ck(self.L.vr_debug_bvh_stats(self.h, _p(out)))
        f = out[4:8].astype(np.uint32).view(np.float32)
        return {"nodes": int(out[0]), "leaves": int(out[1]), "max_leaf": int(out[2]),
                "node_bytes": int(out[3]), "build_ms": float(f[0]), "sah_inner": float(f[1]),
                "sah_leaf": float(f[2]), "morton_alpha": float(f[3])}

    def phase_timing(self, enable):
        self._ck(self.L.vr_debug_phase_timing(self.h, 1 if enable else 0))

    def phase_ms(self):
        ms = np.zeros(3, np.float64)
        n = np.zeros(3, np.int64)
        self._ck(self.L.vr_debug_phase_ms(self.h, _p(ms), _p(n)))
        return {"traverse_ms": float(ms[0]), "shade_ms": float(ms[1]), "other_ms": float(ms[2]),
                "traverse_launches": int(n[0]), "shade_launches": int(n[1])}

    def l2_read_bandwidth(self, nbytes=48 << 20, passes=20):
        """(GB/s read from an L2-resident buffer, L2 size in bytes) -- vr_debug_l2_read_bandwidth"""
        out = np.zeros(2, np.float64)
        self._ck(self.L.vr_debug_l2_read_bandwidth(self.h, int(nbytes), int(passes), _p(out)))
        return float(out[0]), int(out[1])

    def work_counters(self):
        out = np.zeros(5, np.uint64)
        self._ck(self.L.vr_debug_work_counters(self.h, _p(out)))
        return {"node_visits": int(out[0]), "prim_tests": int(out[1]), "nb_tests": int(out[2]),
                "flux_adds": int(out[3]), "sky_finished": int(out[4])}
