"""CPU (-m "not gpu"): pins the oracle (oracle/vr_oracle.c) against every
known-answer value the reference's own tests hold for the hot path
(SURVEY.md section 8c), and -- where oracle/_ref was built from
/root/reference -- against the reference's unmodified code run here.
Nothing in this file touches the CUDA library's compute entry points."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import common
from viennaray_b200 import host, scenes

F = np.float32


def _plane_scene(D, grid_delta, extent, direction, radius, source_dir, bc, offset):
    pts, nrm = scenes.plane_grid(grid_delta, extent, direction)
    s = po.OracleScene(D)
    s.set_disks(pts, nrm, radius)
    s.setup(source_dir, bc, offset)
    return s, pts, nrm


# --------------------------------------------------------------------------------------
# tests/intersectionTest/intersectionTest.cpp:91-92,126-127
# --------------------------------------------------------------------------------------
def test_intersection_known_ids():
    gd = F(0.5)
    r = F(gd * F(0.5 * 1.7320508 * (1 + 1e-5)))
    s, pts, nrm = _plane_scene(3, 0.5, 10.0, (0, 1, 2), r, po.POS_Z, [0, 0, 0], r)
    assert len(pts) == 41 * 41
    d2 = np.array([0.0, 2.0, -1.0])
    d2 /= np.linalg.norm(d2)
    rays = np.array([[0, 0, 2 * r, 0, 0, -1], [0, 9, 2 * r, d2[0], d2[1], d2[2]]], F)
    geom, prim, t, ng = s.intersect(rays)
    assert geom[0] == 1 and prim[0] == 840  # geometryID, primID == 840
    assert geom[1] == 0 and prim[1] == 7    # boundaryID, primID == 7
    assert abs(t[0] - 2 * r) < 1e-6
    if po.have_ref():  # the reference's Boundary/GeometryDisk + substitute intersector
        g2, p2, t2, _ = po.ref_intersect_disks(pts, nrm, r, s.bbox(), po.POS_Z, rays)
        assert (g2 == geom).all() and (p2 == prim).all()
        assert (t2.view(np.uint32) == t.view(np.uint32)).all()


# --------------------------------------------------------------------------------------
# tests/boundaryHit/boundaryHit.cpp:68-76,128-136,188-196
# --------------------------------------------------------------------------------------
BOUNDARY_3D = [
    # plane dirs, source, bc, direction, Ng, primID, expected origin
    ((0, 1, 2), po.POS_Z, [0, 1, 1], (0.5, 0.0, -0.25), (-1, 0, 0), 2, (1.0, 0.5, 0.25)),
    ((0, 2, 1), po.POS_Y, [1, 1, 0], (0.0, -0.25, 0.5), (0, 0, -1), 6, (0.5, 0.25, 1.0)),
    ((0, 1, 2), po.POS_Z, [1, 1, 1], (0.5, 0.0, -0.25), (-1, 0, 0), 2, (-1.0, 0.5, 0.25)),
]


@pytest.mark.parametrize("dirs,src,bc,direction,ng,prim,expect", BOUNDARY_3D)
def test_boundary_hit_3d(dirs, src, bc, direction, ng, prim, expect):
    s, pts, nrm = _plane_scene(3, 0.1, 1.0, dirs, F(0.1), src, bc, F(0.1))
    d = np.array(direction, np.float64)
    dist = np.linalg.norm(d)
    d = (d / dist).astype(F)
    reflect, org, rd, dd = s.boundary_process_hit((0.5, 0.5, 0.5), d, d, ng, prim, dist)
    assert reflect == 1
    assert np.allclose(org, expect, atol=1e-6)
    # VC_TEST_ASSERT_ISCLOSE(rayHit.ray.dir, direction): the returned direction is the ray's
    assert np.allclose(dd, rd, atol=1e-6)
    if bc[s_axis(dirs, src, prim)] == 0:  # reflective: mirrored about the plane normal
        n = np.array(ng, F)
        assert np.allclose(rd, d - 2 * np.dot(d, n) * n, atol=1e-6)
    else:  # periodic: unchanged
        assert np.allclose(rd, d, atol=1e-7)
    if po.have_ref():
        r2, o2, d2 = po.ref_boundary_process_hit(3, s.bbox(), bc, src, (0.5, 0.5, 0.5), d, ng,
                                                 prim, F(dist))
        assert r2 == reflect
        assert (o2.view(np.uint32) == org.view(np.uint32)).all()
        assert (d2.view(np.uint32) == rd.view(np.uint32)).all()


def s_axis(dirs, src, prim):
    _, first, second, _, _ = host.trace_settings(src)
    return first if prim <= 3 else second


# tests/boundaryHit2D/boundaryHit2D.cpp:73-79,122-128 (POS_X) and :185-191,234-240 (POS_Y)
BOUNDARY_2D = [
    ("y", po.POS_X, 0, (1.0, 1.0, 0.0), (-0.5, 1.0, 0.0), (0, -1, 0), 3, (0.5, 2.0, 0.0)),
    ("y", po.POS_X, 1, (1.0, 1.0, 0.0), (-0.5, 1.0, 0.0), (0, -1, 0), 3, (0.5, -2.0, 0.0)),
    ("x", po.POS_Y, 0, (1.0, 1.0, 0.0), (1.0, -0.5, 0.0), (-1, 0, 0), 3, (2.0, 0.5, 0.0)),
    ("x", po.POS_Y, 1, (1.0, 1.0, 0.0), (1.0, -0.5, 0.0), (-1, 0, 0), 3, (-2.0, 0.5, 0.0)),
]


@pytest.mark.parametrize("line,src,cond,origin,direction,ng,prim,expect", BOUNDARY_2D)
def test_boundary_hit_2d(line, src, cond, origin, direction, ng, prim, expect):
    gd, extent = F(0.5), F(2.0)
    vals = []
    v = F(-extent)
    while v <= extent:
        vals.append(v)
        v = F(v + gd)
    pts = np.zeros((len(vals), 3), F)
    nrm = np.zeros((len(vals), 3), F)
    if line == "y":
        pts[:, 1] = vals
        nrm[:, 0] = 1
        bc = [2, cond, 2]
    else:
        pts[:, 0] = vals
        nrm[:, 1] = 1
        bc = [cond, 2, 2]
    s = po.OracleScene(2)
    s.set_disks(pts, nrm, gd)
    s.setup(src, bc, gd)
    d = np.array(direction, np.float64)
    dist = np.linalg.norm(d)
    d = (d / dist).astype(F)
    reflect, org, rd, dd = s.boundary_process_hit(origin, d, d, ng, prim, dist)
    assert reflect == 1
    assert np.allclose(org, expect, atol=1e-6)
    assert np.allclose(dd, rd, atol=1e-6)
    if po.have_ref():
        r2, o2, d2 = po.ref_boundary_process_hit(2, s.bbox(), bc, src, origin, d, ng, prim,
                                                 F(dist))
        assert r2 == reflect and np.allclose(o2, org, atol=1e-7) and np.allclose(d2, rd, atol=1e-7)


# --------------------------------------------------------------------------------------
# tests/createRay/createRay.cpp:55-56 ... 191-192
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("src", [po.POS_X, po.NEG_X, po.POS_Y, po.NEG_Y, po.POS_Z, po.NEG_Z])
def test_create_ray_source_plane_and_direction_sign(src):
    I = common.inputs()
    gd = F(I["sphere3D_gridDelta"])
    s = po.OracleScene(3)
    s.set_disks(I["sphere3D_points"], I["sphere3D_normals"], gd)
    s.setup(src, [0, 0, 0], gd)
    axis, _, _, min_max, pos_neg = host.trace_settings(src)
    rays = s.source_rays(po.Particle(0, 1.0, 2.0, 0.0), s.config(1000, 31), 0, 1000)
    plane = (1.0 + 2 * gd) * (1 if min_max else -1)
    assert np.allclose(rays[:, axis], plane, atol=1e-6)
    assert (np.sign(rays[:, 3 + axis]) == pos_neg).all()
    assert np.allclose(np.linalg.norm(rays[:, 3:], axis=1), 1, atol=1e-5)
    lo, hi = s.bbox()
    for a in range(3):  # origins inside the source rectangle
        assert (rays[:, a] >= lo[a] - 1e-6).all() and (rays[:, a] <= hi[a] + 1e-6).all()


def test_create_ray_tilted_primary_direction():
    I = common.inputs()
    gd = F(I["sphere3D_gridDelta"])
    s = po.OracleScene(3)
    s.set_disks(I["sphere3D_points"], I["sphere3D_normals"], gd)
    s.setup(po.POS_Z, [0, 0, 0], gd)
    pd = np.array([1.0, 1.0, -1.0]) / np.sqrt(3.0)
    rays = s.source_rays(po.Particle(0, 1.0, 2.0, 0.0), s.config(2000, 31, primary_dir=pd), 0, 2000)
    assert (rays[:, 5] < 0).all()
    assert np.allclose(rays[:, 2], 1.0 + 2 * gd, atol=1e-6)
    mean = rays[:, 3:].mean(0)
    assert np.dot(mean / np.linalg.norm(mean), pd) > 0.98  # lobe centred on the primary direction


def test_source_power_cosine_distribution():
    """raySourceRandom.hpp:70-86: cos(theta) = r^(1/(n+1)), i.e. E[cos] = (n+1)/(n+2)."""
    I = common.inputs()
    gd = F(I["sphere3D_gridDelta"])
    s = po.OracleScene(3)
    s.set_disks(I["sphere3D_points"], I["sphere3D_normals"], gd)
    s.setup(po.POS_Z, [0, 0, 0], gd)
    for n in (1.0, 2.0, 100.0):
        rays = s.source_rays(po.Particle(0, 1.0, n, 0.0), s.config(200000, 7), 0, 200000)
        assert abs((-rays[:, 5]).mean() - (n + 1) / (n + 2)) < 2e-3


# --------------------------------------------------------------------------------------
# tests/pointNeighborhood/pointNeighborhood.cpp:51, pointNeighborhood2D.cpp:43,46
# --------------------------------------------------------------------------------------
def test_point_neighborhood_counts_3d():
    pts, nrm = scenes.plane_grid(0.5, 10.0)
    s = po.OracleScene(3)
    s.set_disks(pts, nrm, F(0.5) - F(1e-6))
    off, idx = s.neighbors()
    cnt = np.diff(off)
    lo, hi = pts.min(0), pts.max(0)
    on_x = (pts[:, 0] == lo[0]) | (pts[:, 0] == hi[0])
    on_y = (pts[:, 1] == lo[1]) | (pts[:, 1] == hi[1])
    assert (cnt[on_x & on_y] == 3).all()
    assert (cnt[on_x ^ on_y] == 5).all()
    assert (cnt[~on_x & ~on_y] == 8).all()
    if po.have_ref():
        rc, ri = po.ref_neighbors(3, pts, 2 * (F(0.5) - F(1e-6)))
        assert (rc == cnt).all()
        for i in range(0, len(pts), 37):
            assert sorted(ri[i, :rc[i]]) == list(idx[off[i]:off[i + 1]])


def test_point_neighborhood_counts_2d():
    pts = np.zeros((5, 3), F)
    pts[:, 0] = [-1, -0.5, 0, 0.5, 1]
    nrm = np.zeros_like(pts)
    nrm[:, 1] = 1
    s = po.OracleScene(2)
    s.set_disks(pts, nrm, F(0.5) - F(1e-6))
    off, _ = s.neighbors()
    assert list(np.diff(off)) == [1, 2, 2, 2, 1]


@pytest.mark.parametrize("name", ["disk3D", "disk2D"])
def test_neighbor_lists_equal_reference(name):
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    c = common.case(name)
    s = common.make_oracle(c)
    off, idx = s.neighbors()
    r = host.disk_radius(c["grid_delta"], c["D"])
    rc, ri = po.ref_neighbors(c["D"], c["points"], F(2) * r, cap=64)
    assert (rc == np.diff(off)).all()
    for i in range(len(rc)):
        assert sorted(ri[i, :rc[i]]) == list(idx[off[i]:off[i + 1]])
    if name == "disk3D":  # mean 7.98, max 11
        assert rc.max() == 11 and abs(rc.mean() - 7.98) < 0.01


# --------------------------------------------------------------------------------------
# tests/smoothing/smoothing.cpp:43,50
# --------------------------------------------------------------------------------------
def test_smoothing_known_answer():
    pts = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0], [1, 1, 0], [2, 1, 0]], F)
    nrm = np.array([[0, 0, 1]] * 3 + [[0, 1, 0]] * 3, F)
    s = po.OracleScene(3)
    s.set_disks(pts, nrm, host.disk_radius(1.0, 3))
    out = s.smooth_flux(np.array([1, 1, 1, 0, 0, 0], F))
    assert np.allclose(out, [1, 1, 1, 0, 0, 0], atol=1e-6)


@pytest.mark.parametrize("name", ["disk3D", "trench"])
def test_postprocessing_equals_reference(name):
    """normalizeFlux(MAX) and smoothFlux(k), k = 1, 2, 3, of the oracle against the reference's
    own TraceDisk methods on the same flux vector (rayTraceDisk.hpp:103-193).  MAX is formed in
    double on both sides: bit-equal.  The smoothing sums run over the same neighbour sets in a
    different row order (the oracle's rows are ascending): float tolerance."""
    if not po.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference)")
    c = common.case(name)
    orc = common.make_oracle(c)
    rng = np.random.default_rng(11)
    flux = (rng.random(orc.n) * 50).astype(F)
    areas = po.ref_disk_areas(c["D"], c["points"], c["normals"], c["grid_delta"], c["bc"],
                              c["source_dir"])
    args = (c["D"], c["points"], c["normals"], c["grid_delta"], c["bc"], c["source_dir"])
    ref = po.ref_post_disk(*args, flux, norm=2)
    mine = orc.normalize_flux_max(flux, areas)
    assert (ref.view(np.uint32) == mine.view(np.uint32)).all()
    for k in (1, 2, 3):
        ref = po.ref_post_disk(*args, flux, norm=0, smooth=k)
        mine = orc.smooth_flux(flux, k)
        assert np.allclose(ref, mine, rtol=2e-6, atol=1e-6), "k = %d" % k


# --------------------------------------------------------------------------------------
# tests/rngSeed/rngSeed.cpp:48-51 and tests/traceInterface/traceInterface.cpp:67
# --------------------------------------------------------------------------------------
def test_rng_seed_bitwise_reproducible_and_num_rays():
    c = common.case("plane")  # 21 x 21 plane, sticking 1
    s = common.make_oracle(c)
    assert s.n == 441
    cfg = s.config(441 * 10, 12345 + 0)
    f1, i1 = s.trace(common.oracle_particle(c), cfg)
    f2, i2 = s.trace(common.oracle_particle(c), cfg)
    assert (f1 == f2).all() and i1.as_dict() == i2.as_dict()
    assert i1.numRays == 4410
    f3, _ = s.trace(common.oracle_particle(c), s.config(4410, 12345 + 1))  # runNumber + 1
    assert (f3 != f1).any()
    # sticking 1: every contribution is exactly one unit of weight
    assert (f1 % (1 << 30) == 0).all()
    # shards of the ray-index range add up to the whole job (multi-GPU partition)
    fa, ia = s.trace(common.oracle_particle(c), cfg, 0, 1500)
    fb, ib = s.trace(common.oracle_particle(c), cfg, 1500, 4410)
    assert (fa + fb == f1).all()
    assert ia.totalTraces + ib.totalTraces == i1.totalTraces


# --------------------------------------------------------------------------------------
# the walk itself against the reference's unmodified TraceKernel (oracle/_ref)
# --------------------------------------------------------------------------------------
def _repeat_oracle(s, part, num, seeds):
    out = []
    info = None
    for sd in seeds:
        f, info = s.trace(part, s.config(num, sd))
        out.append(f / po.FLUX_SCALE)
    return np.asarray(out), info


def _first_bounce(orc, rays, prim, t, D, seed=5):
    """Secondary rays leaving the first hit points in random upward directions."""
    rng = np.random.default_rng(seed)
    hitp = rays[:, :3] + rays[:, 3:] * t[:, None]
    nrm = orc.normals()[prim]
    d = rng.normal(size=(len(rays), 3)).astype(F)
    if D == 2:
        d[:, 2] = 0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[(d * nrm).sum(1) < 0] *= -1
    return np.ascontiguousarray(np.concatenate([hitp, d], 1), F)


@pytest.mark.parametrize("name", ["disk3D", "trench", "holes"])
def test_fixed_ray_set_against_reference_intersector(name):
    """BASELINE correctness level 1 on the oracle side: geomID / primID / t of the same 200k
    primary rays + one bounce that tests/test_gpu_parity.py feeds to the CUDA kernels, here
    against the reference's own Boundary + GeometryDisk + rtcIntersect1 call sequence
    (rayTraceKernel.hpp:41-45,156-166) in oracle/_ref.  IDs exact, t 0 ulp."""
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    c = common.case(name)
    s = common.make_oracle(c)
    r = host.disk_radius(c["grid_delta"], c["D"])
    rays = s.source_rays(common.oracle_particle(c), s.config(10**7, 12346), 0, 200000)
    for leg in range(2):
        go, po_, to, ngo = s.intersect(rays)
        gr, pr, tr, ngr = po.ref_intersect_disks(c["points"], c["normals"], r, s.bbox(),
                                                 c["source_dir"], rays)
        assert (go == gr).all(), "geomID mismatch on %d rays" % int((go != gr).sum())
        assert (po_ == pr).all(), "primID mismatch on %d rays" % int((po_ != pr).sum())
        hit = go != 0xFFFFFFFF
        assert hit.mean() > (0.5 if leg == 0 else 0.1)
        assert (to[hit].view(np.uint32) == tr[hit].view(np.uint32)).all()
        bnd = go == 0  # unnormalised boundary normals steer processHit (rayBoundary.hpp:36-38)
        assert (ngo[bnd].view(np.uint32) == ngr[bnd].view(np.uint32)).all()
        keep = go == 1
        rays = _first_bounce(s, rays[keep], po_[keep], to[keep], c["D"])


def _stat_compare(fo, io, fr, ir, num, K, hits=True):
    mo, mr = fo.mean(0), fr.astype(np.float64).mean(0)
    rel_l2 = np.linalg.norm(mo - mr) / np.linalg.norm(mr)
    # expected L2 of pure sampling noise between two K-run means
    var = (fo.var(0, ddof=1) + fr.astype(np.float64).var(0, ddof=1)) / K
    noise = np.sqrt(var.sum()) / np.linalg.norm(mr)
    assert rel_l2 < 1.5 * noise + 1e-3, (rel_l2, noise)
    sel = var > 0
    z = np.abs(mo - mr)[sel] / np.sqrt(var[sel])
    # t-distributed with ~14 dof: P(|z| > 3) ~ 1%; allow 3%
    assert (z > 3).mean() < 0.03, float((z > 3).mean())
    # the random walk has the same shape: traces and hits per ray
    tr_o, tr_r = io.totalTraces / num, ir[1] / num
    assert abs(tr_o - tr_r) / tr_r < 0.01, (tr_o, tr_r)
    if hits:
        assert abs(io.geoHits / num - ir[3] / num) / (ir[3] / num) < 0.01


def _stat_case(name):
    """name[+specular][+tilted]: a config of tests/common.py, optionally with the specular
    particle (rayParticle.hpp:165-204) and a tilted source (raySourceRandom.hpp:88-116)."""
    parts = name.split("+")
    c = dict(common.case(parts[0]))
    primary = None
    if "specular" in parts:
        c.update(kind=1, sticking=0.3, power=5.0, cone=0.0)
    if "tilted" in parts:
        primary = np.array([0.3, 0.1, -1.0]) / np.linalg.norm([0.3, 0.1, -1.0])
    return c, primary


STAT_CASES = [("disk3D", 50), ("triangle3D", 100), ("disk2D", 2000), ("trench_ion", 60),
              ("trench", 60), ("holes", 40), ("trench+specular+tilted", 60),
              ("trench+tilted", 60)]


@pytest.mark.parametrize("name,rays_per_prim", STAT_CASES)
def test_flux_statistical_parity_with_reference_kernel(name, rays_per_prim):
    """BASELINE correctness level 2: per-primitive flux of the oracle against the
    reference CPU tracer at equal ray counts -- relative L2 of the means and the
    per-primitive 3-sigma test over K = 8 seeded repeats per side.  Covers all five
    configs (C4 with both of its particles, C5 reduced) plus the specular particle and the
    tilted source."""
    if not po.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    c, primary = _stat_case(name)
    s = common.make_oracle(c)
    K = 8
    num = s.n * rays_per_prim
    fo, io = [], None
    for k in range(K):
        f, io = s.trace(common.oracle_particle(c), s.config(num, 100 + k, primary_dir=primary))
        fo.append(f / po.FLUX_SCALE)
    fo = np.asarray(fo)
    if c["geo"] == "disk":
        fr, ir, _ = po.ref_trace_disk(c["D"], c["points"], c["normals"], c["grid_delta"], c["bc"],
                                      c["source_dir"], c["kind"], c["sticking"], c["power"],
                                      c["cone"], rays_fixed=num, seed=555, runs=K,
                                      primary_dir=primary)
    else:
        fr, ir, _ = po.ref_trace_triangle(c["verts"], c["tris"], c["grid_delta"], c["bc"],
                                          c["source_dir"], c["kind"], c["sticking"], c["power"],
                                          c["cone"], rays_fixed=num, seed=555, runs=K)
    _stat_compare(fo, io, fr, ir, num, K)


def _material_ids(c):
    """two materials: 'mask' on top (id 1), 'substrate' below the top surface (id 0)"""
    z = c["points"][:, 2 if c["D"] == 3 else 1]
    return (z > z.max() - 1e-3).astype(np.int32)


OPTION_CASES = [
    # name, rays/prim, wdist, mean free path, sticking by material
    ("trench", 60, True, 0.0, None),          # VIENNARAY_USE_WDIST, rayTraceKernel.hpp:258-296
    ("trench", 60, False, 30.0, None),        # getMeanFreePath() > 0, rayTraceKernel.hpp:179-203
    ("trench_ion", 60, False, 30.0, None),
    ("trench", 60, False, 0.0, (0.05, 0.6)),  # materialId-dependent sticking, :310-313
    ("triangle3D", 40, False, 0.0, (0.02, 0.5)),
    ("disk3D", 20, True, 25.0, (0.3, 0.05)),  # all three at once
]


@pytest.mark.parametrize("name,rays_per_prim,wdist,mfp,table", OPTION_CASES)
def test_optional_features_statistical_parity_with_reference(name, rays_per_prim, wdist, mfp,
                                                             table):
    """The oracle's restatement of the optional parts of the loop -- distance-weighted spread,
    mean-free-path scattering, sticking looked up by the hit primitive's materialId -- against
    the reference's own code: oracle/_ref built a second time with -DVIENNARAY_USE_WDIST, and a
    user-side particle in oracle/ref_driver.cpp that overrides getMeanFreePath() and switches
    on the materialId it is handed."""
    if not po.have_ref() or (wdist and not po.have_ref_wdist()):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    c = common.case(name)
    s = common.make_oracle(c)
    mats = None
    part = common.oracle_particle(c)
    part.meanFreePath = mfp
    if table is not None:
        if c["geo"] == "disk":
            mats = _material_ids(c)
        else:  # triangles: by the height of the first vertex
            z = c["verts"][c["tris"][:, 0], 2]
            mats = (z > z.max() - 1e-3).astype(np.int32)
        assert 0 < mats.sum() < len(mats)
        s.set_material_ids(mats)
        part.set_sticking_by_material(table)
    K = 8
    num = s.n * rays_per_prim
    fo, io = [], None
    for k in range(K):
        f, io = s.trace(part, s.config(num, 300 + k, wdist=wdist))
        fo.append(f / po.FLUX_SCALE)
    fo = np.asarray(fo)
    kw = dict(rays_fixed=num, seed=777, runs=K, mean_free_path=mfp, sticking_by_material=table,
              material_ids=mats)
    if c["geo"] == "disk":
        fr, ir, _ = po.ref_trace_disk(c["D"], c["points"], c["normals"], c["grid_delta"], c["bc"],
                                      c["source_dir"], c["kind"], c["sticking"], c["power"],
                                      c["cone"], wdist=wdist, **kw)
    else:
        fr, ir, _ = po.ref_trace_triangle(c["verts"], c["tris"], c["grid_delta"], c["bc"],
                                          c["source_dir"], c["kind"], c["sticking"], c["power"],
                                          c["cone"], **kw)
    _stat_compare(fo, io, fr, ir, num, K)
    if mfp > 0:  # scatter events per ray (TraceInfo.particleHits, rayTraceKernel.hpp:198)
        assert ir[4] > 0
        assert abs(io.particleHits - ir[4]) / ir[4] < 0.02, (io.particleHits, int(ir[4]))
    if table is not None:  # the table is really in use: a constant-sticking run differs
        f0, _ = s.trace(common.oracle_particle(c), s.config(num, 300))
        assert not np.allclose(f0 / po.FLUX_SCALE, fo[0])


# --------------------------------------------------------------------------------------
# tests/createSourceGrid/createSourceGrid.cpp:43-46
# --------------------------------------------------------------------------------------
def test_source_grid_origins_and_direction():
    I = common.inputs()
    gd = F(I["sphere3D_gridDelta"])
    s = po.OracleScene(3)
    s.set_disks(I["sphere3D_points"], I["sphere3D_normals"], gd)
    s.setup(po.POS_Z, [0, 0, 0], gd)
    lo, hi = s.bbox()
    grid = host.create_source_grid(lo, hi, len(I["sphere3D_points"]), gd, host.POS_Z)
    assert 100 < len(grid) <= len(I["sphere3D_points"])
    s.set_source_grid(grid)
    m = len(grid)
    rays = s.source_rays(po.Particle(0, 1.0, 1.0, 0.0), s.config(m, 0), 0, 2 * m)
    assert (rays[:, 5] < 0).all()
    assert np.allclose(rays[:, 2], 1.0 + 2 * gd, atol=1e-6)
    assert np.allclose(rays[:m, 0], grid[:, 0], atol=1e-6)
    assert np.allclose(rays[:m, 1], grid[:, 1], atol=1e-6)
    assert (rays[m:, :3] == rays[:m, :3]).all()  # idx % numPoints
    assert np.allclose(np.linalg.norm(rays[:, 3:], axis=1), 1, atol=1e-5)
    # cos^n source: E[cos theta] = 2/3 for n = 1 (raySourceGrid.hpp:42-51)
    big = s.source_rays(po.Particle(0, 1.0, 1.0, 0.0), s.config(10**6, 5), 0, 200000)
    assert abs((-big[:, 5]).mean() - 2.0 / 3.0) < 3e-3
    s.set_source_grid(None)
    back = s.source_rays(po.Particle(0, 1.0, 1.0, 0.0), s.config(m, 0), 0, 10)
    assert not np.allclose(back[:, 0], grid[:10, 0])
