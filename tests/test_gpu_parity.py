"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer / index results must match exactly; hit distances must
match bit-for-bit too (both sides use unfused IEEE float arithmetic in the same
order); per-primitive flux is compared in its 2^-30 fixed-point form, exactly."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import common
from viennaray_b200 import capi, host

pytestmark = pytest.mark.gpu

SEED = 12345 + 1  # rngSeed + runNumber (rayTraceKernel.hpp:100)


@pytest.fixture(scope="module")
def ctx0():
    c = capi.Context(0)
    yield c
    c.close()


def test_philox_known_answer(ctx0):
    L = po.oracle_lib()
    out = np.zeros(4, np.uint32)
    for args in [(0, 0, 0, 0, 0, 0), (0xffffffff,) * 6,
                 (0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)]:
        L.vro_philox4x32(*args, out.ctypes.data)
        assert (ctx0.debug_philox(*args) == out).all()
    # Random123 known-answer vectors for philox4x32-10
    assert list(ctx0.debug_philox(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c,
                                                         0x9b00dbd8]
    assert list(ctx0.debug_philox(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e,
                                  0x03707344)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_math_bit_exact(ctx0):
    L = po.oracle_lib()
    rng = np.random.default_rng(1)
    u = np.concatenate([rng.random(200000, dtype=np.float32),
                        np.arange(0, 1, 1 / 4096, dtype=np.float32),
                        np.array([0, 0.25, 0.5, 0.75, 0.124999, 0.125, 0.99999994], np.float32)])
    s = np.zeros_like(u)
    c = np.zeros_like(u)
    L.vro_math_sincos2pi(u.ctypes.data, len(u), s.ctypes.data, c.ctypes.data)
    g = ctx0.debug_math(0, u)
    assert (g[:len(u)].view(np.uint32) == s.view(np.uint32)).all()
    assert (g[len(u):].view(np.uint32) == c.view(np.uint32)).all()
    assert np.abs(s - np.sin(2 * np.pi * u.astype(np.float64))).max() < 5e-7
    for e in (0.5, 1.0 / 101.0, 1.0 / 3.0, 0.9):
        o = np.zeros_like(u)
        L.vro_math_pow(u.ctypes.data, np.float32(e), len(u), o.ctypes.data)
        g = ctx0.debug_math(1, u, e)
        assert (g.view(np.uint32) == o.view(np.uint32)).all()
        assert np.abs(o - u.astype(np.float64) ** float(np.float32(e))).max() < 1e-6
    o = np.zeros_like(u)
    L.vro_math_acos(u.ctypes.data, len(u), o.ctypes.data)
    g = ctx0.debug_math(2, u)
    assert (g.view(np.uint32) == o.view(np.uint32)).all()
    assert np.abs(o - np.arccos(u.astype(np.float64))).max() < 1e-6


@pytest.mark.parametrize("kind,D", [(0, 3), (0, 2), (1, 3), (2, 3), (2, 2)])
def test_reflection_bit_exact(ctx0, kind, D):
    L = po.oracle_lib()
    m = 20000
    inc = np.deg2rad(75.0)
    d = np.array([0.0, -np.sin(inc), -np.cos(inc)], np.float32)
    n = np.array([0.0, 0.0, 1.0], np.float32)
    if D == 2:
        d = np.array([np.sin(inc), -np.cos(inc), 0.0], np.float32)
        n = np.array([0.0, 1.0, 0.0], np.float32)
    cone = np.float32(np.deg2rad(85.0))
    o = np.zeros((m, 3), np.float32)
    L.vro_reflect(kind, D, d.ctypes.data, n.ctypes.data, cone, 77, 1000, m, o.ctypes.data)
    g = ctx0.debug_reflect(kind, D, d, n, cone, 77, 1000, m)
    assert (g.view(np.uint32) == o.view(np.uint32)).all()
    assert np.allclose(np.linalg.norm(o, axis=1), 1, atol=1e-5)
    assert (o @ n > -1e-6).all()


CASES = ["disk3D", "triangle3D", "disk2D", "trench", "trench_ion", "holes", "plane", "sphere3D",
         "sphere2D"]


@pytest.fixture(scope="module", params=CASES)
def pair(request):
    c = common.case(request.param)
    orc = common.make_oracle(c)
    ctx, src, st = common.make_gpu(c)
    yield c, orc, ctx, src, st
    ctx.close()


def test_host_setup_matches_oracle(pair):
    c, orc, ctx, src, st = pair
    lo, hi = st["bbox"]
    assert (orc.bbox() == np.stack([lo, hi])).all()
    if c["geo"] == "disk":
        off, idx = orc.neighbors()
        assert (off == st["nb"][0]).all() and (idx == st["nb"][1]).all()
    else:
        assert (orc.normals().view(np.uint32) == st["normals"].view(np.uint32)).all()


def test_source_rays_bit_exact(pair):
    c, orc, ctx, src, st = pair
    m = 50000
    cfg_o = orc.config(10**6, SEED)
    ro = orc.source_rays(common.oracle_particle(c), cfg_o, 1234, m)
    rg = ctx.debug_source_rays(src, common.gpu_particle(c), host.config(10**6, SEED), 1234, m)
    assert (ro.view(np.uint32) == rg.view(np.uint32)).all()


def _first_bounce_rays(orc, rays, prim, t, c, seed=5):
    """Secondary rays leaving the first hit points in random upward directions."""
    rng = np.random.default_rng(seed)
    hitp = rays[:, :3] + rays[:, 3:] * t[:, None]
    nrm = orc.normals()[np.where(prim == 0xFFFFFFFF, 0, prim)]
    d = rng.normal(size=(len(rays), 3)).astype(np.float32)
    if c["D"] == 2:
        d[:, 2] = 0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    flip = (d * nrm).sum(1) < 0
    d[flip] *= -1
    return np.ascontiguousarray(np.concatenate([hitp, d], 1), np.float32)


def test_fixed_ray_set_ids_t_and_neighbour_sets(pair):
    """BASELINE correctness level 1: hit primitive IDs, hit distances and
    neighbour-disk sets for a fixed ray set (primary rays + one bounce)."""
    c, orc, ctx, src, st = pair
    m = 200000 if c["name"] != "disk2D" else 50000
    rays = orc.source_rays(common.oracle_particle(c), orc.config(10**7, SEED), 0, m)
    for leg in range(2):
        go, po_, to, _ = orc.intersect(rays)
        gg, pg, tg, cnt_g, nb_g = ctx.debug_intersect(rays, nb_cap=24)
        assert (go == gg).all(), "geomID mismatch on %d rays" % int((go != gg).sum())
        assert (po_ == pg).all(), "primID mismatch on %d rays" % int((po_ != pg).sum())
        hit = go != 0xFFFFFFFF
        assert (to[hit].view(np.uint32) == tg[hit].view(np.uint32)).all()  # 0 ulp
        if c["geo"] == "disk":
            geo_hit = np.where(go == 1, po_, 0xFFFFFFFF).astype(np.uint32)
            cnt_o, nb_o = orc.neighbor_hits(rays, geo_hit, cap=24)
            assert cnt_o.max() <= 24
            assert (cnt_o == np.where(go == 1, cnt_g, 0)).all()
            sel = go == 1
            assert (np.sort(nb_o[sel], 1) == np.sort(nb_g[sel], 1)).all()
        keep = go == 1
        rays = _first_bounce_rays(orc, rays[keep], po_[keep], to[keep], c)


@pytest.mark.parametrize("shards", [1, 3])
def test_full_walk_flux_bit_exact(pair, shards):
    """Whole Monte Carlo walks: every ray takes the same path on both sides, so
    the fixed-point per-primitive sums and all TraceInfo counters are equal."""
    c, orc, ctx, src, st = pair
    num = 300000 if c["name"] != "disk2D" else 100000
    fo, io = orc.trace(common.oracle_particle(c), orc.config(num, SEED))
    total = np.zeros(orc.n, np.uint64)
    infos = []
    bounds = np.linspace(0, num, shards + 1).astype(np.int64)
    for k in range(shards):  # ray-index shards, as the multi-GPU path splits them
        cfg = host.config(num, SEED, int(bounds[k]), int(bounds[k + 1]))
        ctx.trace_device(src, [common.gpu_particle(c)], cfg, sync=True)
        total += ctx.flux_download_fixed()[0]
        infos.append(ctx.flux_download()[1][0])
    d = io.as_dict()
    assert sum(i.totalRaysTraced for i in infos) == d["totalTraces"]
    assert sum(i.geometryHits for i in infos) == d["geoHits"]
    assert sum(i.nonGeometryHits for i in infos) == d["nonGeoHits"]
    assert sum(i.boundaryHits for i in infos) == d["boundaryHits"]
    assert sum(i.reflections for i in infos) == d["reflections"]
    assert sum(i.raysTerminated for i in infos) == d["raysTerminated"]
    assert (total == fo).all(), "%d of %d primitives differ" % (int((total != fo).sum()), orc.n)


def test_vr_trace_host_api_and_determinism(pair):
    c, orc, ctx, src, st = pair
    cfg = host.config(50000, SEED)
    f1, i1 = ctx.trace(src, [common.gpu_particle(c)], cfg)
    f2, i2 = ctx.trace(src, [common.gpu_particle(c)], cfg)
    assert (f1 == f2).all()  # bitwise, like tests/rngSeed/rngSeed.cpp:48-51
    assert i1[0].numRays == 50000 and i1[0].totalRaysTraced == i2[0].totalRaysTraced
    fo, _ = orc.trace(common.oracle_particle(c), orc.config(50000, SEED))
    assert (f1[0] == fo / po.FLUX_SCALE).all()


def test_multi_particle_label_major(pair):
    c, orc, ctx, src, st = pair
    if c["geo"] != "disk" or c["D"] != 3:
        pytest.skip("multi-particle case is the 3D disk trench")
    parts = [capi.ParticleDesc(0, 0.1, 1.0, 0.0), capi.ParticleDesc(2, 0.5, 100.0, np.deg2rad(85.0))]
    cfg = host.config(100000, SEED)
    f, infos = ctx.trace(src, parts, cfg)
    for k, p in enumerate(parts):
        oc = orc.config(100000, SEED, stream=k)
        fo, io = orc.trace(po.Particle(p.kind, p.sticking, p.sourcePower, p.coneMinAngle), oc)
        assert (f[k] == fo / po.FLUX_SCALE).all()
        assert infos[k].geometryHits == io.geoHits


def test_flux_postprocess_on_device(pair):
    """normalizeFlux(SOURCE) + smoothFlux(1) on the device against the oracle's
    float restatement of rayTraceDisk.hpp:103-193, bit for bit."""
    c, orc, ctx, src, st = pair
    num = 80000
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(num, SEED), sync=True)
    fo, _ = orc.trace(common.oracle_particle(c), orc.config(num, SEED))
    raw = (fo / po.FLUX_SCALE).astype(np.float32)
    assert (ctx.flux_postprocess() == raw).all()
    rng = np.random.default_rng(5)
    areas = (0.5 + rng.random(orc.n)).astype(np.float32)
    lo, hi = st["bbox"]
    _, first, second, _, _ = host.trace_settings(c["source_dir"])
    area = np.float32(hi[first] - lo[first])
    if c["D"] == 3:
        area = np.float32(area * np.float32(hi[second] - lo[second]))
    norm = np.float32(area / np.float32(num))
    expect = orc.normalize_flux_source(raw, areas, num)
    got = ctx.flux_postprocess(0, areas, norm, smooth=False)
    assert (got.view(np.uint32) == expect.view(np.uint32)).all()
    got = ctx.flux_postprocess(0, areas, norm, smooth=True)
    if c["geo"] == "disk":
        expect = orc.smooth_flux(expect)
    assert (got.view(np.uint32) == expect.view(np.uint32)).all()


def test_flux_postprocess_max_and_wide_smoothing(pair):
    """normalizeFlux(MAX) (rayTraceDisk.hpp:110-118 in double, rayTraceTriangle.hpp:99-107) and
    smoothFlux(k) with k > 1 -- a neighbourhood of k * 2 * radius built on the device
    (rayTraceDisk.hpp:160-168) -- against the oracle's restatement, bit for bit."""
    c, orc, ctx, src, st = pair
    num = 60000
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(num, SEED), sync=True)
    fo, _ = orc.trace(common.oracle_particle(c), orc.config(num, SEED))
    raw = (fo / po.FLUX_SCALE).astype(np.float32)
    assert (ctx.flux_postprocess_ex() == raw).all()
    rng = np.random.default_rng(9)
    areas = (0.5 + rng.random(orc.n)).astype(np.float32)
    r = host.disk_radius(c["grid_delta"], c["D"]) if c["geo"] == "disk" else np.float32(0)
    total = float(np.float32(r * r)) * np.pi  # diskRadius_ * diskRadius_ * M_PI
    expect = orc.normalize_flux_max(raw, areas)
    got = ctx.flux_postprocess_ex(0, areas, capi.NORM_MAX, total)
    assert (got.view(np.uint32) == expect.view(np.uint32)).all()
    if c["geo"] != "disk":
        # triangles are not smoothed (rayTraceTriangle.hpp has no neighbourhood)
        assert (ctx.flux_postprocess_ex(0, areas, capi.NORM_MAX, total, 2, 1.0) == expect).all()
        return
    for k in (1, 2, 3):
        e = orc.smooth_flux(expect, k)
        g = ctx.flux_postprocess_ex(0, areas, capi.NORM_MAX, total, k, r)
        assert (g.view(np.uint32) == e.view(np.uint32)).all(), "k = %d" % k
    # SOURCE through the general call equals the first interface
    norm = 0.37
    a = ctx.flux_postprocess(0, areas, norm, smooth=True)
    b = ctx.flux_postprocess_ex(0, areas, capi.NORM_SOURCE, float(np.float32(norm)), 1, r)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()


@pytest.mark.parametrize("name", ["disk3D", "disk2D", "trench", "holes", "plane"])
def test_neighbor_build_on_device(name):
    """PointNeighborhood on the device == host build == oracle, and a trace
    through device-built lists gives the same flux words."""
    c = common.case(name)
    orc = common.make_oracle(c)
    st = common.product_setup(c)
    ctx = capi.Context(0)
    ctx.set_disks(st["xyzr"], st["normals"])  # no lists
    r = host.disk_radius(c["grid_delta"], c["D"])
    ctx.build_neighbors_device(c["D"], c["points"], np.float32(2) * r)
    off, idx = ctx.get_neighbors()
    off_o, idx_o = orc.neighbors()
    assert (off == off_o).all() and (idx == idx_o).all()
    lo, hi = st["bbox"]
    _, first, second, _, _ = host.trace_settings(c["source_dir"])
    cond2 = c["bc"][second] if c["D"] == 3 else capi.BOUNDARY_IGNORE
    ctx.set_boundary(lo, hi, first, second, c["bc"][first], cond2, c["D"])
    ctx.commit()
    src = host.source_desc(lo, hi, c["source_dir"])
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(60000, SEED), sync=True)
    fo, _ = orc.trace(common.oracle_particle(c), orc.config(60000, SEED))
    assert (ctx.flux_download_fixed()[0] == fo).all()
    ctx.close()


@pytest.mark.parametrize("name,power", [("trench", 1.0), ("trench_ion", 100.0), ("disk2D", 1.0),
                                        ("triangle3D", 3.0)])
def test_grid_source(name, power):
    """SourceGrid (raySourceGrid.hpp): origins idx % numPoints, cos^n direction;
    source rays and whole walks bit-equal to the oracle."""
    c = common.case(name)
    c["power"] = power
    orc = common.make_oracle(c)
    ctx, _, st = common.make_gpu(c)
    lo, hi = st["bbox"]
    if c["D"] == 3:  # 37 x 23 origins on the source plane
        gx, gy = np.meshgrid(np.linspace(lo[0] + 1e-4, hi[0] - 1e-4, 37, dtype=np.float32),
                             np.linspace(lo[1] + 1e-4, hi[1] - 1e-4, 23, dtype=np.float32))
        grid = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, hi[2], np.float32)], 1)
    else:
        grid = np.stack([np.linspace(lo[0] + 1e-4, hi[0] - 1e-4, 50, dtype=np.float32),
                         np.full(50, hi[1], np.float32), np.zeros(50, np.float32)], 1)
    orc.set_source_grid(grid)
    ctx.set_source_grid(grid)
    src = host.source_desc(lo, hi, c["source_dir"], use_grid=True)
    m = 30000
    ro = orc.source_rays(common.oracle_particle(c), orc.config(10**6, SEED), 77, m)
    rg = ctx.debug_source_rays(src, common.gpu_particle(c), host.config(10**6, SEED), 77, m)
    assert (ro.view(np.uint32) == rg.view(np.uint32)).all()
    num = 100000
    fo, io = orc.trace(common.oracle_particle(c), orc.config(num, SEED))
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(num, SEED), sync=True)
    assert (ctx.flux_download_fixed()[0] == fo).all()
    assert ctx.flux_download()[1][0].totalRaysTraced == io.totalTraces
    # a grid source without grid points is a state error, not a silent fallback
    ctx.set_source_grid(None)
    with pytest.raises(capi.VrError):
        ctx.trace_device(src, [common.gpu_particle(c)], host.config(100, SEED), sync=True)
    ctx.close()
