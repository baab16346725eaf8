"""-m gpu: the option matrix of the hot path against the oracle, bit for bit --
all six source directions, every pairing of boundary conditions, the three
device particles, tilted primary directions, the hit-count limits
(maxReflections / maxBoundaryHits, rayTraceKernel.hpp:207-211,320-324) and the
2D path, on small scenes (a trench turned towards each source side)."""
import itertools

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import common
from viennaray_b200 import capi, host, scenes

pytestmark = pytest.mark.gpu
SEED = 991
PARTICLES = {"diffuse": (0, 0.3, 1.0, 0.0), "specular": (1, 0.5, 5.0, 0.0),
             "coned": (2, 0.5, 50.0, float(np.deg2rad(80.0)))}


def turned_trench(source_dir):
    """The reduced C4 trench with its opening turned towards the given source side."""
    p, n, gd = scenes.trench(num_slices=24, half_width=5, depth=14, half_extent=16)
    axis, first, second, min_max, _ = host.trace_settings(source_dir)
    sign = 1.0 if min_max else -1.0
    pts = np.zeros_like(p)
    nrm = np.zeros_like(n)
    # trench frame: x along the trench, y across, z up  ->  first, second, source axis
    pts[:, first], pts[:, second], pts[:, axis] = p[:, 0], p[:, 1], sign * p[:, 2]
    nrm[:, first], nrm[:, second], nrm[:, axis] = n[:, 0], n[:, 1], sign * n[:, 2]
    return np.ascontiguousarray(pts), np.ascontiguousarray(nrm), gd


def run_pair(c, num, max_refl=0xFFFFFFFF, max_bhits=1000, primary_dir=None, mfp=0.0, wdist=False,
             materials=None, table=None):
    orc = common.make_oracle(c)
    ctx, src, _ = common.make_gpu(c, primary_dir=primary_dir, material_ids=materials)
    po_part = common.oracle_particle(c)
    po_part.meanFreePath = mfp
    g_part = common.gpu_particle(c)
    g_part.meanFreePath = mfp
    if materials is not None:
        orc.set_material_ids(materials)
    if table is not None:
        po_part.set_sticking_by_material(table)
        g_part.set_sticking_by_material(table)
    fo, io = orc.trace(po_part,
                       orc.config(num, SEED, max_reflections=max_refl, max_boundary_hits=max_bhits,
                                  primary_dir=primary_dir, wdist=wdist))
    ctx.trace_device(src, [g_part],
                     host.config(num, SEED, max_reflections=max_refl, max_boundary_hits=max_bhits,
                                 wdist=wdist), sync=True)
    fg = ctx.flux_download_fixed()[0]
    ig = ctx.flux_download()[1][0]
    launches = ctx.last_launch_count()
    ctx.close()
    d = io.as_dict()
    d["launches"], d["iterations"] = launches
    assert (fg == fo).all(), "%d primitives differ" % int((fg != fo).sum())
    assert (ig.totalRaysTraced, ig.geometryHits, ig.nonGeometryHits, ig.boundaryHits,
            ig.reflections, ig.raysTerminated, ig.particleHits) == \
        (d["totalTraces"], d["geoHits"], d["nonGeoHits"], d["boundaryHits"], d["reflections"],
         d["raysTerminated"], d["particleHits"])
    return d


@pytest.mark.parametrize("source_dir", range(6))
@pytest.mark.parametrize("bc_pair", list(itertools.product(range(3), repeat=2)))
def test_source_sides_and_boundary_conditions(source_dir, bc_pair):
    pts, nrm, gd = turned_trench(source_dir)
    _, first, second, _, _ = host.trace_settings(source_dir)
    bc = [2, 2, 2]
    bc[first], bc[second] = bc_pair
    kind, st, pw, cone = PARTICLES[("diffuse", "specular", "coned")[(source_dir + bc_pair[0]) % 3]]
    c = dict(name="turned", D=3, geo="disk", points=pts, normals=nrm, grid_delta=gd, bc=bc,
             source_dir=source_dir, kind=kind, sticking=st, power=pw, cone=cone)
    d = run_pair(c, 60000)
    assert d["geoHits"] > 0


@pytest.mark.parametrize("pname", list(PARTICLES))
@pytest.mark.parametrize("primary", [(0.3, 0.2, -1.0), (-0.6, 0.0, -0.8)])
def test_tilted_primary_direction(pname, primary):
    pts, nrm, gd = turned_trench(host.POS_Z)
    kind, st, pw, cone = PARTICLES[pname]
    c = dict(name="tilted", D=3, geo="disk", points=pts, normals=nrm, grid_delta=gd, bc=[1, 0, 2],
             source_dir=host.POS_Z, kind=kind, sticking=st, power=pw, cone=cone)
    run_pair(c, 60000, primary_dir=np.asarray(primary, np.float32))


@pytest.mark.parametrize("max_refl,max_bhits", [(0, 1000), (1, 1000), (3, 2), (0xFFFFFFFF, 0),
                                                (0xFFFFFFFF, 1)])
def test_hit_count_limits(max_refl, max_bhits):
    pts, nrm, gd = turned_trench(host.POS_Z)
    c = dict(name="limits", D=3, geo="disk", points=pts, normals=nrm, grid_delta=gd, bc=[0, 1, 2],
             source_dir=host.POS_Z, kind=0, sticking=0.05, power=1.0, cone=0.0)
    d = run_pair(c, 60000, max_refl=max_refl, max_bhits=max_bhits)
    if max_refl < 10:
        assert d["raysTerminated"] > 0


@pytest.mark.parametrize("source_dir", [host.POS_X, host.NEG_X, host.POS_Y, host.NEG_Y])
@pytest.mark.parametrize("cond", range(3))
def test_2d_sources_and_boundaries(source_dir, cond):
    c0 = common.case("disk2D")  # trench profile in x-y, opening towards +y
    pts, nrm = c0["points"].copy(), c0["normals"].copy()
    axis, first, _, min_max, _ = host.trace_settings(source_dir)
    sign = 1.0 if min_max else -1.0
    p2, n2 = np.zeros_like(pts), np.zeros_like(nrm)
    p2[:, first], p2[:, axis] = pts[:, 0], sign * pts[:, 1]
    n2[:, first], n2[:, axis] = nrm[:, 0], sign * nrm[:, 1]
    bc = [2, 2, 2]
    bc[first] = cond
    c = dict(c0, points=np.ascontiguousarray(p2), normals=np.ascontiguousarray(n2), bc=bc,
             source_dir=source_dir, kind=(0, 2, 1)[cond], sticking=0.2, power=(1.0, 20.0, 3.0)[cond],
             cone=float(np.deg2rad(70.0)))
    run_pair(c, 40000)


@pytest.mark.parametrize("bc_pair", [(0, 0), (1, 1), (2, 2), (0, 1)])
@pytest.mark.parametrize("pname", list(PARTICLES))
def test_triangles_matrix(bc_pair, pname):
    c = common.case("triangle3D")
    kind, st, pw, cone = PARTICLES[pname]
    c.update(bc=[bc_pair[0], bc_pair[1], 2], kind=kind, sticking=st, power=pw, cone=cone)
    run_pair(c, 40000)


@pytest.mark.parametrize("mfp", [2.0, 20.0])
@pytest.mark.parametrize("name,bc", [("trench", [1, 1, 2]), ("trench", [0, 2, 2]), ("disk2D", [1, 2, 2]),
                                     ("triangle3D", [0, 0, 2])])
def test_mean_free_path_scattering(name, bc, mfp):
    """rayTraceKernel.hpp:179-203: scatter draws, particleHits counter."""
    c = common.case(name)
    c["bc"] = bc
    d = run_pair(c, 50000, mfp=mfp)
    assert d["particleHits"] > 0


@pytest.mark.parametrize("name,pname", [("trench", "diffuse"), ("trench", "coned"), ("holes", "specular"),
                                        ("disk2D", "diffuse")])
def test_distance_weighted_spread(name, pname):
    """VIENNARAY_USE_WDIST (rayTraceKernel.hpp:258-296)."""
    c = common.case(name)
    kind, st, pw, cone = PARTICLES[pname]
    c.update(kind=kind, sticking=st, power=pw, cone=cone)
    run_pair(c, 60000, wdist=True)
    # and both options together
    run_pair(c, 30000, wdist=True, mfp=10.0)


@pytest.mark.parametrize("name,table", [("trench", (0.05, 0.6)), ("trench_ion", (0.9, 0.1, 0.4)),
                                        ("triangle3D", (0.02, 0.5)), ("disk2D", (0.3,)),
                                        ("holes", (0.01, 1.0))])
def test_sticking_by_material(name, table):
    """The materialId of the hit primitive selects the sticking probability
    (rayTraceKernel.hpp:310-313; the vr_particle_desc table).  IDs outside the table --
    here the negative one and the one past its end -- keep the particle's constant."""
    c = common.case(name)
    n = len(c["points"]) if c["geo"] == "disk" else len(c["tris"])
    mats = (np.arange(n) * 7919 % (len(table) + 2) - 1).astype(np.int32)  # -1 .. len(table)
    d = run_pair(c, 60000, materials=mats, table=table)
    d0 = run_pair(c, 60000, materials=mats)  # IDs without a table: the constant sticking
    assert d["reflections"] != d0["reflections"]
    run_pair(c, 30000, materials=mats, table=table, wdist=(c["geo"] == "disk"), mfp=15.0)


@pytest.mark.parametrize("name", ["trench", "trench_ion", "triangle3D", "sphere2D"])
@pytest.mark.parametrize("pool,tail", [(4096, 0), (4096, 262144), (1 << 16, 0), (0, 0)])
def test_wavefront_schedule_does_not_change_results(monkeypatch, name, pool, tail):
    """The small cases above end up in the single-launch tail kernel almost at once.  Here
    the schedule is forced the other ways (both read at vr_ctx_create): a pool far smaller
    than the job, so that slots are regenerated in place over many wavefront iterations
    before the source runs dry, and VR_TAIL_RAYS=0, so that the traverse / shade / flip
    kernels -- not the tail kernel -- carry every ray to its end.  Flux words and counters
    must still be the oracle's."""
    if pool:
        monkeypatch.setenv("VR_POOL_SLOTS", str(pool))
    monkeypatch.setenv("VR_TAIL_RAYS", str(tail))
    c = common.case(name)
    d = run_pair(c, 150000)
    if pool == 4096:  # at least one wavefront iteration per pool refill
        assert d["iterations"] >= 150000 // 4096
    if tail == 0:  # and the walk's long tail is made of iterations, not of one tail launch
        assert d["iterations"] >= 20


@pytest.mark.parametrize("mfp,wdist", [(3.0, False), (0.0, True)])
def test_wavefront_schedule_extended_shade(monkeypatch, mfp, wdist):
    """The same for the EXT = 1 shade instantiation (mean free path, WDIST)."""
    monkeypatch.setenv("VR_POOL_SLOTS", "4096")
    monkeypatch.setenv("VR_TAIL_RAYS", "0")
    c = common.case("trench")
    run_pair(c, 100000, mfp=mfp, wdist=wdist)


@pytest.mark.parametrize("name", ["trench", "trench_ion", "holes", "disk3D"])
@pytest.mark.parametrize("tail", [0, 262144])
def test_spread_kernel_option(monkeypatch, name, tail):
    """VR_SPREAD_SPLIT=1: the shade kernel queues every geometry hit and spreadKernel does the
    neighbour tests and the flux adds afterwards (the default for scenes that do not fit L2).
    Integer sums: flux words and counters must not change."""
    monkeypatch.setenv("VR_SPREAD_SPLIT", "1")
    monkeypatch.setenv("VR_TAIL_RAYS", str(tail))
    monkeypatch.setenv("VR_POOL_SLOTS", "16384")
    c = common.case(name)
    run_pair(c, 150000)


@pytest.mark.parametrize("lanes", [1, 2, 3, 4])
@pytest.mark.parametrize("pool", [8192, 0])
def test_wavefront_lanes_do_not_change_results(monkeypatch, lanes, pool):
    """VR_LANES: the jobs of a trace -- one per particle, or the pieces of a particle's ray range
    when there are fewer particles than lanes -- run side by side on their own streams, pools
    and host threads (vr_api.cu, vr_ctx::xlane).  Three particles in one call: flux words and
    counters of every particle must be the oracle's for any number of lanes."""
    monkeypatch.setenv("VR_LANES", str(lanes))
    if pool:
        monkeypatch.setenv("VR_POOL_SLOTS", str(pool))
    c = common.case("trench")
    orc = common.make_oracle(c)
    ctx, src, _ = common.make_gpu(c)
    num = 200000
    descs = [PARTICLES["diffuse"], PARTICLES["coned"], PARTICLES["specular"]]
    ctx.trace_device(src, [capi.ParticleDesc(*d) for d in descs], host.config(num, SEED), sync=True)
    fg = ctx.flux_download_fixed()
    ig = ctx.flux_download()[1]
    for k, d in enumerate(descs):
        fo, io = orc.trace(po.Particle(*d), orc.config(num, SEED, stream=k))
        assert (fg[k] == fo).all(), "particle %d: %d primitives differ" % (k, int((fg[k] != fo).sum()))
        assert (ig[k].totalRaysTraced, ig[k].geometryHits, ig[k].boundaryHits, ig[k].reflections) == \
            (io.totalTraces, io.geoHits, io.boundaryHits, io.reflections)
    # a lone particle: its range is cut into pieces when it is long enough for the pool
    ctx.trace_device(src, [capi.ParticleDesc(*descs[0])], host.config(num, SEED), sync=True)
    fo, io = orc.trace(po.Particle(*descs[0]), orc.config(num, SEED))
    assert (ctx.flux_download_fixed()[0] == fo).all()
    assert ctx.flux_download()[1][0].totalRaysTraced == io.totalTraces
    ctx.close()


@pytest.mark.parametrize("name", ["trench", "trench_ion", "holes", "disk3D"])
def test_boundary_test_without_its_shortcut(monkeypatch, name):
    """VR_BOUNDARY_GENERIC: every boundary hit through the candidate planes and the exact
    triangle tests (the path the shortcut of boundaryTest, vr_trace.cu, falls back to inside
    its margins).  Every other test runs with the shortcut; both must give the oracle's flux
    words and counters (periodic and reflective boundaries, sky-map walks)."""
    monkeypatch.setenv("VR_BOUNDARY_GENERIC", "1")
    run_pair(common.case(name), 150000)
