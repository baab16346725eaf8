"""-m gpu: edge cases of the C ABI -- tiny scenes (a leaf-only BVH), empty and
one-ray shards, call-order and argument errors (status codes + messages, never
a silent fallback)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import common
from viennaray_b200 import capi, host

pytestmark = pytest.mark.gpu
SEED = 17


def tiny_case(n):
    pts = np.zeros((n, 3), np.float32)
    pts[:, 0] = np.arange(n) * 0.5
    nrm = np.zeros((n, 3), np.float32)
    nrm[:, 2] = 1
    return dict(name="tiny", D=3, geo="disk", points=pts, normals=nrm, grid_delta=0.5,
                bc=[0, 1, 2], source_dir=host.POS_Z, kind=0, sticking=0.4, power=1.0, cone=0.0)


@pytest.mark.parametrize("n", [1, 2, 4, 5, 8, 10, 11, 17])
def test_tiny_scenes(n):
    """n <= 10 (VR_LEAF_MAX): the whole BVH is one leaf reference; n = 11: the first inner node."""
    c = tiny_case(n)
    orc = common.make_oracle(c)
    ctx, src, _ = common.make_gpu(c)
    fo, io = orc.trace(common.oracle_particle(c), orc.config(20000, SEED))
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(20000, SEED), sync=True)
    assert (ctx.flux_download_fixed()[0] == fo).all()
    assert ctx.flux_download()[1][0].totalRaysTraced == io.totalTraces
    rays = orc.source_rays(common.oracle_particle(c), orc.config(1000, SEED), 0, 1000)
    go, pr, to, _ = orc.intersect(rays)
    gg, pg, tg, _, _ = ctx.debug_intersect(rays)
    assert (go == gg).all() and (pr == pg).all()
    ctx.close()


def test_empty_and_single_ray_shards():
    c = common.case("plane")
    orc = common.make_oracle(c)
    ctx, src, _ = common.make_gpu(c)
    part = common.gpu_particle(c)
    flux, infos = ctx.trace(src, [part], host.config(1000, SEED, 500, 500))  # empty shard
    assert flux.sum() == 0 and infos[0].totalRaysTraced == 0
    flux, infos = ctx.trace(src, [part], host.config(0, SEED))  # no rays at all
    assert flux.sum() == 0
    total = np.zeros(orc.n, np.uint64)
    for k in range(5):  # five one-ray shards == the oracle's five rays
        ctx.trace_device(src, [part], host.config(5, SEED, k, k + 1), sync=True)
        total += ctx.flux_download_fixed()[0]
    fo, _ = orc.trace(common.oracle_particle(c), orc.config(5, SEED))
    assert (total == fo).all()
    ctx.close()


def test_axis_parallel_rays():
    """zero direction components (the slab test's reciprocal) and rays in a primitive's plane"""
    c = common.case("trench")
    orc = common.make_oracle(c)
    ctx, _, _ = common.make_gpu(c)
    rng = np.random.default_rng(2)
    m = 6000
    org = np.stack([rng.uniform(0, 59, m), rng.uniform(-9, 9, m), rng.uniform(-39, 5, m)], 1)
    d = np.zeros((m, 3))
    d[np.arange(m), rng.integers(0, 3, m)] = rng.choice([-1.0, 1.0], m)
    rays = np.ascontiguousarray(np.concatenate([org, d], 1), np.float32)
    go, pr, to, _ = orc.intersect(rays)
    gg, pg, tg, _, _ = ctx.debug_intersect(rays)
    assert (go == gg).all() and (pr == pg).all()
    hit = go != 0xFFFFFFFF
    assert (to[hit].view(np.uint32) == tg[hit].view(np.uint32)).all()
    ctx.close()


def test_call_order_and_argument_errors():
    ctx = capi.Context(0)
    c = common.case("plane")
    st = common.product_setup(c)
    src = host.source_desc(*st["bbox"], c["source_dir"])
    part = common.gpu_particle(c)
    with pytest.raises(capi.VrError) as e:  # nothing set
        ctx.commit()
    assert e.value.code == 3 and "no geometry" in str(e.value)
    ctx.set_disks(st["xyzr"], st["normals"], *st["nb"])
    with pytest.raises(capi.VrError) as e:  # boundary missing
        ctx.commit()
    assert e.value.code == 3
    with pytest.raises(capi.VrError) as e:  # trace before commit
        ctx.trace(src, [part], host.config(10, SEED))
    assert e.value.code == 3 and "not committed" in str(e.value)
    lo, hi = st["bbox"]
    with pytest.raises(capi.VrError) as e:
        ctx.set_boundary(lo, hi, 0, 0, 0, 0, 3)  # same axis twice
    assert e.value.code == 2
    ctx.set_boundary(lo, hi, 0, 1, 0, 0, 3)
    ctx.commit()
    with pytest.raises(capi.VrError) as e:  # unknown particle kind: no host fallback
        ctx.trace(src, [capi.ParticleDesc(7, 0.5, 1.0, 0.0)], host.config(10, SEED))
    assert e.value.code == 4 and "built-in" in str(e.value)
    with pytest.raises(capi.VrError) as e:  # shard outside the job
        ctx.trace(src, [part], host.config(10, SEED, 5, 20))
    assert e.value.code == 2
    # post-processing: before any trace, then with arguments that do not go together
    with pytest.raises(capi.VrError) as e:
        ctx.flux_postprocess_ex()
    assert e.value.code == 3
    ctx.trace(src, [part], host.config(1000, SEED))
    areas = np.ones(len(st["xyzr"]), np.float32)
    with pytest.raises(capi.VrError) as e:  # normalisation without areas
        ctx.flux_postprocess_ex(0, None, capi.NORM_MAX, 1.0)
    assert e.value.code == 2
    with pytest.raises(capi.VrError) as e:  # wider smoothing without the radius
        ctx.flux_postprocess_ex(0, areas, capi.NORM_SOURCE, 1.0, 2, 0.0)
    assert e.value.code == 2
    with pytest.raises(capi.VrError) as e:  # particle index out of range
        ctx.flux_postprocess_ex(3, areas, capi.NORM_SOURCE, 1.0)
    assert e.value.code == 2
    assert ctx.flux_postprocess_ex(0, areas, capi.NORM_SOURCE, 1.0, 3, 0.5).shape == (len(areas),)
    bad = st["nb"][1].copy()
    bad[0] = 10**6
    with pytest.raises(capi.VrError) as e:  # neighbour index out of range
        ctx.set_disks(st["xyzr"], st["normals"], st["nb"][0], bad)
    assert e.value.code == 2
    with pytest.raises(capi.VrError):  # device ordinal out of range
        capi.Context(1000)
    ctx.close()


def test_reuse_context_across_geometries():
    """one context, scene replaced several times (ViennaPS re-traces an evolving surface)"""
    ctx = capi.Context(0)
    for name in ("plane", "trench", "triangle3D", "plane"):
        c = common.case(name)
        st = common.product_setup(c)
        orc = common.make_oracle(c)
        if c["geo"] == "disk":
            ctx.set_disks(st["xyzr"], st["normals"], *st["nb"])
        else:
            ctx.set_triangles(st["verts"], st["tris"], st["normals"])
        lo, hi = st["bbox"]
        _, first, second, _, _ = host.trace_settings(c["source_dir"])
        ctx.set_boundary(lo, hi, first, second, c["bc"][first], c["bc"][second], 3)
        ctx.commit()
        src = host.source_desc(lo, hi, c["source_dir"])
        ctx.trace_device(src, [common.gpu_particle(c)], host.config(30000, SEED), sync=True)
        fo, _ = orc.trace(common.oracle_particle(c), orc.config(30000, SEED))
        assert (ctx.flux_download_fixed()[0] == fo).all(), name
    ctx.close()


@pytest.mark.parametrize("shift", [(5000.0, -3000.0, 800.0), (-1.0e5, 2.0e5, -4.0e4)])
def test_scene_far_from_the_origin(shift):
    """coordinates that are large against the features: node quantisation, the boundary
    pre-check margins and the sky map must stay conservative (results stay bit-equal)"""
    c = common.case("trench")
    c["points"] = (c["points"] + np.asarray(shift, np.float32)).astype(np.float32)
    orc = common.make_oracle(c)
    ctx, src, _ = common.make_gpu(c)
    num = 80000
    fo, io = orc.trace(common.oracle_particle(c), orc.config(num, SEED))
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(num, SEED), sync=True)
    assert (ctx.flux_download_fixed()[0] == fo).all()
    assert ctx.flux_download()[1][0].totalRaysTraced == io.totalTraces
    rays = orc.source_rays(common.oracle_particle(c), orc.config(num, SEED), 0, 20000)
    go, pr, to, _ = orc.intersect(rays)
    gg, pg, tg, _, _ = ctx.debug_intersect(rays)
    assert (go == gg).all() and (pr == pg).all()
    ctx.close()


@pytest.mark.parametrize("name", ["trench", "triangle3D"])
@pytest.mark.parametrize("n", [0, 3, 9, 17, 40])
def test_wide_nodes_option(monkeypatch, name, n):
    """VR_BVH_WIDE=1 (4-wide nodes, read when the scene is committed) changes the order of the
    traversal only: IDs, t and the whole-walk flux words stay those of the oracle.  n > 0:
    tiny scenes whose wide root has unused entries."""
    monkeypatch.setenv("VR_BVH_WIDE", "1")
    c = tiny_case(n) if n else common.case(name)
    if n and name != "trench":
        pytest.skip("tiny scenes are disk scenes")
    orc = common.make_oracle(c)
    ctx, src, _ = common.make_gpu(c)
    rays = 20000 if n else 200000
    fo, io = orc.trace(common.oracle_particle(c), orc.config(rays, SEED))
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(rays, SEED), sync=True)
    assert (ctx.flux_download_fixed()[0] == fo).all()
    assert ctx.flux_download()[1][0].totalRaysTraced == io.totalTraces
    probe = orc.source_rays(common.oracle_particle(c), orc.config(4000, SEED), 0, 4000)
    go, pr, to, _ = orc.intersect(probe)
    gg, pg, tg, _, _ = ctx.debug_intersect(probe)
    assert (go == gg).all() and (pr == pg).all() and (to == tg).all()
    ctx.close()


def test_l2_read_bandwidth_diagnostic():
    """vr_debug_l2_read_bandwidth returns a plausible rate (above HBM's, below 100 TB/s) and
    the device's L2 size; bad arguments are an error, not a crash."""
    ctx = capi.Context(0)
    gbps, l2 = ctx.l2_read_bandwidth(32 << 20, 10)
    assert l2 >= 32 << 20
    assert 3000.0 < gbps < 100000.0, gbps
    with pytest.raises(capi.VrError):
        ctx.l2_read_bandwidth(0, 10)
    ctx.close()


def test_long_neighbour_lists():
    """A plane sampled four times denser than its disk radius asks for: every disk has far more
    than the eight neighbours of the row format (vr_internal.h, nbRow), so the spread walks the
    row and several rounds of the CSR behind it; both reflective boundaries are hit."""
    from viennaray_b200 import scenes
    p, n = scenes.plane_grid(0.125, 2.0)
    c = dict(name="dense", D=3, geo="disk", points=p, normals=n, grid_delta=0.5, bc=[0, 0, 0],
             source_dir=host.POS_Z, kind=0, sticking=0.5, power=1.0, cone=0.0)
    st = common.product_setup(c)
    counts = np.diff(st["nb"][0])
    assert counts.max() > 20 and counts.min() >= 8
    orc = common.make_oracle(c)
    ctx, src, _ = common.make_gpu(c)
    fo, io = orc.trace(common.oracle_particle(c), orc.config(30000, SEED))
    ctx.trace_device(src, [common.gpu_particle(c)], host.config(30000, SEED), sync=True)
    assert (ctx.flux_download_fixed()[0] == fo).all()
    assert ctx.flux_download()[1][0].totalRaysTraced == io.totalTraces
    ctx.close()
