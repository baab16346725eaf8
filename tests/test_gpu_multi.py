"""-m gpu, needs at least two GPUs (skipped otherwise): several devices behind ONE context
(vr_ctx_create_multi) -- the ray-index range sharded inside the library, one NCCL all-reduce
of the result words, no Python in the data path -- give the single-device flux bit for bit,
through the ctypes binding and through the C++ Trace mirror (VIENNARAY_B200_DEVICES)."""
import os
import subprocess

import numpy as np
import pytest

from tests import common
from viennaray_b200 import capi, host

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "viennaray_b200")


def _num_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_num_gpus() < 2, reason="needs two GPUs")


def _setup(ctx, c, st):
    if c["geo"] == "disk":
        ctx.set_disks(st["xyzr"], st["normals"], st["nb"][0], st["nb"][1])
    else:
        ctx.set_triangles(st["verts"], st["tris"], st["normals"])
    lo, hi = st["bbox"]
    _, first, second, _, _ = host.trace_settings(c["source_dir"])
    cond2 = c["bc"][second] if c["D"] == 3 else capi.BOUNDARY_IGNORE
    ctx.set_boundary(lo, hi, first, second, c["bc"][first], cond2, c["D"])
    ctx.commit()
    return host.source_desc(lo, hi, c["source_dir"])


@needs2
@pytest.mark.parametrize("name", ["trench", "trench_ion", "triangle3D", "holes"])
def test_multi_device_context_equals_single_device(name):
    c = common.case(name)
    st = common.product_setup(c)
    devices = list(range(min(_num_gpus(), 8)))
    num = 1_000_003  # not divisible by the device count: slices differ by one ray
    parts = [common.gpu_particle(c), capi.ParticleDesc(1, 0.4, 3.0, 0.0)]
    single = capi.Context(0)
    src = _setup(single, c, st)
    f1, i1 = single.trace(src, parts, host.config(num, 777))
    w1 = single.flux_download_fixed()
    single.close()
    multi = capi.Context(devices)
    assert multi.num_devices() == len(devices)
    src = _setup(multi, c, st)
    fm, im = multi.trace(src, parts, host.config(num, 777))
    wm = multi.flux_download_fixed()
    assert (wm == w1).all()
    assert (fm == f1).all()
    for a, b in zip(i1, im):
        assert (a.numRays, a.totalRaysTraced, a.geometryHits, a.nonGeometryHits, a.boundaryHits,
                a.reflections, a.raysTerminated) == \
            (b.numRays, b.totalRaysTraced, b.geometryHits, b.nonGeometryHits, b.boundaryHits,
             b.reflections, b.raysTerminated)
    # a shard of the job through the multi-device context, and the post-processing path
    f2, _ = multi.trace(src, parts, host.config(num, 777, 1000, 500_000))
    s2 = capi.Context(0)
    src1 = _setup(s2, c, st)
    g2, _ = s2.trace(src1, parts, host.config(num, 777, 1000, 500_000))
    assert (f2 == g2).all()
    if c["geo"] == "disk":
        areas = np.full(multi.n, 1.5, np.float32)
        assert (multi.flux_postprocess(0, areas, 0.25, True) ==
                s2.flux_postprocess(0, areas, 0.25, True)).all()
    s2.close()
    multi.close()


@needs2
def test_cpp_trace_mirror_on_two_gpus(tmp_path):
    exe = str(tmp_path / "test_host_api")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include", "viennaray_b200"),
                    os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"), "-L", LIBDIR,
                    "-lviennaray_b200", "-Wl,-rpath," + LIBDIR, "-o", exe], check=True)
    outs = {}
    for tag, env in (("one", {}), ("two", {"VIENNARAY_B200_DEVICES": "0,1"})):
        d = tmp_path / tag
        d.mkdir()
        res = subprocess.run([exe, "gpu", str(d)], capture_output=True, text=True,
                             env=dict(os.environ, **env))
        assert res.returncode == 0, res.stderr + res.stdout
        outs[tag] = {f: np.fromfile(str(d / f), np.float32) for f in sorted(os.listdir(str(d)))}
    assert outs["one"].keys() == outs["two"].keys() and len(outs["one"]) >= 3
    for f in outs["one"]:
        assert (outs["one"][f] == outs["two"][f]).all(), f


def test_multi_context_argument_errors():
    with pytest.raises(capi.VrError):
        capi.Context([0, 0])  # a device listed twice
    with pytest.raises(capi.VrError):
        capi.Context([])
    one = capi.Context([0])   # one device: an ordinary context
    assert one.num_devices() == 1
    one.close()
