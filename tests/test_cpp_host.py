"""The C++ host mirror of the reference interface (include/viennaray_b200/):
compiled with g++ against the C-ABI library and run the way the reference runs
its own tests.  CPU part: data containers, particles, neighbourhoods, disk
areas, smoothing, error behaviour without a device.  GPU part (-m gpu): traces
through TraceDisk / TraceTriangle and compares the flux bit-for-bit with the
same run through the ctypes binding and the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import common
from viennaray_b200 import capi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "viennaray_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "test_host_api")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include", "viennaray_b200"),
                    os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"), "-L", LIBDIR,
                    "-lviennaray_b200", "-Wl,-rpath," + LIBDIR, "-o", out], check=True)
    return out


def test_cpp_host_cpu_blocks(exe):
    res = subprocess.run([exe, "cpu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "test_host_api cpu: ok" in res.stdout
    # the device-less apply() reported an error instead of tracing on the host
    assert "no CPU path" in res.stderr or "ERROR" in res.stderr


def test_cpp_disk_areas_equal_reference(exe, tmp_path):
    if not po.have_ref():
        pytest.skip("oracle/_ref not built")
    c = common.case("disk3D")
    pts = np.ascontiguousarray(c["points"], np.float32)
    nrm = np.ascontiguousarray(c["normals"], np.float32)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "areas.f64")
    with open(inp, "wb") as f:
        f.write(struct.pack("<If", len(pts), c["grid_delta"]))
        f.write(pts.tobytes())
        f.write(nrm.tobytes())
    subprocess.run([exe, "areas", inp, out], check=True)
    mine = np.fromfile(out, np.float64)
    ref = po.ref_disk_areas(3, pts, nrm, c["grid_delta"], [1, 1, 1], po.POS_Z)
    assert len(mine) == len(ref)
    assert np.allclose(mine, ref, rtol=2e-5, atol=1e-6)
    r = host.disk_radius(c["grid_delta"], 3)
    assert (ref < 0.99 * np.pi * r * r).sum() > 100  # the case does clip disks at the box


@pytest.mark.gpu
def test_cpp_host_traces_match_c_abi_path(exe, tmp_path):
    res = subprocess.run([exe, "gpu", str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr + res.stdout
    c = common.case("plane")

    def via_capi(sticking, bc, seed, max_bhits=1000):
        cc = dict(c, sticking=sticking, bc=bc)
        ctx, src, _ = common.make_gpu(cc)
        flux, info = ctx.trace(src, [common.gpu_particle(cc)],
                               host.config(4410, seed, max_boundary_hits=max_bhits))
        ctx.close()
        return flux[0].astype(np.float32)

    f = np.fromfile(str(tmp_path / "rngSeed_flux.f32"), np.float32)
    assert (f == via_capi(1.0, [0, 0, 0], 12345 + 1)).all()
    f = np.fromfile(str(tmp_path / "traceInterface_raw.f32"), np.float32)
    assert (f == via_capi(0.5, [0, 0, 0], 0 + 1, 10)).all()
    f = np.fromfile(str(tmp_path / "traceInterface_run2.f32"), np.float32)
    assert (f == via_capi(0.5, [0, 0, 0], 0 + 2, 10)).all()
    # and the same numbers from the oracle
    cc = dict(c, sticking=0.5)
    orc = common.make_oracle(cc)
    fo, _ = orc.trace(common.oracle_particle(cc), orc.config(4410, 1, max_boundary_hits=10))
    f = np.fromfile(str(tmp_path / "traceInterface_raw.f32"), np.float32)
    assert (f == (fo / po.FLUX_SCALE).astype(np.float32)).all()
