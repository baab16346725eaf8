"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol that
include/*.h declares, its host-side helpers agree with the oracle, argument
errors are reported through status codes, and -- without a CUDA device -- the
product refuses to run instead of falling back to anything."""
import ctypes as C
import glob
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import common
from viennaray_b200 import capi, host, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def _declared_symbols():
    names = set()
    for path in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names |= set(re.findall(r"\b(vr_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(capi.EXPORTS) == declared  # the ctypes binding covers the whole header


def test_library_is_sm100a_cuda_and_has_no_oracle_dependency():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout
    needed = subprocess.run(["readelf", "-d", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "vr_ref" not in needed
    syms = subprocess.run(["nm", "-D", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "vro_" not in syms


def test_struct_layouts_match_header():
    # sizes the C compiler gives the header's structs
    src = r'''
#include <stdio.h>
#include "viennaray_b200.h"
int main(){printf("%zu %zu %zu %zu\n", sizeof(vr_source_desc), sizeof(vr_particle_desc),
sizeof(vr_config), sizeof(vr_trace_info));return 0;}'''
    exe = "/tmp/vr_sizes_%d" % os.getpid()
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src, text=True, check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    os.remove(exe)
    assert sizes == [C.sizeof(capi.SourceDesc), C.sizeof(capi.ParticleDesc),
                     C.sizeof(capi.Config), C.sizeof(capi.TraceInfo)]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.VrError) as e:
        capi.Context(0)
    assert e.value.code == 1  # VR_ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_null_context_arguments_are_rejected():
    L = capi.lib()
    assert L.vr_scene_commit(None) == 2  # VR_ERR_ARGUMENT
    assert L.vr_trace(None, None, None, 0, None, None, None) == 2
    assert L.vr_ctx_create(0, None) == 2
    L.vr_ctx_destroy(None)  # no-op


@pytest.mark.parametrize("name", ["disk3D", "disk2D", "trench", "holes", "plane"])
def test_build_neighbors_equals_oracle(name):
    c = common.case(name)
    s = common.make_oracle(c)
    off_o, idx_o = s.neighbors()
    r = host.disk_radius(c["grid_delta"], c["D"])
    off, idx = capi.build_neighbors(c["D"], c["points"], F(2) * r)
    assert (off == off_o).all() and (idx == idx_o).all()
    # symmetric relation, no self references, rows ascending
    rows = np.repeat(np.arange(len(off) - 1), np.diff(off))
    assert (rows != idx).all()
    pairs = set(zip(rows.tolist()[:5000], idx.tolist()[:5000]))
    all_pairs = set(zip(rows.tolist(), idx.tolist()))
    assert all((b, a) in all_pairs for a, b in pairs)
    for i in range(0, len(off) - 1, 997):
        row = idx[off[i]:off[i + 1]]
        assert (np.diff(row.astype(np.int64)) > 0).all()


def test_build_neighbors_edge_cases():
    off, idx = capi.build_neighbors(3, np.zeros((0, 3), F), 1.0)
    assert list(off) == [0] and len(idx) == 0
    off, idx = capi.build_neighbors(3, np.zeros((1, 3), F), 1.0)
    assert list(off) == [0, 0]
    # coincident points are each other's neighbours
    off, idx = capi.build_neighbors(3, np.zeros((3, 3), F), 0.5)
    assert list(off) == [0, 2, 4, 6] and list(idx) == [1, 2, 0, 2, 0, 1]
    # per-axis AND euclidean test (rayPointNeighborhood.hpp:287-298)
    pts = np.array([[0, 0, 0], [1, 0, 0], [0.8, 0.8, 0], [0, 0, 1.0000001]], F)
    off, idx = capi.build_neighbors(3, pts, 1.0)
    assert list(idx[off[0]:off[1]]) == [1]


@pytest.mark.parametrize("name", ["disk3D", "triangle3D", "disk2D", "trench", "holes"])
def test_host_setup_matches_oracle(name):
    c = common.case(name)
    s = common.make_oracle(c)
    st = common.product_setup(c)
    lo, hi = st["bbox"]
    assert (s.bbox() == np.stack([lo, hi])).all()
    if c["geo"] == "triangle":
        assert (s.normals().view(np.uint32) == st["normals"].view(np.uint32)).all()


def test_trace_settings_and_bbox_rules():
    # rayUtil.hpp:145-202 and :104-143
    assert host.trace_settings(host.POS_Z) == (2, 0, 1, 1, -1)
    assert host.trace_settings(host.NEG_X) == (0, 1, 2, 0, 1)
    lo, hi = host.adjust_bbox(np.array([-1, -1, -1], F), np.array([1, 1, 1], F), host.POS_Z,
                              0.25, 3)
    assert list(lo) == [-1, -1, -1] and list(hi) == [1, 1, 1.5]
    lo, hi = host.adjust_bbox(np.array([-1, -1, 0], F), np.array([1, 1, 0], F), host.NEG_Y,
                              0.25, 2)
    assert list(lo) == [-1, -1.5, -0.25] and list(hi) == [1, 1, 0.25]
    with pytest.raises(ValueError):
        host.adjust_bbox(lo, hi, host.POS_Z, 0.25, 2)
    # tests/buildBoundary: source plane sits 2 * offset above the geometry
    pts, _ = scenes.plane_grid(0.5, 1.0)
    glo, ghi = host.geometry_bbox(pts, 3)
    lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, 0.5, 3)
    assert hi[2] == 1.0 and lo[2] == 0.0 and lo[0] == -1 and hi[0] == 1


def test_orthonormal_basis():
    b = host.orthonormal_basis([1.0, 1.0, -1.0])
    assert np.allclose(b @ b.T, np.eye(3), atol=1e-6)
    assert np.allclose(b[0], np.array([1, 1, -1]) / np.sqrt(3), atol=1e-6)


def test_scene_generators_sizes():
    p, n, gd = scenes.trench(num_slices=3)
    assert len(p) == 3 * 1001 and gd == 1.0
    assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-6)
    p, n, gd = scenes.hole_array(cells=1, pitch=20, radius=5, depth=6)
    assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-6)
    assert p[:, 2].min() == -6
