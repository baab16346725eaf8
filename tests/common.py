"""Shared set-up for the parity tests: the five BASELINE configs (C4/C5 at
reduced size unless asked otherwise) as plain dicts, and builders that feed the
SAME inputs to the CPU oracle and to the CUDA library through its C ABI."""
import os

import numpy as np

from oracle import pyoracle as po
from viennaray_b200 import capi, host, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_inputs = None


def inputs():
    global _inputs
    if _inputs is None:
        _inputs = dict(np.load(os.path.join(GOLDEN, "inputs.npz")))
    return _inputs


def case(name):
    I = inputs()
    if name == "disk3D":  # C1, examples/disk3D/disk3D.cpp
        return dict(name=name, D=3, geo="disk", points=I["disk3D_points"],
                    normals=I["disk3D_normals"], grid_delta=float(I["disk3D_gridDelta"]),
                    bc=[1, 1, 1], source_dir=host.POS_Z, kind=0, sticking=0.1, power=1.0,
                    cone=0.0)
    if name == "triangle3D":  # C2, examples/triangle3D/triangle3D.cpp
        return dict(name=name, D=3, geo="triangle", verts=I["triangle3D_nodes"],
                    tris=I["triangle3D_triangles"],
                    grid_delta=float(I["triangle3D_gridDelta"]), bc=[0, 0, 0],
                    source_dir=host.POS_Z, kind=0, sticking=0.1, power=1.0, cone=0.0)
    if name == "disk2D":  # C3, examples/disk2D/disk2D.cpp
        return dict(name=name, D=2, geo="disk", points=I["disk2D_points"],
                    normals=I["disk2D_normals"], grid_delta=float(I["disk2D_gridDelta"]),
                    bc=[1, 1, 1], source_dir=host.POS_Y, kind=0, sticking=0.1, power=1.0,
                    cone=0.0)
    if name.startswith("trench"):  # C4 (reduced unless "trench_full")
        full = name == "trench_full"
        p, n, gd = scenes.trench() if full else scenes.trench(num_slices=60, half_width=10,
                                                              depth=40, half_extent=40)
        kind, st, pw, cone = (2, 0.5, 100.0, np.deg2rad(85.0)) if name.endswith("ion") else \
            (0, 0.1, 1.0, 0.0)
        return dict(name=name, D=3, geo="disk", points=p, normals=n, grid_delta=gd, bc=[1, 1, 1],
                    source_dir=host.POS_Z, kind=kind, sticking=st, power=pw, cone=cone)
    if name.startswith("holes"):  # C5 (reduced unless "holes_full")
        p, n, gd = scenes.hole_array() if name == "holes_full" else \
            scenes.hole_array(cells=2, pitch=40, radius=8, depth=48)
        return dict(name=name, D=3, geo="disk", points=p, normals=n, grid_delta=gd, bc=[0, 0, 0],
                    source_dir=host.POS_Z, kind=0, sticking=0.2, power=100.0, cone=0.0)
    if name in ("sphere3D", "sphere2D"):  # tests/Resources/sphereGrid{3D,2D}_R1.dat (createRay etc.)
        D = 3 if name == "sphere3D" else 2
        return dict(name=name, D=D, geo="disk", points=I[name + "_points"],
                    normals=I[name + "_normals"], grid_delta=float(I[name + "_gridDelta"]),
                    bc=[0, 1, 0] if D == 3 else [1, 1, 1],
                    source_dir=host.POS_Z if D == 3 else host.POS_Y, kind=2, sticking=0.3,
                    power=4.0, cone=float(np.deg2rad(60.0)))
    if name == "plane":  # tests/rngSeed, tests/traceInterface geometry
        p, n = scenes.plane_grid(0.5, 5.0)
        return dict(name=name, D=3, geo="disk", points=p, normals=n, grid_delta=0.5, bc=[0, 0, 0],
                    source_dir=host.POS_Z, kind=0, sticking=1.0, power=1.0, cone=0.0)
    raise KeyError(name)


def source_offset(c):
    return host.disk_radius(c["grid_delta"], c["D"]) if c["geo"] == "disk" else \
        np.float32(c["grid_delta"])


def make_oracle(c):
    s = po.OracleScene(c["D"])
    if c["geo"] == "disk":
        s.set_disks(c["points"], c["normals"], host.disk_radius(c["grid_delta"], c["D"]))
    else:
        s.set_triangles(c["verts"], c["tris"])
    s.setup(c["source_dir"], c["bc"], source_offset(c))
    return s


def oracle_particle(c):
    return po.Particle(c["kind"], c["sticking"], c["power"], c["cone"])


def gpu_particle(c):
    return capi.ParticleDesc(c["kind"], c["sticking"], c["power"], c["cone"])


def product_setup(c):
    """Host-side set-up done by the PRODUCT (no oracle involved): returns the
    arrays handed to the C ABI."""
    D = c["D"]
    out = {}
    if c["geo"] == "disk":
        pts = np.ascontiguousarray(c["points"], np.float32).copy()
        nrm = np.ascontiguousarray(c["normals"], np.float32).copy()
        r = host.disk_radius(c["grid_delta"], D)
        off, idx = capi.build_neighbors(D, pts, np.float32(2) * r)
        if D == 2:
            pts[:, 2] = 0
            nrm[:, 2] = 0
        out["xyzr"] = np.concatenate([pts, np.full((len(pts), 1), r, np.float32)], 1)
        out["normals"] = nrm
        out["nb"] = (off, idx)
        glo, ghi = host.geometry_bbox(pts, D)
    else:
        v = np.ascontiguousarray(c["verts"], np.float32)
        out["verts"], out["tris"] = v, np.ascontiguousarray(c["tris"], np.uint32)
        out["normals"] = host.triangle_normals(v, c["tris"])
        glo, ghi = v.min(0), v.max(0)
    lo, hi = host.adjust_bbox(glo, ghi, c["source_dir"], source_offset(c), D)
    out["bbox"] = (lo, hi)
    return out


def make_gpu(c, device=0, primary_dir=None, material_ids=None):
    st = product_setup(c)
    ctx = capi.Context(device)
    if c["geo"] == "disk":
        ctx.set_disks(st["xyzr"], st["normals"], st["nb"][0], st["nb"][1],
                      material_ids=material_ids)
    else:
        ctx.set_triangles(st["verts"], st["tris"], st["normals"], material_ids=material_ids)
    lo, hi = st["bbox"]
    _, first, second, _, _ = host.trace_settings(c["source_dir"])
    cond2 = c["bc"][second] if c["D"] == 3 else capi.BOUNDARY_IGNORE
    ctx.set_boundary(lo, hi, first, second, c["bc"][first], cond2, c["D"])
    ctx.commit()
    src = host.source_desc(lo, hi, c["source_dir"], primary_dir)
    return ctx, src, st
