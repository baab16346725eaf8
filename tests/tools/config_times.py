"""All five BASELINE.json configs through the host-buffer C ABI (upload + device commit +
vr_trace + flux on the host inside the timing), with the reference's CPU kernel
(oracle/_ref: unmodified TraceKernel + substitute intersector) timed beside it on a bounded
sample.  C4 is what bench.py measures; it is included here at 1e8 rays per particle so that the
table has one column.  Output is committed as profiles/r1_all_configs.txt."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle as po  # noqa: E402   (tests/ infrastructure: the CPU reference arm of a comparison)
from tests import common  # noqa: E402
from viennaray_b200 import capi, host  # noqa: E402

CPU_SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
CONFIGS = [("C1 disk3D", "disk3D", None, 2000), ("C2 triangle3D", "triangle3D", None, 2000),
           ("C3 disk2D", "disk2D", None, 2000), ("C4 trench 1M (neutral)", "trench_full", int(1e8), 0),
           ("C5 holes 4M", "holes_full", int(2e8), 0)]
print("%-24s %10s %12s %10s %10s %14s %14s %8s" % ("config", "prims", "rays", "first ms", "step ms",
                                                   "gpu rays/s", "cpu rays/s", "ratio"))
for label, name, fixed, per_point in CONFIGS:
    c = common.case(name)
    st = common.product_setup(c)
    n = len(c["points"]) if c["geo"] == "disk" else len(c["tris"])
    rays = fixed if fixed else per_point * n
    lo, hi = st["bbox"]
    _, first, second, _, _ = host.trace_settings(c["source_dir"])
    cond2 = c["bc"][second] if c["D"] == 3 else capi.BOUNDARY_IGNORE
    src = host.source_desc(lo, hi, c["source_dir"])
    part = [common.gpu_particle(c)]
    # one context reused over the steps, as a time-stepping caller (ViennaPS) does: every
    # step uploads the geometry again, commits (device BVH build), traces, reads the flux
    ctx = capi.Context(0)
    times = []
    for rep in range(4):
        t = time.perf_counter()
        if c["geo"] == "disk":
            ctx.set_disks(st["xyzr"], st["normals"], st["nb"][0], st["nb"][1])
        else:
            ctx.set_triangles(st["verts"], st["tris"], st["normals"])
        ctx.set_boundary(lo, hi, first, second, c["bc"][first], cond2, c["D"])
        ctx.commit()
        flux, info = ctx.trace(src, part, host.config(rays, 12345))
        times.append(time.perf_counter() - t)
    ctx.close()
    t_first, best = times[0], min(times[1:])
    # CPU reference on a bounded sample: slope of two runs (the scene build is inside its timer)
    cpu = float("nan")
    if po.have_ref():
        def run(m):
            if c["geo"] == "disk":
                return po.ref_trace_disk(c["D"], c["points"], c["normals"], c["grid_delta"], c["bc"],
                                         c["source_dir"], c["kind"], c["sticking"], c["power"],
                                         c["cone"], rays_fixed=m, seed=12345)[2]
            return po.ref_trace_triangle(c["verts"], c["tris"], c["grid_delta"], c["bc"],
                                         c["source_dir"], c["kind"], c["sticking"], c["power"],
                                         c["cone"], rays_fixed=m, seed=12345)[2]
        m1 = 100_000
        t1, t2 = run(m1), run(3 * m1)
        r = 2 * m1 / max(t2 - t1, 1e-9)
        m3 = int(min(max(r * CPU_SECONDS, 4 * m1), rays))
        if m3 > 3 * m1:
            t3 = run(m3)
            r = (m3 - m1) / max(t3 - t1, 1e-9)
        cpu = r
    g = rays / best
    print("%-24s %10d %12d %10.1f %10.2f %14.4g %14.4g %8.0f" % (label, n, rays, t_first * 1e3, best * 1e3, g,
                                                          cpu, g / cpu), flush=True)
print("step ms: best of 3 steps on a reused context, each = upload + device commit (BVH) + vr_trace "
      "+ flux to host, one particle; first ms: the first step, which also allocates the ray pools "
      "and loads the kernels; cpu: %d host threads, reference TraceKernel + substitute intersector, "
      "bounded sample" % os.cpu_count())
