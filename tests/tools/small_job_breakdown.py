"""Where the wall time of a small job goes (C1-C3): create / upload + commit / trace / close,
and the steady state of a reused context (what a time-stepping caller sees)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import common
from viennaray_b200 import capi, host
for name in ("disk2D", "triangle3D", "disk3D"):
    c = common.case(name)
    st = common.product_setup(c)
    n = len(c["points"]) if c["geo"] == "disk" else len(c["tris"])
    rays = 2000 * n
    lo, hi = st["bbox"]
    _, first, second, _, _ = host.trace_settings(c["source_dir"])
    cond2 = c["bc"][second] if c["D"] == 3 else capi.BOUNDARY_IGNORE
    src = host.source_desc(lo, hi, c["source_dir"])
    part = [common.gpu_particle(c)]
    t0 = time.perf_counter()
    ctx = capi.Context(0)
    t1 = time.perf_counter()
    print("%-11s create %.1f ms" % (name, (t1 - t0) * 1e3))
    for rep in range(4):
        t1 = time.perf_counter()
        if c["geo"] == "disk":
            ctx.set_disks(st["xyzr"], st["normals"], st["nb"][0], st["nb"][1])
        else:
            ctx.set_triangles(st["verts"], st["tris"], st["normals"])
        ctx.set_boundary(lo, hi, first, second, c["bc"][first], cond2, c["D"])
        ctx.commit()
        ctx.synchronize()
        t2 = time.perf_counter()
        ctx.trace_device(src, part, host.config(rays, 12345), sync=True)
        t3 = time.perf_counter()
        flux, info = ctx.flux_download()
        t4 = time.perf_counter()
        kms = ctx.last_kernel_ms(); lc = ctx.last_launch_count()
        i = info[0]
        print("   step %d: upload+commit %.2f ms, trace_device %.2f ms (kernels %.2f ms, launches %s), flux download %.2f ms; traces/ray %.2f"
              % (rep, (t2 - t1) * 1e3, (t3 - t2) * 1e3, kms, lc, (t4 - t3) * 1e3, i.totalRaysTraced / rays), flush=True)
    t0 = time.perf_counter()
    ctx.close()
    print("   close %.1f ms" % ((time.perf_counter() - t0) * 1e3))
