"""One trace of the C4 trench for profiling: python scripts/profile_step.py RAYS [ion|neutral|both]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from viennaray_b200 import capi, host, scenes
rays = int(float(sys.argv[1])) if len(sys.argv) > 1 else 16_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "neutral"
slices = int(sys.argv[3]) if len(sys.argv) > 3 else 999
points, normals, gd = scenes.trench(num_slices=slices)
n = len(points); r = host.disk_radius(gd, 3)
xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
off, idx = capi.build_neighbors(3, points, np.float32(2) * r)
glo, ghi = host.geometry_bbox(points, 3)
lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
src = host.source_desc(lo, hi, host.POS_Z)
parts = {"neutral": [capi.ParticleDesc(0, 0.1, 1.0, 0.0)],
         "ion": [capi.ParticleDesc(2, 0.5, 100.0, float(np.deg2rad(85.0)))]}
parts["both"] = parts["neutral"] + parts["ion"]
ctx = capi.Context(0)
ctx.set_disks(xyzr, normals, off, idx); ctx.set_boundary(lo, hi, 0, 1, 1, 1, 3); ctx.commit()
for rep in range(2):
    t = time.perf_counter()
    ctx.trace_device(src, parts[which], host.config(rays, 12346), sync=True)
    dt = time.perf_counter() - t
    k, it = ctx.last_launch_count()
    print("rep", rep, "rays", rays, which, "kernel_ms", ctx.last_kernel_ms(), "wall_ms", dt * 1e3,
          "launches", k, "iterations", it, "Mrays/s", len(parts[which]) * rays / ctx.last_kernel_ms() / 1e3)
