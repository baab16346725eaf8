"""Wall-clock breakdown of the host-buffer path: set_disks / commit / trace / download."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from viennaray_b200 import capi, host, scenes
rays = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
points, normals, gd = scenes.trench()
n = len(points); r = host.disk_radius(gd, 3)
xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
t = time.perf_counter(); off, idx = capi.build_neighbors(3, points, np.float32(2) * r)
print("build_neighbors %.1f ms" % ((time.perf_counter() - t) * 1e3))
glo, ghi = host.geometry_bbox(points, 3)
lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
src = host.source_desc(lo, hi, host.POS_Z)
parts = [capi.ParticleDesc(0, 0.1, 1.0, 0.0), capi.ParticleDesc(2, 0.5, 100.0, float(np.deg2rad(85.0)))]
ctx = capi.Context(0)
for rep in range(3):
    t0 = time.perf_counter(); ctx.set_disks(xyzr, normals, off, idx)
    t1 = time.perf_counter(); ctx.set_boundary(lo, hi, 0, 1, 1, 1, 3); ctx.commit()
    t2 = time.perf_counter(); ctx.trace_device(src, parts, host.config(rays, 12346), sync=True)
    t3 = time.perf_counter(); flux, info = ctx.flux_download()
    t4 = time.perf_counter()
    print("rep %d set_disks %.1f  commit %.1f (bvh %.1f)  trace %.1f (kernel %.1f)  download %.1f ms" %
          (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ctx.bvh_stats()["build_ms"], (t3 - t2) * 1e3,
           ctx.last_kernel_ms(), (t4 - t3) * 1e3))
