"""C5 (4M-disk hole array, power-cosine source n = 100, sticking 0.2, reflective boundaries):
one trace + flux post-processing on the device, timed."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
from viennaray_b200 import capi, host, scenes
rays = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
points, normals, gd = scenes.hole_array()
n = len(points); r = host.disk_radius(gd, 3)
xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
glo, ghi = host.geometry_bbox(points, 3)
lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
src = host.source_desc(lo, hi, host.POS_Z)
part = [capi.ParticleDesc(0, 0.2, 100.0, 0.0)]
if os.environ.get("VR_COUNT_WORK") == "1":
    pass
ctx = capi.Context(0)
t = time.perf_counter(); ctx.set_disks(xyzr, normals); ctx.build_neighbors_device(3, points, np.float32(2) * r)
ctx.set_boundary(lo, hi, 0, 1, 0, 0, 3); ctx.commit(); t_setup = time.perf_counter() - t
print("disks", n, "setup (upload + device neighbours + BVH) %.1f ms" % (t_setup * 1e3), ctx.bvh_stats())
for rep in range(2):
    t = time.perf_counter()
    ctx.trace_device(src, part, host.config(rays, 12346), sync=True)
    dt = time.perf_counter() - t
    _, info = ctx.flux_download()
    i = info[0]
    print("rep", rep, "rays", rays, "kernel_ms %.1f" % ctx.last_kernel_ms(), "Mrays/s %.1f" % (rays / ctx.last_kernel_ms() / 1e3),
          "traces/ray %.2f geo/ray %.2f bnd/ray %.3f" % (i.totalRaysTraced / rays, i.geometryHits / rays, i.boundaryHits / rays),
          "launches", ctx.last_launch_count())
t = time.perf_counter()
flux = ctx.flux_postprocess(0, np.full(n, np.pi * r * r, np.float32), float((hi[0]-lo[0])*(hi[1]-lo[1])) / rays, smooth=True)
print("normalise + smooth on device + download: %.1f ms; mean top flux %.3f, min %.4f" % ((time.perf_counter() - t) * 1e3, flux[points[:, 2] == 0].mean(), flux.min()))
if os.environ.get("VR_COUNT_WORK") == "1":
    print(ctx.work_counters())
