"""Per-ray work counters and phase times of the C5 hole array (VR_COUNT_WORK / VR_TIME_KERNELS)."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
os.environ["VR_COUNT_WORK"] = "1"
os.environ["VR_TIME_KERNELS"] = "1"
from viennaray_b200 import capi, host, scenes
rays = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
points, normals, gd = scenes.hole_array()
n = len(points); r = host.disk_radius(gd, 3)
xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
glo, ghi = host.geometry_bbox(points, 3)
lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
src = host.source_desc(lo, hi, host.POS_Z)
ctx = capi.Context(0)
ctx.set_disks(xyzr, normals); ctx.build_neighbors_device(3, points, np.float32(2) * r)
ctx.set_boundary(lo, hi, 0, 1, 0, 0, 3); ctx.commit()
print(ctx.bvh_stats())
p = capi.ParticleDesc(0, 0.2, 100.0, 0.0)
for rep in range(2):
    ctx.phase_timing(True)
    ctx.trace_device(src, [p], host.config(rays, 12346), sync=True)
    w = ctx.work_counters()
    ph = ctx.phase_ms()
    _, info = ctx.flux_download()
    i = info[0]
    print({k: round(v / rays, 3) for k, v in w.items()},
          "traces/ray %.3f geo %.3f miss %.3f" % (i.totalRaysTraced / rays, i.geometryHits / rays, i.nonGeometryHits / rays),
          "nodes/traversal %.2f prims/traversal %.2f" % (w["node_visits"] / (i.totalRaysTraced - w["sky_finished"]),
                                                         w["prim_tests"] / (i.totalRaysTraced - w["sky_finished"])), ph, "kernel_ms", ctx.last_kernel_ms())
