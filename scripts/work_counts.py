"""Per-ray work counters (node visits, primitive tests, ...) of the C4 trench."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
os.environ["VR_COUNT_WORK"] = "1"
from viennaray_b200 import capi, host, scenes
rays = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
points, normals, gd = scenes.trench()
n = len(points); r = host.disk_radius(gd, 3)
xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
off, idx = capi.build_neighbors(3, points, np.float32(2) * r)
glo, ghi = host.geometry_bbox(points, 3)
lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
src = host.source_desc(lo, hi, host.POS_Z)
ctx = capi.Context(0)
ctx.set_disks(xyzr, normals, off, idx); ctx.set_boundary(lo, hi, 0, 1, 1, 1, 3); ctx.commit()
print(ctx.bvh_stats())
for name, p in (("neutral", capi.ParticleDesc(0, 0.1, 1.0, 0.0)),
                ("ion", capi.ParticleDesc(2, 0.5, 100.0, float(np.deg2rad(85.0))))):
    ctx.trace_device(src, [p], host.config(rays, 12346), sync=True)
    w = ctx.work_counters()
    _, info = ctx.flux_download()
    i = info[0]
    print(name, {k: round(v / rays, 3) for k, v in w.items()},
          "traces/ray %.3f geo %.3f miss %.3f" % (i.totalRaysTraced / rays, i.geometryHits / rays,
                                                 i.nonGeometryHits / rays),
          "sky_finished/ray %.3f" % (w["sky_finished"] / rays), "nodes/trace %.2f prims/trace %.2f" % (w["node_visits"] / i.totalRaysTraced,
                                                w["prim_tests"] / i.totalRaysTraced))
