import sys
import numpy as np
sys.path.insert(0, '.')
from viennaray_b200 import capi, host, scenes
for name, gen, bc in (("C4 trench", scenes.trench, 1), ("C5 holes", scenes.hole_array, 0)):
    points, normals, gd = gen()
    n = len(points); r = host.disk_radius(gd, 3)
    xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
    glo, ghi = host.geometry_bbox(points, 3)
    lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
    ctx = capi.Context(0)
    ctx.set_disks(xyzr, normals); ctx.set_boundary(lo, hi, 0, 1, bc, bc, 3); ctx.commit()
    s = ctx.bvh_stats()
    print(name, "sah_inner %.2f sah_leaf %.2f nodes %d build %.2f ms alpha %.2f" % (s["sah_inner"], s["sah_leaf"], s["nodes"], s["build_ms"], s["morton_alpha"]))
    ctx.close()
