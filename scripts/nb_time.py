import sys, time
import numpy as np
sys.path.insert(0, '.')
from viennaray_b200 import capi, host, scenes
for name, gen in (("trench 1M", scenes.trench), ("holes 4M", scenes.hole_array)):
    points, normals, gd = gen()
    n = len(points); r = host.disk_radius(gd, 3)
    xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
    ctx = capi.Context(0)
    ctx.set_disks(xyzr, normals)
    for rep in range(2):
        t = time.perf_counter(); ctx.build_neighbors_device(3, points, np.float32(2) * r); td = time.perf_counter() - t
    t = time.perf_counter(); off, idx = capi.build_neighbors(3, points, np.float32(2) * r); th = time.perf_counter() - t
    o2, i2 = ctx.get_neighbors()
    print(name, n, "device %.1f ms  host %.1f ms  equal %s  mean row %.2f" % (td * 1e3, th * 1e3, bool((o2 == off).all() and (i2 == idx).all()), len(idx) / n))
    ctx.close()
