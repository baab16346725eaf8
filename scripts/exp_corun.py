"""Experiment: do the traverse and the shade kernel of two independent traces overlap on one GPU when
each is capped to part of an SM (VR_TRAV_CAP blocks/SM, VR_SHADE_SMEM_PAD bytes: knobs of commits a29880a..1ab9da5, removed since)?  Runs the C4 step
on one context, then on two contexts from two host threads, and prints the aggregate rates."""
import sys, time, threading
import numpy as np
sys.path.insert(0, '.')
from viennaray_b200 import capi, host, scenes
rays = int(float(sys.argv[1])) if len(sys.argv) > 1 else 128_000_000
nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
points, normals, gd = scenes.trench()
n = len(points); r = host.disk_radius(gd, 3)
xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
off, idx = capi.build_neighbors(3, points, np.float32(2) * r)
glo, ghi = host.geometry_bbox(points, 3)
lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
src = host.source_desc(lo, hi, host.POS_Z)
parts = [capi.ParticleDesc(0, 0.1, 1.0, 0.0), capi.ParticleDesc(2, 0.5, 100.0, float(np.deg2rad(85.0)))]
ctxs = []
for k in range(nctx):
    c = capi.Context(0)
    c.set_disks(xyzr, normals, off, idx); c.set_boundary(lo, hi, 0, 1, 1, 1, 3); c.commit()
    ctxs.append(c)
def run(c, k):
    c.trace_device(src, parts, host.config(rays, 12346 + k), sync=True)
for rep in range(3):
    t = time.perf_counter()
    th = [threading.Thread(target=run, args=(c, k)) for k, c in enumerate(ctxs)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.perf_counter() - t
    print("rep", rep, "contexts", nctx, "rays each", rays, "wall_ms %.1f" % (dt * 1e3),
          "aggregate Mrays/s %.1f" % (nctx * 2 * rays / dt / 1e6), flush=True)
