"""Experiment: how much faster is the traverse kernel when a warp's rays are of one kind?  Times the
production traverse kernel (vr_debug_intersect) on the C4 trench for primary rays alone, for bounce
rays alone, and for the two shuffled together (what the in-place wavefront pool looks like)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from viennaray_b200 import capi, host, scenes

M = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
points, normals, gd = scenes.trench()
n = len(points); r = host.disk_radius(gd, 3)
xyzr = np.concatenate([points, np.full((n, 1), r, np.float32)], 1)
off, idx = capi.build_neighbors(3, points, np.float32(2) * r)
glo, ghi = host.geometry_bbox(points, 3)
lo, hi = host.adjust_bbox(glo, ghi, host.POS_Z, r, 3)
src = host.source_desc(lo, hi, host.POS_Z)
ctx = capi.Context(0)
ctx.set_disks(xyzr, normals, off, idx); ctx.set_boundary(lo, hi, 0, 1, 1, 1, 3); ctx.commit()
part = capi.ParticleDesc(0, 0.1, 1.0, 0.0)
ctx.trace_device(src, [part], host.config(100000, 1), sync=True)  # builds the sky map (and the entry table)


def timeit(rays, label):
    best = 1e9
    for _ in range(3):
        out = ctx.debug_intersect(rays, nb_cap=1)
        best = min(best, ctx.last_kernel_ms())
    print("%-40s %9d rays %8.3f ms  %7.1f Mtrav/s" % (label, len(rays), best, len(rays) / best / 1e3), flush=True)
    return out, best


rng = np.random.default_rng(3)
prim_rays = ctx.debug_source_rays(src, part, host.config(10**9, 12346), 0, M)
(geom, prim, t, _, _), tp = timeit(prim_rays, "primary rays alone")
# bounce rays as the wavefront holds them: diffuse directions off the hit points, those that do not
# leave towards the sky at once (the shade kernel finishes these itself) -- keep downward / grazing
gen = prim_rays
bounces = []
for leg in range(3):
    hit = geom == 1
    hp = gen[hit, :3] + gen[hit, 3:] * t[hit, None]
    nrm = normals[prim[hit]]
    u = rng.normal(size=hp.shape).astype(np.float32)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    d = nrm + u
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = np.ascontiguousarray(np.concatenate([hp, d], 1), np.float32)
    inside = hp[:, 2] < -0.5  # in the trench: the rays the sky map cannot finish
    b = b[inside]
    geom, prim, t, _, _ = ctx.debug_intersect(b, nb_cap=1)
    bounces.append(b)
    gen = b
bo = np.concatenate(bounces)[: int(0.7 * M)]
_, tb = timeit(bo, "bounce rays alone (trench)")
mix = np.concatenate([prim_rays, bo])
rng.shuffle(mix)
_, tm = timeit(mix, "shuffled together")
print("separate %.3f ms, together %.3f ms: %.1f %%" % (tp + tb, tm, 100 * (tm / (tp + tb) - 1)))
