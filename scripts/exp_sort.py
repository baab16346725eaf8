"""Experiment: how much does ray ordering buy the traverse kernel?  Times the
production traverse kernel (vr_debug_intersect) on the C4 trench for the same
rays in generation order and sorted by the Morton code of their origin."""
import sys, time
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


def morton(org, bits=7):
    q = ((org - lo) / (hi - lo) * (2**bits - 1)).clip(0, 2**bits - 1).astype(np.uint64)
    key = np.zeros(len(org), np.uint64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return key


def timeit(rays, label):
    best = 1e9
    for _ in range(3):
        out = ctx.debug_intersect(rays, nb_cap=1)
        best = min(best, ctx.last_kernel_ms())
    print("%-44s %8.3f ms  %7.1f Mtrav/s" % (label, best, len(rays) / best / 1e3), flush=True)
    return out


rng = np.random.default_rng(3)
for name, part in (("neutral", capi.ParticleDesc(0, 0.1, 1.0, 0.0)),
                   ("ion", capi.ParticleDesc(2, 0.5, 100.0, float(np.deg2rad(85.0))))):
    rays = ctx.debug_source_rays(src, part, host.config(10**9, 12346), 0, M)
    geom, prim, t, _, _ = timeit(rays, name + " primary, generation order")
    for bits in (5, 7, 9):
        o = np.argsort(morton(rays[:, :3], bits), kind="stable")
        timeit(rays[o], name + " primary, origin Morton %d bits/axis" % bits)
    # first bounce: diffuse direction about the hit normal
    hit = geom == 1
    hp = rays[hit, :3] + rays[hit, 3:] * t[hit, None]
    nrm = normals[prim[hit]]
    u = rng.normal(size=hp.shape).astype(np.float32)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    d = nrm + u
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = np.ascontiguousarray(np.concatenate([hp, d], 1), np.float32)
    timeit(b, name + " bounce, generation order")
    for bits in (5, 7, 9):
        o = np.argsort(morton(b[:, :3], bits), kind="stable")
        timeit(b[o], name + " bounce, origin Morton %d bits/axis" % bits)
    o = np.argsort(morton(b[:, :3], 7) * np.uint64(8) +
                   ((d[:, 0] > 0) * 4 + (d[:, 1] > 0) * 2 + (d[:, 2] > 0)).astype(np.uint64),
                   kind="stable")
    timeit(b[o], name + " bounce, origin Morton 7 + octant")
